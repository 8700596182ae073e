"""losses.py:65-89 reconstruction_loss on the GPU (fac_reconstruction_loss through facodec_b200.losses) against the
committed fixture of the imported reference and against the oracle restatement run live on this box's CPU.
Tolerance: 2e-5 relative on the scalar and on each of its 13 components (fp32 sums in different orders; the DFT is the
fp32-faithful 3-pass tensor-core class)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
REL = 2e-5


def test_reconstruction_loss_vs_golden(built_lib):
    from facodec_b200 import losses, synth
    gold = np.load(os.path.join(ROOT, "tests", "golden", "recon_loss.npz"))
    x, G_x = synth.synth_loss_pair(int(gold["B"]), int(gold["T"]), int(gold["seed"]))
    L, terms = losses.reconstruction_loss(x.cuda(), G_x.cuda(), return_terms=True)
    torch.cuda.synchronize()
    assert L.dim() == 0
    assert abs(float(L) - float(gold["loss"])) <= REL * abs(float(gold["loss"]))
    t = terms.cpu().numpy()
    assert t.shape == (13,)
    assert np.all(np.abs(t - gold["terms"]) <= REL * np.abs(gold["terms"])), (t, gold["terms"])


@pytest.mark.parametrize("B,T", [(4, 96000), (1, 30011), (3, 1025)])
def test_reconstruction_loss_vs_oracle(built_lib, B, T):
    """4 s utterances (the benchmark shape), a ragged length, and the shortest legal length (T = 1025 > reflect pad 1024)."""
    import warnings
    warnings.simplefilter("ignore")
    from facodec_b200 import losses, synth
    from oracle import facodec_oracle as O
    x, G_x = synth.synth_loss_pair(B, T, seed=5)
    L, terms = losses.reconstruction_loss(x.cuda(), G_x.cuda(), return_terms=True)
    L2 = losses.reconstruction_loss(x[:, 0].cuda(), G_x[:, 0].cuda())          # [B, T] form, second call (arena reuse)
    torch.cuda.synchronize()
    with torch.no_grad():
        Lo, to = O.reconstruction_loss(x, G_x, return_terms=True)
    assert torch.equal(L, L2)
    assert abs(float(L) - float(Lo)) <= REL * abs(float(Lo))
    assert torch.all((terms.cpu() - to).abs() <= REL * to.abs()), (terms.cpu(), to)


def test_reconstruction_loss_error_paths(built_lib):
    import facodec_b200 as fb
    from facodec_b200 import losses, synth
    x, G_x = synth.synth_loss_pair(1, 2000, seed=1)
    with pytest.raises(fb.FacError):
        losses.reconstruction_loss(x, G_x.cuda())                      # CPU tensor: no fallback
    with pytest.raises(fb.FacError):
        losses.reconstruction_loss(x.cuda()[..., :1024], G_x.cuda()[..., :1024])   # not longer than the reflect padding
    with pytest.raises(fb.FacError):
        losses.reconstruction_loss(x.cuda(), G_x.cuda()[..., :1500])


@pytest.mark.parametrize("B,T", [(4, 96000), (2, 5001)])
def test_dac_spectral_losses_vs_oracle(B, T, built_lib):
    """dac/nn/loss.py criteria as train.py:153-164 builds them -- MultiScaleSTFTLoss(), MelSpectrogramLoss(7 scales, pow = 1,
    mag_weight = 0), L1Loss() -- against the oracle's restatement (parity unpinned against audiotools itself: not vendored)."""
    import warnings
    warnings.simplefilter("ignore")
    from facodec_b200 import losses, synth
    from oracle import facodec_oracle as O
    x, y = synth.synth_loss_pair(B, T, seed=9)
    xd, yd = x.cuda(), y.cuda()
    stft = losses.MultiScaleSTFTLoss()
    mel = losses.MelSpectrogramLoss(n_mels=[5, 10, 20, 40, 80, 160, 320], window_lengths=[32, 64, 128, 256, 512, 1024, 2048],
                                    mel_fmin=[0] * 7, mel_fmax=[None] * 7, pow=1.0, mag_weight=0.0, clamp_eps=1e-5)
    mel2 = losses.MelSpectrogramLoss()                                   # defaults: [150, 80] over [2048, 512], pow 2, both terms
    l1 = losses.L1Loss()
    got = [stft(xd, yd), mel(xd, yd), mel2(xd, yd), l1(xd, yd), stft(xd, yd)]
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = [O.multiscale_stft_loss(x, y),
               O.mel_spectrogram_loss(x, y, 24000, n_mels=[5, 10, 20, 40, 80, 160, 320], window_lengths=[32, 64, 128, 256, 512, 1024, 2048],
                                      mel_fmin=[0] * 7, mel_fmax=[None] * 7, pow=1.0, mag_weight=0.0),
               O.mel_spectrogram_loss(x, y, 24000),
               (x - y).abs().mean(),
               O.multiscale_stft_loss(x, y)]
    for name, a, b in zip(("stft", "mel_train", "mel_default", "l1", "stft_again"), got, ref):
        assert a.dim() == 0
        assert abs(float(a) - float(b)) <= 5e-5 * abs(float(b)), (name, float(a), float(b))
