"""CPU: the oracle restatement against (a) the committed golden fixtures generated from the
imported unmodified reference and (b) the imported reference itself when /root/reference exists."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, case_inputs, load_golden, state_dicts
from oracle import facodec_oracle as O
from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Bit-exact in the build container (same CPU, same ATen kernels as when the fixtures were made);
# on another host CPU oneDNN may pick other kernels, so floats get a tight tolerance there.
SAME_HOST = ref_import.available()
ATOL = 0.0 if SAME_HOST else 2e-5


def _close(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, name
    if ATOL == 0.0:
        assert np.array_equal(a, b), f"{name}: max diff {np.abs(a - b).max()}"
    else:
        assert np.abs(a - b).max() <= ATOL * max(1.0, np.abs(b).max()), name


def test_case_table():
    from oracle import make_golden
    assert make_golden.CASES == GOLDEN_CASES
    for name in GOLDEN_CASES:
        assert os.path.exists(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))


def test_synth_is_host_independent():
    """Synthetic checkpoints must be the same bits everywhere (golden fixtures depend on it)."""
    from facodec_b200 import synth
    sd = synth.synth_encoder(1)
    h = hashlib.sha256()
    for k in ("block.0.conv.conv.weight_g", "block.1.block.0.block.1.conv.conv.weight_v", "block.6.alpha"):
        h.update(sd[k].numpy().tobytes())
    assert h.hexdigest() == "db65aa08d774396d996e318b9407e70d4417e942368011e08c72c2d351e27314"
    w = synth.synth_waves(1, 1000, seed=3)
    assert abs(float(w.abs().max()) - 1.0) < 1e-7


@pytest.mark.parametrize("name", ["b2_t7200", "b1_t7000_ragged", "b3_t1500_short", "b2_t6000_fullwaves", "b1_t96000"])
def test_oracle_matches_golden(name):
    c = GOLDEN_CASES[name]
    g = load_golden(name)
    sds = state_dicts(c["wseed"])
    x, kw = case_inputs(c)
    with torch.no_grad():
        z = O.encoder_forward(sds["encoder"], x)
        q = O.quantizer_forward(sds["quantizer"], z, x, n_c=c["n_c"], return_codes=True, **kw)
        y = O.decoder_forward(sds["decoder"], q[0])
    _close(z, g["z"], "z")
    _close(q[0], g["outs"], "outs")
    _close(q[4], g["timbre"], "timbre")
    _close(y, g["y"], "y")
    for k, t in zip(("codes_p", "codes_c", "codes_r"), q[5]):
        if SAME_HOST:
            assert np.array_equal(t.numpy(), g[k]), k
        else:
            assert (t.numpy() != g[k]).mean() < 0.02, k
    assert abs(float(q[2]) - float(g["commitment"])) <= 1e-5 * abs(float(g["commitment"]))
    if "z_p" in g:
        for k, t in zip(("z_p", "z_c", "z_r"), q[1]):
            _close(t, g[k], k)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_oracle_matches_imported_reference():
    """Pins the restatement to the real thing: bit-identical tensors from the unmodified reference."""
    import warnings
    warnings.simplefilter("ignore")
    model = ref_import.build_reference_model(0)
    sds = state_dicts(1)
    for k in ("encoder", "quantizer", "decoder"):
        model[k].load_state_dict(sds[k])
    x, _ = case_inputs(dict(B=2, T=4500, xseed=21))
    with torch.no_grad():
        z = model.encoder(x)
        q = model.quantizer(z, x, n_c=2, return_codes=True)
        y = model.decoder(q[0])
        z2, q2, y2 = O.codec_forward(sds, x, n_c=2)
    assert torch.equal(z, z2) and torch.equal(q[0], q2[0]) and torch.equal(y, y2)
    assert torch.equal(q[4], q2[4])
    for a, b in zip(q[5], q2[5]):
        assert torch.equal(a, b)
    for a, b in zip(q[1], q2[1]):
        assert torch.equal(a, b)
    assert float(q[2]) == float(q2[2]) and float(q[3]) == float(q2[3])


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_fvq_rvq_and_alias_free_match_reference():
    import sys
    import warnings
    warnings.simplefilter("ignore")
    ref_import.import_reference()
    from quantize.rvq import ResidualVQ as RefRVQ
    from alias_free_torch import Activation1d as RefAct
    torch.manual_seed(3)
    rvq = RefRVQ(num_quantizers=4, codebook_size=10, dim=1024, codebook_dim=8, commitment=0.25).eval()
    x = torch.randn(2, 1024, 17)
    layers = []
    for l in rvq.layers:
        layers.append(dict(in_w=l.in_proj.weight.detach(), in_b=l.in_proj.bias.detach(),
                           out_w=l.out_proj.weight.detach(), out_b=l.out_proj.bias.detach(),
                           codebook=l.codebook.weight.detach()))
    with torch.no_grad():
        a = rvq(x)
        b = O.fvq_residual_vq(layers, x)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0]) and torch.equal(a[3], b[3])
    act = RefAct(activation=torch.nn.Identity())
    xx = torch.randn(2, 5, 50)
    assert torch.allclose(act(xx), O.alias_free_act(xx, lambda u: u), atol=0, rtol=0)


def test_reflect_pad_short_branch():
    """encodec.py:96-113: length <= pad => zero-extend, reflect, truncate."""
    x = torch.arange(1, 4, dtype=torch.float32).reshape(1, 1, 3)
    y = O._pad1d_reflect(x, 5, 0)
    assert y.shape[-1] == 8
    assert y.flatten().tolist() == [0.0, 0.0, 0.0, 3.0, 2.0, 1.0, 2.0, 3.0]


# ---------------------------------------------------------------------------------------------------------------
# round 2: voice-conversion path (modules/redecoder.py), predictor heads (modules/quantize.py:29-125), dataset mel
# ---------------------------------------------------------------------------------------------------------------
from conftest import REDEC_CASES  # noqa: E402


def test_redecoder_case_table():
    from oracle import make_golden
    assert make_golden.REDEC_CASES == REDEC_CASES
    for name in REDEC_CASES:
        assert os.path.exists(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))


@pytest.mark.parametrize("name", list(REDEC_CASES))
def test_redecoder_oracle_matches_golden(name):
    from facodec_b200 import synth
    c = REDEC_CASES[name]
    g = load_golden(name)
    src = load_golden(c["src"])
    sds = synth.synth_redecoder_state_dicts(c["wseed"])
    cp, cc, timbre = (torch.from_numpy(src[k]) for k in ("codes_p", "codes_c", "timbre"))
    with torch.no_grad():
        z = O.redecoder_forward(sds["encoder"], cp, cc, timbre, use_p_code=c["use_p"], n_c=c["n_c"])
        y = O.decoder_forward(sds["decoder"], z, causal=False, lstm=0)
    _close(z, g["z"], "z")
    _close(y, g["y"], "y")


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_redecoder_oracle_matches_imported_reference():
    """modules/redecoder.py:35-48 + the non-causal, LSTM-free Decoder of build_model(stage='redecoder'): bit-identical."""
    import warnings
    warnings.simplefilter("ignore")
    from facodec_b200 import synth
    model = ref_import.build_reference_redecoder(0)
    sds = synth.synth_redecoder_state_dicts(2)
    for k in ("encoder", "decoder"):
        model[k].load_state_dict(sds[k])
    g = torch.Generator().manual_seed(5)
    cp = torch.randint(0, 1024, (2, 1, 13), generator=g)
    cc = torch.randint(0, 1024, (2, 2, 13), generator=g)
    timbre = torch.randn(2, 1024, generator=g)
    for use_p, n_c in ((False, 1), (True, 2)):
        with torch.no_grad():
            z = model.encoder(cp, cc, timbre, use_p_code=use_p, n_c=n_c)
            y = model.decoder(z)
            z2 = O.redecoder_forward(sds["encoder"], cp, cc, timbre, use_p_code=use_p, n_c=n_c)
            y2 = O.decoder_forward(sds["decoder"], z2, causal=False, lstm=0)
        assert torch.equal(z, z2) and torch.equal(y, y2)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_predictor_heads_and_snakebeta_match_imported_reference():
    """SnakeBeta (modules/quantize.py:29-88) inside Activation1d, the heads' ResidualUnit (:90-104) and CNNLSTM (:106-125),
    imported unmodified: the restatement is bit-identical (round 1 pinned the alias-free activation with Identity only)."""
    import warnings
    warnings.simplefilter("ignore")
    from facodec_b200 import synth
    ref_import.import_reference()
    from modules.quantize import CNNLSTM, SnakeBeta
    from alias_free_torch import Activation1d as RefAct
    g = torch.Generator().manual_seed(9)
    sb = SnakeBeta(6, alpha_logscale=True)
    with torch.no_grad():
        sb.alpha.copy_(torch.randn(6, generator=g) * 0.3)
        sb.beta.copy_(torch.randn(6, generator=g) * 0.3)
    x = torch.randn(2, 6, 40, generator=g)
    with torch.no_grad():
        assert torch.equal(sb(x), O.snake_beta(x, sb.alpha, sb.beta))
        act = RefAct(activation=sb)
        assert torch.equal(act(x), O.alias_free_act(x, lambda u: O.snake_beta(u, sb.alpha, sb.beta)))
    for (indim, outdim, heads, glob) in ((64, 10, 2, False), (32, 7, 1, True)):
        m = CNNLSTM(indim, outdim, heads, global_pred=glob).eval()
        sd = synth.synth_cnnlstm(3, indim, outdim, heads)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("filter") for k in missing)      # only the registered filter buffers
        xx = torch.randn(2, indim, 33, generator=g)
        with torch.no_grad():
            a = m(xx)
            b = O.cnnlstm_forward(sd, xx, heads, global_pred=glob)
        assert len(a) == len(b) == heads
        for u, v in zip(a, b):
            assert torch.equal(u, v)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_dataset_mel_matches_imported_meldataset():
    """meldataset.py:37-47 preprocess (torchaudio MelSpectrogram with its default sample_rate = 16000) imported unmodified
    (soundfile / librosa stubbed: file I/O only), against the restatement fed with synth's host-independent window and
    16 kHz filterbank."""
    import warnings
    warnings.simplefilter("ignore")
    from facodec_b200 import synth
    ref_import.import_reference()
    import meldataset
    w = synth.synth_waves(1, 5000, seed=3)[0, 0]
    with torch.no_grad():
        ref = meldataset.preprocess(w.numpy())
        fb_ref = meldataset.to_mel.mel_scale.fb
        win_ref = meldataset.to_mel.spectrogram.window
        got_same = O.dataset_mel(w, win_ref, fb_ref)
        assert torch.equal(ref, got_same)
        fb = synth.melscale_fbanks_htk(sample_rate=16000, f_max=8000.0)
        assert float((fb - fb_ref).abs().max()) <= 1e-5      # fp64-then-round vs torchaudio fp32 evaluation
        got = O.dataset_mel(w, synth.hann_window_periodic(1200), fb)
    assert tuple(ref.shape) == (1, 80, 5000 // 300 + 1)
    assert float((got - ref).abs().max()) <= 2e-5


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_dac_code_file_matches_imported_dacfile(tmp_path):
    """dac/model/base.py:15-54: files written here are byte-identical to the reference class's, and each side loads the
    other's (codes, every metadata field)."""
    from facodec_b200 import codefile
    ref_import.import_reference()
    from dac.model.base import DACFile as RefFile
    g = torch.Generator().manual_seed(9)
    codes = [torch.randint(0, 1024, (2, n, 37), generator=g) for n in (1, 2, 3)]
    mine = codefile.from_forward(codes, original_length=37 * 300, input_db=torch.tensor([-23.5, -17.25]))
    ref = RefFile(codes=codefile.pack_codes(codes), chunk_length=37, original_length=37 * 300,
                  input_db=torch.tensor([-23.5, -17.25]), channels=1, sample_rate=24000, padding=True, dac_version="1.0.0")
    pa, pb = mine.save(tmp_path / "mine"), ref.save(tmp_path / "ref")
    assert pa.suffix == ".dac" and open(pa, "rb").read() == open(pb, "rb").read()
    a, b = RefFile.load(pa), codefile.DACFile.load(pb)
    for f in (a, b):
        assert torch.equal(f.codes, codefile.pack_codes(codes))
        assert (f.chunk_length, f.original_length, f.channels, f.sample_rate, f.padding, f.dac_version) == (37, 11100, 1, 24000, True, "1.0.0")
        assert np.array_equal(np.asarray(f.input_db), np.array([-23.5, -17.25], np.float32))
    for u, v in zip(codefile.unpack_codes(b.codes, n_c=2), codes):
        assert torch.equal(u, v)


def _loss_signals(B=2, T=4800, seed=11):
    from facodec_b200 import synth
    return synth.synth_loss_pair(B, T, seed)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
def test_reconstruction_loss_matches_imported_losses_py():
    """losses.py:65-89 imported unmodified (torchaudio is installed) against the restatement: same scalar, bit for bit."""
    import warnings
    warnings.simplefilter("ignore")
    ref_import.import_reference()
    import losses as ref_losses
    x, G_x = _loss_signals()
    with torch.no_grad():
        ref = ref_losses.reconstruction_loss(x, G_x)
        got = O.reconstruction_loss(x, G_x)
    assert ref.dim() == 0 and torch.equal(ref, got)


def test_reconstruction_loss_golden():
    """tests/golden/recon_loss.npz: loss + 13 terms the imported reference modules gave for the seeded pair (oracle/make_golden.py)."""
    import warnings
    warnings.simplefilter("ignore")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "recon_loss.npz"))
    x, G_x = _loss_signals(int(gold["B"]), int(gold["T"]), int(gold["seed"]))
    with torch.no_grad():
        L, terms = O.reconstruction_loss(x, G_x, return_terms=True)
    assert abs(float(L) - float(gold["loss"])) <= 2e-6 * abs(float(gold["loss"]))
    assert np.allclose(terms.numpy(), gold["terms"], rtol=2e-6, atol=0)


FAP_FLAGS = dict(use_gr_content_f0=False, use_gr_prosody_phone=False, use_gr_residual_f0=True, use_gr_residual_phone=True,
                 use_gr_timbre_content=True, use_gr_timbre_prosody=False, use_gr_x_timbre=True, norm_f0=True)   # modules/commons.py:311-322 + config.yml


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not on this box")
@pytest.mark.parametrize("timbre_norm", [True, False])
def test_fa_predictors_match_imported_reference(timbre_norm):
    """FApredictors (modules/quantize.py:456-619) imported unmodified with build_model's flags, both forward variants:
    the restatement over its state_dict is bit-identical, output by output."""
    import warnings
    warnings.simplefilter("ignore")
    ref_import.import_reference()
    from modules.quantize import FApredictors
    torch.manual_seed(4)
    m = FApredictors(in_dim=32, timbre_norm=timbre_norm, use_gr_content_global_f0=True, **FAP_FLAGS).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(6)
    lat = [torch.randn(2, 32, 19, generator=g) for _ in range(3 if timbre_norm else 4)]
    with torch.no_grad():
        if timbre_norm:
            timbre = torch.randn(2, 32, generator=g)
            ref = m(lat, timbre)
            got = O.fa_predictors_forward(sd, lat, timbre, timbre_norm=True, **FAP_FLAGS)
        else:
            ref = m(lat)
            got = O.fa_predictors_forward(sd, lat, None, timbre_norm=False, **FAP_FLAGS)
    for a, b in zip(ref, got):
        assert a.keys() == b.keys()
        for k in a:
            if a[k] is None or b[k] is None:
                assert a[k] is None and b[k] is None, k
            elif a[k].shape[-1] == 1:
                # 1-wide nn.Linear heads (f0 / uv): torch's CPU F.linear takes another kernel for weights that do not require
                # grad (the oracle works on detached state_dict tensors, the module on Parameters): 1-2 ulp apart
                assert float((a[k] - b[k]).abs().max()) <= 5e-7 * max(1.0, float(a[k].abs().max())), k
            else:
                assert torch.equal(a[k], b[k]), k


def test_slaney_mel_filterbank_against_torchaudio():
    """The restated librosa.filters.mel (Slaney scale + area norm; librosa itself is not installed) against torchaudio's
    independent Slaney implementation, for the geometries train.py:155-163 uses."""
    import torchaudio
    for w, nm in ((32, 5), (64, 10), (256, 40), (512, 80), (2048, 320), (2048, 150)):
        mine = O.librosa_mel_filters(24000, w, nm, 0.0, None)
        ref = torchaudio.functional.melscale_fbanks(w // 2 + 1, 0.0, 12000.0, nm, 24000, norm="slaney", mel_scale="slaney").T
        assert mine.shape == ref.shape == (nm, w // 2 + 1)
        assert float((mine - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), (w, nm)


def test_dac_spectral_losses_restatement_properties():
    """dac/nn/loss.py MultiScaleSTFTLoss / MelSpectrogramLoss restatements (parity unpinned: audiotools absent): zero for
    identical signals, the magnitude path equals a direct torch.stft evaluation, train.py's mel configuration runs."""
    from facodec_b200 import synth
    x, y = synth.synth_loss_pair(2, 3000, seed=3)
    assert float(O.multiscale_stft_loss(x, x)) == 0.0
    assert float(O.mel_spectrogram_loss(x, x)) == 0.0
    st = torch.stft(x[:, 0], 512, hop_length=128, window=torch.hann_window(512), return_complex=True)
    sy = torch.stft(y[:, 0], 512, hop_length=128, window=torch.hann_window(512), return_complex=True)
    direct = (st.abs() - sy.abs()).abs().mean()
    got = O.multiscale_stft_loss(x, y, window_lengths=(512,), log_weight=0.0)
    assert abs(float(got) - float(direct)) <= 1e-6 * float(direct)
    L = O.mel_spectrogram_loss(x, y, 24000, n_mels=[5, 10, 20, 40, 80, 160, 320], window_lengths=[32, 64, 128, 256, 512, 1024, 2048],
                               mel_fmin=[0] * 7, mel_fmax=[None] * 7, pow=1.0, mag_weight=0.0)
    assert torch.isfinite(L) and float(L) > 0
