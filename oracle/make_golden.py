"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
*unmodified imported reference* (oracle/ref_import.py) on CPU, in the build
container (needs /root/reference).  Run:  python -m oracle.make_golden

Weights come from facodec_b200.synth.synth_state_dicts(seed) (host-independent
bits) loaded into the reference modules with load_state_dict, exactly as
reconstruct.py:30-37 loads a checkpoint; waves from synth.synth_waves
(PseudoDataset law).  Nothing but the case table, seeds and the reference's
own outputs goes into the fixtures, so any box can regenerate the inputs and
compare.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from facodec_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> (weight seed, wave seed, batch, samples, n_c, full_waves?)
CASES = {
    "b2_t7200": dict(wseed=0, xseed=114514, B=2, T=7200, n_c=2),
    "b1_t96000": dict(wseed=0, xseed=114514, B=1, T=96000, n_c=2),      # BASELINE configs[0]
    "b1_t7000_ragged": dict(wseed=0, xseed=7, B=1, T=7000, n_c=2),      # T % 300 != 0
    "b3_t1500_short": dict(wseed=1, xseed=9, B=3, T=1500, n_c=1),       # reflect-pad short-input branch
    "b2_t6000_fullwaves": dict(wseed=1, xseed=11, B=2, T=6000, n_c=2, full=9000, lens=(9000, 4800)),
}


def run_case(model, c):
    x = synth.synth_waves(c["B"], c["T"], seed=c["xseed"])
    kw = {}
    if "full" in c:
        kw["full_waves"] = synth.synth_waves(c["B"], c["full"], seed=c["xseed"] + 1).squeeze(1)
        kw["wave_lens"] = torch.tensor(c["lens"], dtype=torch.int64)
    with torch.no_grad():
        z = model.encoder(x)
        q = model.quantizer(z, x, n_c=c["n_c"], return_codes=True, **kw)
        y = model.decoder(q[0])
    out = dict(z=z, outs=q[0], z_p=q[1][0], z_c=q[1][1], z_r=q[1][2], commitment=q[2],
               codebook=q[3], timbre=q[4], codes_p=q[5][0], codes_c=q[5][1], codes_r=q[5][2], y=y)
    return {k: v.numpy() for k, v in out.items()}


# Voice-conversion fixtures (reconstruct_redecoder.py:108-122): codes + timbre of a codec fixture -> redecoder.encoder ->
# redecoder.decoder.  name -> (source codec fixture, redecoder weight seed, use_p_code, n_c)
REDEC_CASES = {
    "redec_b2_t7200_vc": dict(src="b2_t7200", wseed=0, use_p=False, n_c=1),      # the call reconstruct_redecoder.py makes
    "redec_b2_t7200_full": dict(src="b2_t7200", wseed=0, use_p=True, n_c=2),      # every embedding table
    "redec_b3_t1500_short": dict(src="b3_t1500_short", wseed=1, use_p=True, n_c=1),   # 5 frames: non-causal short-input pads
}


def run_redec_case(model, c):
    g = dict(np.load(os.path.join(GOLDEN_DIR, c["src"] + ".npz")))
    cp, cc, timbre = (torch.from_numpy(g[k]) for k in ("codes_p", "codes_c", "timbre"))
    with torch.no_grad():
        z = model.encoder(cp, cc, timbre, use_p_code=c["use_p"], n_c=c["n_c"])
        y = model.decoder(z)
    return dict(z=z.numpy(), y=y.numpy())


def vq_margin_report(sd, prefix, latents):
    """top-1 / top-2 gap of every decision of one VectorQuantize (dac/nn/quantize.py:78-94): returns the per-frame margin
    dist[2nd] - dist[1st] of the reference's own distance matrix (fp32).  Used by scripts/vq_margins.py to report how
    far the benchmark batch's decisions are from a tie."""
    import torch.nn.functional as F
    w_in = torch._weight_norm(sd[prefix + ".in_proj.weight_v"], sd[prefix + ".in_proj.weight_g"], 0)
    z_e = F.conv1d(latents, w_in, sd[prefix + ".in_proj.bias"])
    b, d, t = z_e.shape
    enc = F.normalize(z_e.permute(0, 2, 1).reshape(b * t, d))
    cb = F.normalize(sd[prefix + ".codebook.weight"])
    dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cb.t() + cb.pow(2).sum(1, keepdim=True).t()
    top2 = torch.topk(-dist, 2, dim=1).values
    return (top2[:, 0] - top2[:, 1]).reshape(b, t)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    model = ref_import.build_reference_model(0)
    loaded = None
    for name, c in CASES.items():
        if loaded != c["wseed"]:
            sds = synth.synth_state_dicts(c["wseed"])
            for k in ("encoder", "quantizer", "decoder"):
                model[k].load_state_dict(sds[k])
            loaded = c["wseed"]
        out = run_case(model, c)
        # z_p/z_c/z_r are large; keep float32 for the small cases only
        if c["B"] * c["T"] > 20000:
            for k in ("z_p", "z_c", "z_r"):
                out.pop(k)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")
    main_redecoder()


def main_redecoder():
    model = ref_import.build_reference_redecoder(0)
    loaded = None
    for name, c in REDEC_CASES.items():
        if loaded != c["wseed"]:
            sds = synth.synth_redecoder_state_dicts(c["wseed"])
            for k in ("encoder", "decoder"):
                model[k].load_state_dict(sds[k])
            loaded = c["wseed"]
        out = run_redec_case(model, c)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


RECON_LOSS_CASE = dict(B=2, T=4800, seed=11)


def main_recon_loss():
    """tests/golden/recon_loss.npz: losses.py:65-89 reconstruction_loss of the imported reference for one seeded pair, plus
    its 13 components (mse; l1, l2 per scale) formed with the same torchaudio transforms the reference constructs."""
    import warnings
    warnings.simplefilter("ignore")
    from torchaudio.transforms import MelSpectrogram
    ref_import.import_reference()
    import losses as ref_losses
    c = RECON_LOSS_CASE
    x, G_x = synth.synth_loss_pair(c["B"], c["T"], c["seed"])
    with torch.no_grad():
        loss = ref_losses.reconstruction_loss(x, G_x)
        terms = [torch.nn.functional.mse_loss(x, G_x)]
        for i in range(6, 12):
            s = 2 ** i
            melspec = MelSpectrogram(sample_rate=16000, n_fft=max(s, 512), win_length=s, hop_length=s // 4, n_mels=64)
            S_x, S_G = melspec(x), melspec(G_x)
            terms.append((S_x - S_G).abs().mean())
            terms.append((((torch.log(S_x.abs() + 1e-7) - torch.log(S_G.abs() + 1e-7)) ** 2).mean(dim=-2) ** 0.5).mean())
    path = os.path.join(GOLDEN_DIR, "recon_loss.npz")
    np.savez(path, loss=np.float32(loss), terms=torch.stack(terms).numpy(), B=c["B"], T=c["T"], seed=c["seed"])
    print("recon_loss", float(loss), [float(t) for t in terms])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "recon_loss":
        main_recon_loss()
    elif len(sys.argv) > 1 and sys.argv[1] == "redecoder":
        main_redecoder()
    else:
        main()
