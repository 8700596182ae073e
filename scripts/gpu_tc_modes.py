"""Accuracy of the three tensor_cores modes against the golden fixtures / oracle (B200 box)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200 as fb
from facodec_b200 import synth
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN_CASES, case_inputs, load_golden

sds = synth.synth_state_dicts(0)
m = fb.build_model()
for k in ("encoder", "quantizer", "decoder"):
    m[k].load_state_dict(sds[k]); m[k].eval()
codec = fb.Codec(m)
eng = codec.engine
for name in ("b2_t7200", "b1_t96000"):
    c = GOLDEN_CASES[name]; g = load_golden(name)
    x, kw = case_inputs(c)
    xd = x.cuda()
    for mode in (0, 1, 2):
        eng.set_option("tensor_cores", mode)
        z = m.encoder(xd)
        q = m.quantizer(z, xd, n_c=2, return_codes=True)
        y = m.decoder(q[0])
        torch.cuda.synchronize()
        zerr = float(np.abs(z.cpu().numpy() - g["z"]).max() / np.abs(g["z"]).max())
        mism = sum(int((t.cpu().numpy() != g[k]).sum()) for k, t in zip(("codes_p", "codes_c", "codes_r"), q[5]))
        rms = float(np.sqrt(((y.cpu().numpy().astype(np.float64) - g["y"]) ** 2).mean()))
        # decoder alone on golden outs
        y2 = m.decoder(torch.from_numpy(g["outs"]).cuda())
        rms2 = float(np.sqrt(((y2.cpu().numpy().astype(np.float64) - g["y"]) ** 2).mean()))
        print(f"{name} mode {mode}: z relmax {zerr:.2e} code mismatches {mism} y rms {rms:.2e} dec-only rms {rms2:.2e}", flush=True)
# timing at B=32
x = synth.synth_waves(32, 96000).cuda()
for mode in (0, 1, 2):
    eng.set_option("tensor_cores", mode)
    for _ in range(2):
        codec.forward(x, n_c=2)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(3):
        y, codes, _ = codec.forward(x, n_c=2)
    torch.cuda.synchronize()
    dt = (time.time() - t) / 3
    if mode == 0:
        ref_codes = [c.clone() for c in codes]; ref_y = y.clone()
    mism = sum(int((a != b).sum()) for a, b in zip(codes, ref_codes))
    rms = float((y.double() - ref_y.double()).pow(2).mean().sqrt())
    print(f"B=32 mode {mode}: {dt*1e3:.1f} ms/step = {128/dt:.0f} audio-s/s ; vs mode0: code mismatches {mism}/61440, y rms {rms:.2e}", flush=True)
