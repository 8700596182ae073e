"""Per-stage diagnosis on the GPU box: every tapped intermediate of the CUDA path vs the oracle's."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200 as fb
from facodec_b200 import synth
from oracle import facodec_oracle as O

B, T, seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 7200, 0
sds = synth.synth_state_dicts(seed)
m = fb.build_model()
for k in ("encoder", "quantizer", "decoder"):
    m[k].load_state_dict(sds[k]); m[k].eval()
x = synth.synth_waves(B, T)
taps = {}
with torch.no_grad():
    zo = O.encoder_forward(sds["encoder"], x, taps=taps)
    qo = O.quantizer_forward(sds["quantizer"], zo, x, n_c=2, return_codes=True, taps=taps)
    yo = O.decoder_forward(sds["decoder"], qo[0], taps=taps)
eng = m.encoder._engine
eng.sync_weights(torch.device("cuda:0"))
L, h = eng.L, eng.handle
bufs = {}
for name, t in taps.items():
    bufs[name] = torch.zeros(t.numel(), device="cuda")
    L.fac_debug_tap(h, name.encode(), ctypes.c_void_p(bufs[name].data_ptr()), t.numel())


def rep(name, ours, ref):
    d = (ours.double() - ref.double())
    print(f"{name:14s} shape {tuple(ref.shape)} max|ref| {ref.abs().max():.4g} maxerr {d.abs().max():.3e} rms {d.pow(2).mean().sqrt():.3e}")


xd = x.cuda()
# teacher-forced per module
z = m.encoder(xd)
q = m.quantizer(zo.cuda(), xd, n_c=2, return_codes=True)
y = m.decoder(qo[0].cuda())
torch.cuda.synchronize()
for name, t in taps.items():
    Bn, C, Tn = t.shape
    ours = bufs[name].cpu().reshape(Bn, Tn, C).transpose(1, 2)
    rep(name, ours, t)
rep("z", z.cpu(), zo)
rep("timbre(tf)", q[4].cpu(), qo[4])
for n, a, b in zip(("z_p", "z_c", "z_r"), q[1], qo[1]):
    rep(n + "(tf)", a.cpu(), b)
rep("outs(tf)", q[0].cpu(), qo[0])
for n, a, b in zip(("codes_p", "codes_c", "codes_r"), q[5], qo[5]):
    print(n, "mismatch", int((a.cpu() != b).sum()), "of", b.numel())
print("losses", float(q[2]), float(qo[2]))
rep("y(tf)", y.cpu(), yo)
# determinism
z2 = m.encoder(xd); y2 = m.decoder(qo[0].cuda()); q2 = m.quantizer(zo.cuda(), xd, n_c=2, return_codes=True)
torch.cuda.synchronize()
print("rerun equal: z", torch.equal(z, z2), "y", torch.equal(y, y2), "outs", torch.equal(q[0], q2[0]), "timbre", torch.equal(q[4], q2[4]))
