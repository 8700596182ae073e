"""Per-call-site device time of one B=32 x 4 s forward under different residency / fusion options."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200 as fb
from facodec_b200 import synth
sds = synth.synth_state_dicts(0)
m = fb.build_model()
for k in ("encoder", "quantizer", "decoder"):
    m[k].load_state_dict(sds[k]); m[k].eval()
codec = fb.Codec(m); eng = codec.engine
x = synth.synth_waves(32, 96000).cuda()
eng.set_option("tc_occ2_maxn", 0)
L, h = eng.L, eng.handle
base = None
results = {}
for (occ2, fuse) in ((0, 1), (128, 1), (256, 1), (256, 2), (0, 0), (256, 0)):
    eng.set_option("tc_occ2_maxn", occ2); eng.set_option("fuse_resunit", fuse)
    y, codes, timbre = codec.forward(x, n_c=2); codec.forward(x, n_c=2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): codec.forward(x, n_c=2)
    b.record(); torch.cuda.synchronize()
    wall = a.elapsed_time(b) / 3
    if base is None: base = (y.clone(), [c.clone() for c in codes])
    rms = float(((y.double() - base[0].double()) ** 2).mean().sqrt())
    same = all(torch.equal(c, d) for c, d in zip(codes, base[1]))
    L.fac_profile_reset(h); L.fac_profile_enable(h, 1)
    codec.forward(x, n_c=2)
    torch.cuda.synchronize()
    L.fac_profile_enable(h, 0)
    n = L.fac_profile_dump(h, None, 0)
    buf = ctypes.create_string_buffer(n)
    L.fac_profile_dump(h, buf, n)
    rows = [l.split("\t") for l in buf.value.decode().strip().split("\n")]
    tot = sum(float(r[1]) for r in rows)
    print(f"=== occ2={occ2} fuse={fuse}: {wall:.2f} ms/step (profiled sum {tot:.1f}); y rms vs first {rms:.2e}; codes equal {same}")
    results[(occ2, fuse)] = {r[0]: float(r[1]) for r in rows}
keys = sorted(results[(0, 1)].keys(), key=lambda k: -results[(0, 1)][k])
cfgs = list(results.keys())
print("layer".ljust(58) + "".join(f"{str(c):>12}" for c in cfgs))
allk = []
for c in cfgs:
    for k in results[c]:
        if k not in allk: allk.append(k)
allk.sort(key=lambda k: -max(results[c].get(k, 0) for c in cfgs))
for k in allk[:60]:
    print(k[:57].ljust(58) + "".join(f"{results[c].get(k, float('nan')):12.2f}" for c in cfgs))
