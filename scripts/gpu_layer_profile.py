"""Per-call-site device time of one B=32 x 4 s forward (events around every launch)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200.build as _b
if os.environ.get('FAC_LIB_VARIANT'):
    _b.LIB = os.path.join(ROOT, 'facodec_b200', '_C', os.environ['FAC_LIB_VARIANT'])
import facodec_b200 as fb
from facodec_b200 import synth
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sds = synth.synth_state_dicts(0)
m = fb.build_model()
for k in ("encoder", "quantizer", "decoder"):
    m[k].load_state_dict(sds[k]); m[k].eval()
codec = fb.Codec(m); eng = codec.engine
x = synth.synth_waves(32, 96000).cuda()
eng.set_option("tensor_cores", mode)
codec.forward(x, n_c=2); codec.forward(x, n_c=2)
L, h = eng.L, eng.handle
L.fac_profile_reset(h); L.fac_profile_enable(h, 1)
codec.forward(x, n_c=2)
torch.cuda.synchronize()
L.fac_profile_enable(h, 0)
n = L.fac_profile_dump(h, None, 0)
buf = ctypes.create_string_buffer(n)
L.fac_profile_dump(h, buf, n)
rows = [l.split("\t") for l in buf.value.decode().strip().split("\n")]
rows.sort(key=lambda r: -float(r[1]))
tot = sum(float(r[1]) for r in rows)
print(f"mode {mode}: total profiled {tot:.1f} ms")
for r in rows[:45]:
    ms, gf, gb, nl = float(r[1]), float(r[2]), float(r[3]), int(r[4])
    print(f"{ms:8.2f} ms {100*ms/tot:5.1f}%  {gf/ms if ms else 0:7.1f} TFLOP/s {gb/ms*1e3 if ms else 0:7.0f} GB/s  x{nl:3d}  {r[0]}")
