"""Two full B=32 x 4 s forwards (1 warm-up + 1 measured) for ncu captures (never a bench number)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200 as fb
from facodec_b200 import synth
sds = synth.synth_state_dicts(0)
m = fb.build_model()
for k in ("encoder", "quantizer", "decoder"):
    m[k].load_state_dict(sds[k]); m[k].eval()
codec = fb.Codec(m)
x = synth.synth_waves(32, 96000).cuda()
for _ in range(2):
    codec.forward(x, n_c=2)
torch.cuda.synchronize()
print("launches per forward", codec.launch_count())
