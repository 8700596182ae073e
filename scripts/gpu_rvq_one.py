"""ResidualVQ over 2^18 frames (ncu target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200 as fb
rvq = fb.ResidualVQ(num_quantizers=4, codebook_size=10, dim=1024, codebook_dim=8).eval()
x = torch.randn(256, 1024, 1024, device="cuda")
for _ in range(2):
    rvq(x, channels_last=True, return_all=False)
torch.cuda.synchronize()
print("ok")
