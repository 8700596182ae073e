"""One conv_tt_kernel launch of an encoder geometry (ncu target): python scripts/gpu_tt_one.py C K dil T"""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
C, K, dil, T = (int(a) for a in sys.argv[1:5])
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 4
B = 32
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
g = torch.Generator().manual_seed(1)
x = (torch.randn(B, T, C, generator=g) * 0.5).cuda()
w = torch.randn(C, C, K, generator=g) / math.sqrt(C * K)
bias = torch.zeros(C); a1 = torch.ones(C); a2 = torch.ones(C)
y = torch.empty_like(x)
res = x if K == 1 else None
for it in range(2):
    rc = e.L.fac_debug_conv_tc(e.handle, P(x), P(w.contiguous()), P(bias), B, T, C, C, K, dil, 1, (K - 1) * dil, 0, 1, P(a1) if K > 1 else None,
                               P(a2) if K > 1 else None, 0, P(res), P(y), T, mode, None)
    assert rc == 0, e.L.fac_last_error(e.handle)
torch.cuda.synchronize()
print("ok")
