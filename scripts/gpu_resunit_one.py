"""One fused ResidualUnit launch of a decoder geometry (ncu target): python scripts/gpu_resunit_one.py C dil T [mode]
mode 6 = fused, bf16 hi/lo 1x1 + one-pass fp16 k = 7 conv (the default decoder class); 4 = fused bf16 hi/lo."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
C, dil, T = (int(a) for a in sys.argv[1:4])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 6
B = 32
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
g = torch.Generator().manual_seed(1)
x = (torch.randn(B, T, C, generator=g) * 0.5).cuda()
w7 = torch.randn(C, C, 7, generator=g) / math.sqrt(C * 7); w1 = torch.randn(C, C, 1, generator=g) / math.sqrt(C)
b7 = torch.zeros(C); b1 = torch.zeros(C); a1 = torch.ones(C); a2 = torch.ones(C)
y = torch.empty_like(x)
for it in range(2):
    rc = e.L.fac_debug_resunit(e.handle, P(x), P(w7), P(b7), P(w1), P(b1), P(a1), P(a2), B, T, C, dil, mode, P(y), None)
    assert rc == 0, e.L.fac_last_error(e.handle)
torch.cuda.synchronize()
print("ok")
