"""Promoted kernel (conv_tcp) probe CTA: where the worker warps spend their time."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for (B, T, C, K, dil) in ((32, 96000, 64, 7, 3), (32, 48000, 128, 7, 3), (32, 9600, 256, 7, 3), (32, 1920, 512, 7, 3), (32, 96000, 64, 1, 1), (32, 48000, 128, 1, 1)):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, T, C, generator=g) * 0.5).cuda()
    w = torch.randn(C, C, K, generator=g) / math.sqrt(C * K); b = torch.zeros(C); a1 = torch.ones(C)
    y = torch.empty_like(x)
    pl = (K - 1) * dil
    for _ in range(2):
        rc = e.L.fac_debug_conv_tc(e.handle, P(x), P(w), P(b), B, T, C, C, K, dil, 1, pl, 0, 1, P(a1), P(a1) if K == 7 else None, 0, None, P(y), T, 1, None)
    assert rc == 0, e.L.fac_last_error(e.handle)
    out = (ctypes.c_longlong * 8)()
    e.L.fac_debug_tc_phase_clocks(e.handle, out)
    pass
    print(f"C={C} K={K} T={T}: producers done {out[1]-out[0]}  last promote {out[2]-out[0]}  staged {out[3]-out[0]}  epilogue done {out[5]-out[0]} clks")
