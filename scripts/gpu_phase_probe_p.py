"""Role wait totals of the persistent promoted conv kernel (CTA 3) for encoder geometries."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200.build as _b
if os.environ.get('FAC_LIB_VARIANT'):
    _b.LIB = os.path.join(ROOT, 'facodec_b200', '_C', os.environ['FAC_LIB_VARIANT'])
from facodec_b200.modules import Engine
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 1
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for (B, T, C, K, dil) in ((32, 96000, 64, 7, 3), (32, 96000, 64, 1, 1), (32, 48000, 128, 7, 3), (32, 48000, 128, 1, 1), (32, 9600, 256, 7, 1), (32, 1920, 512, 7, 1)):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, T, C, generator=g) * 0.5).cuda()
    w = torch.randn(C, C, K, generator=g) / math.sqrt(C * K)
    bias = torch.zeros(C); a1 = torch.ones(C); a2 = torch.ones(C)
    y = torch.empty_like(x)
    pl = (K - 1) * dil
    res = x if K == 1 else None
    for it in range(2):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        rc = e.L.fac_debug_conv_tc(e.handle, P(x), P(w.contiguous()), P(bias), B, T, C, C, K, dil, 1, pl, 0, 1, P(a1) if K > 1 else None,
                                   P(a2) if K > 1 else None, 0, P(res), P(y), T, MODE, None)
    if rc != 0:
        print(C, K, "rc", rc, e.L.fac_last_error(e.handle)); continue
    out = (ctypes.c_longlong * 8)()
    e.L.fac_debug_tc_phase_clocks(e.handle, out)
    o = [out[i] for i in range(8)]
    n = max(1, o[7])
    print(f"mode {MODE} C={C} K={K} T={T}: tiles {o[7]}  per tile: total {o[0]//n}  prod_wait_a_empty {o[1]//n}  mma_wait_a_full {o[2]//n}  mma_wait_b_full {o[3]//n}  "
          f"mma_wait_acc_free {o[4]//n}  acc_wait_ready {o[5]//n}  acc_epilogue {o[6]//n}")
