"""Role wait totals of conv_tt_kernel (CTA 3) for encoder geometries (fac_set_option tt_probe)."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
e = Engine(); e._ensure(torch.device("cuda:0"))
e.set_option("tt_probe", 1)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for (B, T, C, K, dil) in ((32, 96000, 64, 7, 3), (32, 48000, 128, 7, 3), (32, 9600, 256, 7, 1), (32, 9600, 256, 1, 1), (32, 1920, 512, 7, 1)):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, T, C, generator=g) * 0.5).cuda()
    w = torch.randn(C, C, K, generator=g) / math.sqrt(C * K)
    bias = torch.zeros(C); a1 = torch.ones(C); a2 = torch.ones(C)
    y = torch.empty_like(x)
    pl = (K - 1) * dil
    res = x if K == 1 else None
    for it in range(3):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        rc = e.L.fac_debug_conv_tc(e.handle, P(x), P(w.contiguous()), P(bias), B, T, C, C, K, dil, 1, pl, 0, 1, P(a1) if K > 1 else None,
                                   P(a2) if K > 1 else None, 0, P(res), P(y), T, 4, None)
        t1.record(); torch.cuda.synchronize()
    if rc != 0:
        print(C, K, "rc", rc, e.L.fac_last_error(e.handle)); continue
    out = (ctypes.c_longlong * 8)()
    e.L.fac_debug_tc_phase_clocks(e.handle, out)
    o = [out[i] for i in range(8)]
    gx = (T + 255) // 256; gy = (C + 127) // 128
    ntiles = gx * gy * B
    n = (ntiles - 3 + 147) // 148
    print(f"C={C} K={K} T={T}: {t0.elapsed_time(t1):.3f} ms (incl. host pack), tiles/CTA {n}; per tile cycles: cta {o[0]//n}  prod_wait_empty {o[1]//n}  "
          f"mma_wait_act {o[2]//n}  mma_wait_w {o[3]//n}  mma_wait_drain {o[4]//n}  acc_wait_mma {o[5]//n}  acc_drain {o[6]//n}  acc_epilogue {o[7]//n}")
