"""Role wait totals of conv_tt_kernel (CTA 3) for arbitrary geometries: args "Cin,Cout,K,dil,stride,T[,snake]" ..."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
B = 32
for spec in sys.argv[1:]:
    f = [int(v) for v in spec.split(",")]
    Cin, Cout, K, dil, stride, T = f[:6]
    snake = f[6] if len(f) > 6 else 1
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, T, Cin, generator=g) * 0.5).cuda()
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    bias = torch.zeros(Cout); a1 = torch.ones(Cin); a2 = torch.ones(Cout)
    pl = (K - 1) * dil + 1 - stride
    Tout = (T + pl - ((K - 1) * dil + 1)) // stride + 1
    y = torch.empty(B, Tout, Cout, device="cuda")
    for probe in (0, 1):
        e.set_option("tt_probe", probe)
        for it in range(3):
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            rc = e.L.fac_debug_conv_tc(e.handle, P(x), P(w.contiguous()), P(bias), B, T, Cin, Cout, K, dil, stride, pl, 0, 1,
                                       P(a1) if snake else None, None, 0, None, P(y), Tout, 4, None)
        assert rc == 0, e.L.fac_last_error(e.handle)
    out = (ctypes.c_longlong * 8)()
    e.L.fac_debug_tc_phase_clocks(e.handle, out)
    o = [out[i] for i in range(8)]
    NT = 256
    gx = (Tout + NT - 1) // NT; gy = (Cout + 127) // 128
    n = max(1, (gx * gy * B - 3 + 147) // 148)
    print(f"{spec}: tiles/CTA {n}; per tile cycles: cta {o[0]//n}  prod_wait_empty {o[1]//n}  mma_wait_act {o[2]//n}  mma_wait_w {o[3]//n}  "
          f"mma_wait_drain {o[4]//n}  acc_wait_mma {o[5]//n}  acc_drain {o[6]//n}  acc_epilogue {o[7]//n}")
