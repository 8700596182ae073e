"""Phase timing (clock64) of one probe CTA for ResidualUnit geometries, fused and two-launch."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for (B, T, C, dil) in ((32, 48000, 192, 9), (32, 96000, 96, 9)):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, T, C, generator=g) * 0.5).cuda()
    w7 = torch.randn(C, C, 7, generator=g) / math.sqrt(C * 7); w1 = torch.randn(C, C, 1, generator=g) / math.sqrt(C)
    b7 = torch.zeros(C); b1 = torch.zeros(C); a1 = torch.ones(C); a2 = torch.ones(C)
    y = torch.empty_like(x)
    for mode, wide, dbg in ((6, 1, 0), (6, 0, 0)):
        e.L.fac_set_option(e.handle, b"tc_wide", wide)
        e.L.fac_set_option(e.handle, b"tc_dbg", dbg)
        rc = e.L.fac_debug_resunit(e.handle, P(x), P(w7), P(b7), P(w1), P(b1), P(a1), P(a2), B, T, C, dil, mode, P(y), None)
        if rc != 0:
            print(C, "mode", mode, "rc", rc, e.L.fac_last_error(e.handle)); continue
        out = (ctypes.c_longlong * 8)()
        e.L.fac_debug_tc_phase_clocks(e.handle, out)
        t = [out[i] - out[0] for i in range(6)]
        tr = (ctypes.c_longlong * 80)()
        e.L.fac_debug_tc_trace(e.handle, tr)
        if dbg == 0:
            base = out[0]
            for c in range(C // 16):
                print(f"      chunk {c:2d}: buffer-free {tr[32 + c] - base:6d}  stored {tr[48 + c] - base:6d}  mma-saw {tr[c] - base:6d}  mma-issued {tr[16 + c] - base:6d}  (weights wait {tr[64 + c]})")
        o4 = (ctypes.c_longlong * 4)()
        e.L.fac_debug_tc_producer_clocks(e.handle, o4)
        print(f"    producer thread 0 over {o4[3]} chunks: wait-free-buffer {o4[0]}  wait-loads {o4[1]}  transform+store {o4[2]}")
        print(f"C={C} T={T} mode={mode} wide={wide} dbg={dbg}: produced {t[1]}  gemm1_done {t[2]}  a2_done {t[3]}  gemm2_done {t[4]}  epilogue_done {t[5]}  | MMA warp waited: operands {out[6]} weights {out[7]}  (clks; last launch of the unit)")
