import ctypes, math, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
e = Engine(); e._ensure(torch.device("cuda:0"))
P = lambda t: ctypes.c_void_p(t.data_ptr())
for (B, T, H, bf) in ((32, 320, 1024, 0), (32, 320, 1536, 1)):
    e.set_option('decoder_bf16', bf)   # 0 -> 3-pass (encoder class), 1 -> one fp16 pass (decoder class)
    g = torch.Generator().manual_seed(1)
    lstm = torch.nn.LSTM(H, H, 2)
    ws = [getattr(lstm, f"{n}_l{l}").detach().contiguous() for l in range(2) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    arr = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in ws])
    x = torch.randn(B, T, H, generator=g).cuda(); y = torch.empty_like(x)
    for _ in range(2):
        rc = e.L.fac_debug_slstm(e.handle, P(x), arr, B, T, H, P(y), None)
    out = (ctypes.c_longlong * 4)()
    e.L.fac_debug_lstm_phase_clocks(e.handle, out)
    tot = sum(out)
    print(f"H={H}: per-step clks: barrier wait {out[0]/T:.0f}  k-loop {out[1]/T:.0f}  reduce+gates {out[2]/T:.0f}  publish {out[3]/T:.0f}  total {tot/T:.0f} (~{tot/T/1.965e3:.1f} us)")
