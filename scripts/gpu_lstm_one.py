"""One SLSTM (2 layers) at the benchmark geometry (ncu target): python scripts/gpu_lstm_one.py H bf16"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facodec_b200.modules import Engine
H, bf = int(sys.argv[1]), int(sys.argv[2])
e = Engine(); e._ensure(torch.device("cuda:0"))
e.set_option("decoder_bf16", bf)
P = lambda t: ctypes.c_void_p(t.data_ptr())
B, T = 32, 320
g = torch.Generator().manual_seed(1)
lstm = torch.nn.LSTM(H, H, 2)
ws = [getattr(lstm, f"{n}_l{l}").detach().contiguous() for l in range(2) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
arr = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in ws])
x = torch.randn(B, T, H, generator=g).cuda(); y = torch.empty_like(x)
for _ in range(2):
    rc = e.L.fac_debug_slstm(e.handle, P(x), arr, B, T, H, P(y), None)
    assert rc == 0
torch.cuda.synchronize()
print("ok")
