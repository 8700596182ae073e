"""CPU experiment behind lstm2.cu's downstream precision class: how far does the reconstructed waveform move when the decoder
LSTM's recurrent operands (W_hh and h) are rounded to fp16 / bf16?  Manual LSTM loop (torch, fp32 accumulation) inside the
oracle's decoder; bar = 1e-4 waveform RMS (BASELINE north_star).  python scripts/cpu_lstm_precision.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facodec_b200 import synth
from oracle import facodec_oracle as O

torch.set_num_threads(os.cpu_count() or 1)
sds = synth.synth_state_dicts(0)
x = synth.synth_waves(2, 24000, seed=3)


def lstm_manual(x, sd, prefix, wq, hq):
    xt = x.permute(2, 0, 1)
    T, B, H = xt.shape
    inp = xt
    for l in range(2):
        Wih, Whh = sd[f"{prefix}.weight_ih_l{l}"], wq(sd[f"{prefix}.weight_hh_l{l}"])
        xg = inp @ Wih.t() + sd[f"{prefix}.bias_ih_l{l}"] + sd[f"{prefix}.bias_hh_l{l}"]
        h = torch.zeros(B, H); c = torch.zeros(B, H); outs = []
        for t in range(T):
            i, f, g, o = (xg[t] + hq(h) @ Whh.t()).chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs)
    return (inp + xt).permute(1, 2, 0)


def decoder(sd, z, wq, hq):
    h = O.sconv1d(z, sd, "model.0.conv.conv")
    h = lstm_manual(h, sd, "model.1.lstm", wq, hq)
    for i, s in enumerate((6, 5, 5, 2)):
        p = f"model.{i + 2}"
        h = O.snake(h, sd[f"{p}.block.0.alpha"])
        h = O.sconvtr1d(h, sd, f"{p}.block.1.convtr.convtr", s)
        for j, d in enumerate((1, 3, 9)):
            h = O.residual_unit(h, sd, f"{p}.block.{j + 2}", d)
    h = O.snake(h, sd["model.6.alpha"])
    return torch.tanh(O.sconv1d(h, sd, "model.7.conv.conv"))


with torch.no_grad():
    z = O.encoder_forward(sds["encoder"], x)
    q = O.quantizer_forward(sds["quantizer"], z, x, n_c=2, return_codes=True)
    ident = lambda t: t
    f16 = lambda t: t.half().float()
    b16 = lambda t: t.bfloat16().float()
    y0 = decoder(sds["decoder"], q[0], ident, ident)
    yref = O.decoder_forward(sds["decoder"], q[0])
    print(f"manual loop vs torch LSTM: rms {float(((y0 - yref).double() ** 2).mean().sqrt()):.3e}; waveform rms {float(y0.pow(2).mean().sqrt()):.4f}")
    for name, wq, hq in (("W fp16, h fp32", f16, ident), ("W fp16, h fp16 (lstm2 one-pass class)", f16, f16),
                         ("W fp16, h bf16", f16, b16), ("W bf16, h bf16", b16, b16)):
        y = decoder(sds["decoder"], q[0], wq, hq)
        print(f"{name}: waveform rms error {float(((y - y0).double() ** 2).mean().sqrt()):.3e}  max {float((y - y0).abs().max()):.3e}")
