"""CPU experiment behind the decoder's precision classes (conv_tc.cu): how far does the reconstructed waveform move when the
operands (activations AND weights) of chosen decoder convolutions are rounded to ONE fp16 (or bf16) value instead of the
fp32-faithful hi/lo pairs?  fp32 accumulation, oracle decoder, bar = 1e-4 waveform RMS (BASELINE north_star).
python scripts/cpu_decoder_precision.py [seconds] [seed]"""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facodec_b200 import synth
from oracle import facodec_oracle as O

torch.set_num_threads(os.cpu_count() or 1)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sds = synth.synth_state_dicts(seed)
x = synth.synth_waves(2, int(24000 * secs) // 300 * 300, seed=3 + seed)

ident = lambda t: t
f16 = lambda t: t.half().float()
b16 = lambda t: t.bfloat16().float()


def conv(x, sd, prefix, q, dilation=1):
    """O.sconv1d (stride 1, causal) with both operands passed through q."""
    w = O._wn_weight(sd, prefix)
    k = w.shape[-1]
    pad = (k - 1) * dilation
    xp = O._pad1d_reflect(x, pad, 0) if pad else x
    return F.conv1d(q(xp), q(w), sd[prefix + ".bias"], dilation=dilation)


def convtr(x, sd, prefix, stride, q):
    w = O._wn_weight(sd, prefix)
    k = w.shape[-1]
    y = F.conv_transpose1d(q(x), q(w), sd[prefix + ".bias"], stride=stride)
    return y[..., : y.shape[-1] - (k - stride)]


def lstm(x, sd, prefix, q):
    """SLSTM with the input projections' operands through q; recurrence as torch computes it."""
    xt = x.permute(2, 0, 1)
    inp = xt
    T, B, H = xt.shape
    for l in range(2):
        Wih, Whh = sd[f"{prefix}.weight_ih_l{l}"], sd[f"{prefix}.weight_hh_l{l}"]
        xg = q(inp) @ q(Wih).t() + sd[f"{prefix}.bias_ih_l{l}"] + sd[f"{prefix}.bias_hh_l{l}"]
        h = torch.zeros(B, H); c = torch.zeros(B, H); outs = []
        for t in range(T):
            i, f, g, o = (xg[t] + h @ Whh.t()).chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs)
    return (inp + xt).permute(1, 2, 0)


def decoder(sd, z, cfg):
    """cfg: dict layer-class -> quantizer.  classes: conv0, ih, up{i}, c7_{i}, c1_{i}, out"""
    g = lambda k: cfg.get(k, ident)
    h = conv(z, sd, "model.0.conv.conv", g("conv0"))
    h = lstm(h, sd, "model.1.lstm", g("ih"))
    for i, s in enumerate((6, 5, 5, 2)):
        p = f"model.{i + 2}"
        h = O.snake(h, sd[f"{p}.block.0.alpha"])
        h = convtr(h, sd, f"{p}.block.1.convtr.convtr", s, g(f"up{i}"))
        for j, d in enumerate((1, 3, 9)):
            pp = f"{p}.block.{j + 2}"
            y = O.snake(h, sd[pp + ".block.0.alpha"])
            y = conv(y, sd, pp + ".block.1.conv.conv", g(f"c7_{i}"), dilation=d)
            y = O.snake(y, sd[pp + ".block.2.alpha"])
            y = conv(y, sd, pp + ".block.3.conv.conv", g(f"c1_{i}"))
            h = h + y
    h = O.snake(h, sd["model.6.alpha"])
    return torch.tanh(conv(h, sd, "model.7.conv.conv", g("out")))


def classes(pred):
    names = ["conv0", "ih", "out"] + [f"{k}{i}" for i in range(4) for k in ("up", "c7_", "c1_")]
    return [n for n in names if pred(n)]


with torch.no_grad():
    z = O.encoder_forward(sds["encoder"], x)
    qz = O.quantizer_forward(sds["quantizer"], z, x, n_c=2, return_codes=True)[0]
    y0 = decoder(sds["decoder"], qz, {})
    yref = O.decoder_forward(sds["decoder"], qz)
    err = lambda a, b: float(((a - b).double() ** 2).mean().sqrt())
    print(f"{secs} s x 2 utterances, weight seed {seed}: restated decoder vs oracle rms {err(y0, yref):.2e}; waveform rms {float(y0.pow(2).mean().sqrt()):.4f}")
    per_utt = lambda a, b: [f"{float(((a[i] - b[i]).double() ** 2).mean().sqrt()):.2e}" for i in range(a.shape[0])]
    sets = [
        ("every tensor-core conv, one fp16 pass", classes(lambda n: n != "out")),
        ("every tensor-core conv, one bf16 pass", None),
        ("k=7 convs only", classes(lambda n: n.startswith("c7_"))),
        ("ResidualUnits (k=7 + 1x1) only", classes(lambda n: n.startswith("c7_") or n.startswith("c1_"))),
        ("ResidualUnits of blocks 3+4 (C=192/96)", ["c7_2", "c1_2", "c7_3", "c1_3"]),
        ("k=7 of blocks 3+4", ["c7_2", "c7_3"]),
        ("ResidualUnits of blocks 1+2 (C=768/384)", ["c7_0", "c1_0", "c7_1", "c1_1"]),
        ("transposed convs only", classes(lambda n: n.startswith("up"))),
        ("conv0 + LSTM input projections", ["conv0", "ih"]),
    ]
    for i in range(4):
        sets.append((f"ResidualUnits of block {i + 1} only", [f"c7_{i}", f"c1_{i}"]))
    for name, cl in sets:
        if cl is None:
            cfg = {n: b16 for n in classes(lambda n: n != "out")}
        else:
            cfg = {n: f16 for n in cl}
        y = decoder(sds["decoder"], qz, cfg)
        print(f"{name}: waveform rms error {err(y, y0):.3e}  per utterance {per_utt(y, y0)}  max {float((y - y0).abs().max()):.2e}")
