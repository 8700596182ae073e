"""Top-1 / top-2 margins of every VQ decision of the BASELINE configs[1] batch (32 x 4 s, synthetic weights seed 0), on the
CPU oracle: how far is each of the 61 440 argmin decisions from a tie?  (A decision whose margin is below the fp32 noise
of a re-associated sum is not reproducible even between CPU thread counts.)  python scripts/vq_margins.py [n_utts]"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facodec_b200 import synth
from oracle import facodec_oracle as O
from oracle.make_golden import vq_margin_report

torch.set_num_threads(os.cpu_count() or 1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sds = synth.synth_state_dicts(0)
sq = sds["quantizer"]
x = synth.synth_waves(n, 96000)
margins = {k: [] for k in ("prosody.0", "content.0", "content.1", "residual.0", "residual.1", "residual.2")}
with torch.no_grad():
    for i in range(0, n, 4):
        xw = x[i:i + 4]
        z = O.encoder_forward(sds["encoder"], xw)
        pf = O.mel_preprocess(sq, xw, n_bins=20)
        f0 = O.sconv1d(O.wavenet(sq, O.sconv1d(pf, sq, "melspec_linear.conv.conv")), sq, "melspec_linear2.conv.conv")
        margins["prosody.0"].append(vq_margin_report(sq, "prosody_quantizer.quantizers.0", f0))
        zp = O.residual_vq(sq, "prosody_quantizer", f0, 1)[0]
        res = z
        zc = 0
        for j in range(2):
            margins[f"content.{j}"].append(vq_margin_report(sq, f"content_quantizer.quantizers.{j}", res))
            out = O.vector_quantize(sq, f"content_quantizer.quantizers.{j}", res)[0]
            res = res - out
            zc = zc + out
        res = z - zp - zc
        for j in range(3):
            margins[f"residual.{j}"].append(vq_margin_report(sq, f"residual_quantizer.quantizers.{j}", res))
            res = res - O.vector_quantize(sq, f"residual_quantizer.quantizers.{j}", res)[0]
print(f"top-1 / top-2 margin of -dist (normalised 8-dim codes, dist in [0, 4]); {n} utterances x 320 frames per quantizer")
edges = [0, 1e-7, 3e-7, 1e-6, 3e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e9]
print("quantizer      min        p0.1%      median   | decisions per margin bin " + " ".join(f"<{e:g}" for e in edges[1:-1]) + " >=1e-2")
tot = 0
for k, v in margins.items():
    m = torch.cat([t.flatten() for t in v]).numpy()
    tot += m.size
    hist = np.histogram(m, bins=edges)[0]
    print(f"{k:12s} {m.min():.3e}  {np.percentile(m, 0.1):.3e}  {np.median(m):.3e} | " + " ".join(f"{h:6d}" for h in hist))
print("decisions:", tot)
