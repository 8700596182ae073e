"""B=32 x 4 s forward time + code equality vs the default path for a list of FACODEC option sets: python scripts/gpu_opt_compare.py "a=1,b=2" "c=1" ..."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import facodec_b200 as fb
from facodec_b200 import synth
sds = synth.synth_state_dicts(0)
m = fb.build_model()
for k in ("encoder", "quantizer", "decoder"):
    m[k].load_state_dict(sds[k]); m[k].eval()
codec = fb.Codec(m); eng = codec.engine
x = synth.synth_waves(32, 96000).cuda()
def run(n=5):
    codec.forward(x, n_c=2); codec.forward(x, n_c=2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): out = codec.forward(x, n_c=2)
    b.record(); torch.cuda.synchronize()
    return out, a.elapsed_time(b) / n
DEFAULTS = {"encoder_tt": 1, "lstm_v2": 1, "decoder_lstm_fp16": 1, "overlap_front": 1, "fuse_resunit": 1, "decoder_bf16": 1, "tc_occ2_maxn": 256,
            "tensor_cores": 2, "decoder_conv7_fp16": 1, "tc_wide": 1, "tc_slot_issue": 1, "tt_pair": 1, "tc_groups": 1}
(y0, c0, t0), ms0 = run()
print(f"default: {ms0:.2f} ms/step")
for spec in sys.argv[1:]:
    kv = [s.split("=") for s in spec.split(",") if s]
    # interleaved rounds (spec, default, spec, default, ...): run-to-run drift inside one process is ~0.5 ms
    ms_s, ms_d = [], []
    for r in range(3):
        for k, v in kv: eng.set_option(k, int(v))
        (y, c, t), ms = run(8)
        ms_s.append(ms)
        for k, v in kv: eng.set_option(k, DEFAULTS.get(k, 0))
        ms_d.append(run(8)[1])
    diff = sum(int((a != b).sum()) for a, b in zip(c, c0))
    rms = float(((y.double() - y0.double()) ** 2).mean().sqrt())
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"{spec}: {med(ms_s):.2f} ms/step (rounds {' '.join(f'{m:.2f}' for m in ms_s)}) vs default {med(ms_d):.2f} ({' '.join(f'{m:.2f}' for m in ms_d)}); "
          f"codes differing from default {diff}; waveform rms diff {rms:.2e}")
