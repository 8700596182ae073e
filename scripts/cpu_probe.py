"""Host CPU probe on the GPU box: cores visible, cgroup quota, oracle throughput vs thread count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
os.system("lscpu | grep -E 'Model name|Socket|Core|Thread' | head -5")
from facodec_b200 import synth
from oracle import facodec_oracle as O
sds = synth.synth_state_dicts(0)
for nt in (8, 16, 32, 64, 128):
    if nt > os.cpu_count():
        break
    torch.set_num_threads(nt)
    x = synth.synth_waves(2, 96000)
    O.codec_forward(sds, x[:1], n_c=2)
    t = time.time(); O.codec_forward(sds, x, n_c=2); dt = time.time() - t
    print(f"threads {nt}: B=2 x 4 s in {dt:.2f} s = {8/dt:.2f} audio-s/s", flush=True)
