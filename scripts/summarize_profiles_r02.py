"""profiles/r02/*.csv (ncu exports made on the B200 box) -> profiles/SUMMARY_r02.md + profiles/roofline_r02.json"""
import collections, csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "profiles", "r02")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6566.4, "bf16_tflops_sustained": 1439.1}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}
LIST = "launches_r02f.csv" if os.path.exists(os.path.join(D, "launches_r02f.csv")) else "launches_r02.csv"
rows = list(csv.reader(open(os.path.join(D, LIST))))
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launch = collections.OrderedDict()
for r in rows:
    if len(r) <= ix["Metric Value"] or not r[ix["ID"]].isdigit():
        continue
    L = launch.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")})
    L[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1)
ids = sorted(launch)
half = len(ids) // 2                       # the target runs two forwards: keep the second (warm) one
launch = collections.OrderedDict((i, launch[i]) for i in ids[half:])
agg = collections.OrderedDict()
for L in launch.values():
    a = agg.setdefault(L["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += L.get("gpu__time_duration.sum", 0.0)
    a[2] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
out = ["# ncu summary, round 2 (B200, `--clock-control none`)\n",
       f"Sources: `profiles/r02/{LIST}` (the FINAL build of the round; `launches_r02.csv` is the mid-round list) (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`, two "
       "full B=32 x 4 s forwards, the warm one summarised), `prof_*_r02_raw.csv` / `_details.txt` (`ncu --set full --import-source on`, "
       "one launch each: `scripts/ncu_capture_r02f.sh` for the `_r02f` captures of the final build, `scripts/ncu_capture_r02.sh` for the mid-round "
       "`_r02` ones), `layers_eventtimed_r02f.txt` (CUDA events around every call site of the final build, no profiler; `_r02.txt` mid-round), "
       "`tc_chunk_trace_before_r02.log` / `tc_chunk_trace_r02f.log` (per-chunk clock64 timeline of one conv_tc CTA before / after the issue-loop "
       "rework), `tma_probe_r02.log`, `tt_role_probes_r02.log`, `lstm2_phase_clocks_r02.log`, `mma_probe_r02.log`, `bench_1gpu_r02e/f.json`, "
       "`bench_2gpu_r02e.json`, `bench_vq_r02e.json`, `bench_trainfwd_r02.json`, `gpu_tests_r02g.log`.  Numbers under ncu are cold-cache and "
       "serialised (and the quantizer front runs on a second stream in production): compare SHARES; bench numbers come from bench.py only.\n",
       f"\n## Launch list (one forward, {len(launch)} launches, {tot:.1f} ms under ncu)\n\n| kernel | launches | ms | share | DRAM read+write GB |\n|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % | {v[2] / 1e9:.2f} |")
fam = [k for k in agg if "conv_tc_kernel" in k or "conv_tt_kernel" in k or "conv_tcp_kernel" in k]
cl = sum(agg[k][0] for k in fam); cb = sum(agg[k][2] for k in fam); cms = sum(agg[k][1] for k in fam)
DESC_F = {"tc_fused_c192": "conv_tc_kernel<FUSED, BF16, G1F16, 16 workers>: decoder ResidualUnit C=192, d=3, T=48000 (B=32), final build",
          "tc_fused_c96": "conv_tc_kernel<FUSED, BF16, G1F16, 8 workers>: decoder ResidualUnit C=96, d=3, T=96000, two CTAs per SM, final build",
          "tc_c384k7": "conv_tc_kernel<BF16, G1F16>: decoder conv7 C=384 (one fp16 pass), T=9600, final build",
          "tt_c128k7": "conv_tt_kernel<SNAKE>: encoder conv7 C=128, d=3, T=48000, final build (kernel unchanged)",
          "lstm2_enc": "lstm_rec2_kernel<8, 3-pass>: encoder LSTM layer, final build (kernel unchanged)"}
DESC = {"tt_c128k7": "conv_tt_kernel<SNAKE>: encoder conv7 C=128, d=3, T=48000 (B=32)", "tt_c64k7": "conv_tt_kernel<SNAKE>: encoder conv7 C=64, d=3, T=96000",
        "tt_c512k7": "conv_tt_kernel<SNAKE>: encoder conv7 C=512, T=1920", "tt_c256k1": "conv_tt_kernel<NONE>: encoder 1x1 conv C=256 + residual, T=9600",
        "lstm2_enc": "lstm_rec2_kernel<8, 3-pass>: encoder LSTM layer, H=1024, 320 steps", "lstm2_dec": "lstm_rec2_kernel<12, 1-pass fp16>: decoder LSTM layer, H=1536",
        "rvq": "rvq_kernel: ResidualVQ 4 x 1024 entries, 2^18 frames, 4 frames per warp"}
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dr"), ("dram__bytes_write.sum", "dw"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor"),
        ("sm__inst_executed_pipe_tensor.sum", "tinst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "inst"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2")]
out.append("\n## Full captures\n\n| capture | kernel / layer | grid | time ms | DRAM read+write | DRAM % | tensor pipe active % | issue active % | L1TEX % | L2 % | regs | warp-instr (M) |\n|---|---|---|---|---|---|---|---|---|---|---|---|")
caps = {}
for name, desc, suffix in [(n, d, "r02f") for n, d in DESC_F.items()] + [(n, d + " (mid-round build)", "r02") for n, d in DESC.items()]:
    pth = os.path.join(D, f"prof_{name}_{suffix}_raw.csv")
    if not os.path.exists(pth):
        continue
    name = f"{name}_{suffix}"
    rr = list(csv.reader(open(pth)))
    h = rr[0]; units = rr[1]; val = rr[2]
    m = {}
    for key, short in WANT:
        if key in h:
            i = h.index(key)
            try:
                m[short] = float(val[i].replace(",", "")) * UNIT.get(units[i], 1)
            except ValueError:
                pass
    caps[name] = m
    g = lambda k, f="{:.1f}": f.format(m[k]) if k in m else "-"
    out.append(f"| `{name}` | {desc} | {g('grid', '{:.0f}')} | {g('time', '{:.3f}')} | {(m.get('dr', 0) + m.get('dw', 0)) / 1e9:.3f} GB | {g('dram_pct')} | "
               f"{g('tensor')} | {g('issue')} | {g('l1')} | {g('l2')} | {g('regs', '{:.0f}')} | {m.get('inst', 0) / 1e6:.1f} |")
open(os.path.join(ROOT, "profiles", "SUMMARY_r02.md"), "w").write("\n".join(out) + "\n")
json.dump({"conv_family": {"kernels": fam, "launches_per_forward": cl, "dram_bytes_per_forward": cb, "dram_bytes_per_launch": cb / max(1, cl),
                           "ncu_ms_per_forward": cms, "share_of_forward": cms / tot,
                           "source": "profiles/r02/launches_r02.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                                     "--clock-control none, second forward of scripts/ncu_target.py)"},
           "captures": caps, "peaks": peaks}, open(os.path.join(ROOT, "profiles", "roofline_r02.json"), "w"), indent=1)
print("\n".join(out))
