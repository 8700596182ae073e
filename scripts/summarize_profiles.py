"""profiles/r01/*.csv (ncu exports made on the B200 box) -> profiles/SUMMARY_r01.md + profiles/roofline_r01.json"""
import collections, csv, glob, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "profiles", "r01")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6566.4, "bf16_tflops_sustained": 1439.1}
out = ["# ncu summary, round 1 (B200, `--clock-control none`)\n",
       "Source files: `profiles/r01/*_raw.csv` (`ncu --set full ... --page raw --csv`), `*_details.txt`, "
       "`launches_r01.csv` (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`, two full B=32 x 4 s forwards of 115 launches; the warm one is summarised), `layers_eventtimed_final.txt` (CUDA-event time of every call site, no profiler: `scripts/gpu_layer_profile.py`), `bench_1gpu.json` / `bench_2gpu.json` (bench.py lines), `memcheck_smoke.log` (compute-sanitizer, 0 errors).\n",
       "Captures taken from `scripts/ncu_target.py` (second forward). Numbers under ncu are cold-cache and serialised: "
       "use SHARES, not absolutes; bench numbers come from `bench.py` only.\n"]
# launch list: ncu --csv "long" format, one row per (launch, metric)
allrows = list(csv.reader(open(os.path.join(D, "launches_r01.csv"))))
hdr = next(r for r in allrows if "Kernel Name" in r)
ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launch = collections.OrderedDict()
for r in allrows:
    if len(r) <= ix["Metric Value"] or not r[ix["ID"]].isdigit():
        continue
    L = launch.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1)
    L[r[ix["Metric Name"]]] = v
# the target runs two identical forwards: keep the second (warm) one
ids = sorted(launch)
if len(ids) % 2 == 0 and [launch[i]["name"] for i in ids[:len(ids) // 2]] == [launch[i]["name"] for i in ids[len(ids) // 2:]]:
    launch = collections.OrderedDict((i, launch[i]) for i in ids[len(ids) // 2:])
agg = collections.OrderedDict()
for L in launch.values():
    a = agg.setdefault(L["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += L.get("gpu__time_duration.sum", 0.0)
    a[2] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
out.append(f"\n## Launch list (one forward, {len(launch)} launches, {tot:.1f} ms under ncu)\n\n| kernel | launches | ms | share | DRAM read+write GB |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100*v[1]/tot:.1f} % | {v[2]/1e9:.2f} |")
convk = [k for k in agg if k.startswith("fac::conv_tc_kernel") or k.startswith("fac::conv_tcp_kernel") or k.startswith("conv_tc")]
cl = sum(agg[k][0] for k in convk); cb = sum(agg[k][2] for k in convk); cms = sum(agg[k][1] for k in convk)
out.append("\n## Full captures\n\n| capture | kernel / layer | grid | time ms | DRAM read+write (traffic) | DRAM % | tensor pipe active % | L1TEX % | L2 % | regs | issue-active % |\n|---|---|---|---|---|---|---|---|---|---|---|")
desc = {"tcp_c128k7": "conv_tcp_kernel: encoder conv7 C=128, T=48000",
        "lstm_dec_bf16": "lstm_rec_kernel<12,bf16>: decoder LSTM layer, H=1536, 320 steps",
        "tc_tf32_c64k1": "conv_tc_kernel<0,tf32>: encoder 1x1 conv C=64, T=96000 (short chain, 2 CTAs/SM)",
        "tc_bf16_c384k7": "conv_tc_kernel<0,bf16>: decoder conv7 C=384, T=9600 (2 CTAs/SM)",
        "tc_bf16_convtr192to96": "conv_tc_kernel<0,bf16>: decoder ConvTranspose 192->96 (x2), T 48000->96000 (launch 109)",
        "tc_bf16_c384k1": "conv_tc_kernel<0,bf16>: decoder ResidualUnit 1x1 conv C=384, T=9600 (launch 100)",
        "tc_bf16_convtr1536to768": "conv_tc_kernel<0,bf16>: decoder ConvTranspose 1536->768 (x6), T 320->1920 (launch 91)",
        "tc_fused_bf16_c192": "conv_tc_kernel<1,bf16>: decoder fused ResidualUnit C=192, T=48000 (launch 106)",
        "tc_fused_bf16_c96": "conv_tc_kernel<1,bf16>: decoder fused ResidualUnit C=96, T=96000 (launch 110)",
        "tcp_c64k7": "conv_tcp_kernel (3xTF32, promoted): encoder conv7 C=64, T=96000",
        "tcp_c512k7": "conv_tcp_kernel: encoder conv7 C=512, T=1920",
        "lstm_dec": "lstm_rec_kernel<12>: decoder LSTM layer, H=1536, 320 steps"}
roof = {}
for f in sorted(glob.glob(os.path.join(D, "prof_*_raw.csv"))):
    name = os.path.basename(f)[5:-12]
    rr = list(csv.reader(open(f)))
    d = dict(zip(rr[0], rr[2])); u = dict(zip(rr[0], rr[1]))
    def g(k, default="?"):
        return d.get(k, default)
    def gb(k):
        v = float(g(k, "0").replace(",", "")); unit = u.get(k, "")
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
    t_ms = float(g("gpu__time_duration.sum", "0")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u.get("gpu__time_duration.sum", "ms"), 1)
    traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    roof[name] = {"time_ms_under_ncu": t_ms, "dram_traffic_bytes": traffic,
                  "tensor_active_pct": float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "0")),
                  "dram_pct": float(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "0"))}
    out.append(f"| `{name}` | {desc.get(name, name)} | {g('Grid Size')} | {t_ms:.3f} | {traffic/1e9:.3f} GB | "
               f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','0')):.1f} | "
               f"{float(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','0')):.1f} | "
               f"{float(g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed','0')):.1f} | "
               f"{float(g('lts__throughput.avg.pct_of_peak_sustained_elapsed','0')):.1f} | {g('launch__registers_per_thread')} | "
               f"{float(g('smsp__issue_active.avg.pct_of_peak_sustained_active','0')):.1f} |")
out.append("""
## Reading

* **Traffic vs algorithmic bytes.** Fused decoder ResidualUnit, C = 192, T = 48000, B = 32: algorithmic bytes = x in
  1.18 GB + y out 1.18 GB = 2.36 GB (the residual re-read of x hits L2); measured DRAM traffic 2.32 GB. Encoder 1x1 conv,
  C = 64: in + residual + out = 3 x 0.786 = 2.36 GB, measured 2.33 GB at 50 % of DRAM bandwidth (0.57 ms). Nothing is
  re-read from HBM anywhere on the path; over the whole forward the two conv kernels move 65 GB in 84 launches
  (`roofline_r01.json: conv_family`) against 72 GB of per-layer algorithmic bytes (L2 keeps part of the residuals).
* **Tensor pipe.** Decoder conv7 at C = 384: 71 % tensor-pipe active, 634 GFLOP x 3 bf16 passes in 1.39 ms = 1.37 PFLOP/s
  = 95 % of the measured sustained bf16 peak (MEASURED_PEAKS.json). Encoder conv7 (3xTF32 = 6 bf16-equivalent passes):
  C = 512 / 256 / 128 run at 74-77 % of that peak, C = 64 at 50 %; `sm__pipe_tensor_cycles_active` reads 34-57 % there because
  a kind::tf32 MMA occupies the pipe at half rate per flop. The fused 96/192-channel units reach 31-38 %: their CTAs are
  bound by the worker warps (2.5-5 warps per scheduler, long_scoreboard + fixed-latency `wait` stalls dominate, see below).
* SASS of the dominant kernels contains `UTCHMMA` (tcgen05.mma), `UBLKCP` (bulk TMA), `LDTM` (tcgen05.ld), `UTCBAR`
  (tcgen05.commit), `USETMAXREG`: `cuobjdump -sass facodec_b200/_C/libfacodec_b200.so | grep -E 'UTCHMMA|UBLKCP|LDTM|UTCBAR|USETMAXREG'`.
* **What the source-level captures (`--import-source on`, `--page source`) drove this round** (details in DESIGN.md 4.1):
  - unrolled sinf epilogues thrashed the instruction cache (stall_no_inst 24 %) -> pi-periodic sin^2 polynomial, rolled loops;
  - row-per-lane epilogue stores touched 32 lines per instruction (~3000 cycles per 16 columns) -> shared-memory transpose;
  - every UTCHMMA issued from a divergent region sat in an ELECT / BRA.U.ANY loop (+75 cycles per MMA) -> converged warp + elect.sync;
  - fused C = 192 unit, one CTA per SM: 28.8 % issue-active, 0.35 eligible warps per scheduler, workers stalled on
    long_scoreboard (34 %) and `wait` (27 %), producers 19 % and epilogue warps 8 % of the time waiting for MMAs; weights never
    waited for (b_full spin count 0) -> two-CTA-per-SM plan for every tile that fits 256 TMEM columns, resident GEMM-2 operand,
    SFU sine in the bf16-class kernels, one range check per 4 channels;
  - promoted kernel: accumulator warps waited 67 % of a CTA's life, producers idle during the epilogue -> persistent CTAs;
    ~60k warp-instructions of mbarrier polling per tile; wait-time probes showed the MMA warp busy 85 % of a tile at
    ~(A + B bytes) / 64 B per clock per MMA; register re-balancing (control 48 / producers 56 / accumulators 160) removed the
    spills that a 32-register MMA warp had in its issue loop (conv7 C = 64: 1.74 -> 1.48 ms).
* `profiles/r01_mid/` and `profiles/r01a/` hold the captures taken earlier in the round (before the two-CTA plan, the
  persistent kernel and the bf16 LSTM) for comparison.
""")
open(os.path.join(ROOT, "profiles", "SUMMARY_r01.md"), "w").write("\n".join(out) + "\n")
roof["conv_family"] = {"launches_per_forward": cl, "dram_bytes_per_forward": cb, "dram_bytes_per_launch": cb / max(1, cl),
                       "ms_under_ncu": cms, "share_of_forward_under_ncu": cms / tot if tot else None,
                       "source": "profiles/r01/launches_r01.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                                 "dram__bytes_write.sum --clock-control none, second forward of scripts/ncu_target.py)"}
json.dump(roof, open(os.path.join(ROOT, "profiles", "roofline_r01.json"), "w"), indent=1)
print("\n".join(out[-14:]))
