"""profiles/r01/*.csv (ncu exports made on the B200 box) -> profiles/SUMMARY_r01.md + profiles/roofline_r01.json"""
import collections, csv, glob, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "profiles", "r01")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6566.4, "bf16_tflops_sustained": 1439.1}
out = ["# ncu summary, round 1 (B200, `--clock-control none`)\n",
       "Source files: `profiles/r01/*_raw.csv` (`ncu --set full ... --page raw --csv`), `*_details.txt`, "
       "`launches_r01.csv` (`--metrics gpu__time_duration.sum`, one full B=32 x 4 s forward = 114 launches).\n",
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

* ConvTranspose 192->96 at T=48000, B=32: algorithmic bytes = x in 1.18 GB + y out 1.18 GB = 2.36 GB; measured DRAM
  traffic 2.32 GB, so nothing is re-read from HBM.  The conv kernels sit at 2-25 % of DRAM bandwidth and 19-53 % tensor-pipe
  activity: they are bound by the serial produce -> MMA -> epilogue phases of a CTA (one CTA per SM), not by HBM.
* SASS of the dominant kernels contains `UTCHMMA` (tcgen05.mma), `UBLKCP` (bulk TMA), `LDTM` (tcgen05.ld), `UTCBAR`
  (tcgen05.commit): `cuobjdump -sass facodec_b200/_C/libfacodec_b200.so | grep -E 'UTCHMMA|UBLKCP|LDTM|UTCBAR'`.
* Tuning history that these captures drove (details in DESIGN.md 4.1): I-cache thrash from unrolled sinf epilogues
  (stall_no_inst 24 %) -> pi-periodic sin^2 polynomial + rolled loops; 32-line-per-instruction row stores in the
  epilogue (~3000 cycles per 16-column group) -> shared-memory transpose, 4 lines per instruction; ELECT/BRA.U.ANY
  loop around every UTCHMMA issued from a divergent region (+75 cycles per MMA) -> converged warp + elect.sync.
""")
open(os.path.join(ROOT, "profiles", "SUMMARY_r01.md"), "w").write("\n".join(out) + "\n")
roof["conv_family"] = {"launches_per_forward": cl, "dram_bytes_per_forward": cb, "dram_bytes_per_launch": cb / max(1, cl),
                       "ms_under_ncu": cms, "share_of_forward_under_ncu": cms / tot if tot else None,
                       "source": "profiles/r01/launches_r01.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                                 "dram__bytes_write.sum --clock-control none, second forward of scripts/ncu_target.py)"}
json.dump(roof, open(os.path.join(ROOT, "profiles", "roofline_r01.json"), "w"), indent=1)
print("\n".join(out[-14:]))
