#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches.csv into the committed, text-form evidence under profiles/.
   python scripts/make_profile_summary.py r01"""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
UNITS = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2]


lines = [f"# ncu summaries, {tag} (one launch each, `ncu --set full --clock-control none`, scripts/collect_profiles.sh)", ""]
traffic = {}
names = {"sws_fate": "sws_vscale_rgb24_kernel", "sws_lut": "sws_unscaled_kernel", "idct_put": "idct8x8_put", "idct_add": "idct8x8_add",
         "tx_fft": "tx_fft_kernel", "qpel": "qpel_kernel", "chroma": "chroma_kernel", "esa": "esa_kernel"}
work = {"sws_fate": ("32 x 4K frames", 32 * 37324800), "sws_lut": ("32 x 4K frames", 32 * 37324800),
        "idct_put": ("32 x 48960 blocks", 32 * 48960 * 192), "idct_add": ("32 x 48960 blocks", 32 * 48960 * 256),
        "tx_fft": ("65536 x FFT-1024", 65536 * 16384), "qpel": ("16 x 8160 16x16 blocks", 16 * 8160 * 825), "chroma": ("32 x 8160 8x8 blocks", 32 * 8160 * 177),
        "esa": ("1 4K pair", None)}
for key in names:
    rep = os.path.join(G, f"{tag}_{key}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, vals = raw(rep)
    kn = vals[hdr.index('Kernel Name')]
    lines.append(f"## {key}: `{kn}`  ({work[key][0]})")
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (vals[i], units[i])
            lines.append(f"    {k:85s} {vals[i]:>16s} {units[i]}")
    rd = float(d['dram__bytes_read.sum'][0]) * UNITS.get(d['dram__bytes_read.sum'][1], 1)
    wr = float(d['dram__bytes_write.sum'][0]) * UNITS.get(d['dram__bytes_write.sum'][1], 1)
    t_us = float(d['gpu__time_duration.sum'][0]) * (1000 if d['gpu__time_duration.sum'][1] == 'ms' else 1)
    lines.append(f"    -> dram traffic {(rd + wr) / 1e6:.1f} MB per launch" + (f", algorithmic {work[key][1] / 1e6:.1f} MB, ratio {(rd + wr) / work[key][1]:.3f}" if work[key][1] else ""))
    if work[key][1]:
        lines.append(f"    -> under ncu (cold, serialised): {work[key][1] / t_us / 1e3:.0f} GB/s algorithmic")
    lines.append("")
    traffic[names[key]] = {"dram_bytes_per_launch": rd + wr, "launch": work[key][0], "algorithmic_bytes": work[key][1]}
open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
# bench.py's roofline.traffic: per launch of the BENCH workload (256 frames) scaled from the 32-frame capture
tj = {"sws_vscale_rgb24_kernel": traffic.get("sws_vscale_rgb24_kernel", {}).get("dram_bytes_per_launch", 0) * 8 or None,
      "sws_unscaled_kernel": traffic.get("sws_unscaled_kernel", {}).get("dram_bytes_per_launch", 0) * 8 or None,
      "idct8x8_put": traffic.get("idct8x8_put", {}).get("dram_bytes_per_launch", 0) * 8 or None,
      "idct8x8_add": traffic.get("idct8x8_add", {}).get("dram_bytes_per_launch", 0) * 8 or None,
      "_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at 32 frames, x8 for the 256-frame bench launch"}
json.dump(tj, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
# launch list of the bench command
lc = os.path.join(G, "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# launch list of `python bench.py --steps 2 --warmup 1` under ncu ({tag}): gpu__time_duration.sum per kernel", "",
           "kernel | launches | total ns | share", "---|---|---|---"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"`{k}` | {n} | {t:.0f} | {t / tot:.3f}")
    out += ["", "(cold-cache, serialised replays: compare shares, not absolutes; torch's own fill/copy kernels are the synthetic-input set-up)"]
    open(os.path.join(P, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(lc, os.path.join(P, f"{tag}_launches.csv"))
print(open(os.path.join(P, f"{tag}_ncu_summary.md")).read()[:6000])
