"""Turn the scratch ncu output in gpurun_out/ into the tracked summaries under profiles/.

    python benchmarks/summarize_profiles.py <round-tag>
reads  gpurun_out/bench_launches.csv       ncu --metrics gpu__time_duration.sum launch list of bench.py
       gpurun_out/<tag>_raw_bench.csv      `ncu -i prof.ncu-rep --page raw --csv` of the --set full capture inside bench.py
       gpurun_out/<tag>_raw_probe.csv      the same for benchmarks/probe_gather.py (all-distinct resample probe)
       gpurun_out/<tag>_raw_train.csv      the same for benchmarks/zbwd_bench.py (training-mode kernels), if present
       gpurun_out/<tag>_raw_{dmm,bmnist}.csv  the same for benchmarks/bench_models.py --tasks dmm | bmnist (C3 / C4 kernels), if present
(the .ncu-rep files are exported to CSV on the GPU box: they exceed what gpurun copies back)
writes profiles/<tag>_launch_list.md, profiles/<tag>_ncu_kernels.md, profiles/ncu_bench_traffic.json
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)

# ---------------------------------------------------------------- launch list
rows = [r for r in csv.reader(open(os.path.join(GO, "bench_launches.csv"))) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")[:80]
    agg.setdefault(name, []).append(float(r[14]) / 1e3)
tot = sum(sum(v) for v in agg.values())
ours = sum(sum(v) for k, v in agg.items() if "apgs::" in k)
with open(os.path.join(PR, "%s_launch_list.md" % tag), "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras` (%s)\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv` on a B200; per-launch times are "
            "cold-cache and serialised, so compare SHARES.  %d launches captured, %.1f ms of kernel time, "
            "%.1f %% of it in this repo's kernels (`apgs::*`).\n\n" % (len(rows), tot / 1e3, 100 * ours / tot))
    f.write("| kernel | launches | median us | total ms | share |\n|---|---:|---:|---:|---:|\n")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        v2 = sorted(v)
        f.write("| `%s` | %d | %.1f | %.2f | %.1f %% |\n" % (name, len(v), v2[len(v2) // 2], sum(v) / 1e3, 100 * sum(v) / tot))

# ---------------------------------------------------------------- full captures
WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("smsp__inst_executed.sum", "warp instructions"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA")]


def load(path):
    rr = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rr) if r and r[0] == "ID"][0]
    return rr[hi], rr[hi + 1], [r for r in rr[hi + 2:] if r and r[0].isdigit()]


def val(r, hdr, units, m):
    v, u = float(r[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}.get(u, 1)


traffic = {}
with open(os.path.join(PR, "%s_ncu_kernels.md" % tag), "w") as f:
    f.write("# ncu --set full captures (%s)\n\n`ncu --set full --clock-control none --import-source on -k regex:... <cmd>` on a "
            "B200, exported on the box with `ncu -i ... --page raw --csv`.  One capture per kernel (the first after the skipped "
            "warm-up launches).  DRAM bytes are per launch.\n\n" % tag)
    for part, cmd in (("bench", "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras  (GMM S=1000, B=1024, in-step kernels)"),
                      ("probe", "python benchmarks/probe_gather.py  (all-distinct resample probe: reference layout 1540 B per particle, "
                                "near-uniform weights, L2 flushed)"),
                      ("train", "python benchmarks/zbwd_bench.py  (training-mode kernels at S=1000, B=1024)"),
                      ("dmm", "python benchmarks/bench_models.py --no-cpu --tasks dmm  (C3 DMM S=100, B=256, 7 sweeps)"),
                      ("bmnist", "python benchmarks/bench_models.py --no-cpu --tasks bmnist  (C4 bmnist S=10, B=64, T=10, 5 sweeps)")):
        path = os.path.join(GO, "%s_raw_%s.csv" % (tag, part))
        if not os.path.exists(path):
            continue
        hdr, units, body = load(path)
        f.write("# `%s`\n\n" % cmd)
        seen = set()
        for r in body:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            f.write("## `%s`\n\n| metric | value |\n|---|---|\n" % name)
            for m, label in WANT:
                if m in hdr:
                    f.write("| %s | %s %s |\n" % (label, r[hdr.index(m)], units[hdr.index(m)]))
            dr, dw, us = val(r, hdr, units, "dram__bytes_read.sum"), val(r, hdr, units, "dram__bytes_write.sum"), val(r, hdr, units, "gpu__time_duration.sum")
            f.write("| DRAM bytes / duration | %.0f GB/s |\n\n" % ((dr + dw) / us / 1e3))
            short = ("probe:" if part == "probe" else "") + name.split("(")[0].replace("void ", "").split("<")[0]
            traffic.setdefault(short, {"dram_bytes_per_launch": dr + dw, "duration_us": us, "capture": cmd.split("  ")[0]})
json.dump(traffic, open(os.path.join(PR, "ncu_bench_traffic.json"), "w"), indent=1)
print("wrote profiles for", tag, json.dumps(traffic, indent=1))
