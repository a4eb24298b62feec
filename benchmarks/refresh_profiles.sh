#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash benchmarks/refresh_profiles.sh r2'): the captures behind profiles/<tag>_*.md.
# Every profiled command first runs to completion WITHOUT ncu; numbers printed under ncu are never bench values.
# Afterwards, in the build container:  python benchmarks/summarize_profiles.py <tag>
set -uo pipefail
tag="${1:-r2}"
out=gpurun_out
mkdir -p "$out"
cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$cmd > "$out/${tag}_plain_bench.json" 2> "$out/${tag}_plain_bench.err" || { echo "bench failed without ncu"; exit 1; }
# 1. launch list (cold-cache, serialised per-launch times: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file "$out/bench_launches.csv" $cmd > "$out/${tag}_ncu_list.log" 2>&1
# 2. --set full of this repo's kernels inside the step (one sweep's worth of launches after the warm-up passes)
ncu --set full --clock-control none --import-source on -k regex:'k_gmm|k_gather|k_ancestors' --launch-skip 300 --launch-count 12 -f -o "$out/${tag}_bench" $cmd > "$out/${tag}_ncu_full.log" 2>&1
ncu -i "$out/${tag}_bench.ncu-rep" --page raw --csv > "$out/${tag}_raw_bench.csv" 2>/dev/null
# 3. the all-distinct resample probe (HBM point of the gather)
python benchmarks/probe_gather.py > "$out/${tag}_probe.json" 2>&1 && \
ncu --set full --clock-control none -k regex:'k_gather|k_ancestors' --launch-skip 4 --launch-count 2 -f -o "$out/${tag}_probe" python benchmarks/probe_gather.py > "$out/${tag}_ncu_probe.log" 2>&1
ncu -i "$out/${tag}_probe.ncu-rep" --page raw --csv > "$out/${tag}_raw_probe.csv" 2>/dev/null
# 4. training-mode kernels
python benchmarks/zbwd_bench.py > "$out/${tag}_zbwd.json" 2>&1 && \
ncu --set full --clock-control none -k regex:'k_gmm_z_logq_bwd|k_label_table|k_gmm_z_fused' --launch-skip 8 --launch-count 4 -f -o "$out/${tag}_train" python benchmarks/zbwd_bench.py > "$out/${tag}_ncu_train.log" 2>&1
ncu -i "$out/${tag}_train.ncu-rep" --page raw --csv > "$out/${tag}_raw_train.csv" 2>/dev/null
# 5. the dominant kernels of the other two tasks (C3 DMM, C4 bmnist): launch list + one capture per kernel name
for task in dmm bmnist; do
  mcmd="python benchmarks/bench_models.py --no-cpu --tasks $task"
  $mcmd > "$out/${tag}_${task}.json" 2>&1 || { echo "$task bench failed without ncu"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 1500 --csv --log-file "$out/${task}_launches.csv" $mcmd > "$out/${tag}_ncu_${task}_list.log" 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_dmm|k_template_match|k_linear_tc5|k_linear_reduce|k_coor_heads|k_stn|k_logw_wh' --launch-skip 100 --launch-count 30 -f -o "$out/${tag}_${task}" $mcmd > "$out/${tag}_ncu_${task}.log" 2>&1
  ncu -i "$out/${tag}_${task}.ncu-rep" --page raw --csv > "$out/${tag}_raw_${task}.csv" 2>/dev/null
done
# the reports themselves exceed what gpurun copies back; the CSV exports are what is summarised
rm -f "$out"/*.ncu-rep
ls -la "$out"
