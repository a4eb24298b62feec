set -uo pipefail
tag=r2; out=gpurun_out
for task in dmm bmnist; do
  mcmd="python benchmarks/bench_models.py --no-cpu --tasks $task"
  $mcmd > "$out/${tag}_${task}.json" 2>&1 || { echo "$task bench failed without ncu"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 1500 --csv --log-file "$out/${task}_launches.csv" $mcmd > "$out/${tag}_ncu_${task}_list.log" 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_dmm|k_template_match|k_linear_tc5|k_linear_reduce|k_coor_heads|k_stn|k_logw_wh' --launch-skip 100 --launch-count 30 -f -o "$out/${tag}_${task}" $mcmd > "$out/${tag}_ncu_${task}.log" 2>&1
  ncu -i "$out/${tag}_${task}.ncu-rep" --page raw --csv > "$out/${tag}_raw_${task}.csv" 2>/dev/null
  ncu -i "$out/${tag}_${task}.ncu-rep" --page source --csv -k regex:'k_template_match|k_dmm_local_fused|k_linear_tc5' > "$out/${tag}_src_${task}.csv" 2>/dev/null
done
rm -f "$out"/*.ncu-rep
ls -la $out | tail -15
