set -uo pipefail
out=gpurun_out; tag=r2
python -m pytest tests -m gpu -x -q > $out/f_tests.log 2>&1; echo "tests rc=$?" | tee -a $out/f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/f_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $out/f_bench.json 2> $out/f_bench.err; echo "bench rc=$?"
python bench.py --impl reference > $out/f_ref.json 2> $out/f_ref.err; echo "ref rc=$?"
python benchmarks/linear_bench.py > $out/f_linear.json 2>&1
python benchmarks/match_bench.py > $out/f_match.json 2>&1
task=dmm
mcmd="python benchmarks/bench_models.py --no-cpu --tasks $task"
$mcmd > "$out/${tag}_${task}.json" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 1500 --csv --log-file "$out/${task}_launches.csv" $mcmd > "$out/${tag}_ncu_${task}_list.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dmm|k_mlp_heads' --launch-skip 100 --launch-count 30 -f -o "$out/${tag}_${task}" $mcmd > "$out/${tag}_ncu_${task}.log" 2>&1
ncu -i "$out/${tag}_${task}.ncu-rep" --page raw --csv > "$out/${tag}_raw_${task}.csv" 2>/dev/null
rm -f "$out"/*.ncu-rep
tail -2 $out/f_tests.log
