# Runs ON THE GPU BOX: tests + smoke + both bench arms as the driver runs them, then the C2 launch list / ncu captures.
set -uo pipefail
out=gpurun_out; tag=r2
python -m pytest tests -m gpu -x -q > $out/f_tests.log 2>&1; echo "tests rc=$?" | tee -a $out/f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/f_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $out/f_bench.json 2> $out/f_bench.err; echo "bench rc=$?"
python bench.py --impl reference > $out/f_ref.json 2> $out/f_ref.err; echo "ref rc=$?"
cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$cmd > "$out/${tag}_plain_bench.json" 2> "$out/${tag}_plain_bench.err" || { echo "bench failed without ncu"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file "$out/bench_launches.csv" $cmd > "$out/${tag}_ncu_list.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_gmm|k_gather|k_ancestors' --launch-skip 300 --launch-count 12 -f -o "$out/${tag}_bench" $cmd > "$out/${tag}_ncu_full.log" 2>&1
ncu -i "$out/${tag}_bench.ncu-rep" --page raw --csv > "$out/${tag}_raw_bench.csv" 2>/dev/null
rm -f "$out"/*.ncu-rep
tail -2 $out/f_tests.log
