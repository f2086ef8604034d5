set -x
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py 2> gpurun_out/final/bench.err | tail -1 > gpurun_out/final/bench_line.json; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/final/ref.err | tail -1 > gpurun_out/final/bench_ref_line.json; echo ref rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/r1_cfg2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu.log 2>&1; echo ncu rc=$?
python benchmarks/configs.py > gpurun_out/final/configs_a.json 2> gpurun_out/final/configs_a.err
python benchmarks/configs.py cfg3 cfg4m cfg5 --no-cpu > gpurun_out/final/configs_b.json 2> gpurun_out/final/configs_b.err
python benchmarks/interp_latency.py > gpurun_out/final/interp_latency.json 2>&1
python benchmarks/readme_examples.py > gpurun_out/final/readme.json 2>&1
cat gpurun_out/final/bench_line.json | cut -c1-600
