#!/bin/bash
# round 2, final evidence: full GPU test suite, smoke, the bench line (both arms), ncu launch list of the cfg-3 plan
# and `ncu --set full` summaries of the tcgen05 kernel (text/CSV exported on the box), memcheck of small cases
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 100 --warmup 5 > $O/r2b_bench_line.json 2> $O/r2b_bench.err; tail -2 $O/r2b_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_ref_line.json 2> $O/r2b_ref.err; tail -1 $O/r2b_ref.err
CMD3="python benchmarks/configs.py cfg3 --no-cpu"
$CMD3 > $O/p_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 520 --csv --log-file $O/r2b_cfg3_launches.csv $CMD3 > $O/p_ncu3.log 2>&1
CMD3F="python tools/t32_check.py 21"
$CMD3F > $O/p_plain3f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 5 -c 1 -o /tmp/r2b_t32_site $CMD3F > $O/p_ncu3f.log 2>&1
ncu -i /tmp/r2b_t32_site.ncu-rep --page raw --csv > $O/r2b_tf32_cfg3_site_raw.csv 2>/dev/null
ncu -i /tmp/r2b_t32_site.ncu-rep --page details > $O/r2b_tf32_cfg3_site_details.txt 2>/dev/null
CMD4="python tools/t32_check.py 10"
$CMD4 > $O/p_plain4.log 2>&1 && ncu --set full --clock-control none -k regex:gemm_tf32x3 -s 4 -c 1 -o /tmp/r2b_t32_4096 $CMD4 > $O/p_ncu4.log 2>&1
ncu -i /tmp/r2b_t32_4096.ncu-rep --page raw --csv > $O/r2b_tf32_4096_raw.csv 2>/dev/null
ncu -i /tmp/r2b_t32_4096.ncu-rep --page details > $O/r2b_tf32_4096_details.txt 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/t32_check.py 0 2 6 13 16 > $O/r2b_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r2b_memcheck.log
timeout 300 python tools/t32_trace.py 10 > $O/r2b_trace_4096.log 2>&1
CTN_T32_TRACE_LAUNCH=0 timeout 300 python tools/t32_trace.py 21 > $O/r2b_trace_cfg3_site.log 2>&1
ls -la $O | grep r2b
grep -h "Duration\|Tensor.*Active\|pipe_tensor" $O/r2b_tf32_4096_details.txt | head -5
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2b_bench_line.json').read().strip().splitlines()[-1])
print('value', r['value'], 'e2e', r['e2e']['value'], 'roofline', r['roofline']['frac'], r['clocks'])
for c in r.get('configs', []): print(c['name'], c.get('ms_per_contraction'), c['roofline']['frac'], c.get('errors'))
PY
