#!/bin/bash
# round 2 evidence: bench line, ncu launch lists, `ncu --set full` summaries of the dominant kernels (text/CSV
# exported on the box; the .ncu-rep files are too big to bring back)
O=gpurun_out
mkdir -p $O
python bench.py --steps 50 --warmup 5 > $O/r2_bench_line.json 2> $O/r2_bench.err; tail -2 $O/r2_bench.err
# ---- cfg 2: launch list + full capture of the chain kernel ----
CMD2="python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline"
$CMD2 > $O/p_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_cfg2_launches.csv $CMD2 > $O/p_ncu2.log 2>&1
$CMD2 > $O/p_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 3 -c 1 -o /tmp/r2_chain $CMD2 > $O/p_ncu2f.log 2>&1
ncu -i /tmp/r2_chain.ncu-rep --page raw --csv > $O/r2_cfg2_chain_kernel_raw.csv 2>/dev/null
ncu -i /tmp/r2_chain.ncu-rep --page details > $O/r2_cfg2_chain_kernel_details.txt 2>/dev/null
# ---- cfg 3: launch list of the literal-size plan + full capture of a fused tcgen05 site ----
CMD3="python benchmarks/configs.py cfg3 --no-cpu"
$CMD3 > $O/p_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 520 --csv --log-file $O/r2_cfg3_launches.csv $CMD3 > $O/p_ncu3.log 2>&1
CMD3F="python tools/t32_check.py 21"
$CMD3F > $O/p_plain3f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 5 -c 1 -o /tmp/r2_t32_site $CMD3F > $O/p_ncu3f.log 2>&1
ncu -i /tmp/r2_t32_site.ncu-rep --page raw --csv > $O/r2_tf32_cfg3_site_raw.csv 2>/dev/null
ncu -i /tmp/r2_t32_site.ncu-rep --page details > $O/r2_tf32_cfg3_site_details.txt 2>/dev/null
CMD4="python tools/t32_check.py 10"
$CMD4 > $O/p_plain4.log 2>&1 && ncu --set full --clock-control none -k regex:gemm_tf32x3 -s 4 -c 1 -o /tmp/r2_t32_4096 $CMD4 > $O/p_ncu4.log 2>&1
ncu -i /tmp/r2_t32_4096.ncu-rep --page raw --csv > $O/r2_tf32_4096_raw.csv 2>/dev/null
ncu -i /tmp/r2_t32_4096.ncu-rep --page details > $O/r2_tf32_4096_details.txt 2>/dev/null
# ---- cfg 5 / cfg 4: launch lists ----
CMD5="python benchmarks/configs.py cfg5 --no-cpu"
$CMD5 > $O/p_plain5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 70 --csv --log-file $O/r2_cfg5_launches.csv $CMD5 > $O/p_ncu5.log 2>&1
ls -la $O | tail -25
grep -h "gpu__time_duration\|Duration" $O/r2_tf32_cfg3_site_details.txt | head -3
