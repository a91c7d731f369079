#!/bin/bash
cd /root/repo
O=gpurun_out
python bench.py > $O/r2s_bench_n1.json 2> $O/r2s_bench_n1.err; echo "bench rc=$?"
python profiles/tools_ea_breakdown.py > $O/r2s_ea.log 2>&1
python profiles/tools_bo_step_breakdown.py > $O/r2s_bo.log 2>&1
tail -c 600 $O/r2s_bench_n1.json
