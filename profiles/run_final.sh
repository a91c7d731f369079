#!/bin/bash
# round-end validation on one GPU: smoke, gpu tests, bench (both arms), ncu launch list of the bench command
cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r3f_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r3f_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r3f_tests.log 2>&1; echo "tests rc=$?" >> $O/r3f_tests.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r3f_bench_reference.json 2> $O/r3f_ref.err; echo "ref rc=$?" >> $O/r3f_ref.err
python bench.py > $O/r3f_bench_n1.json 2> $O/r3f_bench_n1.err; rc=$?; echo "bench rc=$rc" >> $O/r3f_bench_n1.err
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r3f_launches.csv python bench.py --steps 2 --warmup 1 > $O/r3f_ncu.log 2>&1
fi
tail -2 $O/r3f_smoke.log; tail -3 $O/r3f_tests.log; tail -1 $O/r3f_ref.err; tail -1 $O/r3f_bench_n1.err
