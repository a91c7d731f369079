#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2q_tests.log 2>&1; echo "tests rc=$?" >> $O/r2q_tests.log
for shape in "1000000 200" "10000 50" "100 50"; do
  echo "== $shape" >> $O/r2q_kt.log
  BOSOT_TIME_F32=1 timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2q_kt.log 2>&1
done
tail -5 $O/r2q_tests.log
