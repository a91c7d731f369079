#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "125000 200" "250000 200" "1000000 200"; do
  for fl in 0 16 64; do
    echo "== $shape BOSOT_EXPERIMENT=$fl" >> $O/r2u_kt.log
    BOSOT_EXPERIMENT=$fl timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2u_kt.log 2>&1
  done
done
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2u_tests.log 2>&1; echo "tests rc=$?" >> $O/r2u_tests.log
tail -4 $O/r2u_tests.log
grep -v "^info" $O/r2u_kt.log
