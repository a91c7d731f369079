#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "1000000 200" "125000 200"; do
  for fl in 0 32; do
    echo "== $shape BOSOT_EXPERIMENT=$fl" >> $O/r2w_kt.log
    BOSOT_EXPERIMENT=$fl timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2w_kt.log 2>&1
  done
done
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2w_tests.log 2>&1; echo "tests rc=$?" >> $O/r2w_tests.log
tail -4 $O/r2w_tests.log
grep -v "^info" $O/r2w_kt.log
