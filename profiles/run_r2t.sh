#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "1000000 200" "125000 200"; do
  echo "== $shape" >> $O/r2t_kt.log
  timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2t_kt.log 2>&1
  echo "== $shape BOSOT_EXPERIMENT=8" >> $O/r2t_kt.log
  BOSOT_EXPERIMENT=8 timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2t_kt.log 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2t_tests.log 2>&1; echo "tests rc=$?" >> $O/r2t_tests.log
tail -4 $O/r2t_tests.log
grep -v "^info" $O/r2t_kt.log
