#!/bin/bash
# tests + kernel timings (A/B against the general exp route with BOSOT_EXPERIMENT=4)
cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2p_tests.log 2>&1; echo "tests rc=$?" >> $O/r2p_tests.log
for shape in "1000000 200" "125000 200" "10000 50" "100 50"; do
  echo "== $shape" >> $O/r2p_kt.log
  timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2p_kt.log 2>&1
  echo "== $shape BOSOT_EXPERIMENT=4" >> $O/r2p_kt.log
  BOSOT_EXPERIMENT=4 timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2p_kt.log 2>&1
done
tail -3 $O/r2p_tests.log
