#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2r_tests.log 2>&1; echo "tests rc=$?" >> $O/r2r_tests.log
for shape in "10000 50" "100 50" "16000 64" "3000 30" "30000 80"; do
  echo "== $shape" >> $O/r2r_kt.log
  timeout 300 python profiles/tools_kernel_times.py $shape 20 >> $O/r2r_kt.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sot_pair_kernel|posterior_ws|pair_init" -c 12 -o $O/r2r_c4 python profiles/tools_kernel_times.py 10000 50 2 > $O/r2r_ncu.log 2>&1
tail -5 $O/r2r_tests.log
