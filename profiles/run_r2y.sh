#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "1000000 200" "125000 200" "10000 50" "100 50"; do
  echo "== $shape" >> $O/r2y_kt.log
  timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r2y_kt.log 2>&1
done
grep -v "^info" $O/r2y_kt.log
