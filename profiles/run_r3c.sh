#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "1000000 200" "125000 200"; do
  for fl in 0 128 256; do
    echo "== $shape BOSOT_EXPERIMENT=$fl" >> $O/r3c_kt.log
    BOSOT_EXPERIMENT=$fl timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r3c_kt.log 2>&1
  done
done
grep "^==\|^pairs\|^check" $O/r3c_kt.log
