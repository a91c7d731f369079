#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "125000 200" "250000 200" "1000000 200" "30000 80"; do
  for fl in 0 64; do
    echo "== $shape BOSOT_EXPERIMENT=$fl" >> $O/r3h_kt.log
    BOSOT_EXPERIMENT=$fl timeout 300 python profiles/tools_kernel_times.py $shape 20 >> $O/r3h_kt.log 2>&1
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r3h_tests.log 2>&1; echo "tests rc=$?" >> $O/r3h_tests.log
tail -4 $O/r3h_tests.log
grep "^==\|^pairs\|^device\|^check" $O/r3h_kt.log
