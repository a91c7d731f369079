#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "1000000 200" "125000 200" "10000 50" "100 50"; do
  for z in "" "1"; do
    echo "== $shape BOSOT_NO_ZERO_COPY=$z" >> $O/r3a_kt.log
    if [ -z "$z" ]; then timeout 300 python profiles/tools_kernel_times.py $shape 20 >> $O/r3a_kt.log 2>&1
    else BOSOT_NO_ZERO_COPY=1 timeout 300 python profiles/tools_kernel_times.py $shape 20 >> $O/r3a_kt.log 2>&1; fi
  done
done
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r3a_tests.log 2>&1; echo "tests rc=$?" >> $O/r3a_tests.log
tail -4 $O/r3a_tests.log
grep "^==\|^e2e\|^device" $O/r3a_kt.log
