#!/bin/bash
cd /root/repo
O=gpurun_out
for shape in "1000000 200" "125000 200" "10000 50" "100 50" "30000 80" "200000 300"; do
  echo "== $shape" >> $O/r3m_kt.log
  timeout 300 python profiles/tools_kernel_times.py $shape 10 >> $O/r3m_kt.log 2>&1
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r3m_tests.log 2>&1; echo "tests rc=$?" >> $O/r3m_tests.log
tail -4 $O/r3m_tests.log
grep "^==\|^pairs\|^device\|^e2e" $O/r3m_kt.log
