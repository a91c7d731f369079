#!/bin/bash
# on-box A/B: coverage of the block-vector cache (largest leaf count of one kind that is tabulated)
cd /root/repo
O=gpurun_out
for c in 5 7 4 3; do
  sed -i "s|^constexpr int kBvMaxCount = [0-9]*,|constexpr int kBvMaxCount = $c,|" bosot_b200/csrc/sot_pairs.cu
  python -c "import __graft_entry__ as g; g.build()" > $O/r3o_build_$c.log 2>&1
  echo "== kBvMaxCount = $c" >> $O/r3o_kt.log
  timeout 300 python profiles/tools_kernel_times.py 1000000 200 10 | grep "^pairs\|^check" >> $O/r3o_kt.log 2>&1
  timeout 300 python profiles/tools_kernel_times.py 125000 200 10 | grep "^pairs" >> $O/r3o_kt.log 2>&1
done
cat $O/r3o_kt.log
