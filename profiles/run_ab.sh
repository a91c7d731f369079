#!/bin/bash
# on-box A/B: warps per SM for the R = 8 kernels (E > 224)
cd /root/repo
O=gpurun_out
for w in 28 32; do
  sed -i "s|return R >= 8 ? [0-9]* : 32;|return R >= 8 ? $w : 32;|" bosot_b200/csrc/sot_pairs.cu
  python -c "import __graft_entry__ as g; g.build()" > $O/r3n_build_$w.log 2>&1
  echo "== R = 8 kernels at $w warps" >> $O/r3n_kt.log
  timeout 300 python profiles/tools_kernel_times.py 200000 256 10 | grep "^pairs" >> $O/r3n_kt.log 2>&1
  timeout 300 python profiles/tools_kernel_times.py 200000 400 10 | grep "^pairs" >> $O/r3n_kt.log 2>&1
done
cat $O/r3n_kt.log
