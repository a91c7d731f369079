#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 300 python profiles/tools_kernel_times.py 1000000 200 3 > $O/r2v_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"sot_pair_kernel|posterior_ws|scan_kernel" -c 6 -o $O/r2v_c5 python profiles/tools_kernel_times.py 1000000 200 1 > $O/r2v_ncu.log 2>&1
tail -3 $O/r2v_ncu.log
