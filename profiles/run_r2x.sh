#!/bin/bash
# two GPUs: multi-GPU parity tests, NVLink scatter tool plain then under ncu (NVLink byte counters of the scatter epilogue)
cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -v > $O/r2x_multi_tests.log 2>&1; echo "tests rc=$?" >> $O/r2x_multi_tests.log
timeout 300 python profiles/tools_nvlink_scatter.py 500000 200 > $O/r2x_nvlink_plain.log 2>&1; rc=$?; echo "plain rc=$rc" >> $O/r2x_nvlink_plain.log
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:posterior_ws --csv --log-file $O/r2x_nvlink_ncu.csv python profiles/tools_nvlink_scatter.py 500000 200 > $O/r2x_nvlink_ncu.log 2>&1
fi
tail -3 $O/r2x_multi_tests.log; tail -2 $O/r2x_nvlink_plain.log; tail -12 $O/r2x_nvlink_ncu.csv | cut -c1-300
