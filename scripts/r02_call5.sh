#!/bin/bash
# per-role busy cycles of the fused coder (LLC_ROLE_TIMING build) by slices per SM, + the new multi-device / pipelined-decode tests
p=gpurun_out/r02_c5
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
VT=LLCOMP_B200_LIB=$PWD/llcomp_b200/lib/vt.so
B="timeout 300 python bench.py --no-e2e --no-cpu --no-decode --steps 1 --warmup 1"
env $VT $B --images 148 > ${p}_t148.log 2>&1;  python scripts/role_stats.py ${p}_t148.log
env $VT $B --images 592 > ${p}_t592.log 2>&1;  python scripts/role_stats.py ${p}_t592.log
env $VT LLCOMP_FUSED_NS=4 $B --images 1024 > ${p}_t1024_old4.log 2>&1;  python scripts/role_stats.py ${p}_t1024_old4.log
env $VT $B --images 1024 > ${p}_t1024.log 2>&1;  python scripts/role_stats.py ${p}_t1024.log
gzip -f ${p}_t*.log
