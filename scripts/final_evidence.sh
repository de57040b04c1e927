#!/bin/bash
# Round-end evidence on one GPU: tests, bench (both arms), ncu launch list, one ncu --set full of the encode kernels,
# the other BASELINE configs.  Everything lands in gpurun_out/ with the prefix $1.
p=${1:-r01_v8}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${p}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${p}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${p}_bench.json 2> gpurun_out/bench.err || exit 1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${p}_bench_reference.json 2>> gpurun_out/bench.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${p}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_frontend_tiled|k_slice_coder_fused" -c 2 -o gpurun_out/${p}_encode -f python bench.py --no-cpu --no-e2e --no-decode --steps 1 --warmup 0 > gpurun_out/ncu_e.log 2>&1
tail -1 gpurun_out/ncu_e.log | cut -c1-150
timeout 900 bash scripts/other_configs.sh gpurun_out/${p}_other_configs.jsonl
tail -1 gpurun_out/${p}_bench.json | cut -c1-300
