#!/bin/bash
# $1 GPUs: multi-device entry points on distinct devices (N >= 2), then the strong-scaling bench at N
N=$1; p=gpurun_out/r02_scale
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -x -q 2>&1 | tail -4; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > ${p}_${N}gpu.json 2> ${p}_${N}gpu.err; tail -c 300 ${p}_${N}gpu.err
python - <<PY
import json
d=json.loads(open("${p}_${N}gpu.json").read().strip().splitlines()[-1])
print("N=$N strong:", d["config"]["workload"]); print("value", round(d["value"],2), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],1), "h2d alone", round(d["e2e"]["h2d_alone_ms"],1), "dec", round(d["decode"]["value"],3), "dec e2e", round(d["decode"]["e2e"]["value"],3), "bpp", round(d["bits_per_pixel"],4), d["round_trip_exact"], d.get("identity",{}).get("identical_to_oracle"), d.get("bpp_vs_single_slice_reference"), d["config"]["slices_per_gpu"])
PY
