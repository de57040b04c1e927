#!/bin/bash
# Decoder forms side by side: GPU tests, then configs[3] decode by form / measurement variant, then configs[1] and configs[2]
p=gpurun_out/r02_dec
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="timeout 600 python bench.py --no-cpu --no-e2e --steps 2 --warmup 1"
run() { # name, env, extra args
  env $2 $B $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", "enc", round(d["value"],2), "dec", round(d["decode"]["value"],3), "GB/s", round(d["decode"]["ms_per_step"],1), "ms", d["round_trip_exact"], d["config"]["slices_per_gpu"])
except Exception as e: print("$1", "failed", e)
PY
}
run c3_v1 LLCOMP_DECODER_V1=1 ""
run c3_chain X=1 ""
run c3_var1 LLCOMP_DECODER_VARIANT=1 ""
run c3_var2 LLCOMP_DECODER_VARIANT=2 ""
run c3_var3 LLCOMP_DECODER_VARIANT=3 ""
run c1_v1 LLCOMP_DECODER_V1=1 "--images 1 --size 4096 --tile 512 --scaling weak"
run c1_chain X=1 "--images 1 --size 4096 --tile 512 --scaling weak"
run c2_v1 LLCOMP_DECODER_V1=1 "--images 1 --size 8192 --channels 1 --noise -1 --tile 512 --scaling weak"
run c2_chain X=1 "--images 1 --size 8192 --channels 1 --noise -1 --tile 512 --scaling weak"
run c2_chain_256 X=1 "--images 1 --size 8192 --channels 1 --noise -1 --tile 256 --scaling weak"
run smooth_chain X=1 "--noise 0"
tail -3 ${p}.err
