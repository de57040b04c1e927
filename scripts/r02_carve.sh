#!/bin/bash
p=gpurun_out/r02_carve
B="timeout 600 python bench.py --no-cpu --steps 2 --warmup 1"
run() { # name, env, extra args
  env $2 $B $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    e=d["decode"].get("e2e") or {}
    print("$1", "enc", round(d["value"],2), "dec", round(d["decode"]["value"],3), round(d["decode"]["ms_per_step"],1), "ms; dec e2e", round(e.get("value",0),3), d["round_trip_exact"], d["config"]["slices_per_gpu"])
except Exception as e: print("$1", "failed", e)
PY
}
run l1 X=1 ""
run maxsmem LLCOMP_DECODER_MAX_CARVEOUT=1 ""
run c1_l1 X=1 "--no-e2e --images 1 --size 4096 --tile 512 --scaling weak"
run c1_maxsmem LLCOMP_DECODER_MAX_CARVEOUT=1 "--no-e2e --images 1 --size 4096 --tile 512 --scaling weak"
run c2_l1 X=1 "--no-e2e --images 1 --size 8192 --channels 1 --noise -1 --tile 256 --scaling weak"
run strips_l1 X=1 "--no-e2e --images 512 --strips 2 --scaling weak"
run strips_maxsmem LLCOMP_DECODER_MAX_CARVEOUT=1 "--no-e2e --images 512 --strips 2 --scaling weak"
tail -3 ${p}.err
