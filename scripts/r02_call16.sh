#!/bin/bash
p=gpurun_out/r02_c16
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15
B="timeout 300 python bench.py --no-cpu --no-decode --no-e2e --steps 3 --warmup 2"
run() { # name, env, extra args
  env $2 $B $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"],2), "GB/s enc", {k["name"]: round(k["ms"],3) for k in d["kernels"]}, d["round_trip_exact"])
except Exception as e: print("$1", "failed", e)
PY
}
run pix_1024 X=1 ""
run pix_592 X=1 "--images 592 --scaling weak --strips 1"
run pix_296 X=1 "--images 296 --scaling weak --strips 1"
run rec_296 LLCOMP_CODER_RECORDS=1 "--images 296 --scaling weak --strips 1"
run pix_148 X=1 "--images 148 --scaling weak --strips 1"
# decode end to end by number of groups
for gname in 1 2 4; do
LLCOMP_GROUPS=$gname timeout 600 python bench.py --no-cpu --steps 2 --warmup 2 > ${p}_dec_g$gname.json 2>> ${p}.err
python - <<PY
import json
d=json.loads(open("${p}_dec_g$gname.json").read().strip().splitlines()[-1])
print("groups $gname: enc e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],1), "dec dev", round(d["decode"]["ms_per_step"],1), "dec e2e", round(d["decode"]["e2e"]["ms_per_step"],1))
PY
done
tail -3 ${p}.err
