#!/bin/bash
# configs[3] decode by measurement variant, then ncu source-level captures of the chain decoder on a small batch
p=gpurun_out/r02_dec2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decoder or alternate" 2>&1 | tail -3
B="timeout 600 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1"
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
run c3_chain X=1 ""
run c3_var4 LLCOMP_DECODER_VARIANT=4 ""
S="--images 1024 --size 256 --scaling weak --strips 1"
run small_chain X=1 "$S"
run small_var4 LLCOMP_DECODER_VARIANT=4 "$S"
CMD="python bench.py --no-cpu --no-e2e --steps 1 --warmup 0 $S"
timeout 300 $CMD > ${p}_plain0.json 2>> ${p}.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slice_decoder_chain -c 1 -f -o ${p}_v0 $CMD > ${p}_ncu_v0.log 2>&1
export LLCOMP_DECODER_VARIANT=4
timeout 300 $CMD > ${p}_plain4.json 2>> ${p}.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slice_decoder_chain -c 1 -f -o ${p}_v4 $CMD > ${p}_ncu_v4.log 2>&1
tail -2 ${p}_ncu_v4.log; tail -3 ${p}.err; ls -la gpurun_out | tail -5
