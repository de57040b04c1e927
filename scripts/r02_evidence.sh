#!/bin/bash
# Round evidence on one B200: GPU test suite, bench (both arms), ncu launch list of the step, --set full captures of the
# front end and the fused coder at full size and of the chain decoder on a small batch (a full-size decoder launch takes
# 1.1 s and ncu replays it ~40 times with source counters).  $1 = output prefix, $2 = "notest" to skip pytest
p=${1:-gpurun_out/r02_ev}
if [ "$2" != "notest" ]; then timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; fi
timeout 900 python bench.py --steps 5 --warmup 3 > ${p}_bench.json 2> ${p}_bench.err; tail -c 400 ${p}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${p}_bench_reference.json 2>> ${p}_bench.err
python - <<PY
import json
d=json.loads(open("${p}_bench.json").read().strip().splitlines()[-1])
r=json.loads(open("${p}_bench_reference.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],1), "dec", round(d["decode"]["value"],3), "dec e2e", round(d["decode"]["e2e"]["value"],3), d["round_trip_exact"], "| reference", round(r["value"],4), "dec", round(r["decode"]["value"],4))
print({k["name"]: round(k["ms"],3) for k in d["kernels"]}); print(d.get("identity")); print(d["roofline"]["frac"], d["clocks"])
PY
CMD="python bench.py --no-cpu --no-e2e --steps 1 --warmup 0"
timeout 300 $CMD > ${p}_ncu_plain.json 2>> ${p}_bench.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file ${p}_ncu_launches.csv $CMD > ${p}_ncu_launches.log 2>&1
timeout 300 $CMD > ${p}_ncu_plain.json 2>> ${p}_bench.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_slice_coder_fused|k_frontend" -c 2 -f -o ${p}_hot $CMD > ${p}_ncu_full.log 2>&1
S="--images 1024 --size 256 --scaling weak --strips 1"
timeout 300 $CMD $S > ${p}_ncu_plain_small.json 2>> ${p}_bench.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slice_decoder_chain -c 1 -f -o ${p}_dec $CMD $S > ${p}_ncu_dec.log 2>&1
tail -2 ${p}_ncu_full.log ${p}_ncu_dec.log; ls -la gpurun_out | tail -12
