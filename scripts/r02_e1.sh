#!/bin/bash
# Non-headline configs (scripts/r02_sweeps.sh) and compute-sanitizer memcheck of every default kernel on a small case
bash scripts/r02_sweeps.sh gpurun_out/r02_sweeps.jsonl 2>&1 | tail -40
timeout 300 python scripts/sanitize_case.py > gpurun_out/r02_sanitize_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck python scripts/sanitize_case.py > gpurun_out/r02_memcheck.log 2>&1
tail -5 gpurun_out/r02_memcheck.log
