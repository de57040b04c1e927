#!/usr/bin/env python
"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source sass` (stdin or file) per kernel:
warp-instructions executed by opcode, stall-sample share of the hottest SASS lines.  usage: ncu_source_summary.py src.csv [units]
`units` (e.g. samples of the launch) turns instruction counts into per-unit figures."""
import csv, sys, collections

path = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
kern, hdr, data = None, None, collections.OrderedDict()
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "Kernel Name": kern = row[1]; data[kern] = []; hdr = None; continue
    if row[0] == "Address": hdr = row; continue
    if kern and hdr: data[kern].append(dict(zip(hdr, row)))
for k, lines in data.items():
    tot_i = sum(int(l["Instructions Executed"]) for l in lines)
    tot_t = sum(int(l["Thread Instructions Executed"]) for l in lines)
    tot_s = sum(int(l["# Samples"]) for l in lines)
    print("=" * 100); print(k)
    print(f"SASS lines {len(lines)}  warp-instructions {tot_i:,}  thread-instructions {tot_t:,}  stall samples {tot_s:,}" +
          (f"  -> {tot_i / units:.2f} warp-instr / {tot_t / units:.2f} thread-instr per unit" if units else ""))
    ops = collections.Counter(); smp = collections.Counter()
    for l in lines:
        toks = l["Source"].split()
        op = next((t for t in toks if not t.startswith("@")), "?").split(".")[0].rstrip(";")
        ops[op] += int(l["Instructions Executed"]); smp[op] += int(l["# Samples"])
    print("by opcode (share of warp-instructions | share of stall samples):")
    for op, n in ops.most_common(24):
        print(f"   {op:12s} {100 * n / max(tot_i, 1):6.2f}%  | {100 * smp[op] / max(tot_s, 1):6.2f}%")
    print("hottest lines (share of samples, executions" + (" per unit" if units else "") + "):")
    for idx, l in sorted(enumerate(lines), key=lambda t: -int(t[1]["# Samples"]))[:30]:
        ex = int(l["Instructions Executed"])
        print(f"   {idx:5d} {100 * int(l['# Samples']) / max(tot_s, 1):5.2f}%  {ex / units if units else ex:12.3f}  {l['Source'].strip()[:90]}")
