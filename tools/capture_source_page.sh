#!/bin/bash
# GPU box: source-level ncu capture of one op's large-n launch; keeps only a per-opcode / top-instruction
# digest (the .ncu-rep and the full source CSV are too large to bring back).
set -u
OP=${1:-LssrBeginOp}
K=${SWEEP_K:-3}
mkdir -p gpurun_out
SWEEP_K=$K SWEEP_ITERS=1 python tools/sweep_kernels.py > /dev/null 2>&1 || exit 1
rm -f /tmp/src_$OP.ncu-rep
SWEEP_K=$K SWEEP_ITERS=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:$OP -c 1 -o /tmp/src_$OP python tools/sweep_kernels.py > gpurun_out/ncu_src_$OP.log 2>&1
ncu -i /tmp/src_$OP.ncu-rep --page source --csv > /tmp/src_$OP.csv 2>/dev/null
python - "$OP" "$K" <<'PY' > gpurun_out/source_digest_${OP}_k$K.txt
import collections, csv, sys
rows = list(csv.reader(open(f"/tmp/src_{sys.argv[1]}.csv")))
print(rows[0][1][:160])
data = rows[2:]
tot = sum(int(r[4]) for r in data)
print(f"warp-stall samples {tot}, SASS instructions {len(data)}")
byop = collections.Counter()
for r in data:
    t = r[1].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    byop[op.split(".")[0]] += int(r[4])
print("samples by opcode:", ", ".join(f"{k} {v / tot:.1%}" for k, v in byop.most_common(12)))
print("top instructions (samples, executions, SASS):")
for r in sorted(data, key=lambda r: -int(r[4]))[:15]:
    print(f"  {int(r[4]):6d} {r[5]:>8s}  {r[1].strip()[:90]}")
PY
cat gpurun_out/source_digest_${OP}_k$K.txt
