#!/bin/bash
# GPU box: refresh the evidence under gpurun_out/ that profiles/ is summarised from.
#   1. bench line (no profiler)                         -> gpurun_out/bench_n1.json
#   2. ncu launch list of a short bench run              -> gpurun_out/launches.csv
#   3. ncu --set full of one launch of each fused kernel at n = 2^26 (k = $SWEEP_K, default 3)
#      exported as raw CSV (the .ncu-rep files are too large to bring back)
#                                                         -> gpurun_out/raw_<Op>_k<K>.csv
# Every ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
K=${SWEEP_K:-3}
if [ -z "${SKIP_BENCH:-}" ]; then python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; fi || { echo "bench failed"; tail -5 gpurun_out/bench_n1.err; exit 1; }
tail -c 600 gpurun_out/bench_n1.json; echo
FAST="--steps 2 --warmup 3 --no-cpu-baseline --no-sweep --no-extra"
timeout 300 python bench.py $FAST > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py $FAST > gpurun_out/ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches.csv)"
SWEEP_K=$K SWEEP_ITERS=3 python tools/sweep_kernels.py > gpurun_out/sweep_k$K.txt 2>&1 || { echo "sweep failed"; exit 1; }
cat gpurun_out/sweep_k$K.txt
for OP in ${OPS:-LssrBeginOp ProposeOp ApplyOp DecideOp EvalOp TrialOp}; do
    rm -f /tmp/prof_$OP.ncu-rep
    SWEEP_K=$K SWEEP_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
        -k regex:$OP -c 1 -o /tmp/prof_$OP python tools/sweep_kernels.py > gpurun_out/ncu_$OP.log 2>&1
    ncu -i /tmp/prof_$OP.ncu-rep --page raw --csv > gpurun_out/raw_${OP}_k$K.csv 2>/dev/null
    echo "$OP: $(wc -c < gpurun_out/raw_${OP}_k$K.csv) bytes"
done
