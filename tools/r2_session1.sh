#!/bin/bash
# Round 2, GPU session 1: parity of the sliced kernels, interpreter vs sliced timing on the bench workload, first device
# runs of BASELINE configs #4 (raft MaxTerm 4 / MaxLogLen 3) and #5 (SSI 4x3, depth-bounded), launch list + ncu capture.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_session1.sh'
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s1.log"
: > "$LOG"
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> "$LOG" 2>&1
step "pytest sliced -m gpu" timeout 600 python -m pytest tests/test_sliced.py -m gpu -q -x
step "b4 interpreter" timeout 300 python tools/fixture_bench.py MCPaxos3_b4 --reps 2
step "b4 sliced scalar" timeout 300 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 2
step "b4 sliced array" env TLAG_SL_FORM=array timeout 300 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 2
step "raft t4l3 interpreter" timeout 300 python tools/fixture_bench.py MCraft_t4l3 --reps 2
step "raft t4l3 sliced" timeout 300 python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 2
step "ssi 4x3 depth 8 interpreter" timeout 300 python tools/fixture_bench.py MCssi_4x3 --max-levels 8 --reps 1
step "ssi 4x3 depth 8 sliced" env TLAG_NO_BUILD=1 timeout 300 python tools/fixture_bench.py MCssi_4x3 --max-levels 8 --sliced --reps 1
step "ssi 2x2 interpreter" timeout 200 python tools/fixture_bench.py MCssi_2x2 --reps 2
step "ssi 2x2 sliced" timeout 200 python tools/fixture_bench.py MCssi_2x2 --sliced --reps 2
step "ncu launch list b3 sliced" timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
     --log-file "$OUT/r2_launches_b3_sliced.csv" python tools/fixture_bench.py MCPaxos3_b3 --sliced --reps 1
step "ncu full b3 sliced (one mid level)" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sl_ -s 300 -c 26 \
     -o "$OUT/r2_sl_b3" -f python tools/fixture_bench.py MCPaxos3_b3 --sliced --reps 1
tail -3 "$LOG"
