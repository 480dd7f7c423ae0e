#!/bin/bash
# Round 2, GPU session 2b: configs #4/#5 on the sliced kernels, per-level times, the contract bench line, launch list and
# a full ncu capture of five slice kernels on a big b4 level (kept small: gpurun_out/ must stay under 64 MiB).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r2_session2.sh'
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s2.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
step "raft t4l3 sliced" timeout 300 python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 2
step "ssi 4x3 depth 10 sliced" timeout 600 python tools/fixture_bench.py MCssi_4x3 --sliced --reps 1
step "ssi 4x3 depth 9 interpreter" timeout 400 python tools/fixture_bench.py MCssi_4x3 --max-levels 9 --reps 1
step "b4 sliced per level" timeout 300 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 1 --waves
step "bench N=1" timeout 1500 python bench.py --steps 3 --warmup 3
step "ncu launch list b4 sliced" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
     --clock-control none -k regex:k_sl_ -c 1200 --csv --log-file "$OUT/r2_launches_b4_sliced.csv" python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 1
step "ncu full b4 sliced (5 kernels of level 27)" timeout 900 ncu --set full --clock-control none -k 'regex:k_sl_(inv_2|inv_3|next_1|next_12|next_16)$' -s 130 -c 5 \
     -o "$OUT/r2_sl_b4_l27" -f python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 1
ls -la "$OUT" >> "$LOG"
tail -3 "$LOG"
