#!/bin/bash
# Round 2, GPU session 7 (1 GPU): what the driver runs at round end -- the whole `pytest -m gpu` suite, smoke(), the
# contract bench line (full: K1, CPU arm, configs #4 / #5) -- plus the per-launch metrics list of one b4 BFS on the
# final kernels (duration, DRAM bytes, instructions: the `traffic` of the roofline).  Logs only; no .ncu-rep files.
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s7.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
step "pytest -m gpu (all)" timeout 1200 python -m pytest tests -m gpu -q --tb=short
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
step "bench N=1 (full line)" timeout 1500 python bench.py --steps 5 --warmup 3
step "ncu launch list b4 sliced (final kernels)" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum \
     --clock-control none -k regex:k_sl_ -c 600 --csv --log-file "$OUT/r2_launches_b4_final.csv" python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 1
step "ncu raft (selected metrics, level 13)" timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
     --clock-control none -k regex:k_sl_ -s 420 -c 35 --csv --log-file "$OUT/r2_raft_l13_metrics.csv" python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 1
step "ncu ssi (selected metrics, level 8)" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
     --clock-control none -k regex:k_sl_ -s 350 -c 50 --csv --log-file "$OUT/r2_ssi_l8_metrics.csv" python tools/fixture_bench.py MCssi_4x3 --max-levels 9 --sliced --reps 1
rm -f "$OUT"/*.ncu-rep
du -sh "$OUT" >> "$LOG"
grep -E "^===|^rc=|passed|failed|\"metric\"" "$LOG" | cut -c1-600
