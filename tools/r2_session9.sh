#!/bin/bash
# Round 2, GPU session 9 (1 GPU): A/B of the probe load (ld.volatile = LDG.STRONG.SYS vs ld.cg) in the seen-set probe, on
# the b4 workload and on K1; array-form occupancy variants for raft / SSI.
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s9.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
step "b4 default" timeout 200 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 3
step "b4 probe ld.cg" env TLAG_CSRC_DIR=/tmp/w/csrc_cg/csrc timeout 200 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 3
step "b4 min_slice 160" env TLAG_SL_MIN_SLICE=160 timeout 200 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 3
step "K1 default" timeout 200 python tools/k1test.py
step "K1 probe ld.cg" env TLAG_LIB=/root/repo/tla_rust_b200/csrc/native/libtlag_probe_cg.so timeout 200 python tools/k1test.py
step "raft occ 6" env TLAG_SL_OCC=6 timeout 200 python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 2
step "raft occ 8" env TLAG_SL_OCC=8 timeout 200 python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 2
step "ssi d9 occ 8" env TLAG_SL_OCC=8 timeout 300 python tools/fixture_bench.py MCssi_4x3 --max-levels 9 --sliced --reps 1
step "ssi d9 default" timeout 300 python tools/fixture_bench.py MCssi_4x3 --max-levels 9 --sliced --reps 1
grep -E "^===|device_s|kernel ms|rc=" "$LOG" | grep -v "^rc=0" | sed 's/"counts_match.*//' | cut -c1-330 | grep -v " 20 22 \| 24 26 \| 26 28 "
