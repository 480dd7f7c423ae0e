#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s5.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
step "pytest 2-GPU parity" timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short
tail -c 5000 "$LOG"
