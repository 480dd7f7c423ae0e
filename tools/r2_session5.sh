#!/bin/bash
# Round 2, GPU session 5 (2 GPUs): 2-GPU parity suite after the rank / retry / rollback fixes, TLC-exact replay on the
# device, tlc -gpus 2.
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s5.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
step "pytest 2-GPU parity" timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q
step "pytest exact replay + traces + cli" timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -m gpu -q -k "exact or trace or make_flow"
step "tlc -gpus 2 race.tla (exit status + report)" bash -c 'd=$(mktemp -d); cp models/demo/* $d/; cd $d; export PATH=/root/repo/bin:$PATH; pcal2tla *tla > /dev/null; tlc -gpus 2 race.tla 2>&1 | tail -12; echo rc_race=${PIPESTATUS[0]}'
step "ncu ssi 4x3 (6 kernels of level 8)" timeout 600 ncu --set full --clock-control none -k regex:k_sl_ -s 350 -c 14 -o "$OUT/r2_sl_ssi_l8" -f python tools/fixture_bench.py MCssi_4x3 --max-levels 9 --sliced --reps 1
step "ncu raft t4l3 (6 kernels of level 13)" timeout 600 ncu --set full --clock-control none -k regex:k_sl_ -s 420 -c 12 -o "$OUT/r2_sl_raft_l13" -f python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 1
ls -la "$OUT" >> "$LOG"
tail -3 "$LOG"
