#!/bin/bash
# One gpurun call that settles everything round 1 left unverified on a device:
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'
#
# Every step runs under its own `timeout`, writes into gpurun_out/, and a failing step does not stop the next.
# Nothing printed under ncu is a bench value.
set -u
OUT=gpurun_out
mkdir -p "$OUT"
step() { echo "=== $1" | tee -a "$OUT/session.log"; shift; ( "$@" ) >> "$OUT/session.log" 2>&1; echo "rc=$?" | tee -a "$OUT/session.log"; }

# 1. the parity suite (includes the xfail-guarded SSI subroutine fixture and the recompiled raft fixtures)
step "pytest -m gpu" timeout 900 python -m pytest tests -m gpu -q -rxX
# 2. SSI (CALL/RET subroutines, 4096 and 8192 frame classes) and the symmetric subroutine models, with timing
step "fixture bench: MCssi" timeout 300 python tools/fixture_bench.py MCssi --reps 2
step "fixture bench: MCssi 3x1, 2x2, 2x2_wide (8192 class)" timeout 600 python tools/fixture_bench.py MCssi_3x1 MCssi_2x2 MCssi_2x2_wide --reps 2
# 2b. model-specialised native builds (prebuilt by __graft_entry__.build() into csrc/native/): parity + timing next to
#     the interpreter on the same fixtures
step "fixture bench: native MCPaxos3_b2 / Containers / MCPaxos3_b4" timeout 900 python tools/fixture_bench.py MCPaxos3_b2 Containers MCPaxos3_b4 --native --reps 2
step "fixture bench: interpreter MCPaxos3_b2 / Containers / MCPaxos3_b4" timeout 900 python tools/fixture_bench.py MCPaxos3_b2 Containers MCPaxos3_b4 --reps 2
step "fixture bench: native raft (prebuilt library; nvcc needs 5 min otherwise)" timeout 900 python tools/fixture_bench.py MCraft_s3_l --native --reps 2
step "fixture bench: native b4, lane form (built here: ~1 min of nvcc)" env TLAG_NATIVE_SCHED=lane timeout 900 python tools/fixture_bench.py MCPaxos3_b4 --native --reps 2
step "fixture bench: native b4, 1 and 4 CTAs/SM" bash -c 'for o in 1 4; do TLAG_NATIVE_OCC=$o timeout 600 python tools/fixture_bench.py MCPaxos3_b4 --native --reps 2; done'
# 3. contract bench at N=1, then the raft workload on its own
step "bench N=1" timeout 900 python bench.py --steps 3 --warmup 3
step "bench N=1 on the native build" env TLAG_NATIVE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-k1
step "fixture bench: raft" timeout 600 python tools/fixture_bench.py MCraft_s3_m MCraft_s3_l --reps 2
# 4. launch list of a short bench (share of the step per kernel) and one full capture of the wave kernel
step "ncu launch list" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
     --log-file "$OUT/r2_launches_bench_b3.csv" python bench.py --steps 1 --warmup 1 --no-k1 --workload MCPaxos3_b3
step "ncu k_wave full" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 12 -c 1 \
     -o "$OUT/r2_k_wave_b3" -f python bench.py --steps 1 --warmup 0 --no-k1 --workload MCPaxos3_b3
tail -5 "$OUT/session.log"
