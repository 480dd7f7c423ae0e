#!/bin/bash
# Round 2, GPU session 3 (2 GPUs): peer-memory exchange on the device -- the 2-GPU parity suite (p2p and NCCL paths, overflow
# retry, cross-rank counterexample), the bench at N=2 on both exchange paths; on GPU 0: grouped / static-select slice
# variants of the b4 workload, the TMA form of K1 next to the staged one.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1800 -- 'bash tools/r2_session3.sh'
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s3.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
nvidia-smi -L >> "$LOG" 2>&1
step "pytest 2-GPU parity" timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x
step "pytest K1 (TMA form) + sliced" timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_sliced.py -m gpu -q -x -k "probe or sliced_kernels"
step "bench N=2 p2p" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-k1
step "bench N=2 nccl" env TLAG_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --no-k1
step "b4 default (min_slice 256, occ 4)" timeout 200 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 3
for v in "0 4" "256 6" "256 8" "512 4" "1024 4"; do set -- $v
  step "b4 min_slice=$1 occ=$2" env TLAG_SL_MIN_SLICE=$1 TLAG_SL_OCC=$2 timeout 200 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 3
done
step "b4 default, no clustering sort" env TLAG_NO_SORT=1 timeout 200 python tools/fixture_bench.py MCPaxos3_b4 --sliced --reps 3
step "raft default" timeout 200 python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 2
step "raft min_slice=0" env TLAG_SL_MIN_SLICE=0 timeout 200 python tools/fixture_bench.py MCraft_t4l3 --sliced --reps 2
step "K1 tma" timeout 300 python tools/k1test.py
step "K1 staged" env TLAG_K1_STAGED=1 timeout 300 python tools/k1test.py
step "tlc -gpus 2 on models/demo (make flow)" bash -c 'd=$(mktemp -d); cp models/demo/* $d/; cd $d; export PATH=/root/repo/bin:$PATH; pcal2tla *tla; tlc -gpus 2 lock.tla; echo rc_lock=$?; tlc -gpus 2 race.tla; echo rc_race=$?'
ls -la "$OUT" >> "$LOG"
tail -3 "$LOG"
