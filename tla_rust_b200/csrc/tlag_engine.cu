// tlag_engine.cu -- sm_100a BFS engine behind include/tlag.h.
//
// One wave (BFS level) = one launch of k_wave: a persistent grid pulls 32-state chunks of
// the frontier; each thread unpacks its state, runs the invariant program, then the
// next-state program.  Every EMIT event is handled warp-synchronously: pack -> 64-bit
// fingerprint -> open-addressed seen-set probe/insert (atomicCAS on 8-byte slots in HBM)
// -> __ballot_sync compaction of the newly discovered states into the tail of the state
// store (which doubles as the BFS queue: the next frontier is the slice appended by this
// wave).  Successor states never round-trip through HBM before the probe (SURVEY.md §8d:
// the fused form of K2->K1->K3->K4).  No tensor cores: integer / hash work only.
//
// Replaces TLC's Worker loop + StateQueue, action evaluator, FPSet and invariant checker
// (all external to /root/reference, which only fixes the CLI contract
// Makefile:6-7 and the output format README.md:267-321).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>   // CUDA-toolkit header library; used only to reorder a level (not on the hot path)

#include "tlag_dev.cuh"
// second copy of the instruction executor without the extension ops (see tlag_vm_exec.inc)
#define TLAG_VM_EXEC_FN tlag_vm_exec_lean
#define TLAG_VM_EXT 0
#include "tlag_vm_exec.inc"
#undef TLAG_VM_EXEC_FN
#undef TLAG_VM_EXT

// Resident CTAs per SM the interpreter wave kernel is compiled for (register cap = 65536 / (512 x this)).
#define TLAG_WAVE_OCC(FRAME, SMEM) ((FRAME) <= 256 ? 4 : ((FRAME) <= 512 ? 2 : ((SMEM) ? 1 : TLAG_BIG_OCC)))

#define TLAG_MAXW 128
#ifndef TLAG_BIG_OCC
#define TLAG_BIG_OCC 4   /* resident CTAs per SM for the big-frame (> 512 words) wave kernels: latency-bound on local memory */
#endif
#define TLAG_BLOCK 512
#define TLAG_MAX_STEPS (1u << 26)

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                       \
      return TLAG_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

struct tlag_engine {
  tlag_model m;
  std::vector<uint64_t> h_code;
  std::vector<uint32_t> h_init;     // retained initial states (for tlag_restart)
  DevParams p;
  uint64_t* d_code = nullptr; int32_t* d_cpool = nullptr; tlag_slot* d_layout = nullptr;
  uint32_t* d_states = nullptr; uint32_t* d_parent = nullptr; uint32_t* d_meta = nullptr;
  unsigned long long* d_table = nullptr; unsigned table_log2 = 0;
  Counters* d_ctr = nullptr;
  uint64_t cap_states = 0;
  uint64_t lo = 0, hi = 0;          // current frontier [lo,hi)
  uint64_t level = 0;               // level of the frontier (1 = initial states)
  uint64_t init_states = 0;
  uint64_t generated = 0;
  uint64_t depth = 0;
  int verdict = TLAG_V_RUNNING; int detail = 0, detail2 = 0; uint64_t viol_idx = 0;
  double dev_seconds = 0;
  uint64_t launches = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 148;
  int frame_class = 0;
  bool uses_ext = false;            // the image contains LEXLT / SFIND / SINS / EMITD (full interpreter needed)
  bool restarting = false;
  double growth_hint = 4.0;
  void* d_sort = nullptr; uint64_t sort_bytes = 0;
  unsigned long long* d_sent = nullptr;
  uint8_t* d_succ = nullptr; uint64_t succ_cap = 0;   // sliced build: per-frontier-state "has a successor" flags
  unsigned long long* d_dig = nullptr;                // tlag_digest accumulators (XOR, SUM)
  // peer-memory exchange (tlag_p2p_*): one cudaMalloc'ed block [meta | inbox x 2 buffers], exported through CUDA IPC
  P2PParams q;
  void* d_p2p = nullptr; uint64_t p2p_bytes = 0;
  uint32_t* d_send_own = nullptr;
  unsigned int* d_tickets = nullptr;
  void* peer_base[16] = {nullptr};
  unsigned long long p2p_seq = 0;
  int p2p_slot = -1;                                   // entry of g_p2p_parked this engine's exchange buffers belong to
  bool p2p_ready = false;
  uint64_t p2p_level_start = 0, p2p_level_generated = 0;   // for tlag_p2p_rollback: store tail before the level, what it added
  // TLAG_F_KEEP_GOING: the first violation is remembered, the search goes on to the fixpoint
  int kg_verdict = 0, kg_detail = 0, kg_detail2 = 0; uint64_t kg_idx = 0;
  std::string err;
  uint32_t* d_scratch = nullptr; uint64_t scratch_words = 0; uint8_t* d_flags = nullptr; uint64_t flags_cap = 0;
};

// ------------------------------------------------------------------ sliced native build
// The slice kernels live in translation units of their own (csrc/native/<tag>_part<k>.cu, generated and compiled in
// parallel: tla_rust_b200/engine.py build_sliced_library); this unit only knows their launchers.
#ifdef TLAG_SLICED_INC
#include TLAG_SLICED_INC     /* the generated constants: TLAG_SL_W, TLAG_SL_*_LIST, program length + FNV */
#ifndef TLAG_SL_BLOCK
#define TLAG_SL_BLOCK 256
#endif
typedef void (*sl_launch_t)(const DevParams*, unsigned long long, unsigned long long, unsigned, cudaStream_t);
#define TLAG_SL_DECL_INV(j) extern "C" void tlag_sl_launch_inv_##j(const DevParams*, unsigned long long, unsigned long long, unsigned, cudaStream_t);
#define TLAG_SL_DECL_NEXT(j) extern "C" void tlag_sl_launch_next_##j(const DevParams*, unsigned long long, unsigned long long, unsigned, cudaStream_t);
TLAG_SL_INV_LIST(TLAG_SL_DECL_INV)
TLAG_SL_NEXT_LIST(TLAG_SL_DECL_NEXT)
#define TLAG_SL_ADDR_INV(j) tlag_sl_launch_inv_##j,
#define TLAG_SL_ADDR_NEXT(j) tlag_sl_launch_next_##j,
static const sl_launch_t kSlInv[] = { TLAG_SL_INV_LIST(TLAG_SL_ADDR_INV) nullptr };
static const sl_launch_t kSlNext[] = { TLAG_SL_NEXT_LIST(TLAG_SL_ADDR_NEXT) nullptr };

// deadlock = a frontier state for which no slice of Next produced a successor
__global__ void k_sl_deadlock(DevParams p, unsigned long long lo, unsigned long long hi) {
  const unsigned long long i = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi && !p.succ_flag[i - lo]) report_min(&p.ctr->viol_deadlock, i << 20);
}
#endif  // TLAG_SLICED_INC

#ifndef TLAG_SLICED_INC
// ------------------------------------------------------------------ wave kernel
// Warp-scheduled interpreter.  A SIMT interpreter that lets every lane follow its own pc pays a
// divergent fetch/decode/dispatch per distinct opcode per step.  Here the warp executes, at every
// step, the ONE instruction at the minimum pc among its running lanes (REDUX min), for exactly the
// lanes that are at that pc: fetch (shared-memory broadcast), decode and dispatch are warp-uniform,
// frame accesses use the same word index in every lane (coalesced local memory), and lanes that
// took different branches re-join as soon as the laggards catch up (forward progress is by
// construction: the lowest pc always advances).
enum { L_RUN = 0, L_EMIT = 1, L_DONE = 2, L_STOP = 3 };
#define TLAG_PC_PARKED 0xFFFFFFFFu

// A lane that is not running parks its pc at TLAG_PC_PARKED (so it never wins the min) and keeps the
// pc to resume from in `rpc`.  Per step: one REDUX min, one shared-memory broadcast fetch, uniform
// decode/dispatch; only lanes whose pc equals the minimum execute.  (v1 of this loop spent 41 of ~73
// SASS instructions per step on bookkeeping and fetched the instruction with a generic per-lane load:
// profiles/r1_k_wave_warpsched_b2_ncu.txt.)
template <bool SMEM, bool LEAN>
__device__ __forceinline__ void warp_vm(const uint64_t* __restrict__ gcode, const uint64_t* scode,
                                        const int32_t* __restrict__ cpool, int32_t* frame, uint32_t& pc,
                                        uint32_t& rpc, int& ev_out, int32_t& info, int32_t& info2) {
  for (;;) {
    const uint32_t pcm = __reduce_min_sync(0xffffffffu, pc);
    if (pcm == TLAG_PC_PARKED) break;
    const uint64_t w = SMEM ? scode[pcm] : __ldg(gcode + pcm);
    if (pc == pcm) {
      const int ev = LEAN ? tlag_vm_exec_lean(w, cpool, frame, &pc, &info, &info2)
                          : tlag_vm_exec(w, cpool, frame, &pc, &info, &info2);
      if (ev >= 0) { ev_out = ev; rpc = pc; pc = TLAG_PC_PARKED; }
    }
  }
}

// MODE 0: fused insert (single GPU).  MODE 1: route successors to per-owner send regions.
// LEAN: interpreter without the extension ops, for models that do not use them (small frames only).
template <int FRAME, int MODE, bool SMEM, bool LEAN = false>
__global__ void __launch_bounds__(TLAG_BLOCK, TLAG_WAVE_OCC(FRAME, SMEM))
k_wave(DevParams p, unsigned long long lo, unsigned long long hi) {
  extern __shared__ uint64_t s_code[];
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < p.code_len; i += blockDim.x) s_code[i] = p.code[i];
    __syncthreads();
  }
  const uint64_t* code = p.code;
  int32_t frame[FRAME];
  uint32_t succ[TLAG_MAXW];
  const unsigned lane = threadIdx.x & 31;
  const int W = p.W;
  unsigned long long gen_local = 0;
  // TLAG_F_EXACT (launched as ONE warp): lane 0 dequeues one state at a time, in index = FIFO order, and the search
  // stops at the first Assert failure / deadlock -- TLC's single worker, on the device
  const bool exact = MODE == 0 && (p.flags & TLAG_F_EXACT) != 0;

  for (;;) {
    if (exact && *(volatile unsigned long long*)&p.ctr->stop) break;
    unsigned long long chunk = 0;
    if (lane == 0) chunk = atomicAdd(&p.ctr->work, 1ULL);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    const unsigned long long first = lo + chunk * (exact ? 1ULL : 32ULL);
    if (first >= hi) break;
    const unsigned long long idx = first + lane;
    const bool active = idx < hi && (!exact || lane == 0);
    if (active) {
      const uint32_t* src = p.states + idx * (unsigned long long)W;
      for (int i = 0; i < W; ++i) succ[i] = src[i];
      tlag_unpack(p.layout, p.n_slots, succ, frame + p.n_off);
    }
    bool trapped = false;
    int st, ev = 0;
    int32_t info = 0, info2 = 0;
    uint32_t pc, rpc = 0;
    // One interpreter loop for both programs (a single copy of the dispatch switch: the first version
    // inlined it twice and stalled on instruction fetch -- ncu stall_no_instruction 14 per issue).
    // Each lane first runs the invariant program on the state it expands; at its HALT it continues
    // with the next-state program (the invariant code sits at lower pcs, so those lanes go first).
    int phase = p.n_inv > 0 ? 0 : 1;
    st = active ? L_RUN : L_DONE;
    pc = active ? (phase == 0 ? p.entry_inv : p.entry_next) : TLAG_PC_PARKED;
    unsigned nsucc = 0;
    for (;;) {
      warp_vm<SMEM, LEAN>(code, s_code, p.cpool, frame, pc, rpc, ev, info, info2);
      int32_t act = 0;
      if (st == L_RUN) {        // this lane stopped on an event
        if (ev == TLAG_EV_EMIT) { st = L_EMIT; ++nsucc; ++gen_local; }
        else if (ev == TLAG_EV_HALT) {
          if (phase == 0) { phase = 1; pc = p.entry_next; } else st = L_DONE;
        }
        else if (ev == TLAG_EV_GEN) { ++nsucc; ++gen_local; pc = rpc; }
        else if (ev == TLAG_EV_INVF) {
          if (phase == 0) report_min(&p.ctr->viol_inv, (idx << 20) | (unsigned)(info & 0xFFFFF));
          pc = rpc;
        }
        else if (ev == TLAG_EV_ASSERT) {
          report_min(&p.ctr->viol_assert, (idx << 20) | (unsigned)(info & 0xFFFFF));
          if (exact) { atomicExch(&p.ctr->stop, 1ULL); st = L_DONE; trapped = true; }   // stop here: nothing after it is generated
          else pc = rpc;
        }
        else {
          report_min(&p.ctr->viol_trap, (idx << 20) | ((unsigned long long)(info & 15) << 16) | (unsigned)(info2 & 0xFFFF));
          st = L_DONE; trapped = true;
        }
      }
      if (__any_sync(0xffffffffu, pc != TLAG_PC_PARKED)) continue;   // resume the lanes that only reported
      bool has = (st == L_EMIT);
      if (!__any_sync(0xffffffffu, has)) break;                 // every lane is done
      act = info;
      unsigned long long fp = 0;
      if (has) {
        int ov;
        if (!LEAN && info2 > 0) {   // EMITD: copy of the parent's packed words, dirty slot ranges re-packed
          const uint32_t* src = p.states + idx * (unsigned long long)W;
          for (int i = 0; i < W; ++i) succ[i] = src[i];
          ov = tlag_pack_ranges(p.layout, p.cpool, info2, frame + p.p_off, succ);
        } else {
          ov = tlag_pack(p.layout, p.n_slots, frame + p.p_off, succ, W);
        }
        if (ov) {
          report_min(&p.ctr->viol_trap, (idx << 20) | (2ULL << 16) | (unsigned)((ov - 1) & 0xFFFF));
          has = false;
        } else {
          fp = tlag_fingerprint(succ, W);
        }
      }
      if (MODE == 0) {
        int ins = has ? seen_insert(p.table, p.mask, fp) : 0;
        if (ins < 0) { atomicExch(&p.ctr->table_full, 1ULL); ins = 0; }
        const bool isnew = ins > 0;
        const unsigned m = __ballot_sync(0xffffffffu, isnew);
        if (m) {
          const int leader = __ffs((int)m) - 1;
          unsigned long long base = 0;
          if ((int)lane == leader) base = atomicAdd(&p.ctr->n_states, (unsigned long long)__popc(m));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (isnew) {
            const unsigned long long pos = base + (unsigned long long)__popc(m & lanemask_lt());
            if (pos < p.cap_states) {
              uint32_t* dst = p.states + pos * (unsigned long long)W;
              for (int i = 0; i < W; ++i) dst[i] = succ[i];
              p.parent[pos] = (uint32_t)idx;
              p.meta[pos] = ((uint32_t)act << 8) | (uint32_t)(p.rank & 0xFF);
            } else {
              atomicExch(&p.ctr->store_overflow, 1ULL);
            }
          }
        }
      } else {
        // route: owner = high bits of the fingerprint scaled to n_ranks (hash-range partition).
        // A successor whose fingerprint this rank has already routed is known to its owner: drop it
        // here (direct-mapped, exact-compare cache: misses only cost a redundant record).
        // (the slot is written only once the record is in the send region: a chunk re-run after a region overflow
        // must find none of its own dropped records in the cache)
        unsigned long long* cslot = (has && p.sent_cache) ? p.sent_cache + (fp & p.sent_mask) : nullptr;
        if (cslot && __ldcv(cslot) == fp) { has = false; cslot = nullptr; }
        int owner = has ? (int)tlag_owner_k(succ, W, (uint32_t)p.n_ranks, p.owner_words) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        if (has) {
          const int leader = __ffs((int)peers) - 1;
          unsigned long long base = 0;
          if ((int)lane == leader) base = atomicAdd(&p.ctr->send_count[owner], (unsigned long long)__popc(peers));
          base = __shfl_sync(peers, base, leader);
          const unsigned long long pos = base + (unsigned long long)__popc(peers & lanemask_lt());
          if (pos < p.region_cap) {
            uint32_t* dst = p.send + ((unsigned long long)owner * p.region_cap + pos) * (unsigned long long)(W + 2);
            for (int i = 0; i < W; ++i) dst[i] = succ[i];
            dst[W] = (uint32_t)idx;
            dst[W + 1] = ((uint32_t)act << 8) | (uint32_t)(p.rank & 0xFF);
            if (cslot) *cslot = fp;
          } else {
            atomicExch(&p.ctr->route_overflow, 1ULL);
          }
        }
      }
      if (st == L_EMIT) { st = L_RUN; pc = rpc; }
    }
    if (active && nsucc == 0 && !trapped && (p.flags & TLAG_F_DEADLOCK_CHECK)) {
      report_min(&p.ctr->viol_deadlock, idx << 20);
      if (exact) atomicExch(&p.ctr->stop, 1ULL);
    }
  }
  // generated counter: warp reduce then one atomic
  for (int o = 16; o > 0; o >>= 1) gen_local += __shfl_down_sync(0xffffffffu, gen_local, o);
  if (lane == 0 && gen_local) atomicAdd(&p.ctr->generated, gen_local);
}

#endif  // !TLAG_SLICED_INC

// ------------------------------------------------------------------ K1 alone: fingerprint + probe/insert
// One thread per candidate state; W words loaded with the widest aligned vector the layout allows.
template <int VEC>
__global__ void __launch_bounds__(256) k_probe(const uint32_t* __restrict__ states, unsigned long long n, int W,
                                               unsigned long long* table, unsigned long long mask,
                                               uint8_t* __restrict__ is_new, Counters* ctr) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* src = states + i * (unsigned long long)W;
  uint64_t h = tlag_fp_init(W);
  if (VEC == 4) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    const int n4 = W / 4;
#pragma unroll 5
    for (int k = 0; k < n4; ++k) {
      const uint4 v = __ldg(s4 + k);
      h = tlag_fp_pair(h, v.x, v.y);
      h = tlag_fp_pair(h, v.z, v.w);
    }
  } else if (VEC == 2) {
    const uint2* s2 = reinterpret_cast<const uint2*>(src);
    for (int k = 0; k < W / 2; ++k) { const uint2 v = __ldg(s2 + k); h = tlag_fp_pair(h, v.x, v.y); }
  } else {
    int k = 0;
    for (; k + 1 < W; k += 2) h = tlag_fp_pair(h, __ldg(src + k), __ldg(src + k + 1));
    if (k < W) h = tlag_fp_tail(h, __ldg(src + k));
  }
  const unsigned long long fp = tlag_fp_final(h, W);
  int ins = seen_insert(table, mask, fp);
  if (ins < 0) { atomicExch(&ctr->table_full, 1ULL); ins = 0; }
  is_new[i] = (uint8_t)ins;
}

// K1, coalesced form: a CTA of 256 threads owns 256 consecutive candidates = one contiguous span of
// 256*W words.  The span is fetched with fully coalesced 128-bit loads (every 32-byte sector is read
// exactly once; ncu on the strided v0 showed 2.4x DRAM read amplification from L1 thrash), staged in
// shared memory with an odd row stride (W|1 words -> conflict-free), then each thread fingerprints its
// own row and probes the table.
#define TLAG_PROBE_ROWS 1   /* rows per thread of the staged form */
__global__ void __launch_bounds__(256) k_probe_staged(const uint32_t* __restrict__ states, unsigned long long n, int W,
                                                      unsigned long long* table, unsigned long long mask,
                                                      uint8_t* __restrict__ is_new, Counters* ctr) {
  extern __shared__ uint32_t s_rows[];
  const int stride = W | 1;
  const unsigned ROWS = 256 * TLAG_PROBE_ROWS;
  const unsigned long long base = (unsigned long long)blockIdx.x * ROWS;
  const unsigned nst = (unsigned)((n - base) < (unsigned long long)ROWS ? (n - base) : (unsigned long long)ROWS);
  const uint32_t* g = states + base * (unsigned long long)W;
  const unsigned total = nst * (unsigned)W;          // words in this CTA's span
  const unsigned total4 = total >> 2;
  const uint4* g4 = reinterpret_cast<const uint4*>(g);
  for (unsigned i = threadIdx.x; i < total4; i += 256) {
    const uint4 v = __ldcs(g4 + i);                  // streaming: each word is used once
    unsigned wi = i << 2;
    unsigned row = wi / (unsigned)W, col = wi - row * (unsigned)W;
    s_rows[row * stride + col] = v.x; if (++col == (unsigned)W) { col = 0; ++row; }
    s_rows[row * stride + col] = v.y; if (++col == (unsigned)W) { col = 0; ++row; }
    s_rows[row * stride + col] = v.z; if (++col == (unsigned)W) { col = 0; ++row; }
    s_rows[row * stride + col] = v.w;
  }
  for (unsigned wi = (total4 << 2) + threadIdx.x; wi < total; wi += 256) {   // ragged tail (< 4 words)
    const unsigned row = wi / (unsigned)W, col = wi - row * (unsigned)W;
    s_rows[row * stride + col] = g[wi];
  }
  __syncthreads();
  unsigned long long fp[TLAG_PROBE_ROWS];
  unsigned long long slot[TLAG_PROBE_ROWS];
  unsigned long long cur[TLAG_PROBE_ROWS];
  bool live[TLAG_PROBE_ROWS];
#pragma unroll
  for (int q = 0; q < TLAG_PROBE_ROWS; ++q) {
    const unsigned row = threadIdx.x + 256u * q;
    live[q] = row < nst;
    const uint32_t* r = s_rows + (live[q] ? row : 0) * stride;
    uint64_t h = tlag_fp_init(W);
    int k = 0;
    for (; k + 1 < W; k += 2) h = tlag_fp_pair(h, r[k], r[k + 1]);
    if (k < W) h = tlag_fp_tail(h, r[k]);
    fp[q] = tlag_fp_final(h, W);
    slot[q] = fp[q] & mask;
  }
  // first probe of every row issued back to back (independent loads in flight), then resolved
#pragma unroll
  for (int q = 0; q < TLAG_PROBE_ROWS; ++q) cur[q] = live[q] ? __ldcv(&table[slot[q]]) : 0ULL;
#pragma unroll
  for (int q = 0; q < TLAG_PROBE_ROWS; ++q) {
    if (!live[q]) continue;
    int ins;
    if (cur[q] == fp[q]) ins = 0;
    else if (cur[q] == 0ULL) {
      const unsigned long long old = atomicCAS(&table[slot[q]], 0ULL, fp[q]);
      ins = (old == 0ULL) ? 1 : (old == fp[q] ? 0 : seen_insert(table, mask, fp[q]));
    } else {
      ins = seen_insert(table, mask, fp[q]);        // collision chain: generic path from the home slot
    }
    if (ins < 0) { atomicExch(&ctr->table_full, 1ULL); ins = 0; }
    is_new[base + threadIdx.x + 256u * q] = (uint8_t)ins;
  }
}

// K1, Blackwell form (W a multiple of 4 with W/4 odd: W = 4, 12, 20 ...).  A persistent CTA streams its tiles of
// 512 rows through a 2-stage shared-memory ring filled by 1-D bulk copies of the TMA engine (cp.async.bulk ...
// mbarrier::complete_tx::bytes -- UBLKCP in SASS): one elected thread arms the stage's mbarrier with the byte count and
// issues the copy, the copy of tile k+1 is in flight while tile k is hashed and probed, and no thread spends issue
// slots on LDG -> STS staging.  Rows are dense in shared memory (pitch = 4 W bytes, no padding) and read back with
// 128-bit LDS: for W/4 odd the 8 rows of a quarter-warp fall into 8 distinct 4-bank groups, so the loads are
// conflict-free.  Every thread owns TWO rows per tile: both fingerprints are computed first, both home-slot loads are
// issued back to back (two independent DRAM round trips in flight per thread), then resolved.
#define TLAG_TMA_ROWS 2
#define TLAG_TMA_STAGES 2
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}

__global__ void __launch_bounds__(256) k_probe_tma(const uint32_t* __restrict__ states, unsigned long long n, int W,
                                                   unsigned long long* table, unsigned long long mask,
                                                   uint8_t* __restrict__ is_new, Counters* ctr) {
  extern __shared__ __align__(128) uint8_t s_ring[];
  __shared__ __align__(8) uint64_t s_full[TLAG_TMA_STAGES];
  constexpr unsigned ROWS = 256 * TLAG_TMA_ROWS;
  const unsigned row_bytes = (unsigned)W * 4u;
  const unsigned tile_bytes = ROWS * row_bytes;
  const unsigned long long n_tiles = (n + ROWS - 1) / ROWS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TLAG_TMA_STAGES; ++s) mbar_init(&s_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](unsigned long long tile, int stage) {
    const unsigned long long first = tile * ROWS;
    const unsigned rows = (unsigned)((n - first) < (unsigned long long)ROWS ? (n - first) : (unsigned long long)ROWS);
    const unsigned bytes = rows * row_bytes;
    mbar_expect_tx(&s_full[stage], bytes);
    bulk_g2s(s_ring + (size_t)stage * tile_bytes, states + first * (unsigned long long)W, bytes, &s_full[stage]);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < TLAG_TMA_STAGES; ++s) {
      const unsigned long long t = blockIdx.x + (unsigned long long)s * gridDim.x;
      if (t < n_tiles) issue(t, s);
    }
  unsigned long long k = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
    const int stage = (int)(k % TLAG_TMA_STAGES);
    mbar_wait(&s_full[stage], (unsigned)((k / TLAG_TMA_STAGES) & 1));
    const unsigned long long first = tile * ROWS;
    const unsigned nst = (unsigned)((n - first) < (unsigned long long)ROWS ? (n - first) : (unsigned long long)ROWS);
    const uint8_t* base = s_ring + (size_t)stage * tile_bytes;
    unsigned long long fp[TLAG_TMA_ROWS], slot[TLAG_TMA_ROWS], cur[TLAG_TMA_ROWS];
    bool live[TLAG_TMA_ROWS];
#pragma unroll
    for (int q = 0; q < TLAG_TMA_ROWS; ++q) {
      const unsigned row = threadIdx.x + 256u * q;
      live[q] = row < nst;
      const uint4* r4 = reinterpret_cast<const uint4*>(base + (size_t)(live[q] ? row : 0) * row_bytes);
      uint64_t h = tlag_fp_init(W);
      for (int j = 0; j < W / 4; ++j) {
        const uint4 v = r4[j];
        h = tlag_fp_pair(h, v.x, v.y);
        h = tlag_fp_pair(h, v.z, v.w);
      }
      fp[q] = tlag_fp_final(h, W);
      slot[q] = fp[q] & mask;
    }
#pragma unroll
    for (int q = 0; q < TLAG_TMA_ROWS; ++q) cur[q] = live[q] ? __ldcv(&table[slot[q]]) : 0ULL;
    __syncthreads();                                   // every row of this stage has been read: refill it
    if (threadIdx.x == 0) {
      const unsigned long long nt = tile + (unsigned long long)TLAG_TMA_STAGES * gridDim.x;
      if (nt < n_tiles) issue(nt, stage);
    }
#pragma unroll
    for (int q = 0; q < TLAG_TMA_ROWS; ++q) {
      if (!live[q]) continue;
      int ins;
      if (cur[q] == fp[q]) ins = 0;
      else if (cur[q] == 0ULL) {
        const unsigned long long old = atomicCAS(&table[slot[q]], 0ULL, fp[q]);
        ins = (old == 0ULL) ? 1 : (old == fp[q] ? 0 : seen_insert(table, mask, fp[q]));
      } else {
        ins = seen_insert(table, mask, fp[q]);
      }
      if (ins < 0) { atomicExch(&ctr->table_full, 1ULL); ins = 0; }
      is_new[first + threadIdx.x + 256u * q] = (uint8_t)ins;
    }
  }
}

// insert routed records (W state words + parent + meta) into this rank's shard
__global__ void __launch_bounds__(256) k_insert_records(DevParams p, const uint32_t* __restrict__ rec,
                                                        unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const int W = p.W;
  uint32_t w[TLAG_MAXW];
  bool has = i < n;
  unsigned long long fp = 0;
  const uint32_t* src = rec + i * (unsigned long long)(W + 2);
  if (has) {
    for (int k = 0; k < W; ++k) w[k] = src[k];
    fp = tlag_fingerprint(w, W);
  }
  int ins = has ? seen_insert(p.table, p.mask, fp) : 0;
  if (ins < 0) { atomicExch(&p.ctr->table_full, 1ULL); ins = 0; }
  const bool isnew = ins > 0;
  const unsigned m = __ballot_sync(0xffffffffu, isnew);
  if (m) {
    const int leader = __ffs((int)m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(&p.ctr->n_states, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (isnew) {
      const unsigned long long pos = base + (unsigned long long)__popc(m & lanemask_lt());
      if (pos < p.cap_states) {
        uint32_t* dst = p.states + pos * (unsigned long long)W;
        for (int k = 0; k < W; ++k) dst[k] = w[k];
        p.parent[pos] = src[W];
        p.meta[pos] = src[W + 1];
      } else {
        atomicExch(&p.ctr->store_overflow, 1ULL);
      }
    }
  }
}

// ------------------------------------------------------------------ peer-memory exchange (multi-GPU)
// One process per GPU; the state space is partitioned by tlag_owner().  Instead of handing the per-owner send regions to
// NCCL (two all_to_alls and three host round trips per chunk in round 1), every rank maps every other rank's INBOX with
// CUDA IPC and the exchange is device code over NVLink / NVSwitch:
//   k_push          copies this rank's send region for owner d into d's inbox (16-byte stores over NVLink), then
//                   publishes {count, seq} in d's memory with a system-scope release;
//   k_insert_inbox  waits (acquire) until every source has published chunk `seq`, inserts the records into this
//                   rank's seen-set shard / state store (the next frontier), then acknowledges the chunk in the
//                   sources' memory so that they may reuse the buffer.
// Inboxes are double-buffered by chunk parity; a whole BFS level is enqueued on the engine's stream (expand kernels,
// push, insert per chunk) and the host synchronises once per level for the termination all-reduce.
// grid: (blocks_per_dst, n_ranks).  Copies min(send_count[dst], region_cap) records to dst's inbox.
__global__ void __launch_bounds__(256) k_push(P2PParams q, const Counters* ctr, unsigned long long seq) {
  const int dst = blockIdx.y;
  __shared__ int s_last;
  P2PMeta* mine_at_dst = q.peer_meta[dst] + q.rank;
  if (threadIdx.x == 0 && seq > 2) {
    // the buffer (seq & 1) at dst was last used for my chunk seq - 2: wait until dst has consumed it
    const unsigned long long* ack = &q.meta[dst].ack;
    while (ld_acquire_sys(ack) + 2 < seq) __nanosleep(200);
  }
  __syncthreads();
  unsigned long long n = ctr->send_count[dst];
  if (n > q.region_cap) n = q.region_cap;
  const unsigned long long words = n * (unsigned long long)q.rec_words;
  const uint32_t* src = q.send + (unsigned long long)dst * q.region_cap * q.rec_words;
  uint32_t* out = q.peer_inbox[dst] + (((seq & 1ULL) * q.n_ranks + q.rank) * q.inbox_cap) * q.rec_words;
  const unsigned long long w4 = words >> 2;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* o4 = reinterpret_cast<uint4*>(out);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < w4;
       i += (unsigned long long)gridDim.x * blockDim.x)
    o4[i] = s4[i];
  if (blockIdx.x == 0)
    for (unsigned long long i = (w4 << 2) + threadIdx.x; i < words; i += blockDim.x) out[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&q.tickets[dst], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    q.tickets[dst] = 0;
    __threadfence_system();
    *(volatile unsigned long long*)&mine_at_dst->count[seq & 1ULL] = n;
    __threadfence_system();
    st_release_sys(&mine_at_dst->seq[seq & 1ULL], seq);
  }
}

// Inserts every source's records of chunk `seq` (grid-stride over sources and records, whole warps).
__global__ void __launch_bounds__(256) k_insert_inbox(DevParams p, P2PParams q, unsigned long long seq) {
  const unsigned lane = threadIdx.x & 31;
  const int W = p.W;
  __shared__ unsigned long long s_n;
  __shared__ int s_last;
  for (int src = 0; src < q.n_ranks; ++src) {
    if (threadIdx.x == 0) {
      const P2PMeta* m = &q.meta[src];
      while (ld_acquire_sys(&m->seq[seq & 1ULL]) != seq) __nanosleep(200);
      s_n = *(volatile const unsigned long long*)&m->count[seq & 1ULL];
    }
    __syncthreads();
    const unsigned long long n = s_n;
    const uint32_t* rec = q.inbox + (((seq & 1ULL) * q.n_ranks + src) * q.inbox_cap) * q.rec_words;
    const unsigned long long n32 = (n + 31ULL) & ~31ULL;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n32;
         i += (unsigned long long)gridDim.x * blockDim.x) {
      uint32_t w[TLAG_MAXW];
      const bool has = i < n;
      unsigned long long fp = 0;
      const uint32_t* r = rec + i * (unsigned long long)(W + 2);
      if (has) {
        for (int k = 0; k < W; ++k) w[k] = __ldcg(r + k);      // written by a peer over NVLink: read through L2
        fp = tlag_fingerprint(w, W);
      }
      int ins = has ? seen_insert(p.table, p.mask, fp) : 0;
      if (ins < 0) { atomicExch(&p.ctr->table_full, 1ULL); ins = 0; }
      const bool isnew = ins > 0;
      const unsigned mk = __ballot_sync(0xffffffffu, isnew);
      if (mk) {
        const int leader = __ffs((int)mk) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(&p.ctr->n_states, (unsigned long long)__popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (isnew) {
          const unsigned long long pos = base + (unsigned long long)__popc(mk & lanemask_lt());
          if (pos < p.cap_states) {
            uint32_t* dst = p.states + pos * (unsigned long long)W;
            for (int k = 0; k < W; ++k) dst[k] = w[k];
            p.parent[pos] = __ldcg(r + W);
            p.meta[pos] = __ldcg(r + W + 1);
          } else {
            atomicExch(&p.ctr->store_overflow, 1ULL);
          }
        }
      }
    }
    __syncthreads();
  }
  // every block is done with every source: the last one acknowledges the chunk to all sources
  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(&q.tickets[16], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x < (unsigned)q.n_ranks) {
    if (threadIdx.x == 0) q.tickets[16] = 0;
    __threadfence_system();
    st_release_sys(&(q.peer_meta[threadIdx.x] + q.rank)->ack, seq);
  }
}

// checksum of checksums over the state store: XOR and SUM (mod 2^64) of all fingerprints
__global__ void k_digest(DevParams p, unsigned long long n, unsigned long long* out2) {
  unsigned long long x = 0, sm = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t w[TLAG_MAXW];
    for (int k = 0; k < p.W; ++k) w[k] = p.states[i * (unsigned long long)p.W + k];
    const unsigned long long fp = tlag_fingerprint(w, p.W);
    x ^= fp; sm += fp;
  }
  for (int o = 16; o > 0; o >>= 1) { x ^= __shfl_down_sync(0xffffffffu, x, o); sm += __shfl_down_sync(0xffffffffu, sm, o); }
  if ((threadIdx.x & 31) == 0) { atomicXor(&out2[0], x); atomicAdd(&out2[1], sm); }
}

// ---- frontier clustering -----------------------------------------------------------------------
// The warp-scheduled interpreter executes the union of its 32 lanes' paths, so its efficiency is
// (mean path length) / (union length).  States that agree on the tail of the packed vector (the last
// variables, for the Paxos models the high bits of the msgs bitset) follow similar paths: measured on
// a 4096-state window of MCPaxos3_b3 level 14, lane utilisation is 0.35 in discovery order, 0.31 in random
// order and 0.54 when sorted by the last packed word.  Each newly discovered level slice is therefore
// reordered (stable data: nothing references those indices yet) by a 64-bit key of its last two words.
__global__ void k_sort_keys(const uint32_t* __restrict__ states, unsigned long long n, int W,
                            unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* s = states + i * (unsigned long long)W;
  const unsigned long long hi = s[W - 1], lo = W >= 2 ? s[W - 2] : 0;
  keys[i] = (hi << 32) | lo;
  idx[i] = (uint32_t)i;
}

__global__ void k_gather_slice(const uint32_t* __restrict__ states, const uint32_t* __restrict__ parent,
                               const uint32_t* __restrict__ meta, const uint32_t* __restrict__ idx,
                               unsigned long long n, int W, uint32_t* __restrict__ o_states,
                               uint32_t* __restrict__ o_parent, uint32_t* __restrict__ o_meta) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long j = idx[i];
  for (int k = 0; k < W; ++k) o_states[i * (unsigned long long)W + k] = states[j * (unsigned long long)W + k];
  o_parent[i] = parent[j];
  o_meta[i] = meta[j];
}

// rebuild the table from the state store after growing it
__global__ void k_rehash(DevParams p, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[TLAG_MAXW];
  for (int k = 0; k < p.W; ++k) w[k] = p.states[i * (unsigned long long)p.W + k];
  seen_insert(p.table, p.mask, tlag_fingerprint(w, p.W));
}

// Exchange buffers outlive the engine that created them: allocating an 8 GB inbox, exporting it and mapping seven peers'
// inboxes costs seconds, and a host process that checks one model after another (bench.py's end-to-end leg does) would
// pay it per job.  tlag_destroy parks the buffers of a p2p engine here; the next engine of the same shape (device, ranks,
// record size, capacity) adopts them -- same IPC handle, peers' mappings still valid, chunk numbering continued (every
// rank has run the same number of chunks, so the inbox protocol needs no reset).
struct P2PParked {
  int device, n_ranks, rank, rec_words; uint64_t cap;
  void* d_p2p; uint64_t bytes; uint32_t* d_send; unsigned int* d_tickets;
  void* peer_base[16]; uint8_t peer_handle[16][64]; bool peer_ok[16];
  unsigned long long seq; bool in_use;
};
static std::vector<P2PParked> g_p2p_parked;

// ------------------------------------------------------------------ host side
// Big device buffers (state store, parent / meta, seen-set, sort scratch, routed-fingerprint cache, send regions) come
// from the device's stream-ordered memory pool with its release threshold lifted: a buffer freed by a growth step or by
// tlag_destroy stays in the pool, and the next allocation of that size -- the next growth step, the next engine of a
// long-lived host process -- is served without going back to the driver.  cudaMalloc / cudaFree of multi-GB buffers
// cost tens of milliseconds each, and a fresh engine performs a dozen of them while its store and table grow (the
// end-to-end leg of bench.py creates and destroys an engine per job).  The peer-memory inbox stays on cudaMalloc: CUDA
// IPC cannot export pool memory.
static cudaError_t dmalloc(tlag_engine* e, void* pp, size_t bytes) {
  static const bool plain = getenv("TLAG_NO_POOL") != nullptr;
  return plain ? cudaMalloc((void**)pp, bytes) : cudaMallocAsync((void**)pp, bytes, e->stream);
}
static void dfree(tlag_engine* e, void* p) {
  static const bool plain = getenv("TLAG_NO_POOL") != nullptr;
  if (!p) return;
  if (plain) cudaFree(p); else cudaFreeAsync(p, e->stream);
}


static const int kFrameClasses[] = {64, 128, 256, 512, 1024, 2048, 4096, 8192};

#ifdef TLAG_SLICED_INC
// One BFS level (or one chunk of it, route mode) = the invariant kernels, the kernels of the slices of Next, and the
// deadlock scan, back to back on the engine's stream.
template <int MODE>
static cudaError_t launch_wave(tlag_engine* e, uint64_t lo, uint64_t hi) {
  const uint64_t n = hi - lo;
  if (n == 0) return cudaSuccess;
  const bool dl = (e->p.flags & TLAG_F_DEADLOCK_CHECK) != 0;
  if (dl) {
    if (n > e->succ_cap) {
      dfree(e, e->d_succ); e->d_succ = nullptr; e->succ_cap = 0;
      uint64_t cap = n + n / 2 + 4096;
      cudaError_t r = dmalloc(e, &e->d_succ, cap);
      if (r != cudaSuccess) return r;
      e->succ_cap = cap;
    }
    cudaError_t r = cudaMemsetAsync(e->d_succ, 0, n, e->stream);
    if (r != cudaSuccess) return r;
  }
  e->p.succ_flag = dl ? e->d_succ : nullptr;
  e->p.route = MODE;
  uint64_t blocks = (n + TLAG_SL_BLOCK - 1) / TLAG_SL_BLOCK;
  const uint64_t maxb = (uint64_t)e->sm_count * 64;
  if (blocks > maxb) blocks = maxb;
  if (e->p.n_inv > 0)
    for (int j = 0; kSlInv[j]; ++j) { kSlInv[j](&e->p, lo, hi, (unsigned)blocks, e->stream); e->launches++; }
  for (int j = 0; kSlNext[j]; ++j) { kSlNext[j](&e->p, lo, hi, (unsigned)blocks, e->stream); e->launches++; }
  if (dl) { k_sl_deadlock<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->p, lo, hi); e->launches++; }
  return cudaGetLastError();
}
#else
template <int MODE>
static cudaError_t launch_wave(tlag_engine* e, uint64_t lo, uint64_t hi) {
  const uint64_t n = hi - lo;
  uint64_t chunks = (n + 31) / 32;
  uint64_t blocks = (chunks + (TLAG_BLOCK / 32) - 1) / (TLAG_BLOCK / 32);
  if (blocks == 0) blocks = 1;
  size_t smem = e->p.code_in_smem ? (size_t)e->p.code_len * 8 : 0;
  void (*fn)(DevParams, unsigned long long, unsigned long long) = nullptr;
  const bool sm = e->p.code_in_smem != 0;
  const bool lean = !e->uses_ext;
  switch (e->frame_class) {
    case 0: fn = lean ? (sm ? k_wave<64, MODE, true, true> : k_wave<64, MODE, false, true>)
                      : (sm ? k_wave<64, MODE, true> : k_wave<64, MODE, false>); break;
    case 1: fn = lean ? (sm ? k_wave<128, MODE, true, true> : k_wave<128, MODE, false, true>)
                      : (sm ? k_wave<128, MODE, true> : k_wave<128, MODE, false>); break;
    case 2: fn = lean ? (sm ? k_wave<256, MODE, true, true> : k_wave<256, MODE, false, true>)
                      : (sm ? k_wave<256, MODE, true> : k_wave<256, MODE, false>); break;
    case 3: fn = lean ? (sm ? k_wave<512, MODE, true, true> : k_wave<512, MODE, false, true>)
                      : (sm ? k_wave<512, MODE, true> : k_wave<512, MODE, false>); break;
    case 4: fn = sm ? k_wave<1024, MODE, true> : k_wave<1024, MODE, false>; break;
    case 5: fn = sm ? k_wave<2048, MODE, true> : k_wave<2048, MODE, false>; break;
    case 6: fn = sm ? k_wave<4096, MODE, true> : k_wave<4096, MODE, false>; break;
    // 8192 words = 32 KB of local memory per thread (SSI at 4 transactions x 3 keys needs 6.9 K words)
    default: fn = sm ? k_wave<8192, MODE, true> : k_wave<8192, MODE, false>; break;
  }
  if (smem > 48 * 1024) {
    cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r != cudaSuccess) return r;
  }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, TLAG_BLOCK, smem) != cudaSuccess || occ < 1) occ = 1;
  const uint64_t maxb = (uint64_t)e->sm_count * (uint64_t)occ;   // persistent grid: one resident wave of CTAs
  if (blocks > maxb) blocks = maxb;
  if (MODE == 0 && (e->p.flags & TLAG_F_EXACT)) {                 // sequential replay: one warp, one state at a time
    fn<<<1, 32, smem, e->stream>>>(e->p, lo, hi);
    e->launches++;
    return cudaGetLastError();
  }
  fn<<<(unsigned)blocks, TLAG_BLOCK, smem, e->stream>>>(e->p, lo, hi);
  e->launches++;
  return cudaGetLastError();
}

#endif

static int alloc_table(tlag_engine* e, unsigned log2) {
  if (e->d_table) dfree(e, e->d_table);
  e->d_table = nullptr;
  const uint64_t slots = 1ULL << log2;
  CK(dmalloc(e, &e->d_table, slots * 8));
  CK(cudaMemsetAsync(e->d_table, 0, slots * 8, e->stream));
  e->table_log2 = log2;
  e->p.table = e->d_table;
  e->p.mask = slots - 1;
  return TLAG_OK;
}

static int grow_table_if_needed(tlag_engine* e, uint64_t expected_states) {
  const uint64_t slots = 1ULL << e->table_log2;
  if (expected_states * 2 <= slots) return TLAG_OK;
  unsigned log2 = e->table_log2;
  while ((1ULL << log2) < expected_states * 4) ++log2;
  Counters hc;
  CK(cudaMemcpyAsync(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  int r = alloc_table(e, log2);
  if (r) return r;
  const uint64_t n = hc.n_states;
  if (n) {
    k_rehash<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->p, n);
    e->launches++;
    CK(cudaGetLastError());
  }
  return TLAG_OK;
}

static int grow_store_if_needed(tlag_engine* e, uint64_t need) {
  if (need <= e->cap_states) return TLAG_OK;
  uint64_t ncap = e->cap_states;
  while (ncap < need) ncap *= (ncap < (1ULL << 25) ? 4 : 2);   // few re-allocations while the store is small
  if (e->m.max_states && ncap > e->m.max_states) ncap = e->m.max_states;
  // parent links are 32-bit indices into this rank's store and 0xFFFFFFFF marks an initial state: a store never holds
  // more than 2^32 - 2 states per GPU (beyond that the wave reports a store overflow instead of wrapping a link)
  if (ncap > 0xFFFFFFFEull) ncap = 0xFFFFFFFEull;
  if (ncap <= e->cap_states) return TLAG_OK;   // at the configured limit; overflow is detected by the kernel
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  const uint64_t per = (uint64_t)e->m.words_per_state * 4 + 8;
  if ((ncap - 0) * per > freeb * 9 / 10) {
    ncap = e->cap_states + (freeb * 8 / 10) / per;
    if (ncap <= e->cap_states) return TLAG_OK;
  }
  uint32_t *ns = nullptr, *np = nullptr, *nm = nullptr;
  CK(dmalloc(e, &ns, ncap * (uint64_t)e->m.words_per_state * 4));
  CK(dmalloc(e, &np, ncap * 4));
  CK(dmalloc(e, &nm, ncap * 4));
  Counters hc;
  CK(cudaMemcpyAsync(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const uint64_t n = hc.n_states;
  CK(cudaMemcpyAsync(ns, e->d_states, n * (uint64_t)e->m.words_per_state * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(np, e->d_parent, n * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(nm, e->d_meta, n * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  dfree(e, e->d_states); dfree(e, e->d_parent); dfree(e, e->d_meta);
  e->d_states = ns; e->d_parent = np; e->d_meta = nm;
  e->cap_states = ncap;
  e->p.states = ns; e->p.parent = np; e->p.meta = nm; e->p.cap_states = ncap;
  return TLAG_OK;
}

static int cluster_slice(tlag_engine* e, uint64_t first, uint64_t count) {
  static const bool off = getenv("TLAG_NO_SORT") != nullptr;
  if (off || count < 8192 || count > 0xFFFFFFFFull || (e->m.flags & TLAG_F_EXACT)) return TLAG_OK;   // exact: FIFO order is the point
  const uint64_t W = e->m.words_per_state;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (uint32_t*)nullptr, (uint32_t*)nullptr, (int)0);
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (uint32_t*)nullptr, (uint32_t*)nullptr, (unsigned long long)count);
  // one scratch allocation: keys in/out, idx in/out, cub temp, gathered slice
  const uint64_t bytes = count * 8 * 2 + count * 4 * 2 + cub_bytes + 256 + count * (W + 2) * 4;
  if (bytes > e->sort_bytes) {
    dfree(e, e->d_sort); e->d_sort = nullptr; e->sort_bytes = 0;
    uint64_t want = bytes * 2;                                  // geometric: levels grow, a re-allocation per level is slow
    if (dmalloc(e, &e->d_sort, want) != cudaSuccess) {
      cudaGetLastError();
      want = bytes;
      if (dmalloc(e, &e->d_sort, want) != cudaSuccess) { cudaGetLastError(); return TLAG_OK; }   // no memory: skip (optimisation only)
    }
    e->sort_bytes = want;
  }
  uint8_t* b = (uint8_t*)e->d_sort;
  unsigned long long* k_in = (unsigned long long*)b; b += count * 8;
  unsigned long long* k_out = (unsigned long long*)b; b += count * 8;
  uint32_t* i_in = (uint32_t*)b; b += count * 4;
  uint32_t* i_out = (uint32_t*)b; b += count * 4;
  uint32_t* g_states = (uint32_t*)b; b += count * W * 4;
  uint32_t* g_parent = (uint32_t*)b; b += count * 4;
  uint32_t* g_meta = (uint32_t*)b; b += count * 4;
  b = (uint8_t*)(((uintptr_t)b + 255) & ~(uintptr_t)255);
  void* cub_tmp = b;
  uint32_t* st = e->d_states + first * W;
  const unsigned blocks = (unsigned)((count + 255) / 256);
  k_sort_keys<<<blocks, 256, 0, e->stream>>>(st, count, (int)W, k_in, i_in);
  CK(cudaGetLastError());
  if (cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k_in, k_out, i_in, i_out, (unsigned long long)count, 0, 64,
                                      e->stream) != cudaSuccess) { e->err = "cub sort failed"; return TLAG_ECUDA; }
  k_gather_slice<<<blocks, 256, 0, e->stream>>>(st, e->d_parent + first, e->d_meta + first, i_out, count, (int)W,
                                                g_states, g_parent, g_meta);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(st, g_states, count * W * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->d_parent + first, g_parent, count * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->d_meta + first, g_meta, count * 4, cudaMemcpyDeviceToDevice, e->stream));
  e->launches += 3;
  return TLAG_OK;
}

#if defined(TLAG_SLICED_INC)
extern "C" const char* tlag_version(void) { return "tlag 0.2 (sm_100a) native sliced"; }
#else
extern "C" const char* tlag_version(void) { return "tlag 0.1 (sm_100a)"; }
#endif

extern "C" const char* tlag_last_error(const tlag_engine* e) { return e ? e->err.c_str() : "null engine"; }

extern "C" uint64_t tlag_kernel_launches(const tlag_engine* e) { return e ? e->launches : 0; }

extern "C" int tlag_create(const tlag_model* m, tlag_engine** out) {
  if (!m || !out) return TLAG_EINVAL;
  *out = nullptr;
  tlag_engine* e = new tlag_engine();
  *out = e;   // returned even on failure so that tlag_last_error works; caller destroys it
  e->m = *m;
  if (m->words_per_state == 0 || m->words_per_state > TLAG_MAXW) { e->err = "words_per_state out of range (1..128)"; return TLAG_EINVAL; }
  if (m->frame_words > 8192) { e->err = "frame_words > 8192 not supported"; return TLAG_EINVAL; }
#if defined(TLAG_SLICED_INC)
  if (m->flags & TLAG_F_EXACT) { e->err = "TLAG_F_EXACT (sequential replay) runs on the interpreter kernel: use libtlag.so"; return TLAG_EINVAL; }
  {  // this library holds ONE model's program as kernels (constant pool folded in): refuse anything else
    uint64_t h = 0xcbf29ce484222325ULL, hc = 0xcbf29ce484222325ULL;
    for (uint32_t i = 0; i < m->code_len; ++i) h = (h ^ m->code[i]) * 0x100000001b3ULL;
    for (uint32_t i = 0; i < m->cpool_len; ++i) hc = (hc ^ (uint64_t)(uint32_t)m->cpool[i]) * 0x100000001b3ULL;
    if (m->code_len != TLAG_NATIVE_CODE_LEN || h != TLAG_NATIVE_CODE_FNV || hc != TLAG_NATIVE_CPOOL_FNV ||
        m->frame_words > TLAG_SL_FRAME || m->words_per_state != TLAG_SL_W) {
      e->err = "this library was compiled for another model's program (native build)";
      return TLAG_EINVAL;
    }
  }
#endif
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { e->err = "no CUDA device available"; return TLAG_ECUDA; }
  CK(cudaSetDevice(m->device));
  // Seen-set probes are random 8-byte accesses: ncu showed every probe pulling a full 128-byte line
  // from DRAM (4 sectors) with the default L2 fetch granularity; 32 B makes a probe cost one sector.
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, m->device));
  e->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, m->device) == cudaSuccess) {
      uint64_t thr = ~0ULL;                       // keep freed buffers mapped (see dmalloc)
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else cudaGetLastError();
  }
  CK(cudaEventCreate(&e->ev0));
  CK(cudaEventCreate(&e->ev1));
  e->frame_class = 7;
  for (int i = 0; i < 8; ++i) if ((int)m->frame_words <= kFrameClasses[i]) { e->frame_class = i; break; }
  e->uses_ext = getenv("TLAG_VM_FULL") != nullptr;
  for (uint32_t i = 0; i < m->code_len && !e->uses_ext; ++i) {
    const uint32_t op = (uint32_t)(m->code[i] & 0xFF);
    if (op == OP_LEXLT || op == OP_SFIND || op == OP_SINS || op == OP_EMITD || op == OP_CALL || op == OP_RET) e->uses_ext = true;
  }
  // program image
  CK(cudaMalloc(&e->d_code, (size_t)(m->code_len ? m->code_len : 1) * 8));
  CK(cudaMemcpy(e->d_code, m->code, (size_t)m->code_len * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&e->d_cpool, (size_t)(m->cpool_len ? m->cpool_len : 1) * 4));
  CK(cudaMemcpy(e->d_cpool, m->cpool, (size_t)m->cpool_len * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&e->d_layout, (size_t)(m->n_slots ? m->n_slots : 1) * sizeof(tlag_slot)));
  CK(cudaMemcpy(e->d_layout, m->layout, (size_t)m->n_slots * sizeof(tlag_slot), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&e->d_ctr, sizeof(Counters)));
  Counters hc;
  memset(&hc, 0, sizeof(hc));
  hc.viol_inv = hc.viol_assert = hc.viol_trap = hc.viol_deadlock = ~0ULL;
  CK(cudaMemcpy(e->d_ctr, &hc, sizeof(hc), cudaMemcpyHostToDevice));
  // state store
  uint64_t cap = m->max_states ? m->max_states : (1ULL << 20);
  if (cap < 1024) cap = 1024;
  if (!m->max_states) cap = 1ULL << 20;
  else if (cap > (1ULL << 22)) cap = 1ULL << 22;   // start modest, grow on demand up to max_states
  e->cap_states = cap;
  CK(dmalloc(e, &e->d_states, cap * (uint64_t)m->words_per_state * 4));
  CK(dmalloc(e, &e->d_parent, cap * 4));
  CK(dmalloc(e, &e->d_meta, cap * 4));
  memset(&e->p, 0, sizeof(e->p));
  e->p.code = e->d_code; e->p.code_len = m->code_len;
  {
    // bytecode image in shared memory when it fits; TLAG_CODE_SMEM_MAX (bytes) lowers the threshold (tuning knob:
    // a large image takes L1 capacity away from the per-thread frames of big models)
    size_t lim = 200 * 1024;
    if (const char* s_ = getenv("TLAG_CODE_SMEM_MAX")) { const long v = atol(s_); if (v >= 0 && (size_t)v < lim) lim = (size_t)v; }
    // Big frames (container models such as raft) are bound by local-memory latency (ncu: long-scoreboard 37 per issue,
    // 16 warps/SM): they run with the image in global memory (read through L1) so that two CTAs fit per SM.
    if (m->frame_words > 512 && !getenv("TLAG_CODE_SMEM_FORCE")) lim = 0;
    e->p.code_in_smem = ((size_t)m->code_len * 8 <= lim) ? 1 : 0;
  }
  e->p.cpool = e->d_cpool; e->p.layout = e->d_layout; e->p.n_slots = (int)m->n_slots;
  e->p.entry_inv = m->entry_inv; e->p.entry_next = m->entry_next;
  e->p.n_off = 0; e->p.p_off = m->unpacked_words; e->p.W = (int)m->words_per_state;
  e->p.states = e->d_states; e->p.parent = e->d_parent; e->p.meta = e->d_meta; e->p.cap_states = cap;
  e->p.ctr = e->d_ctr; e->p.flags = m->flags; e->p.n_inv = (int)m->n_invariants;
  e->p.n_ranks = 1; e->p.rank = 0; e->p.owner_words = 2;
  unsigned log2 = m->table_slots_log2 ? m->table_slots_log2 : 22;
  int r = alloc_table(e, log2);
  if (r) return r;
  CK(cudaStreamSynchronize(e->stream));
  return TLAG_OK;
}

extern "C" void tlag_destroy(tlag_engine* e) {
  if (!e) return;
  cudaFree(e->d_code); cudaFree(e->d_cpool); cudaFree(e->d_layout); cudaFree(e->d_ctr); cudaFree(e->d_dig);
  if (e->stream) {
    dfree(e, e->d_states); dfree(e, e->d_parent); dfree(e, e->d_meta); dfree(e, e->d_table);
    dfree(e, e->d_scratch); dfree(e, e->d_flags); dfree(e, e->d_sort); dfree(e, e->d_sent); dfree(e, e->d_succ);
    cudaStreamSynchronize(e->stream);
  }
  if (e->p2p_slot >= 0) {                              // exchange buffers and peer mappings are parked, not released
    if (e->stream) cudaStreamSynchronize(e->stream);
    g_p2p_parked[e->p2p_slot].seq = e->p2p_seq;
    g_p2p_parked[e->p2p_slot].in_use = false;
  }
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int ensure_scratch(tlag_engine* e, uint64_t words, uint64_t flags) {
  if (words > e->scratch_words) {
    dfree(e, e->d_scratch); e->d_scratch = nullptr; e->scratch_words = 0;
    CK(dmalloc(e, &e->d_scratch, words * 4));
    e->scratch_words = words;
  }
  if (flags > e->flags_cap) {
    dfree(e, e->d_flags); e->d_flags = nullptr; e->flags_cap = 0;
    CK(dmalloc(e, &e->d_flags, flags));
    e->flags_cap = flags;
  }
  return TLAG_OK;
}

static int launch_probe(tlag_engine* e, const uint32_t* d_states, uint64_t n, uint8_t* d_is_new) {
  const int W = (int)e->m.words_per_state;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (n == 0) return TLAG_OK;
  const bool a16 = ((uintptr_t)d_states % 16) == 0, a8 = ((uintptr_t)d_states % 8) == 0;
  // TLAG_K1_TMA=1 selects the TMA-ring form.  Measured (profiles/r2_s3_session.log, 2^27 candidates, W = 20): 7.30 ms
  // against 6.10 ms for the staged form -- the ring's shared memory (2 x 40 KB per CTA) halves the resident threads, and
  // K1 is bound by the latency of the random table probes, not by the row copies the TMA engine takes over.
  static const bool use_tma = getenv("TLAG_K1_TMA") != nullptr;
  if (a16 && W % 4 == 0 && ((W / 4) & 1) && use_tma) {
    const size_t smem = (size_t)TLAG_TMA_STAGES * 256 * TLAG_TMA_ROWS * (size_t)W * 4;
    CK(cudaFuncSetAttribute(k_probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe_tma, 256, smem) != cudaSuccess || occ < 1) occ = 1;
    const uint64_t tiles = (n + 256 * TLAG_TMA_ROWS - 1) / (256 * TLAG_TMA_ROWS);
    uint64_t grid = (uint64_t)e->sm_count * (uint64_t)occ;                // persistent: one resident wave of CTAs
    if (grid > tiles) grid = tiles;
    k_probe_tma<<<(unsigned)grid, 256, smem, e->stream>>>(d_states, n, W, e->d_table, e->p.mask, d_is_new, e->d_ctr);
  } else if (a16 && W >= 2) {
    const size_t smem = (size_t)256 * TLAG_PROBE_ROWS * (size_t)(W | 1) * 4;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_probe_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned sblocks = (unsigned)((n + 256 * TLAG_PROBE_ROWS - 1) / (256 * TLAG_PROBE_ROWS));
    k_probe_staged<<<sblocks, 256, smem, e->stream>>>(d_states, n, W, e->d_table, e->p.mask, d_is_new, e->d_ctr);
  } else if (W % 4 == 0 && a16) k_probe<4><<<blocks, 256, 0, e->stream>>>(d_states, n, W, e->d_table, e->p.mask, d_is_new, e->d_ctr);
  else if (W % 2 == 0 && a8) k_probe<2><<<blocks, 256, 0, e->stream>>>(d_states, n, W, e->d_table, e->p.mask, d_is_new, e->d_ctr);
  else k_probe<1><<<blocks, 256, 0, e->stream>>>(d_states, n, W, e->d_table, e->p.mask, d_is_new, e->d_ctr);
  e->launches++;
  CK(cudaGetLastError());
  return TLAG_OK;
}

extern "C" int tlag_probe_batch_device(tlag_engine* e, uint64_t d_states, uint64_t n, uint64_t d_is_new, float* kernel_ms) {
  if (!e) return TLAG_EINVAL;
  CK(cudaEventRecord(e->ev0, e->stream));
  int r = launch_probe(e, (const uint32_t*)(uintptr_t)d_states, n, (uint8_t*)(uintptr_t)d_is_new);
  if (r) return r;
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaEventSynchronize(e->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  if (kernel_ms) *kernel_ms = ms;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (hc.table_full) { e->err = "seen-set table is full"; return TLAG_ENOMEM; }
  return TLAG_OK;
}

extern "C" int tlag_probe_batch(tlag_engine* e, const uint32_t* states, uint64_t n, uint8_t* is_new) {
  if (!e || (!states && n) || (!is_new && n)) return TLAG_EINVAL;
  if (n == 0) return TLAG_OK;
  const uint64_t W = e->m.words_per_state;
  int r = ensure_scratch(e, n * W, n);
  if (r) return r;
  CK(cudaMemcpyAsync(e->d_scratch, states, n * W * 4, cudaMemcpyHostToDevice, e->stream));
  r = launch_probe(e, e->d_scratch, n, e->d_flags);
  if (r) return r;
  CK(cudaMemcpyAsync(is_new, e->d_flags, n, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (hc.table_full) { e->err = "seen-set table is full"; return TLAG_ENOMEM; }
  return TLAG_OK;
}

extern "C" int tlag_reset_table(tlag_engine* e) {
  if (!e) return TLAG_EINVAL;
  CK(cudaMemsetAsync(e->d_table, 0, (1ULL << e->table_log2) * 8, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TLAG_OK;
}

extern "C" int tlag_seed(tlag_engine* e, const uint32_t* states, uint64_t n) {
  if (!e || (!states && n)) return TLAG_EINVAL;
  if (e->level != 0) { e->err = "tlag_seed after the search started"; return TLAG_ESTATE; }
  const uint64_t W = e->m.words_per_state;
  int r = grow_store_if_needed(e, e->hi + n + 1024);
  if (r) return r;
  r = grow_table_if_needed(e, e->hi + n);
  if (r) return r;
  if (n) {
    // records: W words + parent(-1) + meta(action -1)
    std::vector<uint32_t> rec(n * (W + 2));
    for (uint64_t i = 0; i < n; ++i) {
      memcpy(&rec[i * (W + 2)], states + i * W, W * 4);
      rec[i * (W + 2) + W] = 0xFFFFFFFFu;
      rec[i * (W + 2) + W + 1] = 0xFFFFFF00u;
    }
    r = ensure_scratch(e, n * (W + 2), 0);
    if (r) return r;
    CK(cudaMemcpyAsync(e->d_scratch, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice, e->stream));
    if (e->m.flags & TLAG_F_EXACT) {
      // store order = the caller's order (TLC enumerates the initial states in a fixed order): one warp at a time
      for (uint64_t o = 0; o < n; o += 32) {
        const uint64_t c = n - o < 32 ? n - o : 32;
        k_insert_records<<<1, 32, 0, e->stream>>>(e->p, e->d_scratch + o * (W + 2), c);
      }
    } else {
      k_insert_records<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->p, e->d_scratch, n);
    }
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
  }
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  e->generated += n;
  e->hi = hc.n_states;
  e->init_states = hc.n_states;
  if (n && !e->restarting) e->h_init.insert(e->h_init.end(), states, states + n * W);
  return TLAG_OK;
}

// Forget everything discovered and start again from the retained initial states (bench loops).
extern "C" int tlag_restart(tlag_engine* e) {
  if (!e) return TLAG_EINVAL;
  Counters hc;
  memset(&hc, 0, sizeof(hc));
  hc.viol_inv = hc.viol_assert = hc.viol_trap = hc.viol_deadlock = ~0ULL;
  CK(cudaMemcpyAsync(e->d_ctr, &hc, sizeof(hc), cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemsetAsync(e->d_table, 0, (1ULL << e->table_log2) * 8, e->stream));
  if (e->d_sent) CK(cudaMemsetAsync(e->d_sent, 0, (e->p.sent_mask + 1) * 8, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->lo = e->hi = 0; e->level = 0; e->init_states = 0; e->generated = 0; e->depth = 0;
  e->verdict = TLAG_V_RUNNING; e->detail = e->detail2 = 0; e->viol_idx = 0; e->dev_seconds = 0;
  e->kg_verdict = 0;
  e->restarting = true;
  const uint64_t W = e->m.words_per_state;
  int r = tlag_seed(e, e->h_init.data(), e->h_init.size() / W);
  e->restarting = false;
  return r;
}

static void fill_result(tlag_engine* e, tlag_result* out, uint64_t n_states) {
  memset(out, 0, sizeof(*out));
  out->verdict = e->verdict;
  out->detail = e->detail;
  out->detail2 = e->detail2;
  out->state_idx = e->viol_idx;
  out->generated = e->generated;
  out->distinct = n_states;
  out->queue_left = (e->verdict == TLAG_V_OK) ? 0 : (n_states - e->hi) + 0;
  if ((e->m.flags & TLAG_F_EXACT) && (e->verdict == TLAG_V_ASSERT || e->verdict == TLAG_V_DEADLOCK) && n_states > e->viol_idx)
    out->queue_left = n_states - e->viol_idx - 1;      // discovered, not yet dequeued when the worker stopped
  out->depth = e->depth;
  out->init_states = e->init_states;
  const double n = (double)n_states, g = (double)e->generated;
  out->fp_collision_estimate = n * (g > n ? g - n : 0.0) / 18446744073709551616.0;
  out->device_seconds = e->dev_seconds;
}

// decode violation counters after a wave; sets e->verdict if something was found
static void collect_violations(tlag_engine* e, const Counters& hc) {
  unsigned long long best = ~0ULL; int kind = 0;
  if (hc.viol_trap != ~0ULL && (hc.viol_trap >> 20) < best) { best = hc.viol_trap >> 20; kind = TLAG_V_EVAL_ERROR; }
  if (hc.viol_assert != ~0ULL && (hc.viol_assert >> 20) < best) { best = hc.viol_assert >> 20; kind = TLAG_V_ASSERT; }
  if (hc.viol_inv != ~0ULL && (hc.viol_inv >> 20) < best) { best = hc.viol_inv >> 20; kind = TLAG_V_INVARIANT; }
  if (hc.viol_deadlock != ~0ULL && (hc.viol_deadlock >> 20) < best) { best = hc.viol_deadlock >> 20; kind = TLAG_V_DEADLOCK; }
  if (!kind) return;
  e->verdict = kind;
  e->viol_idx = best;
  if (kind == TLAG_V_EVAL_ERROR) { e->detail = (int)((hc.viol_trap >> 16) & 15); e->detail2 = (int)(hc.viol_trap & 0xFFFF); }
  else if (kind == TLAG_V_ASSERT) e->detail = (int)(hc.viol_assert & 0xFFFFF);
  else if (kind == TLAG_V_INVARIANT) e->detail = (int)(hc.viol_inv & 0xFFFFF);
}

// Undo a wave that ran out of store or table space: the states it appended are dropped (n_states back to `hi`), the
// seen-set is rebuilt from the store (it holds fingerprints of the dropped states), counters and violation slots are
// cleared -- the level is then expanded again from scratch with more room.
static int rollback_wave(tlag_engine* e, uint64_t hi, uint64_t want_states) {
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  hc.n_states = hi; hc.generated = 0; hc.work = 0; hc.store_overflow = 0; hc.table_full = 0;
  hc.viol_inv = hc.viol_assert = hc.viol_trap = hc.viol_deadlock = ~0ULL;
  CK(cudaMemcpy(e->d_ctr, &hc, sizeof(hc), cudaMemcpyHostToDevice));
  int r = grow_store_if_needed(e, want_states + 4096);
  if (r) return r;
  unsigned log2 = e->table_log2;
  while ((1ULL << log2) < want_states * 2 && log2 < 40) ++log2;
  r = alloc_table(e, log2);
  if (r) return r;
  if (hi) {
    k_rehash<<<(unsigned)((hi + 255) / 256), 256, 0, e->stream>>>(e->p, hi);
    e->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(e->stream));
  return TLAG_OK;
}

extern "C" int tlag_step(tlag_engine* e, tlag_wave_stats* out) {
  if (!e) return TLAG_EINVAL;
  if (out) memset(out, 0, sizeof(*out));
  if (e->level == 0) { e->level = 1; e->lo = 0; e->depth = e->hi > 0 ? 1 : 0; }
  if (e->verdict != TLAG_V_RUNNING) { if (out) out->verdict = e->verdict; return TLAG_OK; }
  const uint64_t lo = e->lo, hi = e->hi;
  if (lo >= hi) {
    e->verdict = e->kg_verdict ? e->kg_verdict : TLAG_V_OK;
    if (e->kg_verdict) { e->detail = e->kg_detail; e->detail2 = e->kg_detail2; e->viol_idx = e->kg_idx; }
    if (out) out->verdict = e->verdict;
    return TLAG_OK;
  }
  int r = grow_store_if_needed(e, hi + (uint64_t)((double)(hi - lo) * e->growth_hint) + 4096);
  if (r) return r;
  r = grow_table_if_needed(e, hi + (uint64_t)((double)(hi - lo) * (e->growth_hint > 4.0 ? e->growth_hint * 0.5 : 2.0)));
  if (r) return r;
  Counters hc;
  float ms = 0;
  for (int attempt = 0;; ++attempt) {
    CK(cudaMemsetAsync(&e->d_ctr->work, 0, 8, e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    cudaError_t ce = launch_wave<0>(e, lo, hi);
    if (ce != cudaSuccess) { e->err = std::string("wave launch: ") + cudaGetErrorString(ce); return TLAG_ECUDA; }
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaMemcpyAsync(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    float ms1 = 0;
    CK(cudaEventElapsedTime(&ms1, e->ev0, e->ev1));
    ms += ms1;
    if (!hc.store_overflow && !hc.table_full) break;
    // The head-room heuristics were too tight for this level (its discovered / expanded ratio jumped): n_states kept
    // counting past the capacity, so it says how much room the level needs.  Redo the level with that much, twice over.
    const uint64_t need = hc.n_states > hi ? hc.n_states : hi;
    const uint64_t want = hi + (need - hi) * 2 + 4096;
    if (attempt >= 3 || (e->m.max_states && need > e->m.max_states) || need > 0xFFFFFFFEull) {
      e->err = hc.store_overflow ? "state store overflow: raise max_states (or the model needs more than 2^32 - 2 states per GPU)"
                                 : "seen-set table full";
      return TLAG_ENOMEM;
    }
    r = rollback_wave(e, hi, want);
    if (r) return r;
    if (e->cap_states < need + 1024) { e->err = "state store overflow: not enough device memory for this level"; return TLAG_ENOMEM; }
  }
  e->dev_seconds += ms * 1e-3;
  const uint64_t gen_wave = hc.generated;
  e->generated += gen_wave;
  CK(cudaMemsetAsync(&e->d_ctr->generated, 0, 8, e->stream));
  const uint64_t n_states = hc.n_states;
  collect_violations(e, hc);
  if (e->verdict != TLAG_V_RUNNING && (e->m.flags & TLAG_F_KEEP_GOING)) {
    // keep exploring: remember the first violation (lowest level, lowest index in it), clear the device slots
    if (!e->kg_verdict) { e->kg_verdict = e->verdict; e->kg_detail = e->detail; e->kg_detail2 = e->detail2; e->kg_idx = e->viol_idx; }
    e->verdict = TLAG_V_RUNNING;
    const unsigned long long ones[4] = {~0ULL, ~0ULL, ~0ULL, ~0ULL};
    CK(cudaMemcpyAsync(&e->d_ctr->viol_inv, ones, sizeof(ones), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  if (out) {
    out->level = e->level; out->expanded = hi - lo; out->generated = gen_wave;
    out->discovered = n_states - hi; out->distinct_total = n_states; out->generated_total = e->generated;
    out->kernel_ms = ms;
  }
  if (n_states > hi) e->depth = e->level + 1;
  if (n_states > hi && e->verdict == TLAG_V_RUNNING) {
    int rs = cluster_slice(e, hi, n_states - hi);
    if (rs) return rs;
  }
  {  // head-room for the next level: twice the observed discovered/expanded ratio, at least 4x
    const double ratio = (double)(n_states - hi) / (double)(hi - lo);
    e->growth_hint = ratio * 2.0 > 4.0 ? ratio * 2.0 : 4.0;
  }
  e->lo = hi; e->hi = n_states; e->level += 1;
  if (e->verdict == TLAG_V_RUNNING && e->lo >= e->hi) {
    e->verdict = e->kg_verdict ? e->kg_verdict : TLAG_V_OK;
    if (e->kg_verdict) { e->detail = e->kg_detail; e->detail2 = e->kg_detail2; e->viol_idx = e->kg_idx; }
  }
  if (out) out->verdict = e->verdict;
  return TLAG_OK;
}

extern "C" int tlag_result_now(tlag_engine* e, tlag_result* out) {
  if (!e || !out) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  fill_result(e, out, hc.n_states);
  return TLAG_OK;
}

extern "C" int tlag_run(tlag_engine* e, tlag_result* out) {
  if (!e || !out) return TLAG_EINVAL;
  tlag_wave_stats ws;
  for (;;) {
    int r = tlag_step(e, &ws);
    if (r) return r;
    if (ws.verdict != TLAG_V_RUNNING) break;
  }
  return tlag_result_now(e, out);
}

extern "C" int tlag_read_states(tlag_engine* e, uint64_t first, uint64_t n, uint32_t* states_out) {
  if (!e || (!states_out && n)) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (first + n > hc.n_states) { e->err = "tlag_read_states: range beyond the state store"; return TLAG_EINVAL; }
  const uint64_t W = e->m.words_per_state;
  CK(cudaMemcpy(states_out, e->d_states + first * W, n * W * 4, cudaMemcpyDeviceToHost));
  return TLAG_OK;
}

extern "C" int tlag_digest(tlag_engine* e, uint64_t* xor_out, uint64_t* sum_out) {
  if (!e || !xor_out || !sum_out) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (!e->d_dig) CK(cudaMalloc(&e->d_dig, 16));
  unsigned long long* d2 = e->d_dig;
  CK(cudaMemsetAsync(d2, 0, 16, e->stream));
  if (hc.n_states) {
    k_digest<<<e->sm_count * 8, 256, 0, e->stream>>>(e->p, hc.n_states, d2);
    e->launches++;
    CK(cudaGetLastError());
  }
  unsigned long long h2[2];
  CK(cudaMemcpyAsync(h2, d2, 16, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *xor_out = h2[0]; *sum_out = h2[1];
  return TLAG_OK;
}

extern "C" int tlag_trace(tlag_engine* e, uint64_t state_idx, uint32_t* states_out, int32_t* actions_out, uint32_t* len_inout) {
  if (!e || !len_inout) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (state_idx >= hc.n_states) { e->err = "tlag_trace: no such state"; return TLAG_EINVAL; }
  const uint64_t W = e->m.words_per_state;
  std::vector<uint64_t> chain;
  std::vector<int32_t> acts;
  uint64_t cur = state_idx;
  for (;;) {
    uint32_t par = 0, meta = 0;
    CK(cudaMemcpyAsync(&par, e->d_parent + cur, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&meta, e->d_meta + cur, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    chain.push_back(cur);
    acts.push_back(par == 0xFFFFFFFFu ? -1 : (int32_t)(meta >> 8));
    if (par == 0xFFFFFFFFu) break;
    cur = par;
    if (chain.size() > (1u << 24)) { e->err = "tlag_trace: parent chain too long"; return TLAG_ESTATE; }
  }
  const uint32_t len = (uint32_t)chain.size();
  if (len > *len_inout) { *len_inout = len; e->err = "tlag_trace: buffer too small"; return TLAG_EINVAL; }
  for (uint32_t i = 0; i < len; ++i) {
    const uint64_t s = chain[len - 1 - i];
    if (states_out) CK(cudaMemcpy(states_out + (uint64_t)i * W, e->d_states + s * W, W * 4, cudaMemcpyDeviceToHost));
    if (actions_out) actions_out[i] = acts[len - 1 - i];
  }
  *len_inout = len;
  return TLAG_OK;
}

// One hop of a counterexample chain: the state at idx with its parent index and meta word (action id << 8 | rank that
// expanded the parent).  With several ranks the parent lives in THAT rank's store: the host follows the chain across
// ranks (tla_rust_b200/dist.py: counterexample).
extern "C" int tlag_read_link(tlag_engine* e, uint64_t idx, uint32_t* state_out, uint32_t* parent_out, uint32_t* meta_out) {
  if (!e || !parent_out || !meta_out) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (idx >= hc.n_states) { e->err = "tlag_read_link: no such state"; return TLAG_EINVAL; }
  const uint64_t W = e->m.words_per_state;
  if (state_out) CK(cudaMemcpyAsync(state_out, e->d_states + idx * W, W * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(parent_out, e->d_parent + idx, 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(meta_out, e->d_meta + idx, 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TLAG_OK;
}

// ---- multi-GPU building blocks -----------------------------------------------------
extern "C" int tlag_frontier(tlag_engine* e, uint64_t* first_idx, uint64_t* count) {
  if (!e) return TLAG_EINVAL;
  if (e->level == 0) { e->level = 1; e->lo = 0; e->depth = e->hi > 0 ? 1 : 0; }
  if (first_idx) *first_idx = e->lo;
  if (count) *count = e->hi - e->lo;
  return TLAG_OK;
}

// Which shard of the partitioned state space this engine holds (host-driven exchange: tlag_expand_route /
// tlag_insert_records; tlag_p2p_init sets it for the peer-memory path).  The rank goes into the meta word of every
// state this engine expands a parent of, and successors the rank owns itself are inserted in place.
// How many trailing packed words the ownership hash covers (every rank must use the same value; before any state is
// routed).  2 keeps whole clusters of like states on one rank; W balances models whose tail words carry no entropy.
extern "C" int tlag_set_owner_words(tlag_engine* e, uint32_t k) {
  if (!e || k < 2) return TLAG_EINVAL;
  e->p.owner_words = (int)(k > e->m.words_per_state ? e->m.words_per_state : k);
  return TLAG_OK;
}

extern "C" int tlag_set_rank(tlag_engine* e, uint32_t n_ranks, uint32_t rank) {
  if (!e || n_ranks == 0 || n_ranks > 16 || rank >= n_ranks) return TLAG_EINVAL;
  e->p.n_ranks = (int)n_ranks; e->p.rank = (int)rank;
  return TLAG_OK;
}

extern "C" int tlag_expand_route(tlag_engine* e, uint32_t n_ranks, uint64_t first, uint64_t count, uint64_t d_send,
                                 uint64_t cap_records, uint64_t* counts, tlag_wave_stats* out) {
  if (!e || !counts || n_ranks == 0 || n_ranks > 16) return TLAG_EINVAL;
  if (out) memset(out, 0, sizeof(*out));
  if (e->level == 0) { e->level = 1; e->lo = 0; e->depth = 1; }
  uint64_t lo = e->lo + first, hi = lo + count;
  if (lo > e->hi) lo = e->hi;
  if (hi > e->hi) hi = e->hi;
  if ((uint32_t)e->p.n_ranks != n_ranks) {
    if (n_ranks > 1) { e->err = "tlag_expand_route: call tlag_set_rank(n_ranks, rank) first"; return TLAG_ESTATE; }
    e->p.n_ranks = 1; e->p.rank = 0;
  }
  if (!e->d_sent && n_ranks > 1 && getenv("TLAG_NO_SENT_CACHE") == nullptr) {
    const unsigned lg = 26;                                    // 2^26 x 8 B = 512 MB
    if (dmalloc(e, &e->d_sent, (1ULL << lg) * 8) == cudaSuccess) {
      CK(cudaMemsetAsync(e->d_sent, 0, (1ULL << lg) * 8, e->stream));
      e->p.sent_cache = e->d_sent; e->p.sent_mask = (1ULL << lg) - 1;
    } else { cudaGetLastError(); e->d_sent = nullptr; }
  }
  e->p.send = (uint32_t*)(uintptr_t)d_send;
  e->p.region_cap = cap_records / n_ranks;
  CK(cudaMemsetAsync(&e->d_ctr->work, 0, 8, e->stream));
  CK(cudaMemsetAsync(e->d_ctr->send_count, 0, sizeof(unsigned long long) * 16, e->stream));
  CK(cudaMemsetAsync(&e->d_ctr->route_overflow, 0, 8, e->stream));   // a failed attempt may be retried with
  CK(cudaMemsetAsync(&e->d_ctr->generated, 0, 8, e->stream));        // larger regions: start from clean counters
  CK(cudaEventRecord(e->ev0, e->stream));
  if (hi > lo) {
    cudaError_t ce = launch_wave<1>(e, lo, hi);
    if (ce != cudaSuccess) { e->err = std::string("k_wave<route> launch: ") + cudaGetErrorString(ce); return TLAG_ECUDA; }
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  Counters hc;
  CK(cudaMemcpyAsync(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->dev_seconds += ms * 1e-3;
  if (hc.route_overflow) {
    // The caller re-runs this chunk with larger regions and throws this attempt's buffer away -- including the records
    // that did fit, whose fingerprints are now in the routed-fingerprint cache.  Forget them all (the cache is only a
    // filter: an empty one costs redundant records, a stale one loses states).
    if (e->d_sent) CK(cudaMemsetAsync(e->d_sent, 0, (e->p.sent_mask + 1) * 8, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->err = "send buffer overflow in tlag_expand_route";
    return TLAG_EOVERFLOW;
  }
  e->generated += hc.generated;
  CK(cudaMemsetAsync(&e->d_ctr->generated, 0, 8, e->stream));
  for (uint32_t r = 0; r < n_ranks; ++r) counts[r] = hc.send_count[r];
  collect_violations(e, hc);
  if (out) {
    out->level = e->level; out->expanded = hi - lo; out->generated = hc.generated;
    out->distinct_total = hc.n_states; out->generated_total = e->generated; out->kernel_ms = ms;
    out->verdict = e->verdict;
  }
  return TLAG_OK;
}

extern "C" int tlag_insert_records(tlag_engine* e, uint64_t d_recv, uint64_t n_records, uint32_t, uint64_t* n_new) {
  if (!e) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpyAsync(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const uint64_t before = hc.n_states;
  int r = grow_store_if_needed(e, before + n_records + 1024);
  if (r) return r;
  r = grow_table_if_needed(e, before + n_records);
  if (r) return r;
  if (n_records) {
    CK(cudaEventRecord(e->ev0, e->stream));
    k_insert_records<<<(unsigned)((n_records + 255) / 256), 256, 0, e->stream>>>(e->p, (const uint32_t*)(uintptr_t)d_recv, n_records);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaEventSynchronize(e->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->dev_seconds += ms * 1e-3;
  }
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  if (hc.store_overflow) { e->err = "state store overflow: raise max_states"; return TLAG_ENOMEM; }
  if (hc.table_full) { e->err = "seen-set table full"; return TLAG_ENOMEM; }
  if (n_new) *n_new = hc.n_states - before;
  return TLAG_OK;
}


// ---- peer-memory exchange: set-up and one-call-per-level driver ------------------------------------------------------
static size_t p2p_meta_bytes() { return 16 * sizeof(P2PMeta); }

extern "C" int tlag_p2p_init(tlag_engine* e, uint32_t n_ranks, uint32_t rank, uint64_t cap_records, uint8_t* handle_out64) {
  if (!e || !handle_out64 || n_ranks == 0 || n_ranks > 16 || rank >= n_ranks || cap_records == 0) return TLAG_EINVAL;
  if (e->d_p2p) { e->err = "tlag_p2p_init called twice"; return TLAG_ESTATE; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  const uint64_t rw = e->m.words_per_state + 2;
  cap_records = (cap_records + 3) & ~3ULL;                       // regions stay 16-byte aligned
  for (size_t i = 0; i < g_p2p_parked.size(); ++i) {
    P2PParked& c = g_p2p_parked[i];
    if (c.in_use || c.device != e->m.device || c.n_ranks != (int)n_ranks || c.rank != (int)rank || c.rec_words != (int)rw ||
        c.cap != cap_records) continue;
    c.in_use = true;
    e->p2p_slot = (int)i;
    e->d_p2p = c.d_p2p; e->p2p_bytes = c.bytes; e->d_send_own = c.d_send; e->d_tickets = c.d_tickets; e->p2p_seq = c.seq;
    memset(&e->q, 0, sizeof(e->q));
    e->q.n_ranks = (int)n_ranks; e->q.rank = (int)rank; e->q.rec_words = (int)rw;
    e->q.inbox_cap = cap_records; e->q.region_cap = cap_records;
    e->q.send = e->d_send_own;
    e->q.meta = (P2PMeta*)e->d_p2p;
    e->q.inbox = (uint32_t*)((uint8_t*)e->d_p2p + p2p_meta_bytes());
    e->q.tickets = e->d_tickets;
    e->peer_base[rank] = e->d_p2p;
    e->q.peer_meta[rank] = e->q.meta; e->q.peer_inbox[rank] = e->q.inbox;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->d_p2p));
    memcpy(handle_out64, &h, 64);
    e->p.n_ranks = (int)n_ranks; e->p.rank = (int)rank;
    return TLAG_OK;
  }
  const uint64_t inbox_bytes = 2ULL * n_ranks * cap_records * rw * 4;
  e->p2p_bytes = p2p_meta_bytes() + inbox_bytes;
  CK(cudaMalloc(&e->d_p2p, e->p2p_bytes));
  CK(cudaMemset(e->d_p2p, 0, p2p_meta_bytes()));
  CK(cudaMalloc(&e->d_send_own, (uint64_t)n_ranks * cap_records * rw * 4));
  CK(cudaMalloc(&e->d_tickets, 32 * sizeof(unsigned int)));
  CK(cudaMemset(e->d_tickets, 0, 32 * sizeof(unsigned int)));
  {
    P2PParked c;
    memset(&c, 0, sizeof(c));
    c.device = e->m.device; c.n_ranks = (int)n_ranks; c.rank = (int)rank; c.rec_words = (int)rw; c.cap = cap_records;
    c.d_p2p = e->d_p2p; c.bytes = e->p2p_bytes; c.d_send = e->d_send_own; c.d_tickets = e->d_tickets; c.in_use = true;
    g_p2p_parked.push_back(c);
    e->p2p_slot = (int)g_p2p_parked.size() - 1;
  }
  memset(&e->q, 0, sizeof(e->q));
  e->q.n_ranks = (int)n_ranks; e->q.rank = (int)rank; e->q.rec_words = (int)rw;
  e->q.inbox_cap = cap_records; e->q.region_cap = cap_records;
  e->q.send = e->d_send_own;
  e->q.meta = (P2PMeta*)e->d_p2p;
  e->q.inbox = (uint32_t*)((uint8_t*)e->d_p2p + p2p_meta_bytes());
  e->q.tickets = e->d_tickets;
  e->peer_base[rank] = e->d_p2p;
  e->q.peer_meta[rank] = e->q.meta; e->q.peer_inbox[rank] = e->q.inbox;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->d_p2p));
  memcpy(handle_out64, &h, 64);
  e->p.n_ranks = (int)n_ranks; e->p.rank = (int)rank;
  return TLAG_OK;
}

extern "C" int tlag_p2p_attach(tlag_engine* e, uint32_t peer, const uint8_t* handle64) {
  if (!e || !handle64 || !e->d_p2p || peer >= (uint32_t)e->q.n_ranks) return TLAG_EINVAL;
  if ((int)peer == e->q.rank) return TLAG_OK;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  P2PParked* c = e->p2p_slot >= 0 ? &g_p2p_parked[e->p2p_slot] : nullptr;
  if (c && c->peer_ok[peer] && memcmp(c->peer_handle[peer], handle64, 64) == 0) {
    base = c->peer_base[peer];                                  // the peer parked and re-adopted the same inbox
  } else {
    if (c && c->peer_ok[peer]) { cudaIpcCloseMemHandle(c->peer_base[peer]); c->peer_ok[peer] = false; }
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    if (c) { c->peer_base[peer] = base; memcpy(c->peer_handle[peer], handle64, 64); c->peer_ok[peer] = true; }
  }
  e->peer_base[peer] = base;
  e->q.peer_meta[peer] = (P2PMeta*)base;
  e->q.peer_inbox[peer] = (uint32_t*)((uint8_t*)base + p2p_meta_bytes());
  int ready = 1;
  for (int r = 0; r < e->q.n_ranks; ++r) if (!e->peer_base[r]) ready = 0;
  e->p2p_ready = ready != 0;
  return TLAG_OK;
}

// One BFS level on this rank: the frontier shard in n_chunks chunks of chunk_states states; per chunk the expand
// kernels (successors bucketed by owner), k_push (NVLink stores into the owners' inboxes) and k_insert_inbox (this
// rank's share of the next frontier).  Every rank must call it with the same n_chunks (ranks with a shorter frontier
// run empty chunks: the inbox protocol counts chunks).  One host synchronisation, at the end.
extern "C" int tlag_p2p_level(tlag_engine* e, uint64_t n_chunks, uint64_t chunk_states, uint64_t expect_inbound,
                              tlag_wave_stats* out) {
  if (!e || chunk_states == 0) return TLAG_EINVAL;
  if (!e->p2p_ready && e->q.n_ranks > 1) { e->err = "tlag_p2p_level before every peer was attached"; return TLAG_ESTATE; }
  if (out) memset(out, 0, sizeof(*out));
  if (e->level == 0) { e->level = 1; e->lo = 0; e->depth = e->hi > 0 ? 1 : 0; }
  const uint32_t n_ranks = (uint32_t)e->q.n_ranks;
  // room for what may arrive this level (the caller passes the all-reduced estimate; at least the heuristic)
  uint64_t expect = (uint64_t)((double)(e->hi - e->lo) * e->growth_hint) + 4096;
  if (expect_inbound > expect) expect = expect_inbound;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  const uint64_t before = hc.n_states;
  e->p2p_level_start = before; e->p2p_level_generated = 0;
  int r = grow_store_if_needed(e, before + expect);
  if (r) return r;
  r = grow_table_if_needed(e, before + expect);
  if (r) return r;
  e->p.n_ranks = (int)n_ranks; e->p.rank = e->q.rank;
  if (!e->d_sent && n_ranks > 1 && getenv("TLAG_NO_SENT_CACHE") == nullptr) {
    const unsigned lg = 26;                                    // 2^26 x 8 B = 512 MB
    if (dmalloc(e, &e->d_sent, (1ULL << lg) * 8) == cudaSuccess) {
      CK(cudaMemsetAsync(e->d_sent, 0, (1ULL << lg) * 8, e->stream));
      e->p.sent_cache = e->d_sent; e->p.sent_mask = (1ULL << lg) - 1;
    } else { cudaGetLastError(); e->d_sent = nullptr; }
  }
  e->p.send = e->d_send_own;
  e->p.region_cap = e->q.region_cap;
  CK(cudaMemsetAsync(&e->d_ctr->route_overflow, 0, 8, e->stream));
  CK(cudaMemsetAsync(&e->d_ctr->generated, 0, 8, e->stream));
  CK(cudaEventRecord(e->ev0, e->stream));
  const unsigned push_blocks = 64, ins_blocks = (unsigned)e->sm_count * 4;
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const unsigned long long seq = ++e->p2p_seq;
    uint64_t lo = e->lo + c * chunk_states, hi = lo + chunk_states;
    if (lo > e->hi) lo = e->hi;
    if (hi > e->hi) hi = e->hi;
    CK(cudaMemsetAsync(&e->d_ctr->work, 0, 8, e->stream));
    CK(cudaMemsetAsync(e->d_ctr->send_count, 0, sizeof(unsigned long long) * 16, e->stream));
    if (hi > lo) {
      cudaError_t ce = launch_wave<1>(e, lo, hi);
      if (ce != cudaSuccess) { e->err = std::string("route wave launch: ") + cudaGetErrorString(ce); return TLAG_ECUDA; }
    }
    k_push<<<dim3(push_blocks, n_ranks), 256, 0, e->stream>>>(e->q, e->d_ctr, seq);
    k_insert_inbox<<<ins_blocks, 256, 0, e->stream>>>(e->p, e->q, seq);
    e->launches += 2;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaMemcpyAsync(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->dev_seconds += ms * 1e-3;
  if (hc.route_overflow) { e->err = "send region overflow in tlag_p2p_level: use smaller chunks or larger regions"; return TLAG_EOVERFLOW; }
  if (hc.store_overflow) { e->err = "state store overflow in tlag_p2p_level"; return TLAG_ENOMEM; }
  if (hc.table_full) { e->err = "seen-set table full in tlag_p2p_level"; return TLAG_ENOMEM; }
  e->generated += hc.generated;
  e->p2p_level_generated = hc.generated;
  CK(cudaMemsetAsync(&e->d_ctr->generated, 0, 8, e->stream));
  collect_violations(e, hc);
  if (out) {
    out->level = e->level; out->expanded = e->hi - e->lo; out->generated = hc.generated;
    out->discovered = hc.n_states - before; out->distinct_total = hc.n_states; out->generated_total = e->generated;
    out->kernel_ms = ms; out->verdict = e->verdict;
  }
  {
    const double ratio = (double)(hc.n_states - before) / (double)((e->hi - e->lo) ? (e->hi - e->lo) : 1);
    e->growth_hint = ratio * 2.0 > 4.0 ? ratio * 2.0 : 4.0;
  }
  return TLAG_OK;
}

// Undo the level tlag_p2p_level just ran on this rank (some rank's send regions overflowed: every rank rolls back and
// the host re-runs the level with smaller chunks).  The states the level appended are dropped, the seen-set is rebuilt
// from the store, the routed-fingerprint cache is emptied; the inbox chunk numbering simply continues.
extern "C" int tlag_p2p_rollback(tlag_engine* e) {
  if (!e || !e->d_p2p) return TLAG_EINVAL;
  CK(cudaStreamSynchronize(e->stream));
  int r = rollback_wave(e, e->p2p_level_start, e->p2p_level_start + 4096);
  if (r) return r;
  if (e->d_sent) CK(cudaMemsetAsync(e->d_sent, 0, (e->p.sent_mask + 1) * 8, e->stream));
  CK(cudaMemsetAsync(&e->d_ctr->route_overflow, 0, 8, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->generated -= e->p2p_level_generated;
  e->p2p_level_generated = 0;
  e->verdict = TLAG_V_RUNNING; e->detail = e->detail2 = 0; e->viol_idx = 0;
  return TLAG_OK;
}

extern "C" int tlag_advance_level(tlag_engine* e, tlag_wave_stats* out) {
  if (!e) return TLAG_EINVAL;
  Counters hc;
  CK(cudaMemcpy(&hc, e->d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  const uint64_t n_states = hc.n_states;
  if (out) { memset(out, 0, sizeof(*out)); out->level = e->level; out->discovered = n_states - e->hi; out->distinct_total = n_states; out->generated_total = e->generated; out->verdict = e->verdict; }
  if (n_states > e->hi) {
    e->depth = e->level + 1;
    int rs = cluster_slice(e, e->hi, n_states - e->hi);
    if (rs) return rs;
  }
  e->lo = e->hi; e->hi = n_states; e->level += 1;
  return TLAG_OK;
}
