// tlag_dev.cuh -- device-side definitions shared by the engine (tlag_engine.cu) and by the translation units of a
// sliced native build (tla_rust_b200/compile/sliced.py writes one .cu per group of slices; they are compiled in
// parallel and linked into one library): counters, kernel parameters, the seen-set probe, and the slice runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tlag.h"
#include "tlag_vm.h"

struct Counters {
  unsigned long long n_states;      // tail of the state store
  unsigned long long generated;
  unsigned long long work;          // chunk dispenser
  unsigned long long viol_inv;      // min key: idx<<20 | detail
  unsigned long long viol_assert;
  unsigned long long viol_trap;     // idx<<20 | code<<16 | line
  unsigned long long viol_deadlock;
  unsigned long long store_overflow;
  unsigned long long table_full;
  unsigned long long route_overflow;
  unsigned long long send_count[16];
  unsigned long long stop;          // TLAG_F_EXACT: an error was found, no further state is dequeued
};

struct DevParams {
  const uint64_t* code; uint32_t code_len; int code_in_smem;
  const int32_t* cpool;
  const tlag_slot* layout; int n_slots;
  uint32_t entry_inv, entry_next;
  uint32_t n_off, p_off;
  int W;
  uint32_t* states; uint32_t* parent; uint32_t* meta;
  unsigned long long cap_states;
  unsigned long long* table; unsigned long long mask;
  Counters* ctr;
  uint32_t flags;
  int n_inv;
  uint8_t* succ_flag;               // sliced build: one byte per frontier state, set when a slice produced a successor
  // route mode
  int route;                        // sliced build: successors go to the send regions instead of the local seen-set
  int owner_words;                  // trailing packed words the ownership hash covers (tlag_owner_k; 2 = clustering key)
  uint32_t* send; unsigned long long region_cap; int n_ranks; int rank;
  unsigned long long* sent_cache; unsigned long long sent_mask;   // direct-mapped filter of fingerprints already routed
};

// ---- peer-memory exchange (tlag_engine.cu: k_push / k_insert_inbox) ----
struct P2PMeta {                       // lives in the OWNER's memory, one per source rank
  unsigned long long count[2];         // records the source pushed into buffer (seq & 1)            [written by the source]
  unsigned long long seq[2];           // chunk number whose records are complete in that buffer      [written by the source]
  unsigned long long ack;              // as a source: last chunk of MINE that this peer has consumed [written by the peer]
  unsigned long long pad[3];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct P2PParams {
  int n_ranks, rank, rec_words;
  unsigned long long inbox_cap;        // records per (owner, source, buffer) region
  unsigned long long region_cap;       // records per destination in the local send buffer
  const uint32_t* send;                // local send regions [n_ranks][region_cap][rec_words]
  uint32_t* peer_inbox[16];            // rank r's inbox base as mapped in this process
  P2PMeta* peer_meta[16];              // rank r's meta array as mapped in this process
  uint32_t* inbox; P2PMeta* meta;      // this rank's own
  unsigned int* tickets;               // [0..16): k_push "last block per destination", [16]: k_insert_inbox
};


// ------------------------------------------------------------------ seen-set
// returns 1 if fp was inserted by this call, 0 if already present, -1 if the table is full
__device__ __forceinline__ int seen_insert(unsigned long long* table, unsigned long long mask,
                                           unsigned long long fp) {
  // read first, CAS only on an empty slot: measured 6.1 ms vs 7.9 ms for a CAS-first probe on the
  // 2^27-candidate K1 batch (atomics are throughput-limited at L2; duplicates need no atomic at all)
  unsigned long long i = fp & mask;
  for (unsigned long long probes = 0; probes <= mask; ++probes) {
    const unsigned long long cur = __ldcv(&table[i]);
    if (cur == fp) return 0;
    if (cur == 0ULL) {
      const unsigned long long old = atomicCAS(&table[i], 0ULL, fp);
      if (old == 0ULL) return 1;
      if (old == fp) return 0;
    }
    i = (i + 1) & mask;
  }
  return -1;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ void report_min(unsigned long long* slot, unsigned long long key) {
  atomicMin(slot, key);
}

// ------------------------------------------------------------------ sliced native build
// (tla_rust_b200/compile/sliced.py)  One kernel per slice of the model's program: k_sl_inv_<i> evaluates invariant i,
// k_sl_next_<j> one disjunct of Next, each over the whole frontier, one thread per state.  All warps of a launch run
// the same few KB of straight-line code (the whole-program compiled form of round 1 lost to the interpreter on
// instruction-cache misses: 354 KB of code, every warp somewhere else); divergence inside a slice is the hardware's
// (compiler-placed reconvergence) instead of a min-pc election per block; a successor is packed, fingerprinted and
// inserted where it is produced, with opportunistic warp aggregation of the tail-counter atomics.
// A translation unit that holds slices defines TLAG_SLICED_DEFS (the generated header with the model's constants) before
// including this file, then the slice functions, then TLAG_SL_KERNEL_INV(j) / TLAG_SL_KERNEL_NEXT(j) for its slices.
#ifdef TLAG_SLICED_DEFS
#include TLAG_SLICED_DEFS
struct tlag_sl_cx {
  const DevParams* p;
  unsigned long long idx;
  const uint32_t* src;              // packed words of the state being expanded (EMITD re-packs over a copy)
  unsigned long long gen;
  unsigned nsucc;
  int phase;                        // 0 = invariant slice, 1 = slice of Next
};

#ifndef TLAG_SL_BLOCK
#define TLAG_SL_BLOCK 256
#endif
#ifndef TLAG_SL_OCC
#define TLAG_SL_OCC 4               /* resident CTAs per SM the slice kernels are compiled for (register cap) */
#endif

static __device__ __noinline__ void sl_emit_words(tlag_sl_cx* cx, int aid, const uint32_t* o);
static __device__ __noinline__ void sl_emit_frame(tlag_sl_cx* cx, const int32_t* f, int aid, int dirty);
#define TLAG_SL_EMIT(aid, dirty) sl_emit_frame(cx, f, (aid), (dirty))
#define TLAG_SL_EMITW(aid, o) sl_emit_words(cx, (aid), (o))
#define TLAG_SL_GEN() do { cx->nsucc++; cx->gen++; } while (0)
#define TLAG_SL_ASSERT(id) report_min(&cx->p->ctr->viol_assert, (cx->idx << 20) | (unsigned)((id) & 0xFFFFF))
#define TLAG_SL_INVF(i) do { if (cx->phase == 0) report_min(&cx->p->ctr->viol_inv, (cx->idx << 20) | (unsigned)((i) & 0xFFFFF)); } while (0)
#define TLAG_SL_TRAP(code, line) do { report_min(&cx->p->ctr->viol_trap, (cx->idx << 20) | ((unsigned long long)((code) & 15) << 16) | (unsigned)((line) & 0xFFFF)); cx->nsucc++; } while (0)
#define TLAG_SL_SUBQ static __device__ __noinline__
// Array form with a big frame (container models): the slice is a function of its own that receives the frame as a
// pointer -- inlined, cicc tries scalar replacement of a 2 K-word array over a goto graph and takes many minutes (raft).
#if defined(TLAG_SL_SEG_NOINLINE)
#define TLAG_SL_SEGQ static __device__ __noinline__
#else
#define TLAG_SL_SEGQ static __device__ __forceinline__
#endif
// (the slice functions of this translation unit follow the #include of this header)
#define TLAG_SL_RUNTIME_PART2 1
#endif  // TLAG_SLICED_DEFS
