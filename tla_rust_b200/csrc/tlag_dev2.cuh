// tlag_dev2.cuh -- second half of the slice runtime: needs the generated constants (TLAG_SL_W ...), so it is included
// by a slice translation unit AFTER tlag_dev.cuh; defines the successor handling and the per-slice kernel macros.
#pragma once

// successor already packed: fingerprint -> seen-set (or owner's send region) -> store
static __device__ __noinline__ void sl_emit_words(tlag_sl_cx* cx, int aid, const uint32_t* o) {
  const DevParams& p = *cx->p;
  constexpr int W = TLAG_SL_W;
  cx->nsucc++; cx->gen++;
  uint32_t w[W];
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = o[i];
  const unsigned long long fp = tlag_fingerprint(w, W);
  const unsigned lane = threadIdx.x & 31;
  // route mode: a successor this rank owns itself never leaves the GPU (1/N of the records)
  const int owner = p.route ? (int)tlag_owner_k(w, W, (uint32_t)p.n_ranks, p.owner_words) : p.rank;
  if (!p.route || owner == p.rank) {
    int ins = seen_insert(p.table, p.mask, fp);
    if (ins < 0) { atomicExch(&p.ctr->table_full, 1ULL); ins = 0; }
    if (ins > 0) {
      // warp-aggregated tail allocation over whichever lanes arrived here together
      const unsigned am = __activemask();
      const int leader = __ffs((int)am) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(&p.ctr->n_states, (unsigned long long)__popc(am));
      base = __shfl_sync(am, base, leader);
      const unsigned long long pos = base + (unsigned long long)__popc(am & lanemask_lt());
      if (pos < p.cap_states) {
        uint32_t* dst = p.states + pos * (unsigned long long)W;
#pragma unroll
        for (int i = 0; i < W; ++i) dst[i] = w[i];
        p.parent[pos] = (uint32_t)cx->idx;
        p.meta[pos] = ((uint32_t)aid << 8) | (uint32_t)(p.rank & 0xFF);
      } else {
        atomicExch(&p.ctr->store_overflow, 1ULL);
      }
    }
  } else {
    // A successor whose fingerprint this rank has already routed is known to its owner (direct-mapped exact-compare
    // cache).  The slot is written only once the record is in the send region, so a chunk that is re-run after a
    // region overflow loses nothing.
    unsigned long long* cslot = p.sent_cache ? p.sent_cache + (fp & p.sent_mask) : nullptr;
    if (cslot && __ldcv(cslot) == fp) return;
    const unsigned am = __activemask();
    const unsigned peers = __match_any_sync(am, owner);
    const int leader = __ffs((int)peers) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(&p.ctr->send_count[owner], (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const unsigned long long pos = base + (unsigned long long)__popc(peers & lanemask_lt());
    if (pos < p.region_cap) {
      uint32_t* dst = p.send + ((unsigned long long)owner * p.region_cap + pos) * (unsigned long long)(W + 2);
#pragma unroll
      for (int i = 0; i < W; ++i) dst[i] = w[i];
      dst[W] = (uint32_t)cx->idx;
      dst[W + 1] = ((uint32_t)aid << 8) | (uint32_t)(p.rank & 0xFF);
      if (cslot) *cslot = fp;
    } else {
      atomicExch(&p.ctr->route_overflow, 1ULL);
    }
  }
}

// array form: pack the primed frame (all slots, or the dirty ranges over a copy of the parent) first
static __device__ __noinline__ void sl_emit_frame(tlag_sl_cx* cx, const int32_t* f, int aid, int dirty) {
  const DevParams& p = *cx->p;
  constexpr int W = TLAG_SL_W;
  uint32_t succ[W];
  int ov;
  if (dirty > 0) {
#pragma unroll
    for (int i = 0; i < W; ++i) succ[i] = cx->src[i];
    ov = tlag_pack_ranges(p.layout, p.cpool, dirty, f + TLAG_SL_USZ, succ);
  } else {
    ov = tlag_pack(p.layout, p.n_slots, f + TLAG_SL_USZ, succ, W);
  }
  if (ov) {
    cx->nsucc++; cx->gen++;
    report_min(&p.ctr->viol_trap, (cx->idx << 20) | (2ULL << 16) | (unsigned)((ov - 1) & 0xFFFF));
    return;
  }
  sl_emit_words(cx, aid, succ);
}


// One thread per frontier state (grid-stride).  FN: the slice function of this kernel.
template <int PHASE, typename FN>
__device__ __forceinline__ void sl_run(const DevParams& p, unsigned long long lo, unsigned long long hi, FN fn) {
  constexpr int W = TLAG_SL_W;
  tlag_sl_cx cxs;
  tlag_sl_cx* cx = &cxs;
  cx->p = &p; cx->gen = 0; cx->phase = PHASE;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long idx = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < hi; idx += stride) {
    const uint32_t* src = p.states + idx * (unsigned long long)W;
    cx->idx = idx; cx->nsucc = 0;
    // the slice unpacks the slots it reads (liveness, compile/sliced.py) straight out of these registers
    uint32_t in_[W];
#pragma unroll
    for (int i = 0; i < W; ++i) in_[i] = src[i];
    cx->src = src;                 // (the global row: taking the address of in_ would force it out of registers)
    fn(p.cpool, in_, cx);
    if (PHASE == 1 && cx->nsucc && p.succ_flag) p.succ_flag[idx - lo] = 1;
  }
  // generated counter: warp reduce, one atomic per warp
  unsigned long long g = cx->gen;
  if (PHASE == 1) {
    for (int o = 16; o > 0; o >>= 1) g += __shfl_down_sync(0xffffffffu, g, o);
    if ((threadIdx.x & 31) == 0 && g) atomicAdd(&p.ctr->generated, g);
  }
}

#define TLAG_SL_FARG const uint32_t*
#define TLAG_SL_KERNEL_INV(j)                                                                                     \
  __global__ void __launch_bounds__(TLAG_SL_BLOCK, TLAG_SL_OCC) k_sl_inv_##j(const __grid_constant__ DevParams p, \
                                                                            unsigned long long lo, unsigned long long hi) { \
    sl_run<0>(p, lo, hi, [](const int32_t* cp, TLAG_SL_FARG f, tlag_sl_cx* cx) { tlag_sl_inv_##j(cp, f, cx); });  \
  }
#define TLAG_SL_KERNEL_NEXT(j)                                                                                     \
  __global__ void __launch_bounds__(TLAG_SL_BLOCK, TLAG_SL_OCC) k_sl_next_##j(const __grid_constant__ DevParams p, \
                                                                             unsigned long long lo, unsigned long long hi) { \
    sl_run<1>(p, lo, hi, [](const int32_t* cp, TLAG_SL_FARG f, tlag_sl_cx* cx) { tlag_sl_next_##j(cp, f, cx); });  \
  }

// launcher the engine's translation unit calls (kernels live in the slice units)
#define TLAG_SL_LAUNCH_INV(j)                                                                                      \
  extern "C" void tlag_sl_launch_inv_##j(const DevParams* p, unsigned long long lo, unsigned long long hi, unsigned blocks, \
                                         cudaStream_t st) { k_sl_inv_##j<<<blocks, TLAG_SL_BLOCK, 0, st>>>(*p, lo, hi); }
#define TLAG_SL_LAUNCH_NEXT(j)                                                                                     \
  extern "C" void tlag_sl_launch_next_##j(const DevParams* p, unsigned long long lo, unsigned long long hi, unsigned blocks, \
                                          cudaStream_t st) { k_sl_next_##j<<<blocks, TLAG_SL_BLOCK, 0, st>>>(*p, lo, hi); }
