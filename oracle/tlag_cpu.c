/* ORACLE O2 (test infrastructure + CPU baseline, NOT product): a plain-C, multi-threaded
 * CPU execution of the same fixed-width bytecode the CUDA engine runs.
 *
 * Restates the BFS loop of TLC's ModelChecker/Worker ([ext] tla2tools.jar, not under
 * /root/reference): level-synchronous frontier expansion, 64-bit
 * fingerprint set with CAS insertion, invariants on every expanded state, deadlock =
 * no successor.  The opcode semantics come from the shared header
 * tla_rust_b200/csrc/tlag_vm.h (one ISA definition for both sides); the *semantic*
 * oracle that is independent of the compiler and of that header is oracle/tlc_oracle.py
 * (pinned by README.md:267-321).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Build: make -C oracle   ->  oracle/libtlag_cpu.so
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

#include "../tla_rust_b200/csrc/tlag_vm.h"

#define MAXW 128
#define MAX_STEPS (1ull << 38)   /* runaway-program backstop (InnerSerial needs ~10^9 instructions for one successor) */

typedef struct {
  uint32_t W;
  const uint64_t *code; uint32_t code_len;
  uint32_t entry_inv, entry_next;
  const int32_t *cpool; uint32_t cpool_len;
  const tlag_slot *layout; uint32_t n_slots;
  uint32_t frame_words, unpacked_words, n_invariants;
  uint32_t flags;           /* 1 = deadlock check; 4 = sequential TLC-exact mode: one worker, FIFO order, stop at the
                             * first Assert failure / deadlock the way TLC's single worker does (error-time counts) */
  uint32_t table_log2;
  uint64_t max_states;
} cpu_model;

typedef struct {
  int32_t verdict, detail, detail2, pad;
  uint64_t state_idx, generated, distinct, depth, init_states;
  uint64_t fp_xor, fp_sum;
  double seconds;
  uint64_t n_levels;
  uint64_t level_sizes[4096];
  /* cumulative after each level (index = level - 1): XOR / SUM digest of every stored fingerprint, generated count --
   * one long run pins every depth-bounded prefix of the same model */
  uint64_t level_xor[4096], level_sum[4096], level_generated[4096];
} cpu_result;

/* The counters every worker hammers (work dispenser, tail of the state store) each get a cache line of their own:
 * in the first version they shared one with table/mask/lo/hi, which every thread reads for every state, and the run
 * on the GPU box's 128 host cores was slower than on 8 (false sharing). */
typedef struct {
  cpu_model m;
  uint32_t *states; uint32_t *parent; uint32_t *meta;
  uint64_t cap;
  uint64_t *table; uint64_t mask;
  uint64_t lo, hi;
  _Alignas(64) uint64_t work;
  _Alignas(64) uint64_t n_states;
  _Alignas(64) uint64_t generated;
  _Alignas(64) uint64_t viol_inv;
  uint64_t viol_assert, viol_trap, viol_deadlock;
  int overflow, table_full;
  int stop_now;             /* sequential-exact mode: an error was found, stop immediately */
  _Alignas(64) uint64_t dig_xor;
  uint64_t dig_sum;
  /* persistent pool (tlagcpu_run): job 0 = first-touch the table / store stripes, 1 = expand [lo,hi), 2 = quit */
  int job, n_threads;
  uint64_t table_slots;
  pthread_barrier_t bar;
  _Alignas(64) char tail_pad[64];
} cpu_engine;

static int seen_insert(uint64_t *table, uint64_t mask, uint64_t fp) {
  uint64_t i = fp & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    uint64_t cur = __atomic_load_n(&table[i], __ATOMIC_RELAXED);
    if (cur == fp) return 0;
    if (cur == 0) {
      uint64_t exp = 0;
      if (__atomic_compare_exchange_n(&table[i], &exp, fp, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return 1;
      if (exp == fp) return 0;
    }
    i = (i + 1) & mask;
  }
  return -1;
}

static void atomic_min64(uint64_t *p, uint64_t v) {
  uint64_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}

#ifdef TLAG_SLICED_INC
/* Test hook (tests/test_sliced.py): the sliced native build of ONE model (tla_rust_b200/compile/sliced.py -- one C
 * function per invariant and per disjunct of Next) run inside this engine the way the CUDA engine runs it: every slice
 * starts from the packed state and a frame whose temporaries hold garbage (a slice that relied on another slice's
 * temporaries would show up as a digest mismatch), successors are handled where they are produced. */
typedef struct {
  cpu_engine *e; uint64_t idx; const uint32_t *src;
  uint64_t gen, dx, ds; unsigned nsucc; int trapped; int phase;
} tlag_sl_cx;

static void sl_emit_words(tlag_sl_cx *cx, int aid, const uint32_t *succ) {
  cpu_engine *e = cx->e;
  const int W = (int)e->m.W;
  cx->nsucc++; cx->gen++;
  uint64_t fp = tlag_fingerprint(succ, W);
  int ins = seen_insert(e->table, e->mask, fp);
  if (ins < 0) { e->table_full = 1; return; }
  if (ins > 0) {
    uint64_t pos = __atomic_fetch_add(&e->n_states, 1, __ATOMIC_RELAXED);
    if (pos < e->cap) {
      memcpy(e->states + pos * W, succ, (size_t)W * 4);
      e->parent[pos] = (uint32_t)cx->idx;
      e->meta[pos] = (uint32_t)aid << 8;
      cx->dx ^= fp; cx->ds += fp;
    } else e->overflow = 1;
  }
}

static void sl_emit_frame(tlag_sl_cx *cx, const int32_t *f, int aid, int dirty) {
  const cpu_model *m = &cx->e->m;
  const int W = (int)m->W;
  uint32_t succ[MAXW];
  int ov;
  if (dirty > 0) {
    memcpy(succ, cx->src, (size_t)W * 4);
    ov = tlag_pack_ranges(m->layout, m->cpool, dirty, f + m->unpacked_words, succ);
  } else ov = tlag_pack(m->layout, (int)m->n_slots, f + m->unpacked_words, succ, W);
  if (ov) {
    cx->nsucc++; cx->gen++;
    atomic_min64(&cx->e->viol_trap, (cx->idx << 20) | (2ULL << 16) | (uint32_t)((ov - 1) & 0xFFFF));
    return;
  }
  sl_emit_words(cx, aid, succ);
}

#define TLAG_SL_EMIT(aid, dirty) sl_emit_frame(cx, f, (aid), (dirty))
#define TLAG_SL_EMITW(aid, o) sl_emit_words(cx, (aid), (o))
#define TLAG_SL_GEN() do { cx->nsucc++; cx->gen++; } while (0)
#define TLAG_SL_ASSERT(id) atomic_min64(&cx->e->viol_assert, (cx->idx << 20) | (uint32_t)((id) & 0xFFFFF))
#define TLAG_SL_INVF(i) do { if (cx->phase == 0) atomic_min64(&cx->e->viol_inv, (cx->idx << 20) | (uint32_t)((i) & 0xFFFFF)); } while (0)
#define TLAG_SL_TRAP(code, line) do { atomic_min64(&cx->e->viol_trap, (cx->idx << 20) | ((uint64_t)((code) & 15) << 16) | (uint32_t)((line) & 0xFFFF)); cx->trapped = 1; } while (0)
#define TLAG_SL_SUBQ static __attribute__((noinline))
#define TLAG_SL_SEGQ static __attribute__((noinline))
#define TLAG_SL_POISON 1     /* array form: the frame of every slice starts as garbage, as in a fresh CUDA thread */
#include TLAG_SLICED_INC
typedef void (*sl_fn)(const int32_t *, const uint32_t *, tlag_sl_cx *);
#define SL_ADDR_INV(j) tlag_sl_inv_##j,
#define SL_ADDR_NEXT(j) tlag_sl_next_##j,
static const sl_fn sl_inv_fns[] = { TLAG_SL_INV_LIST(SL_ADDR_INV) NULL };
static const sl_fn sl_next_fns[] = { TLAG_SL_NEXT_LIST(SL_ADDR_NEXT) NULL };

static void *worker(void *arg) {
  cpu_engine *e = (cpu_engine *)arg;
  const cpu_model *m = &e->m;
  tlag_sl_cx cxs; memset(&cxs, 0, sizeof(cxs));
  tlag_sl_cx *cx = &cxs;
  cx->e = e;
  const int W = (int)m->W;
  for (;;) {
    uint64_t chunk = __atomic_fetch_add(&e->work, 1, __ATOMIC_RELAXED);
    uint64_t first = e->lo + chunk * 64;
    if (first >= e->hi) break;
    uint64_t last = first + 64 < e->hi ? first + 64 : e->hi;
    for (uint64_t idx = first; idx < last; ++idx) {
      cx->idx = idx; cx->src = e->states + idx * W; cx->nsucc = 0; cx->trapped = 0;
      for (int ph = 0; ph < 2; ++ph) {
        const sl_fn *fns = ph == 0 ? sl_inv_fns : sl_next_fns;
        cx->phase = ph;
        for (int j = 0; fns[j]; ++j) {
          fns[j](m->cpool, cx->src, cx);
        }
      }
      if (cx->nsucc == 0 && !cx->trapped && (m->flags & 1)) atomic_min64(&e->viol_deadlock, idx << 20);
    }
  }
  __atomic_fetch_add(&e->generated, cx->gen, __ATOMIC_RELAXED);
  __atomic_fetch_xor(&e->dig_xor, cx->dx, __ATOMIC_RELAXED);
  __atomic_fetch_add(&e->dig_sum, cx->ds, __ATOMIC_RELAXED);
  return NULL;
}
#else
static void *worker(void *arg) {
  cpu_engine *e = (cpu_engine *)arg;
  const cpu_model *m = &e->m;
  int32_t *frame = (int32_t *)calloc(m->frame_words + 8, 4);
  uint32_t succ[MAXW];
  uint64_t gen = 0, dx = 0, ds = 0;
  const int W = (int)m->W;
  for (;;) {
    uint64_t chunk = __atomic_fetch_add(&e->work, 1, __ATOMIC_RELAXED);
    uint64_t first = e->lo + chunk * 64;
    if (first >= e->hi) break;
    uint64_t last = first + 64 < e->hi ? first + 64 : e->hi;
    for (uint64_t idx = first; idx < last; ++idx) {
      tlag_unpack(m->layout, (int)m->n_slots, e->states + idx * W, frame);
      int trapped = 0;
      if (m->n_invariants) {
        uint32_t pc = m->entry_inv;
        for (;;) {
          int32_t info = 0, info2 = 0;
          int ev = tlag_vm_run(m->code, m->cpool, frame, &pc, &info, &info2, MAX_STEPS);
          if (ev == TLAG_EV_HALT) break;
          if (ev == TLAG_EV_INVF) { atomic_min64(&e->viol_inv, (idx << 20) | (uint32_t)(info & 0xFFFFF)); continue; }
          if (ev == TLAG_EV_ASSERT) { atomic_min64(&e->viol_assert, (idx << 20) | (uint32_t)(info & 0xFFFFF)); continue; }
          atomic_min64(&e->viol_trap, (idx << 20) | ((uint64_t)(ev == TLAG_EV_STEPS ? 15 : (info & 15)) << 16) | (uint32_t)(info2 & 0xFFFF));
          trapped = 1;
          break;
        }
      }
      uint32_t pc = m->entry_next;
      unsigned nsucc = 0;
      while (!trapped) {
        int32_t info = 0, info2 = 0;
        int ev = tlag_vm_run(m->code, m->cpool, frame, &pc, &info, &info2, MAX_STEPS);
        if (ev == TLAG_EV_HALT) break;
        if (ev == TLAG_EV_GEN) { ++nsucc; ++gen; continue; }
        if (ev == TLAG_EV_ASSERT) {
          atomic_min64(&e->viol_assert, (idx << 20) | (uint32_t)(info & 0xFFFFF));
          if (m->flags & 4) { e->stop_now = 1; break; }
          continue;
        }
        if (ev == TLAG_EV_INVF) continue;
        if (ev == TLAG_EV_EMIT) {
          ++nsucc; ++gen;
          int ov;
          if (info2 > 0) {      /* EMITD: re-pack only the dirty slot ranges over the parent's packed words */
            memcpy(succ, e->states + idx * W, (size_t)W * 4);
            ov = tlag_pack_ranges(m->layout, m->cpool, info2, frame + m->unpacked_words, succ);
          } else ov = tlag_pack(m->layout, (int)m->n_slots, frame + m->unpacked_words, succ, W);
          if (ov) { atomic_min64(&e->viol_trap, (idx << 20) | (2ULL << 16) | (uint32_t)((ov - 1) & 0xFFFF)); continue; }
          uint64_t fp = tlag_fingerprint(succ, W);
          int ins = seen_insert(e->table, e->mask, fp);
          if (ins < 0) { e->table_full = 1; continue; }
          if (ins > 0) {
            uint64_t pos = __atomic_fetch_add(&e->n_states, 1, __ATOMIC_RELAXED);
            if (pos < e->cap) {
              memcpy(e->states + pos * W, succ, (size_t)W * 4);
              e->parent[pos] = (uint32_t)idx;
              e->meta[pos] = (uint32_t)info << 8;
              dx ^= fp; ds += fp;
            } else e->overflow = 1;
          }
          continue;
        }
        atomic_min64(&e->viol_trap, (idx << 20) | ((uint64_t)(ev == TLAG_EV_STEPS ? 15 : (info & 15)) << 16) | (uint32_t)(info2 & 0xFFFF));
        trapped = 1;
      }
      if (e->stop_now) break;
      if (nsucc == 0 && !trapped && (m->flags & 1)) {
        atomic_min64(&e->viol_deadlock, idx << 20);
        if (m->flags & 4) { e->stop_now = 1; break; }
      }
    }
    if (e->stop_now) break;
  }
  __atomic_fetch_add(&e->generated, gen, __ATOMIC_RELAXED);
  __atomic_fetch_xor(&e->dig_xor, dx, __ATOMIC_RELAXED);
  __atomic_fetch_add(&e->dig_sum, ds, __ATOMIC_RELAXED);
  free(frame);
  return NULL;
}
#endif

/* Persistent worker of tlagcpu_run: pinned to one allowed CPU, parked on a barrier between levels (the first version
 * created and joined n_threads threads per BFS level and let the kernel place them: the same run measured 0.2, 0.4 and
 * 1.3 M states/s on nominally identical 128-core hosts). */
typedef struct { cpu_engine *e; int tid; int cpu; } pool_arg;

static void *pool_worker(void *arg) {
  pool_arg *pa = (pool_arg *)arg;
  cpu_engine *e = pa->e;
  if (pa->cpu >= 0) {
    cpu_set_t cs; CPU_ZERO(&cs); CPU_SET(pa->cpu, &cs);
    pthread_setaffinity_np(pthread_self(), sizeof(cs), &cs);
  }
  for (;;) {
    pthread_barrier_wait(&e->bar);
    const int job = e->job;
    if (job == 2) break;
    if (job == 0) {
      /* first touch: every worker faults in its own stripes of the table and the store, so the pages are spread over
       * the NUMA nodes of the threads that will hit them at random (and no page fault is left inside the timed BFS) */
      const uint64_t nt = (uint64_t)e->n_threads, t = (uint64_t)pa->tid;
      const uint64_t stripe = 1ull << 18;                                  /* 2 MB of table per stripe */
      for (uint64_t s0 = t * stripe; s0 < e->table_slots; s0 += nt * stripe) {
        uint64_t s1 = s0 + stripe < e->table_slots ? s0 + stripe : e->table_slots;
        memset(e->table + s0, 0, (s1 - s0) * 8);
      }
      const uint64_t words = e->cap * (uint64_t)e->m.W, wstripe = 1ull << 19;
      for (uint64_t s0 = t * wstripe; s0 < words; s0 += nt * wstripe) {
        uint64_t s1 = s0 + wstripe < words ? s0 + wstripe : words;
        memset(e->states + s0, 0, (s1 - s0) * 4);
      }
    } else {
      worker(e);
    }
    pthread_barrier_wait(&e->bar);
  }
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* Full BFS on the host cores.  Returns 0 or a negative error code (-2 memory).
 * stop_after_states: stop at the end of the level that reaches that many distinct states (0 = never);
 * max_levels: stop once that many levels exist (levels 1..max_levels-1 are expanded; 0 = no bound) -- the depth-bounded
 * prefix used for state spaces that do not end (SSI at 4 transactions x 3 keys). */
int tlagcpu_run2(const cpu_model *m, const uint32_t *init, uint64_t n_init, int n_threads,
                 uint64_t stop_after_states, uint64_t max_levels, cpu_result *out, uint32_t *states_out,
                 uint64_t states_out_cap) {
  cpu_engine *ep = NULL;
  if (posix_memalign((void **)&ep, 64, sizeof(cpu_engine))) return -2;
  cpu_engine *e = ep;
  memset(e, 0, sizeof(*e));
  memset(out, 0, sizeof(*out));
  e->m = *m;
  const int W = (int)m->W;
  e->cap = m->max_states ? m->max_states : (1ULL << 22);
  unsigned lg = m->table_log2 ? m->table_log2 : 24;
  while ((1ULL << lg) < e->cap * 2) ++lg;
  e->table_slots = 1ULL << lg;
  e->states = (uint32_t *)malloc(e->cap * (size_t)W * 4);
  e->parent = (uint32_t *)malloc(e->cap * 4);
  e->meta = (uint32_t *)malloc(e->cap * 4);
  e->table = (uint64_t *)malloc(e->table_slots * 8);
  if (!e->states || !e->parent || !e->meta || !e->table) { free(e->states); free(e->parent); free(e->meta); free(e->table); free(ep); return -2; }
  e->mask = e->table_slots - 1;
  e->viol_inv = e->viol_assert = e->viol_trap = e->viol_deadlock = ~0ULL;
  if (n_threads < 1 || (m->flags & 4)) n_threads = 1;
  e->n_threads = n_threads;
  /* pool: one thread per requested worker, pinned round-robin over the CPUs this process may use */
  cpu_set_t allowed;
  int cpus[1024], n_cpus = 0;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
    for (int c = 0; c < 1024 && c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus[n_cpus++] = c;
  pthread_barrier_init(&e->bar, NULL, (unsigned)n_threads + 1);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  pool_arg *pa = (pool_arg *)malloc(sizeof(pool_arg) * (size_t)n_threads);
  for (int t = 0; t < n_threads; ++t) {
    pa[t].e = e; pa[t].tid = t; pa[t].cpu = (n_cpus && n_threads > 1) ? cpus[t % n_cpus] : -1;
    pthread_create(&th[t], NULL, pool_worker, &pa[t]);
  }
  e->job = 0;                                   /* first touch (outside the timed region, like the GPU's cudaMalloc) */
  pthread_barrier_wait(&e->bar); pthread_barrier_wait(&e->bar);
  double t0 = now_s();
  for (uint64_t i = 0; i < n_init; ++i) {
    uint64_t fp = tlag_fingerprint(init + i * W, W);
    e->generated++;
    if (seen_insert(e->table, e->mask, fp) > 0) {
      memcpy(e->states + e->n_states * W, init + i * W, (size_t)W * 4);
      e->parent[e->n_states] = 0xFFFFFFFFu;
      e->meta[e->n_states] = 0xFFFFFF00u;
      e->n_states++;
      e->dig_xor ^= fp; e->dig_sum += fp;
    }
  }
  out->init_states = e->n_states;
  uint64_t lo = 0, hi = e->n_states, level = 1, depth = hi ? 1 : 0;
  out->level_sizes[0] = hi;
  out->level_xor[0] = e->dig_xor; out->level_sum[0] = e->dig_sum; out->level_generated[0] = e->generated;
  out->n_levels = 1;
  int verdict = 0, rc = 0;
  while (lo < hi) {
    if (max_levels && level >= max_levels) break;
    e->lo = lo; e->hi = hi; e->work = 0;
    e->job = 1;
    pthread_barrier_wait(&e->bar); pthread_barrier_wait(&e->bar);
    if (e->overflow || e->table_full) { rc = -2; break; }
    uint64_t best = ~0ULL; int kind = 0;
    if (e->viol_trap != ~0ULL && (e->viol_trap >> 20) < best) { best = e->viol_trap >> 20; kind = 4; }
    if (e->viol_assert != ~0ULL && (e->viol_assert >> 20) < best) { best = e->viol_assert >> 20; kind = 2; }
    if (e->viol_inv != ~0ULL && (e->viol_inv >> 20) < best) { best = e->viol_inv >> 20; kind = 1; }
    if (e->viol_deadlock != ~0ULL && (e->viol_deadlock >> 20) < best) { best = e->viol_deadlock >> 20; kind = 3; }
    if (e->n_states > hi || e->stop_now) depth = level + 1;
    if (out->n_levels < 4096) {
      out->level_xor[out->n_levels] = e->dig_xor; out->level_sum[out->n_levels] = e->dig_sum;
      out->level_generated[out->n_levels] = e->generated;
      out->level_sizes[out->n_levels++] = e->n_states - hi;
    }
    lo = hi; hi = e->n_states; level++;
    if (kind) {
      verdict = kind; out->state_idx = best;
      if (kind == 4) { out->detail = (int)((e->viol_trap >> 16) & 15); out->detail2 = (int)(e->viol_trap & 0xFFFF); }
      else if (kind == 2) out->detail = (int)(e->viol_assert & 0xFFFFF);
      else if (kind == 1) out->detail = (int)(e->viol_inv & 0xFFFFF);
      break;
    }
    if (stop_after_states && e->n_states >= stop_after_states) break;
  }
  out->seconds = now_s() - t0;
  e->job = 2;
  pthread_barrier_wait(&e->bar);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  pthread_barrier_destroy(&e->bar);
  free(th); free(pa);
  if (rc == 0) {
    out->verdict = verdict;
    out->generated = e->generated;
    out->distinct = e->n_states;
    out->depth = depth;
    out->fp_xor = e->dig_xor; out->fp_sum = e->dig_sum;
    if (states_out) {
      uint64_t n = e->n_states < states_out_cap ? e->n_states : states_out_cap;
      memcpy(states_out, e->states, n * (size_t)W * 4);
    }
  }
  free(e->states); free(e->parent); free(e->meta); free(e->table); free(ep);
  return rc;
}

int tlagcpu_run(const cpu_model *m, const uint32_t *init, uint64_t n_init, int n_threads,
                uint64_t stop_after_states, cpu_result *out, uint32_t *states_out, uint64_t states_out_cap) {
  return tlagcpu_run2(m, init, n_init, n_threads, stop_after_states, 0, out, states_out, states_out_cap);
}

uint64_t tlagcpu_fingerprint(const uint32_t *w, int W) { return tlag_fingerprint(w, W); }

void tlagcpu_digest(const uint32_t *states, uint64_t n, int W, uint64_t *out2) {
  uint64_t x = 0, s = 0;
  for (uint64_t i = 0; i < n; ++i) { uint64_t fp = tlag_fingerprint(states + i * (uint64_t)W, W); x ^= fp; s += fp; }
  out2[0] = x; out2[1] = s;
}

/* K1 on the host: fingerprint + probe/insert over a batch, n_threads workers, returns seconds. */
typedef struct { const uint32_t *states; uint64_t lo, hi; int W; uint64_t *table; uint64_t mask; uint8_t *is_new; } probe_job;

static void *probe_worker(void *arg) {
  probe_job *j = (probe_job *)arg;
  for (uint64_t i = j->lo; i < j->hi; ++i) {
    uint64_t fp = tlag_fingerprint(j->states + i * (uint64_t)j->W, j->W);
    int ins = seen_insert(j->table, j->mask, fp);
    j->is_new[i] = (uint8_t)(ins > 0);
  }
  return NULL;
}

double tlagcpu_probe_batch(const uint32_t *states, uint64_t n, int W, unsigned table_log2, int n_threads, uint8_t *is_new) {
  uint64_t *table = (uint64_t *)calloc(1ULL << table_log2, 8);
  if (!table) return -1.0;
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  probe_job *jobs = (probe_job *)malloc(sizeof(probe_job) * (size_t)n_threads);
  double t0 = now_s();
  for (int t = 0; t < n_threads; ++t) {
    jobs[t] = (probe_job){states, n * t / n_threads, n * (t + 1) / n_threads, W, table, (1ULL << table_log2) - 1, is_new};
    pthread_create(&th[t], NULL, probe_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  double dt = now_s() - t0;
  free(th); free(jobs); free(table);
  return dt;
}

/* pack / unpack helpers exposed for host-side tests of the layout */
int tlagcpu_pack(const tlag_slot *lay, int nslots, const int32_t *st, uint32_t *out, int W) { return tlag_pack(lay, nslots, st, out, W); }
void tlagcpu_unpack(const tlag_slot *lay, int nslots, const uint32_t *in, int32_t *st) { tlag_unpack(lay, nslots, in, st); }

/* ---- stateful shard engine: the CPU stand-in for tlag_expand_route / tlag_insert_records /
 * tlag_advance_level, used by the world_size-2 gloo tests of the multi-GPU host logic. ---------- */
typedef struct {
  cpu_engine e;
  uint64_t lo, hi, level, depth, init_states;
  int verdict, detail;
  uint64_t viol_idx;       /* first violating state of this shard (index in its store) */
  uint32_t rank;           /* written into the meta word of routed records (whose store holds the parent) */
} cpu_shard;

cpu_shard *tlagcpu_shard_create(const cpu_model *m) {
  cpu_shard *s = (cpu_shard *)calloc(1, sizeof(cpu_shard));
  s->e.m = *m;
  s->e.cap = m->max_states ? m->max_states : (1ULL << 20);
  unsigned lg = 8;
  while ((1ULL << lg) < s->e.cap * 2) ++lg;
  s->e.states = (uint32_t *)malloc(s->e.cap * (size_t)m->W * 4);
  s->e.parent = (uint32_t *)malloc(s->e.cap * 4);
  s->e.meta = (uint32_t *)malloc(s->e.cap * 4);
  s->e.table = (uint64_t *)calloc(1ULL << lg, 8);
  s->e.mask = (1ULL << lg) - 1;
  s->e.viol_inv = s->e.viol_assert = s->e.viol_trap = s->e.viol_deadlock = ~0ULL;
  return s;
}

void tlagcpu_shard_destroy(cpu_shard *s) {
  if (!s) return;
  free(s->e.states); free(s->e.parent); free(s->e.meta); free(s->e.table); free(s);
}

/* records: W words + parent + meta */
uint64_t tlagcpu_shard_insert(cpu_shard *s, const uint32_t *rec, uint64_t n) {
  const int W = (int)s->e.m.W;
  uint64_t n_new = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t *r = rec + i * (uint64_t)(W + 2);
    uint64_t fp = tlag_fingerprint(r, W);
    if (seen_insert(s->e.table, s->e.mask, fp) > 0) {
      uint64_t pos = s->e.n_states++;
      memcpy(s->e.states + pos * W, r, (size_t)W * 4);
      s->e.parent[pos] = r[W]; s->e.meta[pos] = r[W + 1];
      ++n_new;
    }
  }
  return n_new;
}

void tlagcpu_shard_seed(cpu_shard *s, const uint32_t *init, uint64_t n) {
  const int W = (int)s->e.m.W;
  uint32_t rec[MAXW + 2];
  for (uint64_t i = 0; i < n; ++i) {
    memcpy(rec, init + i * W, (size_t)W * 4);
    rec[W] = 0xFFFFFFFFu; rec[W + 1] = 0xFFFFFF00u;
    tlagcpu_shard_insert(s, rec, 1);
  }
  s->e.generated += n;
  s->hi = s->e.n_states;
  s->init_states = s->e.n_states;
}

/* expand frontier [lo,hi): successors bucketed by owner = floor(fp * n_ranks / 2^64) into
 * send[owner * region_cap ...]; counts[r] records per rank.  Returns verdict kind (0 none). */
void tlagcpu_shard_frontier(cpu_shard *s, uint64_t *out2) {
  if (s->level == 0) { s->level = 1; s->lo = 0; s->depth = s->hi ? 1 : 0; }
  out2[0] = s->lo; out2[1] = s->hi - s->lo;
}

int tlagcpu_shard_expand_route(cpu_shard *s, uint32_t n_ranks, uint64_t first, uint64_t count, uint32_t *send,
                               uint64_t cap_records, uint64_t *counts, uint64_t *generated_out) {
  const cpu_model *m = &s->e.m;
  const int W = (int)m->W;
  const uint64_t region = cap_records / n_ranks;
  int32_t *frame = (int32_t *)calloc(m->frame_words + 8, 4);
  uint32_t succ[MAXW];
  uint64_t gen = 0;
  if (s->level == 0) { s->level = 1; s->lo = 0; s->depth = s->hi ? 1 : 0; }
  for (uint32_t r = 0; r < n_ranks; ++r) counts[r] = 0;
  int kind = 0;
  uint64_t c_lo = s->lo + first, c_hi = c_lo + count;
  if (c_lo > s->hi) c_lo = s->hi;
  if (c_hi > s->hi) c_hi = s->hi;
  for (uint64_t idx = c_lo; idx < c_hi; ++idx) {
    tlag_unpack(m->layout, (int)m->n_slots, s->e.states + idx * W, frame);
    if (m->n_invariants) {
      uint32_t pc = m->entry_inv;
      for (;;) {
        int32_t info = 0, info2 = 0;
        int ev = tlag_vm_run(m->code, m->cpool, frame, &pc, &info, &info2, MAX_STEPS);
        if (ev == TLAG_EV_HALT) break;
        if (ev == TLAG_EV_INVF) { if (!kind) { kind = 1; s->detail = info; if (!s->verdict) s->viol_idx = idx; } continue; }
        if (!kind) { kind = 4; if (!s->verdict) s->viol_idx = idx; }
        break;
      }
    }
    uint32_t pc = m->entry_next;
    unsigned nsucc = 0;
    for (;;) {
      int32_t info = 0, info2 = 0;
      int ev = tlag_vm_run(m->code, m->cpool, frame, &pc, &info, &info2, MAX_STEPS);
      if (ev == TLAG_EV_HALT) break;
      if (ev == TLAG_EV_GEN) { ++nsucc; ++gen; continue; }
      if (ev == TLAG_EV_ASSERT) { if (!kind) { kind = 2; s->detail = info; if (!s->verdict) s->viol_idx = idx; } continue; }
      if (ev == TLAG_EV_EMIT) {
        ++nsucc; ++gen;
        int ov;
        if (info2 > 0) {
          memcpy(succ, s->e.states + idx * W, (size_t)W * 4);
          ov = tlag_pack_ranges(m->layout, m->cpool, info2, frame + m->unpacked_words, succ);
        } else ov = tlag_pack(m->layout, (int)m->n_slots, frame + m->unpacked_words, succ, W);
        if (ov) { if (!kind) kind = 4; continue; }
        uint32_t owner = tlag_owner(succ, W, n_ranks);
        if (counts[owner] >= region) { free(frame); return -5; }
        uint32_t *dst = send + (owner * region + counts[owner]) * (uint64_t)(W + 2);
        memcpy(dst, succ, (size_t)W * 4);
        dst[W] = (uint32_t)idx; dst[W + 1] = ((uint32_t)info << 8) | (s->rank & 0xFF);
        counts[owner]++;
        continue;
      }
      if (!kind) kind = 4;
      break;
    }
    if (nsucc == 0 && (m->flags & 1) && !kind) { kind = 3; if (!s->verdict) s->viol_idx = idx; }
  }
  free(frame);
  s->e.generated += gen;
  if (generated_out) *generated_out = gen;
  if (kind && !s->verdict) s->verdict = kind;
  return kind;
}

void tlagcpu_shard_advance(cpu_shard *s) {
  if (s->e.n_states > s->hi) s->depth = s->level + 1;
  s->lo = s->hi; s->hi = s->e.n_states; s->level += 1;
}

void tlagcpu_shard_result(cpu_shard *s, uint64_t *out4) {
  out4[0] = s->e.generated; out4[1] = s->e.n_states; out4[2] = s->depth; out4[3] = (uint64_t)s->verdict;
}

void tlagcpu_shard_set_rank(cpu_shard *s, uint32_t rank) { s->rank = rank; }

/* violation of this shard (verdict, detail, state index) and one hop of a parent chain, for the cross-rank trace */
void tlagcpu_shard_violation(cpu_shard *s, uint64_t *out3) { out3[0] = (uint64_t)s->verdict; out3[1] = (uint64_t)s->detail; out3[2] = s->viol_idx; }
int tlagcpu_shard_read_link(cpu_shard *s, uint64_t idx, uint32_t *state_out, uint32_t *parent_out, uint32_t *meta_out) {
  if (idx >= s->e.n_states) return -1;
  memcpy(state_out, s->e.states + idx * s->e.m.W, (size_t)s->e.m.W * 4);
  *parent_out = s->e.parent[idx]; *meta_out = s->e.meta[idx];
  return 0;
}

uint32_t tlagcpu_owner(const uint32_t *w, int W, uint32_t n_ranks) { return tlag_owner(w, W, n_ranks); }

