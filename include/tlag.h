/* tlag.h -- C ABI of the B200 explicit-state BFS engine (boundary B1 of SURVEY.md §8b).
 *
 * The reference (spacejam/tla-rust) contains no FFI, plugin or operator API at all: its
 * only integration point is the process boundary `tlc FILE.tla...` (Makefile:6-7,
 * README.md:260-263), behind which the external TLC does parse -> BFS -> report.  This
 * header is the thin C-ABI the north star asks for between a host front end (which
 * parses .tla/.cfg and lowers Next / invariants to fixed-width bytecode) and the
 * sm_100a kernels that replace TLC's Worker BFS loop + StateQueue, next-state
 * evaluator, FPSet and invariant checker.  A Rust host binds it with a
 * plain `extern "C"` block; nothing here mentions torch or C++.
 *
 * Ownership: the caller owns every buffer it passes in or receives results in; the
 * engine copies inputs during the call.  The engine owns all device memory.  No
 * pointer returned by the engine outlives tlag_destroy (tlag_last_error is valid until
 * the next call on that engine).  Calls on one engine are not re-entrant.
 * Errors: negative return codes + tlag_last_error.  A model-checking verdict
 * (invariant violated, Assert failed, deadlock, evaluation trap) is NOT an error: it is
 * reported in tlag_result.verdict with return code 0.
 */
#ifndef TLAG_H
#define TLAG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tlag_engine tlag_engine;

enum { TLAG_OK = 0, TLAG_EINVAL = -1, TLAG_ENOMEM = -2, TLAG_ECUDA = -3, TLAG_ENCCL = -4,
       TLAG_EOVERFLOW = -5, TLAG_EEVAL = -6, TLAG_ESTATE = -7 };

/* verdicts (what `tlc` prints: README.md:267-321) */
enum { TLAG_V_OK = 0, TLAG_V_INVARIANT = 1, TLAG_V_ASSERT = 2, TLAG_V_DEADLOCK = 3,
       TLAG_V_EVAL_ERROR = 4, TLAG_V_RUNNING = 5 };

enum { TLAG_F_DEADLOCK_CHECK = 1u,   /* report states without successors            */
       TLAG_F_KEEP_GOING     = 2u,   /* do not stop a run at the first violation    */
       TLAG_F_EXACT          = 4u }; /* TLC-exact replay: the search runs as ONE sequential worker on the device (states
                                      * dequeued in FIFO order, successors in program order) and stops AT the first
                                      * Assert failure / deadlock, so the counts at that moment, the state reported and
                                      * its parent chain are the ones TLC's single worker prints (README.md:267-321).
                                      * Interpreter kernel only; meant for re-running a small model after an error. */

typedef struct {
  uint32_t words_per_state;      /* W: packed state vector width in u32 words (1..128)        */
  const uint64_t *code;          /* bytecode image (compile/bytecode.py), host-owned, copied   */
  uint32_t code_len;             /* in 64-bit instructions                                     */
  uint32_t entry_inv;            /* pc of the invariant program (runs on each expanded state)  */
  uint32_t entry_next;           /* pc of the next-state program (EMITs successors)            */
  const int32_t *cpool;          /* constant pool (tables), copied                             */
  uint32_t cpool_len;
  const int32_t *layout;         /* n_slots x {frame_off, width_bits, bias}: packed layout     */
  uint32_t n_slots;
  uint32_t frame_words;          /* per-thread VM frame size in words (<= 8192)                */
  uint32_t unpacked_words;       /* words of one unpacked state (primed copy follows it)       */
  uint32_t n_invariants;
  uint32_t n_actions;
  uint32_t table_slots_log2;     /* seen-set capacity (8-byte slots); grows by rehash if 0<load*/
  uint64_t max_states;           /* capacity of the state store (0 = size from free memory)    */
  uint32_t flags;                /* TLAG_F_*                                                   */
  int32_t  device;               /* CUDA device ordinal                                        */
} tlag_model;

typedef struct {
  uint64_t level;                /* BFS level just expanded (initial states are level 1)       */
  uint64_t expanded;             /* states expanded in this wave                               */
  uint64_t generated;            /* successors generated in this wave (incl. duplicates)       */
  uint64_t discovered;           /* new distinct states found in this wave                     */
  uint64_t distinct_total;
  uint64_t generated_total;
  float    kernel_ms;            /* device time of the wave kernel (CUDA events)               */
  int32_t  verdict;              /* TLAG_V_RUNNING while the search continues                  */
} tlag_wave_stats;

typedef struct {
  int32_t  verdict;              /* TLAG_V_*                                                   */
  int32_t  detail;               /* invariant index / assert id / trap code                    */
  int32_t  detail2;              /* trap source line                                           */
  int32_t  reserved;
  uint64_t state_idx;            /* index (discovery order) of the offending state             */
  uint64_t generated, distinct, queue_left, depth, init_states;
  double   fp_collision_estimate;/* n*(g-n)/2^64, as TLC's "calculated (optimistic)"           */
  double   device_seconds;       /* sum of wave-kernel times                                   */
} tlag_result;

int  tlag_create(const tlag_model *m, tlag_engine **out);
/* n initial states, W words each (host memory).  Duplicates are dropped; every one counts
 * as "generated" like TLC's "Finished computing initial states" line (testout2:3). */
int  tlag_seed(tlag_engine *e, const uint32_t *states, uint64_t n);
int  tlag_step(tlag_engine *e, tlag_wave_stats *out);          /* one BFS level; blocks       */
int  tlag_run(tlag_engine *e, tlag_result *out);               /* to fixpoint / violation     */
int  tlag_result_now(tlag_engine *e, tlag_result *out);        /* counters so far             */
/* Counterexample reconstruction: walks parent links from state_idx back to an initial
 * state.  states_out receives len*W words (initial state first), actions_out len action
 * ids (entry 0 is -1).  *len_inout: capacity in, length out. */
int  tlag_trace(tlag_engine *e, uint64_t state_idx, uint32_t *states_out, int32_t *actions_out,
                uint32_t *len_inout);
int  tlag_read_states(tlag_engine *e, uint64_t first, uint64_t n, uint32_t *states_out);
/* One hop of a parent chain: state idx (W words), its parent's index and the meta word = action id << 8 | rank whose
 * store holds the parent (multi-GPU: the host follows the chain from rank to rank).  parent 0xFFFFFFFF = initial state. */
int  tlag_read_link(tlag_engine *e, uint64_t idx, uint32_t *state_out, uint32_t *parent_out, uint32_t *meta_out);
/* Checksum of checksums over every stored state: XOR and SUM (mod 2^64) of the 64-bit fingerprints
 * (size-independent parity property for state spaces too large to read back). */
int  tlag_digest(tlag_engine *e, uint64_t *xor_out, uint64_t *sum_out);
/* K1 alone (fingerprint + seen-set probe/insert) on caller-provided HOST states: unit tests
 * and the e2e leg of the roofline bench.  is_new[i] = 1 iff state i was not in the set. */
int  tlag_probe_batch(tlag_engine *e, const uint32_t *states, uint64_t n, uint8_t *is_new);
/* Same on DEVICE-resident buffers (d_states: n*W u32, d_is_new: n bytes); returns the kernel
 * time measured with CUDA events on the engine's stream. */
int  tlag_probe_batch_device(tlag_engine *e, uint64_t d_states, uint64_t n, uint64_t d_is_new,
                             float *kernel_ms);
int  tlag_reset_table(tlag_engine *e);                         /* empty the seen-set           */
int  tlag_restart(tlag_engine *e);     /* forget all discovered states; re-seed the retained initial states */
uint64_t tlag_kernel_launches(const tlag_engine *e);           /* kernels launched so far      */
void tlag_destroy(tlag_engine *e);
const char *tlag_last_error(const tlag_engine *e);
const char *tlag_version(void);

/* ---- multi-GPU building blocks (one engine per rank; the host exchanges the buffers,
 *      e.g. torch.distributed all_to_all_single over NCCL/NVLink) ---------- */
/* Expand the current frontier WITHOUT inserting: successor records (W state words + parent
 * index + action id, record = W+2 u32) are bucketed by owner rank = fingerprint >> (64-log2 R)
 * ... into d_send (device, capacity cap_records), counts[r] records per rank (host out). */
int  tlag_frontier(tlag_engine *e, uint64_t *first_idx, uint64_t *count);   /* current frontier slice */
/* Which shard this engine holds (before the first tlag_expand_route of a multi-rank run): successors the rank owns
 * itself are inserted in place, and the rank is recorded in the meta word of the states whose parent it expands. */
int  tlag_set_rank(tlag_engine *e, uint32_t n_ranks, uint32_t rank);
/* Trailing packed words the ownership hash covers (same on every rank, before anything is routed): 2 (default) keeps
 * clusters of like states on one rank, words_per_state spreads models whose tail words carry little entropy. */
int  tlag_set_owner_words(tlag_engine *e, uint32_t k);
/* `first`/`count` select a sub-range of the frontier (chunked exchange with bounded buffers). */
int  tlag_expand_route(tlag_engine *e, uint32_t n_ranks, uint64_t first, uint64_t count, uint64_t d_send,
                       uint64_t cap_records, uint64_t *counts, tlag_wave_stats *out);
/* Insert received records (device buffer) into this rank's seen-set shard / state store. */
int  tlag_insert_records(tlag_engine *e, uint64_t d_recv, uint64_t n_records, uint32_t src_rank_unused,
                         uint64_t *n_new);
int  tlag_advance_level(tlag_engine *e, tlag_wave_stats *out);  /* frontier <- newly inserted  */

/* ---- peer-memory exchange (NVLink / NVSwitch; no NCCL on the data path) ---------------------------------------
 * tlag_p2p_init allocates this rank's inbox (2 buffers x n_ranks regions of cap_records records) and send regions
 * and returns a 64-byte CUDA IPC handle of the inbox; the host all-gathers the handles (any transport) and calls
 * tlag_p2p_attach for every peer.  tlag_p2p_level then runs one whole BFS level on this rank: per chunk of
 * chunk_states frontier states the expand kernels bucket successors by owner, k_push stores the buckets into the
 * owners' inboxes and publishes {count, chunk number} with a system-scope release, k_insert_inbox waits for every
 * source's chunk, inserts the records (this rank's share of the next frontier) and acknowledges.  All ranks must
 * pass the same n_chunks.  expect_inbound: upper estimate of the records this rank may receive (store head-room).
 * Follow with tlag_advance_level; termination is the host's all-reduce of `discovered`. */
int  tlag_p2p_init(tlag_engine *e, uint32_t n_ranks, uint32_t rank, uint64_t cap_records, uint8_t *handle_out64);
int  tlag_p2p_attach(tlag_engine *e, uint32_t peer, const uint8_t *handle64);
int  tlag_p2p_level(tlag_engine *e, uint64_t n_chunks, uint64_t chunk_states, uint64_t expect_inbound,
                    tlag_wave_stats *out);
/* TLAG_EOVERFLOW from tlag_p2p_level on ANY rank (a send region was too small for a chunk): every rank calls
 * tlag_p2p_rollback -- the level's appended states are dropped, the seen-set rebuilt -- and the level is run again
 * with smaller chunks. */
int  tlag_p2p_rollback(tlag_engine *e);

#ifdef __cplusplus
}
#endif
#endif
