/*
 * b200sv -- C ABI of the B200-native complex128 statevector engine.
 *
 * This is the drop-in boundary for blueqat's dense statevector path: every entry point below replaces a
 * piece of blueqat/backends/numpy_backend.py (reference file:line given per function).  The host side
 * (qaqarot_b200/backend.py, loaded through ctypes) keeps the reference's plugin interface
 * (`Backend.run(gates, n_qubits, shots=, initial=, returns=, ...)`, backendbase.py:89-91) and calls down
 * here for every pass over amplitudes.  No torch types, no C++ types: plain pointers and sizes.
 *
 * Conventions
 *   - A state is 2^n complex128 amplitudes, interleaved (re, im), little endian in the qubit index:
 *     bit k of the amplitude index is qubit k (numpy_backend.py:82-83).
 *   - Every function returns 0 on success; otherwise b200_last_error() describes the failure
 *     (thread local).  Functions are synchronous on return unless stated; one state must not be used
 *     from two host threads at once.
 *   - Host pointers are caller owned.  `b200_state*` handles are opaque and own their device memory
 *     unless created with b200_state_wrap().
 */
#ifndef B200SV_H
#define B200SV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_state b200_state;

/* ---- op program ------------------------------------------------------------------------------- */

enum {
  B200_OP_MAT   = 1,  /* 2x2 on `target` where all `cmask` bits are 1: gate_x/y/h/rx/ry/u/mat1/cx/cu/crx/cry/ccx
                         (numpy_backend.py:75-106,119-177,244-338,385-444).  reals: m00 m01 m10 m11 as (re,im) */
  B200_OP_DIAG  = 2,  /* amp *= (bit(target) ? d1 : d0) where all `cmask` bits are 1: gate_z/s/t/phase/rz/cz/
                         crz/cphase/ccz (:109-116,180-241,341-382).  reals: d0 d1 as (re,im)                */
  B200_OP_X     = 3,  /* pair swap on `target` under `cmask` (x / cx / ccx as pure permutations)          */
  B200_OP_SWAP  = 4,  /* exchange qubits `target` and `aux` (swap gate; reference falls back to 3 cx,
                         gate.py:1172-1177)                                                                */
  B200_OP_PERMUTE = 5, /* in-place involution of index bits: `woff/wlen` locate wlen words p | q << 8, disjoint
                         transpositions (p, q).  The planner relabels swap gates and emits at most two of
                         these at the end of a program instead of 3 CX sweeps per swap (gate.py:1172-1177)   */
  B200_OP_TILE  = 16, /* fused tile-resident pass: `woff/wlen` locate its descriptor in `words`            */
  B200_OP_SCALE = 17, /* multiply the whole state by reals[0..1]                                           */
  B200_OP_INIT  = 18  /* state := |cmask> (prepare(), numpy_backend.py:52-62).  Directly before a TILE op it costs
                         nothing: that pass synthesises its input instead of reading it (no memset sweep, and the
                         pass only writes); `cmask` carries the basis index including the shard bits             */
};

typedef struct b200_op {
  int32_t  kind;    /* B200_OP_*                                            */
  int32_t  target;  /* target qubit                                         */
  int32_t  aux;     /* second qubit (SWAP)                                  */
  int32_t  wlen;    /* TILE / PERMUTE: descriptor length in 64-bit words    */
  uint64_t cmask;   /* control mask over qubit indices (0 = uncontrolled)   */
  uint64_t woff;    /* TILE: descriptor offset into `words`                 */
  uint64_t roff;    /* offset of this op's payload in `reals` (doubles)     */
} b200_op;

/* ---- library ---------------------------------------------------------------------------------- */

const char* b200_last_error(void);
int b200_device_count(int* out);
int b200_version(void);                 /* 104; bumped whenever the TILE descriptor format changes */

/* ---- state lifetime (replaces _NumPyBackendContext.__init__/prepare, numpy_backend.py:36-62) ---- */

int b200_state_create(int device, int n_qubits, b200_state** out);
/* Adopt caller-owned device memory (e.g. a torch tensor's data_ptr) of 2^n_qubits complex128. */
int b200_state_wrap(int device, int n_qubits, void* device_amps, b200_state** out);
int b200_state_destroy(b200_state* st);
int b200_state_n_qubits(const b200_state* st);
void* b200_state_device_ptr(const b200_state* st);

int b200_state_set_basis(b200_state* st, uint64_t index);                       /* prepare(): |index>      */
int b200_state_upload(b200_state* st, const double* host, uint64_t offset, uint64_t count);   /* initial= */
int b200_state_download(const b200_state* st, double* host, uint64_t offset, uint64_t count); /* returns   */
int b200_state_copy(b200_state* dst, const b200_state* src);                    /* ctx.cache (:564-568)    */
/* A state that is one shard of a larger register: `offset` carries the index bits above n_qubits (the shard
   number shifted left by n_qubits).  Controls, phases and Pauli-Z signs on those qubits read them from here. */
int b200_state_set_index_offset(b200_state* st, uint64_t offset);

/* ---- unitary evolution (replaces _run_inner's gate loop, numpy_backend.py:545-570) -------------- */

int b200_apply_program(b200_state* st, const b200_op* ops, int n_ops,
                       const uint64_t* words, size_t n_words, const double* reals, size_t n_reals);
/* The same, piece by piece: the objective loop of vqe.py:67-83 re-plans the circuit for every angle vector, and the
   device can run the first passes while the host is still planning the later ones.  A pass takes everything it
   needs from the host arrays at launch (kernel parameters), so a piece may be handed over as soon as it is encoded.
   flags: B200_APPLY_ASYNC = return without waiting for the device (b200_sync / any reduction waits),
          B200_APPLY_CONTINUE = a further piece of a program whose first piece started b200_last_program_ms's timer. */
#define B200_APPLY_ASYNC 1
#define B200_APPLY_CONTINUE 2
int b200_apply_program_ex(b200_state* st, const b200_op* ops, int n_ops,
                          const uint64_t* words, size_t n_words, const double* reals, size_t n_reals, int flags);
/* Device time of the last b200_apply_program (CUDA events on the state's stream), in milliseconds. */
int b200_last_program_ms(b200_state* st, double* ms);
/* Per-op device times of the last program (only recorded while profiling is on). Returns count in *n. */
int b200_set_profiling(b200_state* st, int on);
int b200_last_op_times(b200_state* st, float* ms, int32_t* kinds, int cap, int* n);
/* Number of kernels this library has launched on this state since creation. */
int b200_launch_count(const b200_state* st, uint64_t* out);

/* ---- measurement (replaces gate_measure / gate_reset, numpy_backend.py:447-497) ------------------ */

int b200_prob_zero(b200_state* st, int target, double* p0);          /* ||psi[bit target = 0]||^2        */
/* mode 0 = measure: keep `outcome` half, zero the other, scale by 1/sqrt(p).
   mode 1 = reset  : as measure, then move the |1> half onto |0> when outcome = 1.                      */
int b200_collapse(b200_state* st, int target, int outcome, double p, int mode);
int b200_norm2(b200_state* st, double* out);
int b200_scale(b200_state* st, double re, double im);
/* First amplitude with |a| > eps in index order (utils.ignore_global_phase, utils.py:130-144).
   *index = UINT64_MAX when there is none. */
int b200_first_above(b200_state* st, double eps, uint64_t* index, double* re, double* im);
/* The same search for a shard whose wires sit on permuted index bits (sharded states keep the layout their swaps and
   exchanges left): *key = the smallest WIRE-order index of an amplitude of this shard with |a| > eps, where bit p of
   (index_offset | local index) is wire wire_of_bit[p] (a permutation of 0..n_bits-1, n_bits <= 40 covering the whole
   register); UINT64_MAX when there is none.  The caller takes the minimum over the shards and reads the amplitude
   with b200_state_download. */
int b200_first_above_wire(b200_state* st, double eps, const int32_t* wire_of_bit, int n_bits, uint64_t* key);

/* Batched terminal-measurement sampler.  For each shot s, walks the measured qubits in `order`
   (the reference's target_iter order), comparing uniforms[s*m + j] with the conditional probability
   P(bit_j = 0 | earlier outcomes) exactly as repeated gate_measure calls would, without collapsing
   the state.  out_bits[s] gets bit j set when qubit order[j] measured 1.  m <= 64.                     */
int b200_sample(b200_state* st, const int32_t* order, int m, const double* uniforms, int n_shots,
                uint64_t* out_bits);
/* One level of that descent, for callers that combine partial sums themselves (a state sharded over several
   GPUs adds the per-shard sums before comparing with the uniforms).  For every group g -- the amplitudes whose
   index equals group_bits[g] on `fixed_mask` -- sums[2g] = sum |amp|^2 with bit `qubit` = 0 and sums[2g+1]
   with bit = 1; qubit < 0: sums[2g] is the whole group.                                                   */
int b200_cond_probs(b200_state* st, uint64_t fixed_mask, int qubit, const uint64_t* group_bits, int n_groups,
                    double* sums);
/* Marginal distribution of the low k index bits (8 <= k <= 13): table[v] = sum of |amp[i]|^2 over the i with
   i mod 2^k = v, from one contiguous read of the state.  The first levels of the descent when the first measured
   qubits sit on index bits 0..k-1 (b200_sample uses it itself; sharded callers add the per-shard tables).     */
int b200_marginal_low(b200_state* st, int k, double* table);

/* ---- Pauli expectation (replaces Expr.to_matrix + sparse_expectation, pauli.py:893-935,
        vqe.py:355-366).  One group = all terms sharing the same X/Y flip mask `xmask`:
        sum_i conj(psi[i ^ xmask]) * psi[i] * sum_k (cre_k + i cim_k) * (-1)^popcount(i & zymask_k).       */
int b200_expect_group(b200_state* st, uint64_t xmask, int n_terms, const uint64_t* zymask,
                      const double* cre, const double* cim, double* out_re, double* out_im);

int b200_sync(b200_state* st);

/* ---- peer-to-peer exchange for states sharded over several GPUs, one process each (no reference
        counterpart: blueqat has no distributed path) ----------------------------------------- */

/* CUDA IPC handle (64 bytes) of a state created by b200_state_create; the partner process maps it. */
int b200_state_ipc_export(const b200_state* st, unsigned char* handle64);
int b200_ipc_open(int device, const unsigned char* handle64, void** peer_amps);
int b200_ipc_close(void* peer_amps);
/* Exchange local index bit `lbit` with the global bit in which the two shards differ (`my_bit` = this shard's
   value of it): for the pairs i in [start, start + count) of the 2^(n-1) index patterns without bit lbit,
   amp[i with lbit = 1 - my_bit] <-> peer_amps[i with lbit = my_bit], by loads / stores over NVLink, synchronous.
   The two partners call it on disjoint pair ranges; the caller fences both processes before and after. */
int b200_peer_swap(b200_state* st, void* peer_amps, int lbit, int my_bit, uint64_t start, uint64_t count);
/* k local index bits `lbits` (ascending) exchanged with k global bits in ONE step: called once per partner shard (the
   2^k - 1 shards that differ from this one in a subset of those global bits).  For the patterns i in [start, start +
   count) of the 2^(n-k) index patterns without the k bits, amp[i | mine_pat] <-> peer_amps[i | peer_pat], where
   mine_pat holds the partner's values of the global bits on the local bits and peer_pat this shard's.  Asynchronous on
   the state's stream unless sync != 0; partners take disjoint ranges; the caller fences all processes before and after. */
int b200_peer_swap_bits(b200_state* st, void* peer_amps, const int32_t* lbits, int k, uint64_t mine_pat, uint64_t peer_pat,
                        uint64_t start, uint64_t count, int sync);

/* ---- measurement support for bench.py / profiling ------------------------------------------------ */

/* Record CUDA event `slot` (0..15) on the state's stream; elapsed device time between two recorded slots. */
int b200_event_record(b200_state* st, int slot);
int b200_event_elapsed_ms(b200_state* st, int slot_a, int slot_b, double* ms);

/* Page-locked host memory for upload/download staging (the reference returns a host ndarray;
   a pinned destination lets b200_state_download run at PCIe rate). */
int b200_host_alloc(size_t bytes, void** out);
int b200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* B200SV_H */
