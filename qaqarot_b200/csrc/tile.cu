// tile.cu -- the fused tile-resident multi-gate pass (B200_OP_TILE), sm_100a.
//
// One pass = one read and one write of the state (32 * 2^n bytes, the HBM-bound unit of SURVEY.md 8(d)), with
// as many gates as the planner could legally bring onto the tile applied on chip in between.
//
// Data layout.  A pass owns A index bits `tile[0..A)` (ascending physical positions; tile[i] = i for i < c so
// that 2^c amplitudes = 2^c * 16 B are contiguous in HBM).  Tile number t in [0, 2^(n-A)) has its bits spread
// over the remaining index positions.
//
// Data movement (round 2; tools/tma_probe.cu and profiles/r02_tma_probe.json are the measurements behind it).
// One persistent CTA per SM: two groups of 2^(A-4) compute threads and a ring of three tile buffers in shared
// memory.  The state is described to the TMA unit as a rank-5 tensor of doubles whose
// dimensions are the runs of tile bits (each followed by the non-tile bits above it, see build_tile_map), so a
// whole tile is ONE cp.async.bulk.tensor per direction (row-granular cp.async.bulk copies retire at one per ~31
// cycles per SM whatever their size and cannot carry 256-byte rows).  The groups take alternate tiles (tile j of
// the CTA: group j % 2, buffer j % 3), so one computes while the other waits or stores and a third tile is always
// in flight: a finished tile leaves through a TMA store issued by warp 0 of its group, and as soon as that store
// has read the buffer the same warp requests the tile that uses the buffer next (j + 3, for the OTHER group).
// There is no producer warp: 2 x 256 threads leave 128 registers per thread (a 17th warp rounds the allocation up
// to 640 threads and 96 registers).  mbarriers: full[stage][use & 1] with the TMA transaction count -- two per
// stage because the groups use a stage alternately and a parity wait is only valid for a waiter that consumed the
// barrier's previous phase itself (tools/tma_probe.cu has the story); no `empty` barriers are needed because the
// thread that refills a buffer is the one that waited for its store.
// Shared-memory slot of local index l is swz(l) = l ^ ((l >> 3) & 7), the TMA unit's SWIZZLE_128B pattern for
// 16-byte elements: GF(2)-linear, so swz(thread part | register part) = swz(thread) ^ swz(register), and a
// quarter warp's eight 16-byte accesses fall into eight different 16-byte bank groups whenever the thread bits
// mapped to lane bits 0..2 are one each of the local bit pairs (0,3), (1,4), (2,5) (the planner orders them so
// and keeps both bits of a pair out of one round's register set when it can).
//
// Rounds.  In a round every thread holds 2^R amplitudes in registers: the R *register bits* (local bit
// positions order[0..R)) enumerate them, the thread index supplies the other A-R local bits through
// order[R..A).  Non-diagonal gates act on register bits only, so a 2x2 gate is pure register arithmetic
// (complex128 FMA chains, target position resolved to a compile-time template by uniform bit tests);
// controls and phases may involve any bit: register bits become compile-time lane-invariant predicates,
// thread bits a per-thread predicate, bits outside the tile a per-tile (CTA-uniform) predicate.
//
// Dispatch (second half of round 2).  launch_tile_pass decodes the descriptor below ON THE HOST into PassConst -- round
// headers, one 16-byte record per gate record, the coefficient payload, the fan directory -- and passes it as a kernel
// parameter: the kernel reads nothing else about the pass (no descriptor upload, a program can be launched pass by
// pass).  Records and coefficients are read with uniform loads into uniform registers; see PassConst and the notes on
// the primitives for what that requires of the code.
//
// Descriptor (uint64 words, produced by qaqarot_b200/planner.py::encode_pass; `reals` = float64 pool):
//   D[0] = A | c<<8 | R<<16 | n_rounds<<24 | n_fans<<32 | flags<<40   D[1], D[2] = tile bit positions, 8 bits each
//          flags 1, 2 (first / last round straight from / to HBM, round 1): accepted and ignored -- with the TMA
//          pipeline a store from the register tile is slower than the TMA store (profiles/r02_tma_probe.json)
//   D[3] = fan directory offset | rounds offset << 32      D[4] = first payload real | payload length << 32
//          (everything the pass reads from `reals` is that one range; each CTA keeps a copy in shared memory)
//   fan directory: n_fans word offsets; fan record: n_entries | roff<<32, then n_entries masks.
//       reals[roff] = w0, then one complex per entry.  Per tile: fix[f] = w0 * prod{w_e : mask_e subset of index}
//   round: n_gates | n_words<<16, order[0..8), order[8..16), then 3 words per gate record:
//       G0 = target kind | j<<8 | diag kind<<16 | fan<<24 | thread_mask<<32 | reg_mask<<48,  G1 = fixed-bit mask,
//       G2 = target payload offset | diag payload offset << 32  (into reals)
//     A record runs where the fixed-bit mask (per tile) and the thread mask (per thread) are satisfied:
//     target stage on register bit j:  1 = complex 2x2 (8 reals, row major re/im), 2 = real 2x2 (same layout),
//                                      3 = pair swap; pairs whose other register bits satisfy reg_mask
//                                      6 = layer: for each register bit k in reg_mask (bits 0..3) the in-place
//                                          shears a0 += z a1; a1 += x a0; a0 += y a1 (skipped when y = 0) on the
//                                          pairs of bit k, payload 4 x (z, x, y, 0); j unused.  z = y = -tan(t/2),
//                                          x = sin t: the rotation by t.  z = -tan t, x = sin t cos t, y = 0, |t| <=
//                                          pi/4: the rotation by t up to cos t on every amplitude and 1 / cos^2 t
//                                          on those with bit k set, which the planner multiplies into the diag
//                                          stage of the same record (planner.shear_params); every 1-qubit unitary
//                                          is phase . rotation . phase (planner.split_rotation)
//     diag stage:  1 / 2 = fan  amp *= fix[fan] * lane_t[lane] * warp_t[warp] (* reg_t[rho] for kind 2) on the
//                  amplitudes with register bit j set (j = R: all); tables 32 + 2^(A-R-5) (+ 2^R) complex
//                  3 = phase  amp *= reals[offset] where rho satisfies reg_mask (never fused with a target stage)
//                  4 = fan without entries: amp *= reals[offset] on register bit j (a 1-qubit phase, no fan record)
//                  5 = fan whose entries all lie outside the tile: amp *= fix[fan], no tables
//                  6 = table: amp[rho] *= table[variant][rho] (2^m tables of 2^R complex, 17 entries apart), any diagonal on
//                      the register bits;
//                      variant = selector bits p1 | p2 << 1 | p3 << 2 with p1 = reg_mask >> 4 & 63, p2 = reg_mask >> 10
//                      & 63, p3 = fan byte; a selector < 48 is an index bit outside the tile, 48 + l the tile-local
//                      thread bit l, 63 none (m = selectors in use, packed first) -- fans and phases whose other
//                      qubits are not register bits ride along as variants
//     A 2x2 followed by a fan on the same register bit (the H + controlled-phase ladder of a QFT) is one record.
#include "common.cuh"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>

namespace b200 {

constexpr int kMaxFans = 64;
constexpr int T_NONE = 0, T_MATC = 1, T_MATR = 2, T_X = 3, T_LAYER = 6;
constexpr int D_NONE = 0, D_FANT = 1, D_FAN = 2, D_PHASE = 3, D_W = 4, D_FIX = 5, D_TAB = 6;

__host__ __device__ __forceinline__ uint32_t swz(uint32_t l) { return l ^ ((l >> 3) & 7u); }

// ---- mbarrier / TMA primitives (PTX) -----------------------------------------------------------------------------
// Everything that runs INSIDE the tile loop is a plain `asm` (not `asm volatile`) with a "memory" clobber and a dummy
// result (always 0) that the caller ORs into a token which ends up in a shared-memory address.  The clobber keeps the
// statement ordered against every load and store and forbids merging or hoisting it; the consumed result keeps it
// alive.  Why not `asm volatile`: with a volatile asm (or a __syncthreads / __syncwarp / __shfl intrinsic) anywhere
// inside a loop nest, NVVM + ptxas stop treating the loop counters of that nest as warp-uniform, and the record loop
// below then reads its records and coefficients with per-thread LDC into vector registers instead of LDCU into
// uniform registers (measured on tools/ur_probe.cu: 96 of 96 DFMA / DMUL take a UR operand without, 0 of 96 with).
__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(count));
}
__device__ __forceinline__ uint32_t mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  uint32_t d;
  asm("{\n\tmbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %2;\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : "r"(sptr(b)), "r"(bytes) : "memory");
  return d;
}
__device__ __forceinline__ uint32_t mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t d;
  asm("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : "r"(sptr(b)), "r"(parity) : "memory");
  return d;
}
__device__ __forceinline__ uint32_t tma_load5(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, const int (&c)[5]) {
  uint32_t d;
  asm("{\n\tcp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2, {%4, %5, %6, %7, %8}], [%3];\n\t"
      "mov.u32 %0, 0;\n\t}"
      : "=r"(d) : "r"(sptr(smem_dst)), "l"(tm), "r"(sptr(bar)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
  return d;
}
__device__ __forceinline__ uint32_t tma_store5(const CUtensorMap* tm, const void* smem_src, const int (&c)[5]) {
  uint32_t d;
  asm("{\n\tcp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%1, {%3, %4, %5, %6, %7}], [%2];\n\tmov.u32 %0, 0;\n\t}"
      : "=r"(d) : "l"(tm), "r"(sptr(smem_src)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
  return d;
}
__device__ __forceinline__ uint32_t bulk_commit() {
  uint32_t d;
  asm("{\n\tcp.async.bulk.commit_group;\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : : "memory");
  return d;
}
template <int N>
__device__ __forceinline__ uint32_t bulk_wait_read() {
  uint32_t d;
  asm("{\n\tcp.async.bulk.wait_group.read %1;\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : "n"(N) : "memory");
  return d;
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t fence_async() {
  uint32_t d;
  asm("{\n\tfence.proxy.async.shared::cta;\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : : "memory");
  return d;
}
// A uniform value on its way into per-thread arithmetic goes through this opaque move first.  Without it ptxas gives
// up the uniform register for the value AND for everything it was derived from (the loop counter it indexes with, the
// record it was read from ...): one shared-memory address computed from the round number was enough to turn the whole
// record loop back into per-thread code.
__device__ __forceinline__ uint32_t vreg(uint32_t x) {
  uint32_t y;
  asm("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
template <int NT>
__device__ __forceinline__ uint32_t group_sync(uint32_t g) {
  uint32_t d;
  asm("{\n\tbar.sync %1, %2;\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : "r"(1u + g), "n"(NT) : "memory");
  return d;
}

__host__ __device__ constexpr int ctz_c(int i) { return (i & 1) ? 0 : (i & 2) ? 1 : (i & 4) ? 2 : (i & 8) ? 3 : 4; }

__device__ __forceinline__ int byte_of(uint64_t lo, uint64_t hi, int i) {
  return (int)(((i < 8 ? lo : hi) >> (8 * (i & 7))) & 0xFF);
}

// ---- register-resident gate bodies (J = register bit, compile time) ---------------------------------------
// The pair / amplitude updates are written as inline PTX that reads and writes the *same* registers ("+d").
// Plain C++ makes NVVM emit each result into a fresh virtual register; at the join after every gate the whole
// 2^R-amplitude register tile is then copied back (64 moves per gate at R = 4 -- as many instructions as the
// arithmetic itself, seen as IMAD.MOV in the first ncu source view).  In-place PTX leaves nothing to copy.

// (a0, a1) <- (m00 a0 + m01 a1, m10 a0 + m11 a1), complex coefficients m = (re, im) as m[0..8)
__device__ __forceinline__ void pair_mat_c(c128& a0, c128& a1, const double (&m)[8], const double (&nm)[4]) {
  asm("{\n\t.reg .f64 t0, t1, t2, t3;\n\t"
      "mul.f64 t0, %4, %0;\n\t"            // m00r a0x
      "mul.f64 t1, %4, %1;\n\t"            // m00r a0y
      "mul.f64 t2, %8, %0;\n\t"            // m10r a0x
      "mul.f64 t3, %8, %1;\n\t"            // m10r a0y
      "fma.rn.f64 t0, %12, %1, t0;\n\t"    // - m00i a0y
      "fma.rn.f64 t1, %5, %0, t1;\n\t"     // + m00i a0x
      "fma.rn.f64 t2, %14, %1, t2;\n\t"    // - m10i a0y
      "fma.rn.f64 t3, %9, %0, t3;\n\t"     // + m10i a0x
      "fma.rn.f64 t0, %6, %2, t0;\n\t"     // + m01r a1x
      "fma.rn.f64 t1, %6, %3, t1;\n\t"     // + m01r a1y
      "fma.rn.f64 t2, %10, %2, t2;\n\t"    // + m11r a1x
      "fma.rn.f64 t3, %10, %3, t3;\n\t"    // + m11r a1y
      "fma.rn.f64 %0, %13, %3, t0;\n\t"    // - m01i a1y
      "fma.rn.f64 %1, %7, %2, t1;\n\t"     // + m01i a1x
      "fma.rn.f64 t2, %15, %3, t2;\n\t"    // - m11i a1y
      "fma.rn.f64 %3, %11, %2, t3;\n\t"    // + m11i a1x
      "mov.f64 %2, t2;\n\t}"
      : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y)
      : "d"(m[0]), "d"(m[1]), "d"(m[2]), "d"(m[3]), "d"(m[4]), "d"(m[5]), "d"(m[6]), "d"(m[7]),
        "d"(nm[0]), "d"(nm[1]), "d"(nm[2]), "d"(nm[3]));
}
// same with real coefficients
__device__ __forceinline__ void pair_mat_r(c128& a0, c128& a1, double m00, double m01, double m10, double m11) {
  asm("{\n\t.reg .f64 t0, t1, t2, t3;\n\t"
      "mul.f64 t0, %4, %0;\n\t"
      "mul.f64 t1, %4, %1;\n\t"
      "mul.f64 t2, %6, %0;\n\t"
      "mul.f64 t3, %6, %1;\n\t"
      "fma.rn.f64 %0, %5, %2, t0;\n\t"
      "fma.rn.f64 %1, %5, %3, t1;\n\t"
      "fma.rn.f64 %2, %7, %2, t2;\n\t"
      "fma.rn.f64 %3, %7, %3, t3;\n\t}"
      : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y)
      : "d"(m00), "d"(m01), "d"(m10), "d"(m11));
}
// a <- a * f   (nfy = -f.y).  Both products of a.y first, then a.y and a.x are overwritten in that order: four FP64
// instructions and no copy (the first version computed a.x into a temporary and moved it back: 2 more issue slots per
// amplitude, in records that are ~75 % issue-bound).
__device__ __forceinline__ void amp_mul(c128& a, double fx, double fy, double nfy) {
  asm("{\n\t.reg .f64 t0, t1;\n\t"
      "mul.f64 t0, %1, %4;\n\t"          // a.y * -f.y
      "mul.f64 t1, %1, %2;\n\t"          // a.y *  f.x
      "fma.rn.f64 %1, %0, %3, t1;\n\t"   // a.y = a.x f.y + a.y f.x
      "fma.rn.f64 %0, %0, %2, t0;\n\t}"  // a.x = a.x f.x - a.y f.y
      : "+d"(a.x), "+d"(a.y)
      : "d"(fx), "d"(fy), "d"(nfy));
}

// a0 <-> a1.  Also opaque in-place PTX: a C++ swap is a renaming in SSA form, and the register permutation it
// leaves at the join defeated the copy coalescing of the *whole* tile (2030 -> 491 moves without it).
__device__ __forceinline__ void pair_swap(c128& a0, c128& a1) {
  asm("{\n\t.reg .f64 t0, t1;\n\t"
      "mov.f64 t0, %0;\n\tmov.f64 t1, %1;\n\tmov.f64 %0, %2;\n\tmov.f64 %1, %3;\n\t"
      "mov.f64 %2, t0;\n\tmov.f64 %3, t1;\n\t}"
      : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y));
}

// COND = false: no control among the register bits (the common case) -- straight-line code, no predicates.
template <int R, int J, bool COND>
__device__ __forceinline__ void mat_c(c128 (&v)[1 << R], const double* __restrict__ mp, uint32_t rm) {
  double m[8], nm[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = mp[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) nm[i] = -m[2 * i + 1];
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    if (rho & (1 << J)) continue;
    if (!COND || (rho & rm) == rm) pair_mat_c(v[rho], v[rho | (1 << J)], m, nm);
  }
}
template <int R, int J, bool COND>
__device__ __forceinline__ void mat_r(c128 (&v)[1 << R], const double* __restrict__ m, uint32_t rm) {
  const double m00 = m[0], m01 = m[2], m10 = m[4], m11 = m[6];
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    if (rho & (1 << J)) continue;
    if (!COND || (rho & rm) == rm) pair_mat_r(v[rho], v[rho | (1 << J)], m00, m01, m10, m11);
  }
}
template <int R, int J, bool COND>
__device__ __forceinline__ void perm_x(c128 (&v)[1 << R], uint32_t rm) {
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    if (rho & (1 << J)) continue;
    if (!COND || (rho & rm) == rm) pair_swap(v[rho], v[rho | (1 << J)]);
  }
}

// Fan on the amplitudes whose register bit K is set (K = R: every amplitude): per-thread factor only ...
template <int R, int K>
__device__ __forceinline__ void fant_apply(c128 (&v)[1 << R], c128 wt) {
  const double nwy = -wt.y;
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    if (K < R && !((rho >> K) & 1)) continue;
    amp_mul(v[rho], wt.x, wt.y, nwy);
  }
}
// ... or with a further factor per register amplitude.
template <int R, int K>
__device__ __forceinline__ void fan_apply(c128 (&v)[1 << R], c128 wt, const double2* __restrict__ reg_t) {
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    if (K < R && !((rho >> K) & 1)) continue;
    const c128 f = cmul(wt, reg_t[rho]);
    amp_mul(v[rho], f.x, f.y, -f.y);
  }
}

// The two stages of a record are dispatched by two switches on keys computed once per CTA.  The compiler lowers
// each switch to a compare tree, and tools/record_cost.py shows that the tree is what a record costs before it does
// any arithmetic (~750 cycles per tile per CTA; one switch over all fused combinations was slower still), so the
// two ops that carry most of the work of a layered circuit are tested first and do several gates per dispatch:
//   LAYER  up to four 2x2 gates, one per register bit, each as three shears      (key1 = kKeyLayer)
//   TAB    any diagonal on the register bits: amp[rho] *= table[rho]              (key2 = kKeyTab)
//   key1 = 16 j + target kind + 8 (register-bit controls present)     (0: no target stage)
//   key2 = 8 j + diag kind, j = 4 for "every amplitude"               (0: no diag stage)
constexpr uint32_t kKeyLayer = 0xF0u, kKeyTab = 0xF1u;

// (a0, a1) <- E(y) F(x) E(z) (a0, a1) on the pairs of register bit J, for the J in `mask`; payload 4 x (z, x, y, -).
// The third update runs for the J in `mask3` only (y != 0, found once per CTA when the records are decoded: a test of
// the loaded y made every rotation wait for its payload before the first shear could issue).
template <int R>
__device__ __forceinline__ void layer_apply(c128 (&v)[1 << R], const double* __restrict__ pay, uint32_t mask, uint32_t mask3) {
#pragma unroll
  for (int J = 0; J < R; ++J) {
    if (!((mask >> J) & 1u)) continue;
    const double2 zx = reinterpret_cast<const double2*>(pay)[2 * J];
#pragma unroll
    for (int rho = 0; rho < (1 << R); ++rho) {
      if (rho & (1 << J)) continue;
      c128& a0 = v[rho];
      c128& a1 = v[rho | (1 << J)];
      asm("fma.rn.f64 %0, %4, %2, %0;\n\tfma.rn.f64 %1, %4, %3, %1;\n\t"
          "fma.rn.f64 %2, %5, %0, %2;\n\tfma.rn.f64 %3, %5, %1, %3;"
          : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(zx.x), "d"(zx.y));
    }
    if ((mask3 >> J) & 1u) {
      const double y = pay[4 * J + 2];
#pragma unroll
      for (int rho = 0; rho < (1 << R); ++rho) {
        if (rho & (1 << J)) continue;
        c128& a0 = v[rho];
        c128& a1 = v[rho | (1 << J)];
        asm("fma.rn.f64 %0, %4, %2, %0;\n\tfma.rn.f64 %1, %4, %3, %1;"
            : "+d"(a0.x), "+d"(a0.y) : "d"(a1.x), "d"(a1.y), "d"(y));
      }
    }
  }
}
// amp[rho] *= tab[rho]: the table of this thread's variant.  Variants lie kTabPitch = 17 entries apart, so every load
// has a compile-time offset from one base, and threads of a warp that picked different variants read different
// 16-byte bank groups ((variant + rho) mod 8).
constexpr uint32_t kTabPitch = 17;
template <int R>
__device__ __forceinline__ void tab_apply(c128 (&v)[1 << R], const double2* __restrict__ tab) {
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    const double2 f = tab[rho];
    amp_mul(v[rho], f.x, f.y, -f.y);
  }
}

// Dispatch on (register bit, kind) by BIT TESTS on uniform values: never a `switch`, and never a chain of four or more
// equality tests of one value either (the compiler turns that back into a switch) -- a jump table is an indirect branch
// through a per-thread register, and the key it is computed from then stops being a uniform value.
template <int R, int J>
__device__ __forceinline__ void run_target_j(c128 (&v)[1 << R], uint32_t kind, uint32_t rm, const double* __restrict__ pay) {
  // kind = T_MATC (1), T_MATR (2), T_X (3), + 8 with controls among the register bits
  if (kind & 8u) {
    if (!(kind & 2u)) mat_c<R, J, true>(v, pay, rm);
    else if (!(kind & 1u)) mat_r<R, J, true>(v, pay, rm);
    else perm_x<R, J, true>(v, rm);
  } else {
    if (!(kind & 2u)) mat_c<R, J, false>(v, pay, 0);
    else if (!(kind & 1u)) mat_r<R, J, false>(v, pay, 0);
    else perm_x<R, J, false>(v, 0);
  }
}
template <int R>
__device__ __forceinline__ void run_target(c128 (&v)[1 << R], uint32_t key1, uint32_t rm, const double* __restrict__ pay) {
  const uint32_t kind = key1 & 15u;
  if (key1 & 0x20u) {
    if (key1 & 0x10u) run_target_j<R, (3 < R ? 3 : 0)>(v, kind, rm, pay);
    else run_target_j<R, (2 < R ? 2 : 0)>(v, kind, rm, pay);
  } else {
    if (key1 & 0x10u) run_target_j<R, 1>(v, kind, rm, pay);
    else run_target_j<R, 0>(v, kind, rm, pay);
  }
}
// A fan (per-thread factor wt, optionally times a factor per register amplitude) on the amplitudes with register bit j
// set (j = 4: every amplitude)
template <int R>
__device__ __forceinline__ void run_fan(c128 (&v)[1 << R], uint32_t j, bool with_reg, c128 wt, const double2* __restrict__ reg_t) {
#define B200_F_CASE(J)                                            \
  {                                                               \
    if (with_reg) fan_apply<R, (J < R ? J : R)>(v, wt, reg_t);    \
    else fant_apply<R, (J < R ? J : R)>(v, wt);                   \
  }
  if (j & 4u) B200_F_CASE(4)
  else if (j & 2u) {
    if (j & 1u) B200_F_CASE(3) else B200_F_CASE(2)
  } else {
    if (j & 1u) B200_F_CASE(1) else B200_F_CASE(0)
  }
#undef B200_F_CASE
}
// The same for a factor that is the same for every thread (kind D_W: a 1-qubit phase), read from the constant bank
template <int R>
__device__ __forceinline__ void run_phase1(c128 (&v)[1 << R], uint32_t j, const double* __restrict__ pay) {
  const double wx = pay[0], wy = pay[1];
#define B200_F_CASE(J)                                            \
  {                                                               \
    _Pragma("unroll")                                             \
    for (int rho = 0; rho < (1 << R); ++rho)                      \
      if (J >= R || ((rho >> (J < R ? J : 0)) & 1)) amp_mul(v[rho], wx, wy, -wy);  \
  }
  if (j & 4u) B200_F_CASE(4)
  else if (j & 2u) {
    if (j & 1u) B200_F_CASE(3) else B200_F_CASE(2)
  } else {
    if (j & 1u) B200_F_CASE(1) else B200_F_CASE(0)
  }
#undef B200_F_CASE
}
// amp *= w on the amplitudes whose register bits satisfy rm (kind D_PHASE)
template <int R>
__device__ __forceinline__ void run_phase_mask(c128 (&v)[1 << R], uint32_t rm, const double* __restrict__ pay) {
  const double wx = pay[0], wy = pay[1];
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho)
    if (((uint32_t)rho & rm) == rm) amp_mul(v[rho], wx, wy, -wy);
}

// ---- pre-decoded pass (built once per CTA in shared memory) ------------------------------------------------------
// Decoding the planner's descriptor per gate, per thread, per tile cost ~60 instructions a gate in the first
// version (ncu source view: a quarter of all issue slots).  The gate list is the same for every tile, so each
// CTA decodes it once into 16-byte records and per-thread / per-round tables, and the tile loop only reads them.
constexpr int kMaxGates = 160;     // per pass
constexpr int kMaxRounds = 6;     // planner.PlanConfig.max_rounds
constexpr int kSpreadTabs = 5;     // 6 bits of the tile number each: up to 30 non-tile index bits
constexpr int kStages = 3;         // tile buffers in the ring
constexpr int kGroups = 2;         // compute groups taking alternate tiles
constexpr int kMaxR = 4;           // register bits per round the library is built for (2^R amplitudes per thread)

constexpr int kMaxPayload = 2560;  // float64 payload of one pass kept in shared memory (planner.MAX_PAYLOAD)
constexpr int kMaxFanWords = 512;  // fan directory + records of the pass's fans (planner keeps a pass below this)
// record (one uint4), three shapes told apart by x & 3 -- the two that carry a layered circuit or a QFT are decoded
// with a handful of instructions, everything else takes the general path:
//   1  LT   rotations + table          x = 1 | rotation mask<<2 | has table<<8      z = selectors p1 | p2<<8 | p3<<16
//   2  LF   rotations + fan            x = 2 | rotation mask<<2 | j<<6 (4: every amplitude) | diag kind<<9 | fan<<12
//   0  general (conditions, 2x2, pair swap, phase)
//                                      x = key1<<2 | key2<<10 | fan<<18 | fixed mask bits 32..39<<24
//                                      z = thread mask | reg mask<<16                w = fixed mask bits 0..31
//   all: y = target payload | diag payload<<16 (complex units, relative to the pass's payload)

// What a record needs in PER-THREAD arithmetic (shared-memory addresses of per-thread tables, thread-bit masks, table
// selectors) does not come from the uniform record: a uniform value that meets per-thread values in an address or a
// shift makes the compiler drop the uniform register for it and for everything upstream (the record, the loop counter).
// Those fields travel in a second stream of 32-bit words in shared memory, read through a per-thread cursor that runs
// in step with the uniform record loop:
// (4-word items start on a 16-byte boundary; the stream is padded in front of them)
//   round:   4 words  srb01, srb23 (slot masks of the register bits, 16 bits each), byte offset of the round's `thr`
//            row in TileSmem, 0
//   LT record with table variants:  1 word   p1 | p2 << 6 | p3 << 12 | table offset (complex units) << 18
//   LF record (fan kinds 1, 2, 5):  1 word   fan | table offset (complex units) << 8
//   general record:                 4 words  thread mask | reg mask << 16, fixed mask bits 0..31,
//                                            fixed mask bits 32..39 | fan << 8, diag payload offset (complex units)
constexpr int kMaxVsWords = 8 * kMaxRounds + 4 * kMaxGates + 8;

struct RoundConst {                // 32 bytes; 32-bit fields read as one uint4 (16-bit parameter loads are not uniform loads)
  uint32_t ng_first;               // n_gates | first << 16: records [first, first + n_gates)
  uint32_t srb01, srb23;           // slot XOR masks of the register bits, swz(1 << order[k]), 16 bits each
  uint32_t pad;
  uint64_t o_lo, o_hi;             // order[0..16) of the round, 8 bits each
};

// Everything about the pass that is the same for every thread: a KERNEL PARAMETER (constant bank), decoded on the host
// by launch_tile_pass.  The record loop and the gate payload are then read with uniform loads (LDCU) into uniform
// registers: the dispatch runs on the uniform datapath with uniform branches (no BSSY / BSYNC pairs, no per-thread
// decode), and rotation coefficients and table entries are UR operands of the DFMA / DMUL instructions -- no LDS, no
// vector registers, no address arithmetic (round 2; before, half of a layer pass's warp time sat on shared-memory
// scoreboards and dispatch branches, profiles/r02c_layers30_ncu_full.md).
struct PassConst {
  uint32_t n_rounds, n_fans, dir_off, rounds_off;
  uint32_t pay_lo, pay_len, n_vs, pad1;
  RoundConst rh[kMaxRounds];
  uint4 rec[kMaxGates + 1];
  uint64_t tb_lo, tb_hi;           // tile bit positions, 8 bits each (descriptor words 1, 2)
  uint32_t vs[kMaxVsWords];        // the per-thread side of the records (copied to shared memory), see below
  uint64_t fanw[kMaxFanWords];     // fan directory and records (descriptor words [dir_off, rounds_off))
  double pay[kMaxPayload];         // the pass's matrices, phases, tables (the same range of `reals`)
};
static_assert(sizeof(PassConst) + 512 <= 32764, "kernel parameters are limited to 32 764 bytes");

// How a tile maps to the tensor the TMA unit sees (kernel parameter, built by build_tile_map).
struct TileMapDev {
  int32_t shift[5];                // index bit where field i starts (field 0 starts at the re/im bit: shift[0] = 0)
  int32_t width[5];                // index bits the field spans (0: unused dimension)
  int32_t n_sub_bits;              // top local bits enumerated by separate copies (more than 5 runs of tile bits)
  int32_t sub_pos[9];              // ... their index positions
  uint32_t sub_amps;               // amplitudes per copy
};

template <int A, int RB>
struct TileSmem {
  c128 stage[kStages][1 << A];     // 1024-byte aligned (SWIZZLE_128B atoms): first member of the dynamic allocation
  c128 fix[kGroups][kMaxFans];     // per-tile fan products, one set per compute group
  uint64_t spread[kSpreadTabs][64];
  double pay[kMaxPayload];         // copy of PassConst::pay for the reads indexed per thread (fan tables, table variants)
  uint64_t fanw[kMaxFanWords];     // fan directory and records (entry counts, payload offsets, fixed-bit masks)
  alignas(16) uint32_t vs[kMaxVsWords];   // copy of PassConst::vs
  // per round: the local index bits a thread contributes, as (lane part)[32] | (warp part)[up to 32]
  uint16_t thr[kMaxRounds][64];
  uint64_t full[kStages][2];       // the tile's TMA load has landed; index = use of the stage & 1
};

// ---- the kernel ------------------------------------------------------------------------------------------------
// FULL = false leaves out the general target stages (complex / real 2x2, pair swap, their controlled forms): a pass
// of 1-qubit unitaries and phases needs none of them, and the kernel without their 24 unrolled bodies runs every
// record faster (round 1: 45.2 -> 41.0 ms on the 30-qubit QFT).
template <int A, int RB, bool FULL, bool TIMING>
__global__ void __launch_bounds__(kGroups * (1 << (A - RB)), 1)
k_tile(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileMapDev tm, const __grid_constant__ PassConst P,
       uint64_t n_tiles, uint64_t index_offset, uint32_t flags, uint64_t init_tile, uint32_t init_local,
       unsigned long long* __restrict__ phase_cycles) {
  constexpr int R = RB;            // register bits per round (2^R amplitudes per thread)
  constexpr int NT = 1 << (A - R), PER = 1 << R, NTB = A - R;
  constexpr int NWARP_T = (NTB > 5) ? (1 << (NTB - 5)) : 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TileSmem<A, RB>& S = *reinterpret_cast<TileSmem<A, RB>*>(smem_raw);

  uint32_t tid;                    // (volatile: ptxas otherwise re-reads SR_TID.X inside the record loop)
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const uint64_t tb_lo = P.tb_lo, tb_hi = P.tb_hi;
  const int n_rounds = (int)P.n_rounds, n_fans = (int)P.n_fans;
  const uint32_t dir_off = P.dir_off, rounds_off = P.rounds_off, pay_lo = P.pay_lo, pay_len = P.pay_len;
  // flags bit 1: the state is a basis vector that nobody has written yet -- every tile is all zeros except
  // tile `init_tile`, whose local index `init_local` holds 1 (B200_OP_INIT folded into this pass)
  const bool init = (flags & 2u) != 0;
  const int n_sub = 1 << tm.n_sub_bits;

  // ---- one-time setup ---------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&S.full[s][0], 1);
      mbar_init(&S.full[s][1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // (the pass arrives entirely as kernel parameters: nothing is uploaded, and a program can be launched pass by pass
  // while the host is still planning the next one)
  for (uint32_t i = tid; i < pay_len; i += blockDim.x) S.pay[i] = P.pay[i];
  for (uint32_t i = tid; i < rounds_off - dir_off; i += blockDim.x) S.fanw[i] = P.fanw[i];
  for (uint32_t i = tid; i < P.n_vs; i += blockDim.x) S.vs[i] = P.vs[i];
  for (uint32_t idx = tid; idx < kSpreadTabs * 64; idx += blockDim.x) {
    uint64_t x = (uint64_t)(idx & 63u) << (6 * (idx >> 6));
    for (int i = 0; i < A; ++i) x = insert_zero(x, byte_of(tb_lo, tb_hi, i));
    S.spread[idx >> 6][idx & 63u] = x;
  }
  for (int rd = 0; rd < n_rounds; ++rd) {
    const uint64_t o_lo = P.rh[rd].o_lo, o_hi = P.rh[rd].o_hi;
    for (uint32_t t = tid; t < 64u; t += blockDim.x) {        // t < 32: thread bits 0..4 (lane), else bits 5.. (warp)
      const uint32_t bits = t < 32u ? t : (t - 32u) << 5;
      uint32_t lt = 0;
      for (int b = 0; b < NTB; ++b) lt |= ((bits >> b) & 1u) << byte_of(o_lo, o_hi, R + b);
      S.thr[rd][t] = (uint16_t)lt;
    }
  }
  __syncthreads();

  auto spread_tile = [&](uint64_t t) -> uint64_t {
    uint64_t b = S.spread[0][t & 63u];
#pragma unroll
    for (int j = 1; j < kSpreadTabs; ++j) b |= S.spread[j][(t >> (6 * j)) & 63u];
    return b;
  };
  // tensor coordinates (in doubles) of sub-copy `sub` of the tile whose non-tile index bits are `base`
  auto coords = [&](uint64_t base, int sub, int (&c)[5]) {
    for (int k = 0; k < tm.n_sub_bits; ++k)
      if ((sub >> k) & 1) base |= 1ull << tm.sub_pos[k];
    c[0] = (int)((base & ((1ull << tm.width[0]) - 1ull)) << 1);
#pragma unroll
    for (int i = 1; i < 5; ++i) c[i] = (int)((base >> tm.shift[i]) & ((1ull << tm.width[i]) - 1ull));
  };

  // ---- compute groups ----------------------------------------------------------------------------------------
  // The group number is the same for all threads of a warp; taken through a warp reduction it is a value the compiler
  // KNOWS to be uniform (REDUX writes a uniform register), so the tile loop, the round loop and the record loop below
  // are uniform control flow and may use the uniform datapath.
  // (gv: the same number as a per-thread value, for addresses)
  const uint32_t gv = tid / NT, g = __reduce_max_sync(0xFFFFFFFFu, gv), t = tid % NT, lane = t & 31u, warp = t >> 5;
  const uint32_t lane_off = lane << 4, warp_off = (32u + (NWARP_T > 1 ? warp : 0u)) << 4;      // into a fan's tables
  c128* const fixp = S.fix[gv];
  const unsigned char* const vsb = reinterpret_cast<const unsigned char*>(S.vs);
  const unsigned char* const thrb_l = reinterpret_cast<const unsigned char*>(&S.thr[0][lane]);
  const unsigned char* const thrb_w = reinterpret_cast<const unsigned char*>(&S.thr[0][32u + warp]);
  unsigned char* const pay_b = reinterpret_cast<unsigned char*>(S.pay);
  // The sub-copies of a tile are issued by DIFFERENT warps of its group: a TMA instruction is a uniform-datapath
  // op, so the lanes of one warp that issue one each are served one after the other (~110 cycles per copy: with all
  // 32 copies of a scattered tile on warp 0, issuing cost the group 3-7 000 cycles per tile while its other warps
  // waited at the barrier).  Thread (warp w, lane l) takes sub-copies l * NW + w, + NT, ...
  constexpr uint32_t NW = NT / 32;
  const int my_sub = (int)(lane * NW + warp);
  uint32_t tok = 0;                // ORs the (zero) results of the synchronisation statements; see the primitives above
  // thread 0 of a group announces tile j of this CTA on its barrier ...
  auto expect = [&](uint32_t j) {
    if (blockIdx.x + (uint64_t)j * gridDim.x < n_tiles) tok |= mbar_expect_tx(&S.full[j % kStages][(j / kStages) & 1u], 16u << A);
  };
  // ... and, after a barrier of the group, every thread requests its sub-copies into the tile's buffer
  auto request = [&](uint32_t j) {
    const uint64_t tile = blockIdx.x + (uint64_t)j * gridDim.x;
    if (tile >= n_tiles || my_sub >= n_sub) return;
    const uint32_t s = j % kStages, u = j / kStages;
    uint64_t* const fb = &S.full[s][u & 1u];
    const uint64_t base = spread_tile(tile);
    for (int sub = my_sub; sub < n_sub; sub += NT) {
      int c[5];
      coords(base, sub, c);
      tok |= tma_load5(S.stage[s] + (size_t)sub * tm.sub_amps, &tmap, fb, c);
    }
  };
  auto store_tile = [&](const c128* buf, uint64_t base) {
    for (int sub = my_sub; sub < n_sub; sub += NT) {
      int c[5];
      coords(base, sub, c);
      tok |= tma_store5(&tmap, buf + (size_t)sub * tm.sub_amps, c);
    }
    tok |= bulk_commit();
  };
  if (!init && g == 0) {            // prologue: the first three tiles
    if (t == 0)
      for (uint32_t j = 0; j < (uint32_t)kStages; ++j) expect(j);
    tok |= group_sync<NT>(0);
    for (uint32_t j = 0; j < (uint32_t)kStages; ++j) request(j);
  }
  if (init) {
    // A pass that starts from a basis state: every tile but one is all zeros and stays all zeros (the pass is
    // linear), so those tiles are TMA stores from a buffer of zeros that is written once -- the pass moves 16 * 2^n
    // bytes and computes one tile.
#pragma unroll
    for (int i = 0; i < PER; ++i) S.stage[g][t + (uint32_t)i * NT] = make_double2(0.0, 0.0);
    tok |= fence_async();
    tok |= group_sync<NT>(g);
  }
  bool store_pending = false;       // this group's previous tile has been handed to the TMA unit, not yet waited for
  // B200_TILE_TIMING=1: thread 0 of every group accumulates the cycles it spends in each phase of a tile
  long long tk0 = TIMING ? clock64() : 0;
  auto tick = [&](int phase) {
    if (TIMING && t == 0) {
      const long long now = clock64();
      atomicAdd(phase_cycles + phase, (unsigned long long)(now - tk0));
      tk0 = now;
    }
  };
  for (uint32_t j = g;; j += kGroups) {
    const uint64_t tile = blockIdx.x + (uint64_t)j * gridDim.x;
    if (tile >= n_tiles) break;
    if (init && tile != init_tile) {
      store_tile(S.stage[g], spread_tile(tile));
      tok |= bulk_wait_read<4>();
      tick(8);
      continue;
    }
    // the one computed tile of an initial pass has the third buffer to itself
    const uint32_t s = init ? 2u : j % kStages, u = j / kStages;
    const uint64_t fixedidx = spread_tile(tile) | index_offset;
    const uint32_t fx_lo = (uint32_t)fixedidx, fx_hi = (uint32_t)(fixedidx >> 32);
    // per-tile fixed-bit products of the fans, while the tile is still in flight: one thread per fan (no warp shuffles
    // inside the tile loop, see the note on the primitives; a pass has at most 64 fans of ~20 entries)
    for (int f = (int)t; f < n_fans; f += NT) {
      const uint32_t off = (uint32_t)S.fanw[f] - dir_off;
      const uint64_t rec = S.fanw[off];
      const int ne = (int)(rec & 0xFFFFFFFFu);
      const double2* __restrict__ w = reinterpret_cast<const double2*>(S.pay + ((uint32_t)(rec >> 32) - pay_lo));
      c128 p = w[0];
      for (int e = 0; e < ne; ++e) {
        const uint64_t m = S.fanw[off + 1 + e];
        if ((fixedidx & m) == m) p = cmul(p, w[1 + e]);
      }
      fixp[f] = p;
    }
    tick(0);                        // fan products
    // announce the tile that will use the previous tile's buffer next (j + 1, for the OTHER group); its sub-copies are
    // requested below, thread by thread
    if (store_pending && t == 0 && !init) expect(j - kGroups + kStages);
    tick(1);
    // (this barrier also keeps the threads of a group within one tile of each other: a parity wait cannot tell phase
    // u from u + 2)
    tok |= group_sync<NT>(g);
    tick(2);                        // group barrier
    unsigned char* smb = reinterpret_cast<unsigned char*>(S.stage[s]);
    if (init) {
#pragma unroll
      for (int i = 0; i < PER; ++i) *reinterpret_cast<c128*>(smb + ((t + (uint32_t)i * NT) << 4)) = make_double2(0.0, 0.0);
      tok |= group_sync<NT>(g);
      if (t == 0) S.stage[s][swz(init_local)] = make_double2(1.0, 0.0);
      tok |= group_sync<NT>(g);
    } else {
      tok |= mbar_wait(&S.full[s][u & 1u], (u >> 1) & 1u);
    }
    smb += tok;                     // (0: keeps the statements above alive and in front of the loads below)
    tick(3);                        // wait for the tile
    c128 v[PER];
    uint32_t vc = 0;                // cursor into the per-thread stream (bytes)
    for (int rd = 0; rd < n_rounds; ++rd) {
      const uint32_t ngf = P.rh[rd].ng_first;
      const int ng = (int)(ngf & 0xFFFFu), first = (int)(ngf >> 16);
      // the per-thread side of the round: slot masks of the register bits, this thread's local index bits
      vc = (vc + 15u) & ~15u;       // (4-word items start on 16 bytes)
      const uint4 vh = *reinterpret_cast<const uint4*>(vsb + vc);
      vc += 16u;
      const uint32_t srb[5] = {(vh.x & 0xFFFFu) << 4, (vh.x >> 16) << 4, (vh.y & 0xFFFFu) << 4, (vh.y >> 16) << 4, 0u};
      const uint32_t lt = (uint32_t)*reinterpret_cast<const uint16_t*>(thrb_l + vh.z) |
                          (uint32_t)*reinterpret_cast<const uint16_t*>(thrb_w + vh.z);
      const uint32_t stb = swz(lt) << 4;

      // Gray-code walk over the 2^R register amplitudes: one XOR per shared-memory address
      {
        uint32_t so = stb;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          if (i) so ^= srb[ctz_c(i)];
          v[i ^ (i >> 1)] = *reinterpret_cast<const c128*>(smb + so);
        }
      }
      // Sub-copy i of a tile is stored and loaded by the same thread and occupies the same part of the buffer, so a
      // thread may refill its part as soon as ITS store has read it -- no barrier.  (Draining 64 KB takes the TMA unit
      // ~4 000 cycles, so there is still a wait here; leaving the refill until later in the round -- a test inside the
      // record loop -- cost more in the loop than the wait: 208 against 183 ms on the 30-qubit layer circuit.)
      if (store_pending) {
        tok |= bulk_wait_read<0>();
        if (!init) request(j - kGroups + kStages);
        store_pending = false;
      }
      tick(4);                      // round: registers <- shared memory (issue), refill of the previous buffer
      // The records and their payload are kernel parameters: `q`, the payload offsets and every coefficient that does
      // not depend on the thread are uniform-register values (see PassConst).
      uint4 q = P.rec[first];
      for (int gi = first; gi < first + ng; ++gi) {
        const uint4 nq = P.rec[gi + 1];             // (the next record is on its way while this one runs)
        const uint32_t op = q.x & 3u;
        const uint32_t off_t = (q.y & 0xFFFFu) * 2u, off_d = (q.y >> 16) * 2u;      // in doubles, into P.pay / S.pay
        const double* __restrict__ pay_t = P.pay + off_t;
        const double* __restrict__ pay_d = P.pay + off_d;
        if (op != 0u) {
          const uint32_t rot = (q.x >> 2) & 15u;
          if (rot) layer_apply<R>(v, pay_t, rot, (q.x >> 20) & 15u);
          if (op == 1u) {
            if (q.x & 0x100u) {
              if (!(q.x & 0x200u)) {
                tab_apply<R>(v, reinterpret_cast<const double2*>(pay_d));           // one table for everybody: UR operands
              } else {
                // one table per combination of (up to) three selector bits: index bits outside the tile (constant per
                // tile) or thread bits of the tile (48 + local position); 63 = none, which reads as 0
                const uint32_t w = *reinterpret_cast<const uint32_t*>(vsb + vc);
                vc += 4u;
                const uint32_t p1 = w & 63u, p2 = (w >> 6) & 63u, p3 = (w >> 12) & 63u;
                const uint64_t sel = fixedidx | ((uint64_t)lt << 48);
                const uint32_t var = (uint32_t)((sel >> p1) & 1ull) | ((uint32_t)((sel >> p2) & 1ull) << 1) |
                                     ((uint32_t)((sel >> p3) & 1ull) << 2);
                tab_apply<R>(v, reinterpret_cast<const double2*>(pay_b + ((w >> 18) << 4)) + var * kTabPitch);
              }
            }
          } else {
            const uint32_t dk = (q.x >> 9) & 7u, jf = (q.x >> 6) & 7u;
            if (dk == D_W) {
              run_phase1<R>(v, jf, pay_d);                               // the factor itself, the same for every thread
            } else {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(vsb + vc);
              vc += 4u;
              c128 wt = fixp[w & 63u];
              if (!(dk & 4u)) {                                          // (not D_FIX: thread tables)
                const unsigned char* const tb = pay_b + ((w >> 8) << 4);
                wt = cmul(wt, cmul(*reinterpret_cast<const double2*>(tb + lane_off),
                                   *reinterpret_cast<const double2*>(tb + warp_off)));
              }
              run_fan<R>(v, jf, (dk & 2u) != 0u && !(dk & 4u), wt, reinterpret_cast<const double2*>(pay_d) + 32 + NWARP_T);
            }
          }
        } else {
          vc = (vc + 15u) & ~15u;
          const uint4 w = *reinterpret_cast<const uint4*>(vsb + vc);
          vc += 16u;
          const uint32_t qz = w.x, qw = w.y, fmh = w.z & 0xFFu, fanv = w.z >> 8;
          // rm: the register-bit mask as a uniform value (branches) and rmv as a per-thread one (predicates: a few
          // instructions predicated by a uniform condition cost the uniform registers just like the jump tables)
          const uint32_t lm = qz & 0xFFFFu, rm = q.z >> 16, rmv = qz >> 16;
          // fixed-bit condition: constant per tile; thread-bit condition: per thread
          const bool active = ((fx_lo & qw) == qw) && ((fx_hi & fmh) == fmh) && ((lt & lm) == lm);
          if (active) {
            const uint32_t key1 = (q.x >> 2) & 0xFFu, key2 = (q.x >> 10) & 0xFFu;
            if (key1 == kKeyLayer) {
              uint32_t m3 = 0;
#pragma unroll
              for (int b = 0; b < 4; ++b) m3 |= (pay_t[4 * b + 2] != 0.0 ? 1u : 0u) << b;
              layer_apply<R>(v, pay_t, rm & 15u, m3);
            }
            else if (FULL && key1 != 0u) run_target<R>(v, key1, rmv, pay_t);
            if (key2 == kKeyTab) {
              const uint32_t p1 = (rmv >> 4) & 63u, p2 = (rmv >> 10) & 63u, p3 = fanv;
              const uint64_t sel = fixedidx | ((uint64_t)lt << 48);
              const uint32_t var = (uint32_t)((sel >> p1) & 1ull) | ((uint32_t)((sel >> p2) & 1ull) << 1) |
                                   ((uint32_t)((sel >> p3) & 1ull) << 2);
              tab_apply<R>(v, reinterpret_cast<const double2*>(pay_b + (w.w << 4)) + var * kTabPitch);
            } else if (key2 != 0u) {
              // kinds: fan with thread tables 1, + register table 2, phase under a register mask 3, plain factor 4, fan
              // without tables 5 -- told apart by bit tests (see run_target)
              const uint32_t dk = key2 & 7u, jd = key2 >> 3;
              if (dk & 4u) {
                if (dk & 1u) run_fan<R>(v, jd, false, fixp[fanv], nullptr);
                else run_phase1<R>(v, jd, pay_d);
              } else if ((dk & 1u) && (dk & 2u)) {
                run_phase_mask<R>(v, rmv, pay_d);
              } else {
                const unsigned char* const tb = pay_b + (w.w << 4);
                const c128 wt = cmul(fixp[fanv], cmul(*reinterpret_cast<const double2*>(tb + lane_off),
                                                      *reinterpret_cast<const double2*>(tb + warp_off)));
                run_fan<R>(v, jd, (dk & 2u) != 0u, wt, reinterpret_cast<const double2*>(pay_d) + 32 + NWARP_T);
              }
            }
          }
        }
        q = nq;
      }
      tick(5);                      // round: gates
      {
        // the 16 slot addresses are derived again (kept live across the gates they cost 16 registers of the 128)
        uint32_t lt2;
        asm("mov.u32 %0, %1;" : "=r"(lt2) : "r"(lt));
        uint32_t so = swz(lt2) << 4;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          if (i) so ^= srb[ctz_c(i)];
          *reinterpret_cast<c128*>(smb + so) = v[i ^ (i >> 1)];
        }
      }
      if (rd == n_rounds - 1) tok |= fence_async();      // the TMA store below reads what this thread just wrote
      tick(6);                      // round: shared memory <- registers
      tok |= group_sync<NT>(g);
      smb += tok;
      tick(7);                      // round: barrier
    }
    // the finished tile leaves through the TMA unit (every warp of the group issues its share of the sub-copies)
    store_tile(S.stage[s], fixedidx ^ index_offset);
    store_pending = true;
    tick(8);                        // issue the stores
  }
  bulk_wait_all();
  if (tok != 0u && phase_cycles != nullptr) phase_cycles[15] = tok;      // never true: tok is 0 (keeps the last statements alive)
}

// ---- host side: the tile as a tensor box -------------------------------------------------------------------------
// Field i of the tensor covers index bits [shift_i, shift_(i+1)): its low bits are tile bits (the box), the rest is
// fixed per tile (the coordinate).  Field 0 is (re/im, index bits 0..2) = 16 doubles = one 128-byte swizzle line and
// extends over the non-tile bits up to the next run.  Runs of tile bits longer than 8 are split (box dimensions are
// limited to 256 elements).  More than 5 fields: the top runs are enumerated by separate copies, each of which lands
// on a contiguous, 1 KB aligned part of the shared-memory tile.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

static int build_tile_map(b200_state* st, const int* tile, int a, TileMapDev& tm, CUtensorMap& map) {
  const int n = st->n;
  B200_REQUIRE(a >= 4 && tile[0] == 0 && tile[1] == 1 && tile[2] == 2, "TILE descriptor: index bits 0..2 must be tile bits");
  int starts[16], lens[16], nch = 0;
  for (int i = 3; i < a;) {
    int j = i + 1;
    while (j < a && tile[j] == tile[j - 1] + 1 && j - i < 8) ++j;
    starts[nch] = tile[i];
    lens[nch++] = j - i;
    i = j;
  }
  const int keep = nch < 4 ? nch : 4;
  tm = TileMapDev{};
  for (int k = keep; k < nch; ++k)
    for (int b = 0; b < lens[k]; ++b) {
      B200_REQUIRE(tm.n_sub_bits < 9, "TILE descriptor: tile bits too scattered for the tensor map");
      tm.sub_pos[tm.n_sub_bits++] = starts[k] + b;
    }
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  int shift[6];
  shift[0] = 0;
  box[0] = 16;
  for (int k = 0; k < keep; ++k) {
    shift[1 + k] = starts[k];
    box[1 + k] = 1u << lens[k];
  }
  const int rank = 1 + keep;
  shift[rank] = n;
  uint32_t box_doubles = 1;
  for (int i = 0; i < 5; ++i) {
    if (i < rank) {
      tm.shift[i] = shift[i];
      tm.width[i] = shift[i + 1] - shift[i];
      // in doubles: field 0 also holds the re/im bit
      gdim[i] = 1ull << (tm.width[i] + (i == 0 ? 1 : 0));
      if (i > 0) gstride[i - 1] = 16ull << shift[i];
    } else {
      tm.shift[i] = 0;
      tm.width[i] = 0;
      gdim[i] = 1;
      box[i] = 1;
      gstride[i - 1] = 16ull << n;
    }
    box_doubles *= box[i];
  }
  tm.sub_amps = box_doubles / 2;
  EncodeTiledFn enc = encode_tiled();
  B200_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, st->amp, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d) for a 2^%d tile on %d qubits", (int)r, a, n);
    set_error(msg);
    return 1;
  }
  return 0;
}

static_assert(sizeof(TileSmem<12, 4>) <= 232448,
              "the 2^12 tile's shared memory exceeds the 227 KB a CTA can have");

template <int A, int RB, bool FULL>
static int launch(b200_state* st, const PassConst& pc, const int* tile, bool init,
                  uint64_t init_tile, uint32_t init_local) {
  constexpr int NT = 1 << (A - RB);
  const size_t smem = sizeof(TileSmem<A, RB>);
  static bool configured[64] = {};
  if (!configured[st->device & 63]) {
    B200_CUDA(cudaFuncSetAttribute(k_tile<A, RB, FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(k_tile<A, RB, FULL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[st->device & 63] = true;
  }
  TileMapDev tm;
  CUtensorMap map;
  if (build_tile_map(st, tile, A, tm, map)) return 1;
  const uint64_t n_tiles = st->dim >> A;
  // persistent: one CTA per SM, each walking tiles blockIdx, blockIdx + grid, ...
  const unsigned grid = (unsigned)(n_tiles < (uint64_t)kSMs ? n_tiles : (uint64_t)kSMs);
  static const bool timing = getenv("B200_TILE_TIMING") != nullptr;
  if (!timing) {
    k_tile<A, RB, FULL, false><<<grid, kGroups * NT, smem, st->stream>>>(map, tm, pc, n_tiles, st->index_offset,
                                                                     init ? 2u : 0u, init_tile, init_local, nullptr);
    st->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  constexpr int kPhases = 9;
  unsigned long long* d_cyc = nullptr;
  B200_CUDA(cudaMalloc(&d_cyc, kPhases * sizeof(unsigned long long)));
  B200_CUDA(cudaMemsetAsync(d_cyc, 0, kPhases * sizeof(unsigned long long), st->stream));
  k_tile<A, RB, FULL, true><<<grid, kGroups * NT, smem, st->stream>>>(map, tm, pc, n_tiles, st->index_offset,
                                                                  init ? 2u : 0u, init_tile, init_local, d_cyc);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  unsigned long long h[kPhases];
  B200_CUDA(cudaMemcpyAsync(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  B200_CUDA(cudaFree(d_cyc));
  double tot = 0;
  for (int i = 0; i < kPhases; ++i) tot += (double)h[i];
  fprintf(stderr, "[k_tile A=%d tiles=%llu ctas=%u copies/tile=%d] cycles per tile per group:", A, (unsigned long long)n_tiles,
          grid, 1 << tm.n_sub_bits);
  static const char* names[kPhases] = {"fan_products", "store_wait+request", "group_barrier", "wait_tile", "rd_load",
                                       "rd_gates", "rd_store", "rd_barrier", "store_issue"};
  for (int i = 0; i < kPhases; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)h[i] / (double)n_tiles);
  fprintf(stderr, " total=%.0f\n", tot / (double)n_tiles);
  return 0;
}

int launch_tile_pass(b200_state* st, const uint64_t* h_desc, int wlen, const double* h_reals, size_t n_reals,
                     uint64_t init_index) {
  B200_REQUIRE(wlen >= 5, "TILE descriptor too short");
  const uint64_t h0 = h_desc[0];
  const int A = (int)(h0 & 0xFF), c = (int)((h0 >> 8) & 0xFF), Rd = (int)((h0 >> 16) & 0xFF);
  const int n_rounds = (int)((h0 >> 24) & 0xFF), n_fans = (int)((h0 >> 32) & 0xFF);
  char msg[200];
  if (A > st->n || c > A || c < 3 || Rd != kMaxR || n_fans > kMaxFans || n_rounds > kMaxRounds || st->n - A > 6 * kSpreadTabs) {
    snprintf(msg, sizeof msg, "TILE descriptor not executable: A=%d c=%d R=%d fans=%d rounds=%d on %d qubits", A, c, Rd,
             n_fans, n_rounds, st->n);
    set_error(msg);
    return 1;
  }
  int tile[16];
  int prev = -1;
  for (int i = 0; i < A; ++i) {
    const int b = (int)(((i < 8 ? h_desc[1] : h_desc[2]) >> (8 * (i & 7))) & 0xFF);
    if (b <= prev || b >= st->n || (i < c && b != i)) {
      set_error("TILE descriptor: tile bits must be ascending, below n_qubits, and 0..c-1 first");
      return 1;
    }
    tile[i] = prev = b;
  }
  const uint64_t dir_off = h_desc[3] & 0xFFFFFFFFu, rounds_off = h_desc[3] >> 32;
  B200_REQUIRE(dir_off + (uint64_t)n_fans <= (uint64_t)wlen && rounds_off <= (uint64_t)wlen,
               "TILE descriptor: offsets out of range");
  B200_REQUIRE(rounds_off >= dir_off && rounds_off - dir_off <= (uint64_t)kMaxFanWords,
               "TILE descriptor: fan records too long");
  B200_REQUIRE((h_desc[4] >> 32) <= (uint64_t)kMaxPayload && (h_desc[4] & 0xFFFFFFFFu) + (h_desc[4] >> 32) <= n_reals,
               "TILE descriptor: payload range too long or outside the reals pool");
  // The pass as the kernel wants it (PassConst): round headers, one 16-byte record per gate record of the descriptor,
  // and the payload.  (The kernel decoded the descriptor itself, once per CTA, until the records moved into the constant
  // bank.)
  static thread_local PassConst pc;
  const uint32_t pay_lo = (uint32_t)(h_desc[4] & 0xFFFFFFFFu), pay_len = (uint32_t)(h_desc[4] >> 32);
  pc.n_rounds = (uint32_t)n_rounds;
  pc.n_fans = (uint32_t)n_fans;
  pc.dir_off = (uint32_t)dir_off;
  pc.rounds_off = (uint32_t)rounds_off;
  pc.pay_lo = pay_lo;
  pc.pay_len = pay_len;
  pc.pad1 = 0;
  for (uint32_t i = 0; i < pay_len; ++i) pc.pay[i] = h_reals[pay_lo + i];
  pc.tb_lo = h_desc[1];
  pc.tb_hi = h_desc[2];
  for (uint64_t i = dir_off; i < rounds_off; ++i) pc.fanw[i - dir_off] = h_desc[i];
  auto in_payload = [&](uint64_t off, uint32_t len) { return off >= pay_lo && off + len <= (uint64_t)pay_lo + pay_len; };
  uint64_t pos = rounds_off, gates = 0;
  uint32_t nvs = 0;               // words of the per-thread stream
  bool full = false;              // does any record need a general target stage (kinds 1..3)?
  for (int rd = 0; rd < n_rounds; ++rd) {
    B200_REQUIRE(pos + 3 <= (uint64_t)wlen, "TILE descriptor: round header out of range");
    const uint64_t ng = h_desc[pos] & 0xFFFF, nw = h_desc[pos] >> 16;
    B200_REQUIRE(nw >= 3 + 3 * ng && pos + nw <= (uint64_t)wlen, "TILE descriptor: round length out of range");
    B200_REQUIRE(gates + ng <= (uint64_t)kMaxGates, "TILE descriptor: too many gates");
    RoundConst& rc = pc.rh[rd];
    B200_REQUIRE(nvs + 8 <= (uint32_t)kMaxVsWords, "TILE descriptor: too many records");
    rc.ng_first = (uint32_t)ng | ((uint32_t)gates << 16);
    rc.pad = 0;
    rc.o_lo = h_desc[pos + 1];
    rc.o_hi = h_desc[pos + 2];
    uint32_t srb[4] = {0, 0, 0, 0};
    for (int k = 0; k < Rd; ++k) srb[k] = swz(1u << ((rc.o_lo >> (8 * k)) & 0xFF)) & 0xFFFFu;
    rc.srb01 = srb[0] | (srb[1] << 16);
    rc.srb23 = srb[2] | (srb[3] << 16);
    while (nvs & 3u) pc.vs[nvs++] = 0;                       // 4-word items start on 16 bytes
    pc.vs[nvs++] = rc.srb01;
    pc.vs[nvs++] = rc.srb23;
    pc.vs[nvs++] = (uint32_t)rd * 128u;                    // the round's row of TileSmem::thr, in bytes
    pc.vs[nvs++] = 0;
    for (uint64_t gi = 0; gi < ng; ++gi) {
      const uint64_t g0 = h_desc[pos + 3 + 3 * gi], fm = h_desc[pos + 4 + 3 * gi], roffs = h_desc[pos + 5 + 3 * gi];
      const uint32_t tk = (uint32_t)(g0 & 0xFF), j = (uint32_t)((g0 >> 8) & 0xFF), dk = (uint32_t)((g0 >> 16) & 0xFF);
      const uint32_t fan = (uint32_t)((g0 >> 24) & 0xFF);
      const uint32_t lm = (uint32_t)((g0 >> 32) & 0xFFFF), rm = (uint32_t)((g0 >> 48) & 0xFFFF);
      full |= tk == T_MATC || tk == T_MATR || tk == T_X;
      const bool has_t = !(tk == T_NONE || tk == T_X), has_d = dk != D_NONE;
      B200_REQUIRE(!has_t || in_payload(roffs & 0xFFFFFFFFu, tk == T_LAYER ? 16 : 8), "TILE descriptor: target payload outside the pass's range");
      B200_REQUIRE(!has_d || in_payload(roffs >> 32, 2), "TILE descriptor: diag payload outside the pass's range");
      const uint32_t pt = has_t ? (((uint32_t)roffs - pay_lo) >> 1) : 0u;
      const uint32_t pd = has_d ? (((uint32_t)(roffs >> 32) - pay_lo) >> 1) : 0u;
      const uint32_t jj = j < (uint32_t)Rd ? j : 4u;            // 4 = no register target ("every amplitude")
      const bool plain = lm == 0 && fm == 0 && (tk == T_NONE || tk == T_LAYER);
      uint32_t rot = 0, rot3 = 0;                                // rotations of a layer; those with a third update
      if (tk == T_LAYER) {
        rot = rm & 15u;
        for (uint32_t b = 0; b < 4u; ++b)
          if (((rot >> b) & 1u) && h_reals[(uint32_t)roffs + 4u * b + 2u] != 0.0) rot3 |= 1u << b;
      }
      uint4 rec;
      rec.y = pt | (pd << 16);
      B200_REQUIRE(nvs + 8 <= (uint32_t)kMaxVsWords, "TILE descriptor: too many records");
      if (plain && (dk == D_NONE || dk == D_TAB)) {
        const uint32_t p1 = (rm >> 4) & 63u, p2 = (rm >> 10) & 63u, p3 = fan & 63u;
        const bool variants = dk == D_TAB && !(p1 == 63u && p2 == 63u && p3 == 63u);
        rec.x = 1u | (rot << 2) | (dk == D_TAB ? 0x100u : 0u) | (variants ? 0x200u : 0u) | (rot3 << 20);
        rec.z = rec.w = 0;
        if (variants) pc.vs[nvs++] = p1 | (p2 << 6) | (p3 << 12) | (pd << 18);
      } else if (plain && (dk == D_FANT || dk == D_FAN || dk == D_W || dk == D_FIX)) {
        rec.x = 2u | (rot << 2) | (jj << 6) | (dk << 9) | ((fan & 63u) << 12) | (rot3 << 20);
        rec.z = rec.w = 0;
        if (dk != D_W) pc.vs[nvs++] = (fan & 63u) | (pd << 8);
      } else {
        while (nvs & 3u) pc.vs[nvs++] = 0;
        pc.vs[nvs++] = lm | (rm << 16);
        pc.vs[nvs++] = (uint32_t)fm;
        pc.vs[nvs++] = (uint32_t)(fm >> 32) | ((fan & 63u) << 8);
        pc.vs[nvs++] = pd;
        const uint32_t key1 = tk == T_NONE ? 0u : tk == T_LAYER ? kKeyLayer : 16u * jj + tk + (rm != 0 ? 8u : 0u);
        const uint32_t key2 = dk == D_NONE ? 0u : dk == D_TAB ? kKeyTab : 8u * jj + dk;
        rec.x = (key1 << 2) | (key2 << 10) | ((fan & 63u) << 18) | ((uint32_t)(fm >> 32) << 24);
        rec.z = lm | (rm << 16);
        rec.w = (uint32_t)fm;
      }
      pc.rec[gates + gi] = rec;
    }
    gates += ng;
    pos += nw;
  }
  B200_REQUIRE(pos == (uint64_t)wlen && gates <= (uint64_t)kMaxGates, "TILE descriptor: bad round lengths or too many gates");
  pc.rec[gates] = make_uint4(0, 0, 0, 0);
  pc.n_vs = nvs;
  const bool init = init_index != kNoInit;
  uint64_t init_tile = ~0ull;
  uint32_t init_local = 0;
  if (init && (init_index & ~(st->dim - 1)) == st->index_offset) {     // the 1 lives on this shard
    bool in_tile[64] = {};
    for (int i = 0; i < A; ++i) {
      in_tile[tile[i]] = true;
      init_local |= (uint32_t)((init_index >> tile[i]) & 1ull) << i;
    }
    init_tile = 0;
    for (int b = 0, k = 0; b < st->n; ++b)
      if (!in_tile[b]) init_tile |= ((init_index >> b) & 1ull) << k++;
  }
  // (k_tile<12, 3, ...>, 8 amplitudes per thread and 2 x 512 threads at 64 registers, was built and measured in round 2:
  // twice the warps but a third more rounds per pass -- 210 against 163 ms on the 30-qubit layer circuit, 26.7 against
  // 22.2 ms on the QFT.  The rounds' shared-memory traffic is what a pass pays for, not latency that more warps hide.)
  switch (A) {
    case 9:  return launch<9, 4, true>(st, pc, tile, init, init_tile, init_local);
    case 10: return launch<10, 4, true>(st, pc, tile, init, init_tile, init_local);
    case 11: return launch<11, 4, true>(st, pc, tile, init, init_tile, init_local);
    // the production tile size also exists without the general target stages (see k_tile)
    case 12: return full ? launch<12, 4, true>(st, pc, tile, init, init_tile, init_local)
                         : launch<12, 4, false>(st, pc, tile, init, init_tile, init_local);
    default:
      snprintf(msg, sizeof msg, "TILE descriptor: tile size 2^%d is not built (9..12)", A);
      set_error(msg);
      return 1;
  }
}

}  // namespace b200
