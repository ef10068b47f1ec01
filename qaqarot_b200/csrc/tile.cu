// tile.cu -- the fused tile-resident multi-gate pass (B200_OP_TILE), sm_100a.
//
// One pass = one read and one write of the state (32 * 2^n bytes, the HBM-bound unit of SURVEY.md 8(d)), with
// as many gates as the planner could legally bring onto the tile applied on chip in between.
//
// Data layout.  A pass owns A index bits `tile[0..A)` (ascending physical positions; tile[i] = i for i < c so
// that 2^c amplitudes = 2^c * 16 B are contiguous in HBM).  Tile number t in [0, 2^(n-A)) has its bits spread
// over the remaining index positions; a CTA copies the 2^A amplitudes of its tile to shared memory with
// 16-byte cp.async (coalesced rows), runs the pass's rounds, and streams the tile back with 16-byte stores.
// Shared-memory slot of local index l is swz(l) = l ^ xor of l's higher 3-bit groups folded into bits 0..2:
// GF(2)-linear, so swz(thread part | register part) = swz(thread) ^ swz(register), and a quarter warp's eight
// 16-byte accesses fall into eight different 16-byte bank groups whenever the thread bits mapped to lane bits
// 0..2 have distinct positions mod 3 (the planner orders them so).
//
// Rounds.  In a round every thread holds 2^R amplitudes in registers: the R *register bits* (local bit
// positions order[0..R)) enumerate them, the thread index supplies the other A-R local bits through
// order[R..A).  Non-diagonal gates act on register bits only, so a 2x2 gate is pure register arithmetic
// (complex128 FMA chains, target position resolved to a compile-time template by a warp-uniform switch);
// controls and phases may involve any bit: register bits become compile-time lane-invariant predicates,
// thread bits a per-thread predicate, bits outside the tile a per-tile (CTA-uniform) predicate.
//
// Descriptor (uint64 words, produced by qaqarot_b200/planner.py::encode_pass; `reals` = float64 pool):
//   D[0] = A | c<<8 | R<<16 | n_rounds<<24 | n_fans<<32 | flags<<40   D[1], D[2] = tile bit positions, 8 bits each
//          flags: 2 = the last round stores straight to HBM (its register bits avoid the c row bits, so a half
//          warp still covers one contiguous 2^c * 16 B row); 1 = same for the first round's loads: accepted and
//          ignored (measured neutral; the code was removed to keep the kernel small)
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
//                                      6 = layer: for each register bit k in reg_mask (bits 0..3) the three
//                                          in-place shears a0 += z a1; a1 += x a0; a0 += y a1 on the pairs of bit
//                                          k, payload 4 x (z, x, y, 0); j unused.  With z = y = -tan(t/2), x = sin t
//                                          a rotation by t: the planner writes every 1-qubit unitary as
//                                          phase . rotation . phase (planner.split_shears)
//     diag stage:  1 / 2 = fan  amp *= fix[fan] * lane_t[lane] * warp_t[warp] (* reg_t[rho] for kind 2) on the
//                  amplitudes with register bit j set (j = R: all); tables 32 + 2^(A-R-5) (+ 2^R) complex
//                  3 = phase  amp *= reals[offset] where rho satisfies reg_mask (never fused with a target stage)
//                  4 = fan without entries: amp *= reals[offset] on register bit j (a 1-qubit phase, no fan record)
//                  5 = fan whose entries all lie outside the tile: amp *= fix[fan], no tables
//                  6 = table: amp[rho] *= table[rho][variant] (2^R x 2^m complex), any diagonal on the register bits;
//                      variant = selector bits p1 | p2 << 1 | p3 << 2 with p1 = reg_mask >> 4 & 63, p2 = reg_mask >> 10
//                      & 63, p3 = fan byte; a selector < 48 is an index bit outside the tile, 48 + l the tile-local
//                      thread bit l, 63 none (m = selectors in use, packed first) -- fans and phases whose other
//                      qubits are not register bits ride along as variants
//     A 2x2 followed by a fan on the same register bit (the H + controlled-phase ladder of a QFT) is one record.
#include "common.cuh"

#include <cstdio>
#include <cstdlib>

namespace b200 {

constexpr int kMaxFans = 64;
constexpr int T_NONE = 0, T_MATC = 1, T_MATR = 2, T_X = 3, T_LAYER = 6;
constexpr int D_NONE = 0, D_FANT = 1, D_FAN = 2, D_PHASE = 3, D_W = 4, D_FIX = 5, D_TAB = 6;

__device__ __forceinline__ uint32_t swz(uint32_t l) {
  return l ^ ((l >> 3) & 7u) ^ ((l >> 6) & 7u) ^ ((l >> 9) & 7u) ^ ((l >> 12) & 7u);
}
__host__ __device__ constexpr uint32_t swz_c(uint32_t l) {
  return l ^ ((l >> 3) & 7u) ^ ((l >> 6) & 7u) ^ ((l >> 9) & 7u) ^ ((l >> 12) & 7u);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
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

// (a0, a1) <- E(y) F(x) E(z) (a0, a1) on the pairs of register bit J, for the J in `mask`; payload 4 x (z, x, y, -)
template <int R>
__device__ __forceinline__ void layer_apply(c128 (&v)[1 << R], const double* __restrict__ pay, uint32_t mask) {
#pragma unroll
  for (int J = 0; J < R; ++J) {
    if (!((mask >> J) & 1u)) continue;
    const double z = pay[4 * J], x = pay[4 * J + 1], y = pay[4 * J + 2];
#pragma unroll
    for (int rho = 0; rho < (1 << R); ++rho) {
      if (rho & (1 << J)) continue;
      c128& a0 = v[rho];
      c128& a1 = v[rho | (1 << J)];
      asm("fma.rn.f64 %0, %4, %2, %0;\n\tfma.rn.f64 %1, %4, %3, %1;\n\t"
          "fma.rn.f64 %2, %5, %0, %2;\n\tfma.rn.f64 %3, %5, %1, %3;\n\t"
          "fma.rn.f64 %0, %6, %2, %0;\n\tfma.rn.f64 %1, %6, %3, %1;"
          : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(z), "d"(x), "d"(y));
    }
  }
}
// amp[rho] *= tab[rho * stride]: the table of this thread's variant (entries of one rho for all variants are adjacent,
// so the threads of a warp that picked different variants read different banks)
template <int R>
__device__ __forceinline__ void tab_apply(c128 (&v)[1 << R], const double2* __restrict__ tab, uint32_t stride) {
#pragma unroll
  for (int rho = 0; rho < (1 << R); ++rho) {
    const double2 f = tab[rho * stride];
    amp_mul(v[rho], f.x, f.y, -f.y);
  }
}

template <int R>
__device__ __forceinline__ void run_target(c128 (&v)[1 << R], uint32_t key1, uint32_t rm, const double* __restrict__ pay) {
  switch (key1) {
#define B200_T_CASES(J)                                                                           \
    case 16 * J + T_MATC: mat_c<R, (J < R ? J : 0), false>(v, pay, 0); break;                       \
    case 16 * J + T_MATR: mat_r<R, (J < R ? J : 0), false>(v, pay, 0); break;                       \
    case 16 * J + T_X: perm_x<R, (J < R ? J : 0), false>(v, 0); break;                              \
    case 16 * J + T_MATC + 8: mat_c<R, (J < R ? J : 0), true>(v, pay, rm); break;                   \
    case 16 * J + T_MATR + 8: mat_r<R, (J < R ? J : 0), true>(v, pay, rm); break;                   \
    case 16 * J + T_X + 8: perm_x<R, (J < R ? J : 0), true>(v, rm); break;
    B200_T_CASES(0) B200_T_CASES(1) B200_T_CASES(2) B200_T_CASES(3)
#undef B200_T_CASES
    default: break;
  }
}
template <int R, int NWARP_T>
__device__ __forceinline__ void run_diag(c128 (&v)[1 << R], uint32_t key2, uint32_t rm, const double* __restrict__ pay,
                                         c128 wt) {
  const double2* __restrict__ reg_t = reinterpret_cast<const double2*>(pay) + 32 + NWARP_T;
  switch (key2) {
#define B200_D_CASES(J)                                                                           \
    case 8 * J + D_FANT: case 8 * J + D_W: case 8 * J + D_FIX:                                      \
      fant_apply<R, (J < R ? J : R)>(v, wt); break;                                                 \
    case 8 * J + D_FAN: fan_apply<R, (J < R ? J : R)>(v, wt, reg_t); break;
    B200_D_CASES(0) B200_D_CASES(1) B200_D_CASES(2) B200_D_CASES(3) B200_D_CASES(4)
#undef B200_D_CASES
    case 8 * 4 + D_PHASE: {
      const double wx = pay[0], wy = pay[1];
#pragma unroll
      for (int rho = 0; rho < (1 << R); ++rho)
        if ((rho & rm) == rm) amp_mul(v[rho], wx, wy, -wy);
      break;
    }
    default: break;
  }
}

// ---- pre-decoded pass (built once per CTA in shared memory) ------------------------------------------------------
// Decoding the planner's descriptor per gate, per thread, per tile cost ~60 instructions a gate in the first
// version (ncu source view: a quarter of all issue slots).  The gate list is the same for every tile, so each
// CTA decodes it once into 32-byte records and per-thread / per-round tables, and the tile loop only reads them.
constexpr int kMaxGates = 160;     // per pass
constexpr int kMaxRounds = 8;
constexpr int kSpreadTabs = 5;     // 6 bits of the tile number each: up to 30 non-tile index bits

constexpr int kMaxPayload = 2560;  // float64 payload of one pass kept in shared memory (planner.MAX_PAYLOAD)
constexpr int kMaxFanWords = 512;  // directory + records of the pass's fans (planner keeps a pass below this)
// record (one uint4): x = key1 | key2<<8 | fan<<16 | fixed mask bits 32..39<<24
//                     y = thread mask | reg mask<<16    z = target payload | diag payload<<16 (complex units, relative)
//                     w = fixed mask bits 0..31

struct RoundHdr {                  // 16 bytes
  uint16_t n_gates, first;         // records [first, first + n_gates)
  uint16_t srb[4];                 // slot XOR masks of the register bits: swz(1 << order[k])
  uint32_t pad;
};

template <int A, int R>
struct TileSmem {
  c128 tile[1 << A];
  c128 fix[2][kMaxFans];           // per-tile fan products, two generations: no barrier separates the last round of
                                   // a tile from the head of the next one when that round stores straight to HBM
  uint64_t spread[kSpreadTabs][64];
  uint4 rec[kMaxGates + 1];
  double pay[kMaxPayload];         // the pass's matrices, phases and fan tables
  uint64_t fanw[kMaxFanWords];     // fan directory and records (entry counts, payload offsets, fixed-bit masks)
  RoundHdr rh[kMaxRounds];
  uint32_t thr[kMaxRounds][1 << (A - R)];   // per round, per thread: local index bits | swizzled byte offset / 16 << 16
  uint32_t round_off[kMaxRounds];
  // direct global <-> register access of the first / last round (when its register bits avoid the row bits):
  uint64_t gthr[1 << (A - R)];     // per thread: where its thread bits sit in the index space
  uint64_t ghb[4];                 // per register bit: its index bit as a mask
};

// ---- the kernel ------------------------------------------------------------------------------------------------
// FULL = false leaves out the general target stages (complex / real 2x2, pair swap, their controlled forms): a pass
// of 1-qubit unitaries and phases needs none of them, and the kernel without their 24 unrolled bodies (4100 instead
// of 6900 instructions) runs every record ~0.03 ms faster at 28 qubits -- measured 45.2 -> 41.0 ms on the 30-qubit QFT.
template <int A, int R, int MINB, bool TIMING, bool FULL>
__global__ void __launch_bounds__(1 << (A - R), MINB)
k_tile(c128* __restrict__ amp, uint64_t n_tiles, uint64_t index_offset, const uint64_t* __restrict__ desc,
       const double* __restrict__ reals, unsigned long long* __restrict__ phase_cycles, uint32_t flags,
       uint64_t init_tile, uint32_t init_local) {
  static_assert(R == 3 || R == 4, "the record dispatch handles 3 or 4 register bits");
  constexpr int NT = 1 << (A - R), PER = 1 << R, NTB = A - R;
  constexpr int NWARP_T = (NTB > 5) ? (1 << (NTB - 5)) : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<A, R>& S = *reinterpret_cast<TileSmem<A, R>*>(smem_raw);
  c128* const sm = S.tile;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t h0 = desc[0], tb_lo = desc[1], tb_hi = desc[2];
  const int n_rounds = (int)((h0 >> 24) & 0xFF), n_fans = (int)((h0 >> 32) & 0xFF);
  const uint32_t dir_off = (uint32_t)(desc[3] & 0xFFFFFFFFu), rounds_off = (uint32_t)(desc[3] >> 32);
  const uint32_t pay_lo = (uint32_t)(desc[4] & 0xFFFFFFFFu), pay_len = (uint32_t)(desc[4] >> 32);
  // flags bit 1: the state is a basis vector that nobody has written yet -- every tile is all zeros except
  // tile `init_tile`, whose local index `init_local` holds 1 (B200_OP_INIT folded into this pass)
  const bool init = (flags & 2u) != 0;
  // (flag 1, "first round loads straight from HBM", is accepted and ignored: measured neutral, and its code cost
  // every record ~0.01 ms -- see the note on kernel size at k_tile)
  const bool direct_out = ((h0 >> 41) & 1) != 0;

  // ---- one-time setup ----------------------------------------------------------------------------------------
  if (tid == 0) {
    uint32_t pos = rounds_off, first = 0;
    for (int rd = 0; rd < n_rounds; ++rd) {
      const uint64_t rh = desc[pos];
      S.round_off[rd] = pos;
      S.rh[rd].n_gates = (uint16_t)(rh & 0xFFFF);
      S.rh[rd].first = (uint16_t)first;
      first += (uint32_t)(rh & 0xFFFF);
      pos += (uint32_t)(rh >> 16);
    }
  }
  for (uint32_t i = tid; i < pay_len; i += NT) S.pay[i] = reals[pay_lo + i];
  for (uint32_t i = dir_off + tid; i < rounds_off; i += NT) S.fanw[i - dir_off] = desc[i];
  for (uint32_t idx = tid; idx < kSpreadTabs * 64; idx += NT) {
    uint64_t x = (uint64_t)(idx & 63u) << (6 * (idx >> 6));
    for (int i = 0; i < A; ++i) x = insert_zero(x, byte_of(tb_lo, tb_hi, i));
    S.spread[idx >> 6][idx & 63u] = x;
  }
  __syncthreads();
  for (int rd = 0; rd < n_rounds; ++rd) {
    const uint64_t* __restrict__ rp = desc + S.round_off[rd];
    const uint64_t o_lo = rp[1], o_hi = rp[2];
    uint32_t lt = 0;
    for (int b = 0; b < NTB; ++b) lt |= ((tid >> b) & 1u) << byte_of(o_lo, o_hi, R + b);
    S.thr[rd][tid] = lt | (swz(lt) << 16);
    if (rd == n_rounds - 1) {      // the last round may store straight to HBM (flag 2)
      uint64_t go = 0;
      for (int b = 0; b < A; ++b) go |= (uint64_t)((lt >> b) & 1u) << byte_of(tb_lo, tb_hi, b);
      S.gthr[tid] = go;
      if (tid < R) S.ghb[tid] = 1ull << byte_of(tb_lo, tb_hi, byte_of(o_lo, o_hi, tid));
    }
    if (tid < R) S.rh[rd].srb[tid] = (uint16_t)swz(1u << byte_of(o_lo, o_hi, tid));
    const int ng = S.rh[rd].n_gates, first = S.rh[rd].first;
    for (int gi = tid; gi < ng; gi += NT) {
      const uint64_t g0 = rp[3 + 3 * gi], fm = rp[4 + 3 * gi], roffs = rp[5 + 3 * gi];
      const uint32_t tk = (uint32_t)(g0 & 0xFF), j = (uint32_t)((g0 >> 8) & 0xFF), dk = (uint32_t)((g0 >> 16) & 0xFF);
      const uint32_t fan = (uint32_t)((g0 >> 24) & 0xFF);
      const uint32_t lm = (uint32_t)((g0 >> 32) & 0xFFFF), rm = (uint32_t)((g0 >> 48) & 0xFFFF);
      const uint32_t pt = tk == T_NONE || tk == T_X ? 0u : (((uint32_t)roffs - pay_lo) >> 1);
      const uint32_t pd = dk == D_NONE ? 0u : (((uint32_t)(roffs >> 32) - pay_lo) >> 1);
      const uint32_t jj = j < (uint32_t)R ? j : 4u;            // 4 = no register target ("every amplitude")
      const uint32_t key1 = tk == T_NONE ? 0u : tk == T_LAYER ? kKeyLayer : 16u * jj + tk + (rm != 0 ? 8u : 0u);
      const uint32_t key2 = dk == D_NONE ? 0u : dk == D_TAB ? kKeyTab : 8u * jj + dk;
      S.rec[first + gi] = make_uint4(key1 | (key2 << 8) | (fan << 16) | ((uint32_t)(fm >> 32) << 24),
                                     lm | (rm << 16), pt | (pd << 16), (uint32_t)fm);
    }
  }
  // copy phases: thread handles local indices l = i * NT + tid; off_t / hb[] deposit l into index positions
  uint64_t off_t = 0;
#pragma unroll
  for (int b = 0; b < NTB; ++b) off_t |= (uint64_t)((tid >> b) & 1u) << byte_of(tb_lo, tb_hi, b);
  const uint32_t sw_t = swz(tid) << 4;              // byte offset of this thread's slot for i = 0
  __syncthreads();

  auto spread_tile = [&](uint64_t t) -> uint64_t {
    uint64_t b = S.spread[0][t & 63u];
#pragma unroll
    for (int j = 1; j < kSpreadTabs; ++j) b |= S.spread[j][(t >> (6 * j)) & 63u];
    return b;
  };

  // B200_TILE_TIMING=1: warp 0 / lane 0 of every CTA accumulates the cycles it spends in each phase of a tile
  long long tk0 = TIMING ? clock64() : 0;
  auto tick = [&](int phase) {
    if (TIMING && tid == 0) {
      const long long now = clock64();
      atomicAdd(phase_cycles + phase, (unsigned long long)(now - tk0));
      tk0 = now;
    }
  };
  // flags bit 2 (with a direct last round): the next tile's cp.async is issued as soon as every warp has taken
  // the last round's amplitudes out of shared memory, so its HBM latency hides under that round's gates
  const bool overlap = (flags & 4u) != 0 && direct_out && !init;
  bool loaded = false;
  uint32_t parity = 0;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1u) {
    c128* const fixp = S.fix[parity];
    const uint64_t base = spread_tile(tile);
    const uint64_t fixedidx = base | index_offset;
    const uint32_t fx_lo = (uint32_t)fixedidx, fx_hi = (uint32_t)(fixedidx >> 32);
    c128* __restrict__ g = amp + base + off_t;
    c128 v[PER];
    if (init) {
#pragma unroll
      for (int i = 0; i < PER; ++i)
        *reinterpret_cast<c128*>(reinterpret_cast<unsigned char*>(sm) + (sw_t ^ (swz_c((uint32_t)i * NT) << 4))) =
            make_double2(0.0, 0.0);
    } else if (!loaded) {
      uint64_t hb[R];
#pragma unroll
      for (int k = 0; k < R; ++k) hb[k] = 1ull << byte_of(tb_lo, tb_hi, NTB + k);
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        uint64_t off = 0;
#pragma unroll
        for (int k = 0; k < R; ++k)
          if ((i >> k) & 1) off |= hb[k];
        cp_async16(reinterpret_cast<unsigned char*>(sm) + (sw_t ^ (swz_c((uint32_t)i * NT) << 4)), g + off);
      }
    }

    tick(8);                      // issue: loads
    // per-tile fixed-bit products of the fans, while the copies are in flight (one warp per fan)
    for (int f = (int)warp; f < n_fans; f += (NT >> 5)) {
      const uint32_t off = (uint32_t)S.fanw[f] - dir_off;
      const uint64_t rec = S.fanw[off];
      const int ne = (int)(rec & 0xFFFFFFFFu);
      const double2* __restrict__ w = reinterpret_cast<const double2*>(S.pay + ((uint32_t)(rec >> 32) - pay_lo));
      c128 p = make_double2(1.0, 0.0);
      for (int e = (int)lane; e < ne; e += 32) {
        const uint64_t m = S.fanw[off + 1 + e];
        if ((fixedidx & m) == m) p = cmul(p, w[1 + e]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        c128 q;
        q.x = __shfl_xor_sync(0xffffffffu, p.x, o);
        q.y = __shfl_xor_sync(0xffffffffu, p.y, o);
        p = cmul(p, q);
      }
      if (lane == 0) fixp[f] = cmul(p, w[0]);
    }
    tick(0);                      // fan products
    if (!init) cp_async_wait_all();
    __syncthreads();              // tile (or fan products) ready; also: every warp has left the previous tile
    if (init && tile == init_tile) {
      if (tid == 0) sm[swz(init_local)] = make_double2(1.0, 0.0);
      __syncthreads();
    }
    tick(1);                      // wait for the tile + barrier

    for (int rd = 0; rd < n_rounds; ++rd) {
      const uint4 hq = *reinterpret_cast<const uint4*>(&S.rh[rd]);
      const int ng = (int)(hq.x & 0xFFFFu), first = (int)(hq.x >> 16);
      const uint32_t srb[5] = {(hq.y & 0xFFFFu) << 4, (hq.y >> 16) << 4, (hq.z & 0xFFFFu) << 4, (hq.z >> 16) << 4, 0u};
      const uint32_t my = S.thr[rd][tid];
      const uint32_t lt = my & 0xFFFFu, stb = (my >> 16) << 4;
      unsigned char* const smb = reinterpret_cast<unsigned char*>(sm);

      // Gray-code walk over the 2^R register amplitudes: one XOR per shared-memory address
      {
        uint32_t s = stb;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          if (i) s ^= srb[ctz_c(i)];
          v[i ^ (i >> 1)] = *reinterpret_cast<const c128*>(smb + s);
        }
      }
      if (overlap && rd == n_rounds - 1) {
        __syncthreads();          // every warp holds its amplitudes: the shared-memory tile is free
        loaded = tile + gridDim.x < n_tiles;
        if (loaded) {
          const c128* __restrict__ gn = amp + spread_tile(tile + gridDim.x) + off_t;
          uint64_t hb[R];
#pragma unroll
          for (int k = 0; k < R; ++k) hb[k] = 1ull << byte_of(tb_lo, tb_hi, NTB + k);
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            uint64_t off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k)
              if ((i >> k) & 1) off |= hb[k];
            cp_async16(reinterpret_cast<unsigned char*>(sm) + (sw_t ^ (swz_c((uint32_t)i * NT) << 4)), gn + off);
          }
        }
      }
      tick(2);                    // round: registers <- shared memory
      // the next record is fetched while the current one runs
      uint4 q = S.rec[first];
      for (int gi = first; gi < first + ng; ++gi) {
        const uint4 nq = S.rec[gi + 1];
        const uint32_t lm = q.y & 0xFFFFu, rm = q.y >> 16, fmh = q.x >> 24;
        // fixed-bit condition: CTA-uniform; thread-bit condition: per thread
        const bool active = ((fx_lo & q.w) == q.w) && ((fx_hi & fmh) == fmh) && ((lt & lm) == lm);
        if (active) {
          const uint32_t key1 = q.x & 0xFFu, key2 = (q.x >> 8) & 0xFFu;
          const double* __restrict__ pay_t = S.pay + 2 * (q.z & 0xFFFFu);
          const double* __restrict__ pay_d = S.pay + 2 * (q.z >> 16);
          if (key1 == kKeyLayer) layer_apply<R>(v, pay_t, rm);       // (rm: which register bits carry a gate)
          else if (FULL && key1 != 0u) run_target<R>(v, key1, rm, pay_t);
          if (key2 == kKeyTab) {
            // one table per combination of (up to) three selector bits: index bits outside the tile (CTA-uniform per
            // tile) or thread bits of the tile (48 + local position); 63 = none, which reads as 0
            const uint32_t p1 = (rm >> 4) & 63u, p2 = (rm >> 10) & 63u, p3 = (q.x >> 16) & 63u;
            const uint64_t sel = fixedidx | ((uint64_t)lt << 48);
            const uint32_t var = (uint32_t)((sel >> p1) & 1ull) | ((uint32_t)((sel >> p2) & 1ull) << 1) |
                                 ((uint32_t)((sel >> p3) & 1ull) << 2);
            const uint32_t m = (p1 != 63u) + (p2 != 63u) + (p3 != 63u);
            tab_apply<R>(v, reinterpret_cast<const double2*>(pay_d) + var, 1u << m);
          } else if (key2 != 0u) {
            const uint32_t dk = key2 & 7u;
            c128 wt = make_double2(1.0, 0.0);
            if (dk == D_FANT || dk == D_FAN) {
              const double2* __restrict__ tab = reinterpret_cast<const double2*>(pay_d);
              wt = cmul(fixp[(q.x >> 16) & 0xFFu], cmul(tab[lane], tab[32 + (NWARP_T > 1 ? warp : 0)]));
            } else if (dk == D_FIX) {
              wt = fixp[(q.x >> 16) & 0xFFu];
            } else if (dk == D_W) {
              wt = *reinterpret_cast<const double2*>(pay_d);             // the factor itself
            }
            run_diag<R, NWARP_T>(v, key2, rm, pay_d, wt);
          }
        }
        q = nq;
      }

      tick(3);                    // round: gates
      if (direct_out && rd == n_rounds - 1) {
        c128* __restrict__ go = amp + base + S.gthr[tid];
        uint64_t hb[R];
#pragma unroll
        for (int k = 0; k < R; ++k) hb[k] = S.ghb[k];
#pragma unroll
        for (int rho = 0; rho < PER; ++rho) {
          uint64_t off = 0;
#pragma unroll
          for (int k = 0; k < R; ++k)
            if ((rho >> k) & 1) off |= hb[k];
          __stcg(go + off, v[rho]);
        }
        tick(4);
      } else {
        uint32_t s = stb;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          if (i) s ^= srb[ctz_c(i)];
          *reinterpret_cast<c128*>(smb + s) = v[i ^ (i >> 1)];
        }
        tick(4);                  // round: shared memory <- registers
        __syncthreads();
      }
      tick(5);                    // round: barrier
    }

    if (!direct_out) {
      uint64_t hb[R];
#pragma unroll
      for (int k = 0; k < R; ++k) hb[k] = 1ull << byte_of(tb_lo, tb_hi, NTB + k);
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        uint64_t off = 0;
#pragma unroll
        for (int k = 0; k < R; ++k)
          if ((i >> k) & 1) off |= hb[k];
        __stcs(g + off, *reinterpret_cast<const c128*>(reinterpret_cast<unsigned char*>(sm) + (sw_t ^ (swz_c((uint32_t)i * NT) << 4))));
      }
    }
    tick(6);                      // stream the tile back
    // the next tile's cp.async must not overwrite slots still being read (a direct first round writes shared
    // memory only after the barrier at the head of the next tile)
    if (!overlap) __syncthreads();
    tick(7);                      // end-of-tile barrier
  }
}

template <int A, int R, int MINB, bool FULL>
static int launch(b200_state* st, const uint64_t* d_desc, const double* d_reals, bool init, uint64_t init_tile,
                  uint32_t init_local) {
  constexpr int NT = 1 << (A - R);
  const size_t smem = sizeof(TileSmem<A, R>);
  static bool configured[64] = {};
  if (!configured[st->device & 63]) {
    B200_CUDA(cudaFuncSetAttribute(k_tile<A, R, MINB, false, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(k_tile<A, R, MINB, true, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[st->device & 63] = true;
  }
  const uint64_t n_tiles = st->dim >> A;
  // persistent: exactly the resident CTAs, each walking tiles blockIdx, blockIdx + grid, ... (it requests its
  // next tile while the last round of the current one still runs)
  static const int ctas_env = getenv("B200_TILE_CTAS_PER_SM") ? atoi(getenv("B200_TILE_CTAS_PER_SM")) : 0;   // experiment
  const uint64_t cap = (uint64_t)kSMs * (ctas_env > 0 && ctas_env < MINB ? ctas_env : MINB);
  const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
  static const bool timing = getenv("B200_TILE_TIMING") != nullptr;
  unsigned long long* d_cyc = nullptr;
  if (timing) {
    B200_CUDA(cudaMalloc(&d_cyc, 9 * sizeof(unsigned long long)));
    B200_CUDA(cudaMemsetAsync(d_cyc, 0, 9 * sizeof(unsigned long long), st->stream));
  }
  static const uint32_t flags = getenv("B200_TILE_NO_OVERLAP") ? 0u : 4u;     // (bit 0 was the L2 prefetch: removed)
  if (timing)
    k_tile<A, R, MINB, true, FULL><<<grid, NT, smem, st->stream>>>(st->amp, n_tiles, st->index_offset, d_desc, d_reals, d_cyc,
                                                             flags | (init ? 2u : 0u), init_tile, init_local);
  else
    k_tile<A, R, MINB, false, FULL><<<grid, NT, smem, st->stream>>>(st->amp, n_tiles, st->index_offset, d_desc, d_reals,
                                                              nullptr, flags | (init ? 2u : 0u), init_tile, init_local);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  if (timing) {
    unsigned long long h[9];
    B200_CUDA(cudaMemcpyAsync(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost, st->stream));
    B200_CUDA(cudaStreamSynchronize(st->stream));
    B200_CUDA(cudaFree(d_cyc));
    double tot = 0;
    for (int i = 0; i < 9; ++i) tot += (double)h[i];
    fprintf(stderr, "[k_tile A=%d R=%d tiles=%llu ctas=%u] cycles per tile per CTA:", A, R, (unsigned long long)n_tiles, grid);
    static const char* names[9] = {"fan_products", "wait", "rd_load", "rd_gates", "rd_store", "rd_barrier", "tile_store",
                                   "end_barrier", "issue_loads"};
    for (int i = 0; i < 9; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)h[i] / (double)n_tiles);
    fprintf(stderr, " total=%.0f\n", tot / (double)n_tiles);
  }
  return 0;
}

int launch_tile_pass(b200_state* st, const uint64_t* d_desc, const uint64_t* h_desc, int wlen,
                     const double* d_reals, size_t n_reals, uint64_t init_index) {
  B200_REQUIRE(wlen >= 5, "TILE descriptor too short");
  const uint64_t h0 = h_desc[0];
  const int A = (int)(h0 & 0xFF), c = (int)((h0 >> 8) & 0xFF), R = (int)((h0 >> 16) & 0xFF);
  const int n_rounds = (int)((h0 >> 24) & 0xFF), n_fans = (int)((h0 >> 32) & 0xFF);
  char msg[200];
  if (A > st->n || c > A || (R != 4 && !(R == 3 && A == 11)) || n_fans > kMaxFans || n_rounds > kMaxRounds ||
      st->n - A > 6 * kSpreadTabs) {
    snprintf(msg, sizeof msg, "TILE descriptor not executable: A=%d c=%d R=%d fans=%d rounds=%d on %d qubits", A, c, R,
             n_fans, n_rounds, st->n);
    set_error(msg);
    return 1;
  }
  int prev = -1;
  for (int i = 0; i < A; ++i) {
    const int b = (int)(((i < 8 ? h_desc[1] : h_desc[2]) >> (8 * (i & 7))) & 0xFF);
    if (b <= prev || b >= st->n || (i < c && b != i)) {
      set_error("TILE descriptor: tile bits must be ascending, below n_qubits, and 0..c-1 first");
      return 1;
    }
    prev = b;
  }
  const uint64_t dir_off = h_desc[3] & 0xFFFFFFFFu, rounds_off = h_desc[3] >> 32;
  B200_REQUIRE(dir_off + (uint64_t)n_fans <= (uint64_t)wlen && rounds_off <= (uint64_t)wlen,
               "TILE descriptor: offsets out of range");
  B200_REQUIRE(rounds_off >= dir_off && rounds_off - dir_off <= (uint64_t)kMaxFanWords,
               "TILE descriptor: fan records too long");
  B200_REQUIRE((h_desc[4] >> 32) <= (uint64_t)kMaxPayload && (h_desc[4] & 0xFFFFFFFFu) + (h_desc[4] >> 32) <= n_reals,
               "TILE descriptor: payload range too long or outside the reals pool");
  uint64_t pos = rounds_off, gates = 0;
  bool full = false;              // does any record need a general target stage (kinds 1..3)?
  for (int rd = 0; rd < n_rounds; ++rd) {
    B200_REQUIRE(pos + 3 <= (uint64_t)wlen, "TILE descriptor: round header out of range");
    const uint64_t ng = h_desc[pos] & 0xFFFF, nw = h_desc[pos] >> 16;
    B200_REQUIRE(nw >= 3 + 3 * ng && pos + nw <= (uint64_t)wlen, "TILE descriptor: round length out of range");
    for (uint64_t g = 0; g < ng; ++g) {
      const int tk = (int)(h_desc[pos + 3 + 3 * g] & 0xFF);
      full |= tk == T_MATC || tk == T_MATR || tk == T_X;
    }
    gates += ng;
    pos += nw;
  }
  B200_REQUIRE(pos == (uint64_t)wlen && gates <= (uint64_t)kMaxGates, "TILE descriptor: bad round lengths or too many gates");
  const bool init = init_index != kNoInit;
  uint64_t init_tile = ~0ull;
  uint32_t init_local = 0;
  if (init && (init_index & ~(st->dim - 1)) == st->index_offset) {     // the 1 lives on this shard
    bool in_tile[64] = {};
    for (int i = 0; i < A; ++i) {
      const int b = (int)(((i < 8 ? h_desc[1] : h_desc[2]) >> (8 * (i & 7))) & 0xFF);
      in_tile[b] = true;
      init_local |= (uint32_t)((init_index >> b) & 1ull) << i;
    }
    init_tile = 0;
    for (int b = 0, k = 0; b < st->n; ++b)
      if (!in_tile[b]) init_tile |= ((init_index >> b) & 1ull) << k++;
  }
  switch (A) {
    case 9:  return launch<9, 4, 8, true>(st, d_desc, d_reals, init, init_tile, init_local);
    case 10: return launch<10, 4, 8, true>(st, d_desc, d_reals, init, init_tile, init_local);
    case 11: return R == 3 ? launch<11, 3, 3, true>(st, d_desc, d_reals, init, init_tile, init_local)
                           : launch<11, 4, 4, true>(st, d_desc, d_reals, init, init_tile, init_local);
    // the production tile size also exists without the general target stages (see k_tile)
    case 12: return full ? launch<12, 4, 2, true>(st, d_desc, d_reals, init, init_tile, init_local)
                         : launch<12, 4, 2, false>(st, d_desc, d_reals, init, init_tile, init_local);
    case 13: return launch<13, 4, 1, true>(st, d_desc, d_reals, init, init_tile, init_local);
    default:
      snprintf(msg, sizeof msg, "TILE descriptor: tile size 2^%d is not built (9..13)", A);
      set_error(msg);
      return 1;
  }
}

}  // namespace b200
