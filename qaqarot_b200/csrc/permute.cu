// permute.cu -- in-place index-bit involution (B200_OP_PERMUTE), sm_100a.
//
// The planner treats `swap` gates (which the reference expands to three CX sweeps, gate.py:1172-1177) as a
// relabelling of which index bit a qubit lives on, and owes the state one physical bit permutation at the end
// of a program (a QFT's final bit reversal is the textbook case).  Any permutation is the product of at most
// two involutions, and an involution of index bits -- a set of disjoint transpositions (p_i, q_i) -- is applied
// here in ONE pass over the state (32 * 2^n bytes, every amplitude read once and written once):
//
//   tile bits T = the low k bits (so rows of 2^k * 16 B stay contiguous in HBM) + the partners of those that
//   are transposed + enough further low bits to fill a 2^A tile; transpositions inside T are done through
//   shared memory (XOR-swizzled slots, as in tile.cu); transpositions entirely outside T pair tile t with
//   tile t' (the tile number with those bit pairs exchanged): a CTA loads both tiles and writes each one,
//   permuted, to the other's location.  Tiles with t' < t are left to their partner; tiles that map to
//   themselves under an identity inner permutation are skipped without touching memory.
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace b200 {

constexpr int kInvMaxA = 10;       // tile of up to 2^10 amplitudes (16 KiB), two tiles per CTA
constexpr int kInvThreads = 256;

struct InvDesc {
  int a;                      // tile bits
  int n_outer;                // transpositions between non-tile bits
  unsigned char tile[kInvMaxA];        // physical positions, ascending
  unsigned char src_local[kInvMaxA];   // destination local bit i takes source local bit src_local[i]
  unsigned char ou[32], ov[32];        // outer pairs as bit positions of the (compressed) tile number
  int inner_identity;
};

__device__ __forceinline__ uint32_t inv_swz(uint32_t l) {
  return l ^ ((l >> 3) & 7u) ^ ((l >> 6) & 7u) ^ ((l >> 9) & 7u);
}

constexpr int kInvEPT = (1 << kInvMaxA) / kInvThreads;     // amplitudes per thread per tile (at most)

__global__ void __launch_bounds__(kInvThreads) k_involution(c128* __restrict__ amp, uint64_t n_tiles,
                                                            const __grid_constant__ InvDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* const sa = smem_raw;
  unsigned char* const sb = smem_raw + (16u << d.a);
  const uint32_t n_local = 1u << d.a;
  // Everything that depends only on the local index is computed once per thread: where local index l sits in
  // the index space (off), its shared-memory slot when written, and the slot its *source* element was written to.
  uint64_t off[kInvEPT];
  uint32_t wslot[kInvEPT], rslot[kInvEPT];
#pragma unroll
  for (int k = 0; k < kInvEPT; ++k) {
    const uint32_t l = threadIdx.x + k * kInvThreads;
    uint64_t o = 0;
    uint32_t ls = 0;
#pragma unroll
    for (int i = 0; i < kInvMaxA; ++i)
      if (i < d.a) {
        o |= (uint64_t)((l >> i) & 1u) << d.tile[i];
        ls |= ((l >> i) & 1u) << d.src_local[i];          // involution: source bit <-> destination bit
      }
    off[k] = o;
    wslot[k] = inv_swz(l) << 4;
    rslot[k] = inv_swz(ls) << 4;
  }
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    uint64_t t2 = t;
#pragma unroll 1
    for (int i = 0; i < d.n_outer; ++i) {
      const uint64_t bu = (t >> d.ou[i]) & 1ull, bv = (t >> d.ov[i]) & 1ull;
      if (bu != bv) t2 ^= (1ull << d.ou[i]) | (1ull << d.ov[i]);
    }
    if (t2 < t || (t2 == t && d.inner_identity)) continue;         // CTA-uniform
    uint64_t base = t, base2 = t2;
#pragma unroll
    for (int i = 0; i < kInvMaxA; ++i)
      if (i < d.a) {
        base = insert_zero(base, d.tile[i]);
        base2 = insert_zero(base2, d.tile[i]);
      }
    const bool two = t2 != t;
    c128 x[kInvEPT], y[kInvEPT];
#pragma unroll
    for (int k = 0; k < kInvEPT; ++k)
      if (threadIdx.x + k * kInvThreads < n_local) {
        x[k] = __ldcg(amp + base + off[k]);
        if (two) y[k] = __ldcg(amp + base2 + off[k]);
      }
#pragma unroll
    for (int k = 0; k < kInvEPT; ++k)
      if (threadIdx.x + k * kInvThreads < n_local) {
        *reinterpret_cast<c128*>(sa + wslot[k]) = x[k];
        if (two) *reinterpret_cast<c128*>(sb + wslot[k]) = y[k];
      }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kInvEPT; ++k)
      if (threadIdx.x + k * kInvThreads < n_local) {
        __stcg(amp + base2 + off[k], *reinterpret_cast<const c128*>(sa + rslot[k]));
        if (two) __stcg(amp + base + off[k], *reinterpret_cast<const c128*>(sb + rslot[k]));
      }
    __syncthreads();
  }
}

// pairs[i] = p | q << 8: disjoint transpositions of index bits.
int launch_permute(b200_state* st, const uint64_t* pairs, int n_pairs) {
  const int n = st->n;
  int partner[64];
  for (int i = 0; i < 64; ++i) partner[i] = i;
  for (int i = 0; i < n_pairs; ++i) {
    const int p = (int)(pairs[i] & 0xFF), q = (int)((pairs[i] >> 8) & 0xFF);
    if (p >= n || q >= n || p == q || partner[p] != p || partner[q] != q) {
      set_error("PERMUTE: transpositions must be disjoint pairs of distinct qubits below n_qubits");
      return 1;
    }
    partner[p] = q;
    partner[q] = p;
  }
  // tile: low k bits, their partners, then further low bits (with partners) while it fits
  const int k = std::min(4, n);
  bool in_tile[64] = {};
  int count = 0;
  auto add = [&](int b) { if (!in_tile[b]) { in_tile[b] = true; ++count; } };
  for (int b = 0; b < k; ++b) { add(b); add(partner[b]); }
  for (int b = k; b < n && count < kInvMaxA; ++b) {
    if (in_tile[b]) continue;
    const int need = partner[b] == b || in_tile[partner[b]] ? 1 : 2;
    if (count + need > kInvMaxA) continue;
    add(b);
    add(partner[b]);
  }
  InvDesc d;
  memset(&d, 0, sizeof d);
  d.a = count;
  int loc[64];
  for (int b = 0, i = 0; b < n; ++b)
    if (in_tile[b]) { d.tile[i] = (unsigned char)b; loc[b] = i++; }
  d.inner_identity = 1;
  for (int i = 0; i < d.a; ++i) {
    d.src_local[i] = (unsigned char)loc[partner[d.tile[i]]];
    if (d.src_local[i] != i) d.inner_identity = 0;
  }
  // outer pairs in tile-number coordinates (non-tile bits, compressed ascending)
  int comp[64], c = 0;
  for (int b = 0; b < n; ++b) comp[b] = in_tile[b] ? -1 : c++;
  for (int b = 0; b < n; ++b)
    if (!in_tile[b] && partner[b] > b) {
      if (d.n_outer == 32) { set_error("PERMUTE: too many transpositions"); return 1; }
      d.ou[d.n_outer] = (unsigned char)comp[b];
      d.ov[d.n_outer] = (unsigned char)comp[partner[b]];
      ++d.n_outer;
    }
  if (d.n_outer == 0 && d.inner_identity) return 0;
  const uint64_t n_tiles = st->dim >> d.a;
  const size_t smem = sizeof(c128) * 2 * (size_t(1) << d.a);
  const uint64_t cap = (uint64_t)kSMs * 16;
  const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, cap);
  k_involution<<<grid, kInvThreads, smem, st->stream>>>(st->amp, n_tiles, d);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  return 0;
}

// ---- direct peer-to-peer exchange of one index bit (sharded states, one process per GPU) ----------------------
// Exchanging local index bit `lbit` with a global bit means: this rank (global bit b) trades its amplitudes whose
// bit lbit is 1-b with the partner's amplitudes whose bit lbit is b, all other bits equal.  `peer` is the partner
// GPU's shard mapped through CUDA IPC, reached over NVLink by plain 16-byte loads and stores (runs of 2^lbit
// amplitudes are contiguous on both sides).  The two partners call this on disjoint halves of the pair range, so
// every pair is swapped exactly once and both NVLink directions carry the same traffic; nothing is staged.
__global__ void __launch_bounds__(256) k_peer_swap(c128* __restrict__ mine, c128* __restrict__ peer, uint64_t start,
                                                   uint64_t count, int lbit, int my_bit) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t mbit = (uint64_t)(1 - my_bit) << lbit, pbit = (uint64_t)my_bit << lbit;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte transfers in flight per thread in each direction
  for (; i + 3 * stride < count; i += 4 * stride) {
    c128 x[4], y[4];
    uint64_t k0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) k0[k] = insert_zero(start + i + k * stride, lbit);
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = __ldcg(peer + (k0[k] | pbit));
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __ldcg(mine + (k0[k] | mbit));
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcg(peer + (k0[k] | pbit), x[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcg(mine + (k0[k] | mbit), y[k]);
  }
  for (; i < count; i += stride) {
    const uint64_t k0 = insert_zero(start + i, lbit);
    const c128 y = __ldcg(peer + (k0 | pbit)), x = __ldcg(mine + (k0 | mbit));
    __stcg(peer + (k0 | pbit), x);
    __stcg(mine + (k0 | mbit), y);
  }
}

// The same for k local bits at once (a remap of k global wires in one step, SURVEY 8e "general k: all-to-all"): with
// the partner that differs from this shard in a subset d of the k global bits, the block of this shard whose local
// bits equal the PARTNER's global bits (`mine_pat`) trades places with the partner's block whose local bits equal
// THIS shard's global bits (`peer_pat`), all other bits equal.  One call per partner, 2^k - 1 partners, each block
// 2^(n-k) amplitudes: (1 - 2^-k) of the shard leaves, against k / 2 for k single-bit exchanges.
__global__ void __launch_bounds__(256) k_peer_swap_bits(c128* __restrict__ mine, c128* __restrict__ peer, uint64_t start,
                                                        uint64_t count, BitPositions bp, uint64_t mine_pat,
                                                        uint64_t peer_pat) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < count; i += 4 * stride) {
    c128 x[4], y[4];
    uint64_t k0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) k0[k] = spread(start + i + k * stride, bp);
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = __ldcg(peer + (k0[k] | peer_pat));
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __ldcg(mine + (k0[k] | mine_pat));
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcg(peer + (k0[k] | peer_pat), x[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcg(mine + (k0[k] | mine_pat), y[k]);
  }
  for (; i < count; i += stride) {
    const uint64_t k0 = spread(start + i, bp);
    const c128 y = __ldcg(peer + (k0 | peer_pat)), x = __ldcg(mine + (k0 | mine_pat));
    __stcg(peer + (k0 | peer_pat), x);
    __stcg(mine + (k0 | mine_pat), y);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_peer_swap_bits(b200_state* st, void* peer_amps, const int32_t* lbits, int k, uint64_t mine_pat,
                                   uint64_t peer_pat, uint64_t start, uint64_t count, int sync) {
  B200_REQUIRE(st != nullptr && (peer_amps != nullptr || count == 0) && lbits != nullptr, "b200_peer_swap_bits: null argument");
  B200_REQUIRE(k >= 1 && k <= 8 && k <= st->n, "b200_peer_swap_bits: 1..8 local bits");
  BitPositions bp;
  bp.n = k;
  uint64_t mask = 0;
  for (int j = 0; j < k; ++j) {
    B200_REQUIRE(lbits[j] >= 0 && lbits[j] < st->n && (j == 0 || lbits[j] > lbits[j - 1]),
                 "b200_peer_swap_bits: local bits must be ascending and inside the shard");
    bp.pos[j] = lbits[j];
    mask |= 1ull << lbits[j];
  }
  B200_REQUIRE((mine_pat & ~mask) == 0 && (peer_pat & ~mask) == 0 && mine_pat != peer_pat,
               "b200_peer_swap_bits: the two block patterns must differ and lie on the local bits");
  const uint64_t block = st->dim >> k;
  B200_REQUIRE(start <= block && count <= block - start, "b200_peer_swap_bits: range exceeds the block");
  B200_CUDA(cudaSetDevice(st->device));
  if (count) {
    k_peer_swap_bits<<<kSMs * 8, 256, 0, st->stream>>>(st->amp, static_cast<c128*>(peer_amps), start, count, bp, mine_pat,
                                                       peer_pat);
    st->launches++;
    B200_CUDA(cudaGetLastError());
  }
  if (sync) B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

extern "C" int b200_state_ipc_export(const b200_state* st, unsigned char* handle64) {
  B200_REQUIRE(st != nullptr && handle64 != nullptr, "b200_state_ipc_export: null argument");
  B200_REQUIRE(st->owns, "b200_state_ipc_export: only states created by b200_state_create can be exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  B200_CUDA(cudaSetDevice(st->device));
  cudaIpcMemHandle_t h;
  B200_CUDA(cudaIpcGetMemHandle(&h, st->amp));
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int b200_ipc_open(int device, const unsigned char* handle64, void** peer_amps) {
  B200_REQUIRE(handle64 != nullptr && peer_amps != nullptr, "b200_ipc_open: null argument");
  B200_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200_CUDA(cudaIpcOpenMemHandle(peer_amps, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int b200_ipc_close(void* peer_amps) {
  if (peer_amps) B200_CUDA(cudaIpcCloseMemHandle(peer_amps));
  return 0;
}

extern "C" int b200_peer_swap(b200_state* st, void* peer_amps, int lbit, int my_bit, uint64_t start, uint64_t count) {
  B200_REQUIRE(st != nullptr && peer_amps != nullptr, "b200_peer_swap: null argument");
  B200_REQUIRE(lbit >= 0 && lbit < st->n && (my_bit == 0 || my_bit == 1), "b200_peer_swap: bad bit");
  B200_REQUIRE(start <= st->dim / 2 && count <= st->dim / 2 - start, "b200_peer_swap: pair range exceeds half the state");
  if (count == 0) return 0;
  B200_CUDA(cudaSetDevice(st->device));
  const unsigned grid = kSMs * 8;
  k_peer_swap<<<grid, 256, 0, st->stream>>>(st->amp, static_cast<c128*>(peer_amps), start, count, lbit, my_bit);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}
