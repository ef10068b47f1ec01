// sample.cu -- batched terminal-measurement sampler (sm_100a).
//
// Replaces the reference's shot loop for circuits that end in measurements: numpy re-executes
// gate_measure per shot and per measured qubit (numpy_backend.py:447-477, 620-623), each a full pass with a
// collapse.  Here all shots descend the measurement tree together, one level per measured qubit in the
// reference's target_iter order.  At level j the shots that share the same earlier outcomes form a group;
// one kernel computes, for every group, S0 = ||psi[group, q_j = 0]||^2 and S1 = ||psi[group, q_j = 1]||^2 over
// the *uncollapsed* state, and the host compares the shot's own uniform with p0 = S0 / (S0 + S1), exactly the
// comparison `rand < p_zero` the reference makes on its renormalised state.  The work per level is
// min(2^j, shots) * 2^(n-j) amplitudes, so the whole descent costs about log2(shots) + 2 passes, independent
// of the number of shots.  Sums are two-stage and fixed-order, hence run-to-run deterministic.
#include "common.cuh"

#include <algorithm>
#include <numeric>

namespace b200 {

struct FreeRuns {          // maximal runs of free (not yet fixed, not current) index bits, ascending
  int n;
  unsigned char start[40];
  unsigned char len[40];
};

__device__ __forceinline__ uint64_t deposit(uint64_t k, const FreeRuns& fr) {
  uint64_t i = 0;
  for (int r = 0; r < fr.n; ++r) {
    const int len = fr.len[r];
    i |= (k & ((1ull << len) - 1ull)) << fr.start[r];
    k >>= len;
  }
  return i;
}

constexpr int kSampThreads = 256;

__global__ void __launch_bounds__(kSampThreads)
k_cond_prob(const c128* __restrict__ amp, uint64_t count, FreeRuns fr, const uint64_t* __restrict__ group_bits,
            uint64_t tbit, int bx, double* __restrict__ partial) {
  const int g = blockIdx.x / bx, b = blockIdx.x % bx;
  const uint64_t fixed = group_bits[g];
  double s0 = 0.0, s1 = 0.0;
  const uint64_t stride = (uint64_t)bx * blockDim.x;
  for (uint64_t k = (uint64_t)b * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i = deposit(k, fr) | fixed;
    const c128 a0 = amp[i];
    s0 += a0.x * a0.x + a0.y * a0.y;
    if (tbit) {                      // (no target qubit: S0 is the whole group and S1 stays 0)
      const c128 a1 = amp[i | tbit];
      s1 += a1.x * a1.x + a1.y * a1.y;
    }
  }
  __shared__ double sh0[kSampThreads / 32], sh1[kSampThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh0[w] = s0; sh1[w] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = 0; i < kSampThreads / 32; ++i) { t0 += sh0[i]; t1 += sh1[i]; }
    partial[2 * (size_t)blockIdx.x] = t0;
    partial[2 * (size_t)blockIdx.x + 1] = t1;
  }
}

__global__ void k_group_finish(const double* __restrict__ partial, int n_groups, int bx, double* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  double t0 = 0.0, t1 = 0.0;
  for (int b = 0; b < bx; ++b) {
    t0 += partial[2 * ((size_t)g * bx + b)];
    t1 += partial[2 * ((size_t)g * bx + b) + 1];
  }
  out[2 * (size_t)g] = t0;
  out[2 * (size_t)g + 1] = t1;
}

// Marginal distribution of the LOW k index bits: table[v] = sum of |amp[i]|^2 over the i whose low k bits are v.
// This is what the first levels of a descent need when the first measured qubits sit on index bits 0..k-1 (m[:] on a
// state in its natural layout) -- and the case in which k_cond_prob is at its worst: its groups then are strided
// subsets (one 16-byte amplitude every 2^k * 16 bytes), a quarter of the DRAM efficiency of a contiguous read or
// less; at 36 qubits on 8 GPUs the descent spent ~0.4 s that way.  Here a block reads a CONTIGUOUS slab, thread t the
// amplitudes base + t + 256 s: the bin of an amplitude is its index mod 2^k, so thread t only ever touches the bins
// congruent to t mod 256 -- private to it, no atomics, a fixed summation order, one bank per lane.
constexpr int kMargThreads = 256;
constexpr int kMargMaxBits = 13;        // 2^13 bins of 8 bytes = 64 KB of shared memory per block

__global__ void __launch_bounds__(kMargThreads)
k_marginal_low(const c128* __restrict__ amp, uint64_t per_block, int k, double* __restrict__ partial) {
  extern __shared__ double bins[];
  const uint32_t nb = 1u << k, mask = nb - 1u;
  for (uint32_t i = threadIdx.x; i < nb; i += kMargThreads) bins[i] = 0.0;
  __syncthreads();
  const c128* __restrict__ p = amp + (uint64_t)blockIdx.x * per_block;     // per_block is a multiple of 1024
  for (uint64_t off = threadIdx.x; off < per_block; off += 4 * kMargThreads) {
    c128 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = __ldcs(p + off + (uint64_t)u * kMargThreads);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t b = (uint32_t)(off + (uint64_t)u * kMargThreads) & mask;    // (the slab starts on a multiple of 2^k)
      bins[b] += a[u].x * a[u].x + a[u].y * a[u].y;
    }
  }
  __syncthreads();
  double* __restrict__ out = partial + (size_t)blockIdx.x * nb;
  for (uint32_t i = threadIdx.x; i < nb; i += kMargThreads) out[i] = bins[i];
}

__global__ void k_marginal_finish(const double* __restrict__ partial, int n_blocks, uint32_t nb, double* __restrict__ table) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nb) return;
  double t = 0.0;
  for (int b = 0; b < n_blocks; ++b) t += partial[(size_t)b * nb + v];
  table[v] = t;
}

// The three device arrays of one reduction live in one grow-only allocation owned by the state (a descent makes up
// to 64 reductions; allocating and freeing per level cost three cudaMalloc / cudaFree pairs each, and cudaFree
// synchronises the device).
static int sampler_scratch(b200_state* st, size_t bytes, unsigned char** out) {
  if (bytes > st->cap_samp) {
    if (st->d_samp) {
      B200_CUDA(cudaStreamSynchronize(st->stream));
      B200_CUDA(cudaFree(st->d_samp));
      st->d_samp = nullptr;
      st->cap_samp = 0;
    }
    size_t cap = 1 << 20;
    while (cap < bytes) cap <<= 1;
    B200_CUDA(cudaMalloc(&st->d_samp, cap));
    st->cap_samp = cap;
  }
  *out = static_cast<unsigned char*>(st->d_samp);
  return 0;
}

}  // namespace b200

using namespace b200;

// One level of the descent: for every group g (amplitudes whose index has exactly the bits group_bits[g] among
// `fixed_mask`), S0 = sum |amp|^2 over the group with bit `qubit` = 0 and S1 with bit = 1 (qubit < 0: S0 is the
// whole group, S1 = 0).  sums[2g], sums[2g+1].  All masks are over this state's own n index bits.
static int cond_probs(b200_state* st, uint64_t fixed_mask, int qubit, const uint64_t* group_bits, int n_groups,
                      double* sums) {
  const int max_blocks = kSMs * 16;
  const uint64_t tbit = qubit >= 0 ? 1ull << qubit : 0ull;
  const size_t sz_bits = sizeof(uint64_t) * (size_t)n_groups, sz_out = sizeof(double) * 2 * (size_t)n_groups;
  const size_t sz_partial = sizeof(double) * 2 * ((size_t)n_groups + (size_t)max_blocks);
  unsigned char* base = nullptr;
  if (sampler_scratch(st, sz_bits + sz_out + sz_partial, &base)) return 1;
  struct { void* p; } d_bits{base}, d_out{base + sz_bits}, d_partial{base + sz_bits + sz_out};
  FreeRuns fr;
  fr.n = 0;
  int free_bits = 0;
  for (int b = 0; b < st->n;) {
    if ((fixed_mask | tbit) >> b & 1) { ++b; continue; }
    int e = b;
    while (e < st->n && !((fixed_mask | tbit) >> e & 1)) ++e;
    fr.start[fr.n] = (unsigned char)b;
    fr.len[fr.n] = (unsigned char)(e - b);
    ++fr.n;
    free_bits += e - b;
    b = e;
  }
  const uint64_t count = 1ull << free_bits;
  uint64_t want = (count + (uint64_t)kSampThreads * 4 - 1) / ((uint64_t)kSampThreads * 4);
  int bx = (int)std::min<uint64_t>(want, (uint64_t)std::max(1, max_blocks / n_groups));
  if (bx < 1) bx = 1;
  B200_CUDA(cudaMemcpyAsync(d_bits.p, group_bits, sizeof(uint64_t) * n_groups, cudaMemcpyHostToDevice, st->stream));
  k_cond_prob<<<n_groups * bx, kSampThreads, 0, st->stream>>>(st->amp, count, fr, (const uint64_t*)d_bits.p, tbit, bx,
                                                             (double*)d_partial.p);
  B200_CUDA(cudaGetLastError());
  k_group_finish<<<(n_groups + 127) / 128, 128, 0, st->stream>>>((const double*)d_partial.p, n_groups, bx,
                                                                 (double*)d_out.p);
  B200_CUDA(cudaGetLastError());
  st->launches += 2;
  B200_CUDA(cudaMemcpyAsync(sums, d_out.p, sizeof(double) * 2 * n_groups, cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

static int marginal_low(b200_state* st, int k, double* table) {
  const uint32_t nb = 1u << k;
  // slabs: a multiple of 2^k and of 4 * 256 amplitudes each, at most 512 of them
  uint64_t per_block = std::max<uint64_t>((uint64_t)nb, 4ull * kMargThreads);
  while (st->dim / per_block > 512) per_block <<= 1;
  const int n_blocks = (int)(st->dim / per_block);
  const size_t sz_partial = sizeof(double) * (size_t)n_blocks * nb, sz_table = sizeof(double) * nb;
  unsigned char* base = nullptr;
  if (sampler_scratch(st, sz_partial + sz_table, &base)) return 1;
  double* d_partial = reinterpret_cast<double*>(base);
  double* d_table = reinterpret_cast<double*>(base + sz_partial);
  static bool configured[64] = {};
  if (!configured[st->device & 63]) {
    B200_CUDA(cudaFuncSetAttribute(k_marginal_low, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) << kMargMaxBits)));
    configured[st->device & 63] = true;
  }
  k_marginal_low<<<n_blocks, kMargThreads, sizeof(double) * nb, st->stream>>>(st->amp, per_block, k, d_partial);
  B200_CUDA(cudaGetLastError());
  k_marginal_finish<<<(nb + 255) / 256, 256, 0, st->stream>>>(d_partial, n_blocks, nb, d_table);
  B200_CUDA(cudaGetLastError());
  st->launches += 2;
  B200_CUDA(cudaMemcpyAsync(table, d_table, sz_table, cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

extern "C" int b200_marginal_low(b200_state* st, int k, double* table) {
  B200_REQUIRE(st != nullptr && table != nullptr, "b200_marginal_low: null argument");
  B200_REQUIRE(k >= 8 && k <= kMargMaxBits && k <= st->n - 2, "b200_marginal_low: 8..13 low bits of a state with at least 2 more");
  B200_CUDA(cudaSetDevice(st->device));
  return marginal_low(st, k, table);
}

extern "C" int b200_cond_probs(b200_state* st, uint64_t fixed_mask, int qubit, const uint64_t* group_bits,
                               int n_groups, double* sums) {
  B200_REQUIRE(st != nullptr && group_bits != nullptr && sums != nullptr, "b200_cond_probs: null argument");
  B200_REQUIRE(n_groups > 0, "b200_cond_probs: no groups");
  B200_REQUIRE(qubit < st->n, "b200_cond_probs: qubit out of range");
  B200_REQUIRE((fixed_mask & ~(st->dim - 1)) == 0, "b200_cond_probs: fixed mask out of range");
  B200_REQUIRE(qubit < 0 || !((fixed_mask >> qubit) & 1), "b200_cond_probs: qubit is already fixed");
  B200_CUDA(cudaSetDevice(st->device));
  return cond_probs(st, fixed_mask, qubit, group_bits, n_groups, sums);
}

extern "C" int b200_sample(b200_state* st, const int32_t* order, int m, const double* uniforms, int n_shots,
                           uint64_t* out_bits) {
  B200_REQUIRE(st != nullptr && out_bits != nullptr, "b200_sample: null argument");
  B200_REQUIRE(m >= 0 && m <= 64, "b200_sample: at most 64 measured qubits per call");
  B200_REQUIRE(n_shots >= 0, "b200_sample: negative shot count");
  B200_REQUIRE(m == 0 || (order != nullptr && uniforms != nullptr), "b200_sample: null order/uniforms");
  for (int j = 0; j < m; ++j) B200_REQUIRE(order[j] >= 0 && order[j] < st->n, "b200_sample: qubit out of range");
  std::fill(out_bits, out_bits + n_shots, 0ull);
  if (m == 0 || n_shots == 0) return 0;
  B200_CUDA(cudaSetDevice(st->device));

  std::vector<uint64_t> prefix(n_shots, 0ull);        // outcomes so far, as index bits
  std::vector<int> idx(n_shots), group_of(n_shots);
  std::vector<uint64_t> group_bits;
  std::vector<double> sums;
  uint64_t fixed_mask = 0;

  // The first K levels (K <= 10 distinct qubits) cost ONE pass: the marginal distribution of those K qubits is
  // read off a single conditional-probability reduction with all 2^(K-1) prefixes as groups, and the
  // conditionals of the levels above it are sums of its entries.  Level by level, each of the first
  // log2(shots) levels would be a full read of the state.
  int j0 = 0;
  // ... and when the first K measured qubits are exactly the index bits 0..K-1 (in any order), K up to 13, the table comes
  // from one CONTIGUOUS read of the state (k_marginal_low) instead of 2^(K-1) strided groups.
  {
    int K = 0;
    uint64_t seen = 0;
    for (int t = 0, kk = 0; t < m && t < kMargMaxBits && !((seen >> order[t]) & 1ull); ++t) {
      seen |= 1ull << order[t];
      ++kk;
      if (seen == (1ull << kk) - 1ull) K = kk;          // the first kk qubits are the low kk bits
    }
    if (K >= 11 && K <= st->n - 2) {
      std::vector<double> table((size_t)1 << K);
      if (marginal_low(st, K, table.data())) return 1;
      std::vector<std::vector<double>> node(K + 1);
      node[K].resize((size_t)1 << K);
      for (uint32_t p = 0; p < (1u << K); ++p) {         // bit t of p = outcome of order[t] = index bit order[t]
        uint32_t v = 0;
        for (int t = 0; t < K; ++t) v |= ((p >> t) & 1u) << order[t];
        node[K][p] = table[v];
      }
      for (int j = K - 1; j >= 0; --j) {
        node[j].resize((size_t)1 << j);
        for (int p = 0; p < (1 << j); ++p) node[j][p] = node[j + 1][p] + node[j + 1][p | (1 << j)];
      }
      for (int s = 0; s < n_shots; ++s) {
        int p = 0;
        for (int j = 0; j < K; ++j) {
          const double s0 = node[j + 1][p], s1 = node[j + 1][p | (1 << j)];
          const double p0 = s0 / (s0 + s1);
          if (!(uniforms[(size_t)s * m + j] < p0)) {
            p |= 1 << j;
            out_bits[s] |= 1ull << j;
            prefix[s] |= 1ull << order[j];
          }
        }
      }
      fixed_mask = (1ull << K) - 1ull;
      j0 = K;
    }
  }
  if (j0 == 0) {
    int K = 0;
    uint64_t seen = 0;
    while (K < m && K < 10 && !((seen >> order[K]) & 1ull)) seen |= 1ull << order[K++];
    if (K >= 2) {
      const int n_groups = 1 << (K - 1);
      uint64_t mask0 = 0;
      for (int t = 0; t < K - 1; ++t) mask0 |= 1ull << order[t];
      group_bits.resize(n_groups);
      for (int g = 0; g < n_groups; ++g) {
        uint64_t b = 0;
        for (int t = 0; t < K - 1; ++t)
          if ((g >> t) & 1) b |= 1ull << order[t];
        group_bits[g] = b;
      }
      sums.resize(2 * (size_t)n_groups);
      if (cond_probs(st, mask0, order[K - 1], group_bits.data(), n_groups, sums.data())) return 1;
      // node[j][p]: probability of the first j outcomes p (bit t of p = outcome of order[t])
      std::vector<std::vector<double>> node(K + 1);
      node[K].resize((size_t)1 << K);
      for (int g = 0; g < n_groups; ++g) {
        node[K][g] = sums[2 * (size_t)g];
        node[K][g | (1 << (K - 1))] = sums[2 * (size_t)g + 1];
      }
      for (int j = K - 1; j >= 0; --j) {
        node[j].resize((size_t)1 << j);
        for (int p = 0; p < (1 << j); ++p) node[j][p] = node[j + 1][p] + node[j + 1][p | (1 << j)];
      }
      for (int s = 0; s < n_shots; ++s) {
        int p = 0;
        for (int j = 0; j < K; ++j) {
          const double s0 = node[j + 1][p], s1 = node[j + 1][p | (1 << j)];
          const double p0 = s0 / (s0 + s1);
          if (!(uniforms[(size_t)s * m + j] < p0)) {
            p |= 1 << j;
            out_bits[s] |= 1ull << j;
            prefix[s] |= 1ull << order[j];
          }
        }
      }
      fixed_mask = mask0 | (1ull << order[K - 1]);
      j0 = K;
    }
  }

  for (int j = j0; j < m; ++j) {
    const int q = order[j];
    const uint64_t tbit = 1ull << q;
    if (fixed_mask & tbit) {
      // measured before: the state is already collapsed on this qubit, p_zero is 1 or 0.
      for (int s = 0; s < n_shots; ++s) {
        const double p0 = (prefix[s] & tbit) ? 0.0 : 1.0;
        if (!(uniforms[(size_t)s * m + j] < p0)) out_bits[s] |= 1ull << j;
      }
      continue;
    }
    // group the shots by their outcomes so far
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return prefix[a] < prefix[b]; });
    group_bits.clear();
    for (int r = 0; r < n_shots; ++r) {
      const int s = idx[r];
      if (r == 0 || prefix[s] != group_bits.back()) group_bits.push_back(prefix[s]);
      group_of[s] = (int)group_bits.size() - 1;
    }
    const int n_groups = (int)group_bits.size();
    sums.resize(2 * (size_t)n_groups);
    if (cond_probs(st, fixed_mask, q, group_bits.data(), n_groups, sums.data())) return 1;
    for (int s = 0; s < n_shots; ++s) {
      const int g = group_of[s];
      const double s0 = sums[2 * (size_t)g], s1 = sums[2 * (size_t)g + 1];
      const double p0 = s0 / (s0 + s1);
      if (!(uniforms[(size_t)s * m + j] < p0)) {
        out_bits[s] |= 1ull << j;
        prefix[s] |= tbit;
      }
    }
    fixed_mask |= tbit;
  }
  return 0;
}
