// b200sv.cu -- C ABI + the single-gate, reduction and measurement kernels (sm_100a).
//
// Every kernel here is HBM-bound integer-indexed complex128 work: one coalesced 16-byte load/store per
// amplitude, pair / subspace enumeration by zero-bit insertion so that controlled and diagonal gates touch
// only the amplitudes the reference's boolean masks select (numpy_backend.py:75-497).  The fused
// multi-gate pass lives in tile.cu.
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <algorithm>

namespace b200 {

static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }

bool cuda_ok(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return true;
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %s (%d) at %s:%d in `%s`", cudaGetErrorString(e), (int)e, file, line, what);
  g_error = buf;
  return false;
}

struct M2 { double v[8]; };   // m00 m01 m10 m11, each (re, im)
struct D2 { double v[4]; };   // d0 d1, each (re, im)

constexpr int kThreads = 256;

static inline int grid_for(uint64_t count, int per_thread = 1) {
  uint64_t blocks = (count + (uint64_t)kThreads * per_thread - 1) / ((uint64_t)kThreads * per_thread);
  uint64_t cap = (uint64_t)kSMs * 32;            // 32 waves-worth of 256-thread CTAs, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- unitary single-gate kernels ---------------------------------------------------------------

// 2x2 on `tbit`, restricted to the subspace where the `ones` bits are set.  `bp` holds the sorted
// positions of target and control bits; k enumerates the remaining n - bp.n bits.
__global__ void __launch_bounds__(kThreads) k_mat(c128* __restrict__ amp, uint64_t count, BitPositions bp,
                                                  uint64_t ones, uint64_t tbit, M2 m) {
  const c128 m00 = make_double2(m.v[0], m.v[1]), m01 = make_double2(m.v[2], m.v[3]);
  const c128 m10 = make_double2(m.v[4], m.v[5]), m11 = make_double2(m.v[6], m.v[7]);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i0 = spread(k, bp) | ones, i1 = i0 | tbit;
    const c128 a0 = amp[i0], a1 = amp[i1];
    amp[i0] = cfma(m01, a1, cmul(m00, a0));
    amp[i1] = cfma(m11, a1, cmul(m10, a0));
  }
}

__global__ void __launch_bounds__(kThreads) k_x(c128* __restrict__ amp, uint64_t count, BitPositions bp,
                                                uint64_t ones, uint64_t tbit) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i0 = spread(k, bp) | ones, i1 = i0 | tbit;
    const c128 a0 = amp[i0], a1 = amp[i1];
    amp[i0] = a1;
    amp[i1] = a0;
  }
}

// amp[i] *= d for the subspace where all `ones` bits (controls and target) are set: z/s/t/phase/cz/cphase/ccz.
__global__ void __launch_bounds__(kThreads) k_phase(c128* __restrict__ amp, uint64_t count, BitPositions bp,
                                                    uint64_t ones, c128 d) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i = spread(k, bp) | ones;
    amp[i] = cmul(amp[i], d);
  }
}

// amp[i] *= bit(i, target) ? d1 : d0 where the control bits are set: rz / crz.
__global__ void __launch_bounds__(kThreads) k_diag2(c128* __restrict__ amp, uint64_t count, BitPositions bp,
                                                    uint64_t ones, uint64_t tbit, D2 d) {
  const c128 d0 = make_double2(d.v[0], d.v[1]), d1 = make_double2(d.v[2], d.v[3]);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i = spread(k, bp) | ones;
    amp[i] = cmul(amp[i], (i & tbit) ? d1 : d0);
  }
}

__global__ void __launch_bounds__(kThreads) k_swapbits(c128* __restrict__ amp, uint64_t count, BitPositions bp,
                                                       uint64_t abit, uint64_t bbit) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t base = spread(k, bp), i = base | abit, j = base | bbit;
    const c128 x = amp[i], y = amp[j];
    amp[i] = y;
    amp[j] = x;
  }
}

__global__ void __launch_bounds__(kThreads) k_scale(c128* __restrict__ amp, uint64_t count, c128 s) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    amp[i] = cmul(amp[i], s);
}

__global__ void k_set_one(c128* amp, uint64_t index) { amp[index] = make_double2(1.0, 0.0); }

// ---- deterministic two-stage reductions ------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of (a, b); result valid in thread 0.
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[kThreads / 32], sb[kThreads / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = (l < kThreads / 32) ? sa[l] : 0.0;
    b = (l < kThreads / 32) ? sb[l] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// partial[b] = sum over this block's share of |amp[i]|^2 for i with bit `target` == 0 (or all i if target<0)
__global__ void __launch_bounds__(kThreads) k_prob(const c128* __restrict__ amp, uint64_t count, int target,
                                                   double* __restrict__ partial) {
  double acc = 0.0, unused = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i = target >= 0 ? insert_zero(k, target) : k;
    const c128 a = amp[i];
    acc += a.x * a.x + a.y * a.y;
  }
  block_sum2(acc, unused);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Four amplitudes per thread share every term's mask and coefficient: per amplitude the term loop is an AND, a POPC and
// one (HAS_IM: two) FP64 additions, and the shared-memory broadcasts of the one-amplitude loop are divided by four
// (28-qubit QAOA energy, 42 Z.Z terms: the whole run(hamiltonian=) call 26.4 -> 25.5 ms).  The order of every sum is
// the one of the one-amplitude loop (bit-identical results).
template <bool HAS_IM>
__global__ void __launch_bounds__(kThreads) k_expect(const c128* __restrict__ amp, uint64_t dim, uint64_t offset,
                                                     uint64_t xmask,
                                                     int nt, const uint64_t* __restrict__ zy,
                                                     const double* __restrict__ cre, const double* __restrict__ cim,
                                                     double* __restrict__ partial) {
  extern __shared__ unsigned char smem_raw[];
  uint64_t* s_zy = reinterpret_cast<uint64_t*>(smem_raw);
  double* s_re = reinterpret_cast<double*>(s_zy + nt);
  double* s_im = s_re + nt;
  for (int k = threadIdx.x; k < nt; k += blockDim.x) { s_zy[k] = zy[k]; s_re[k] = cre[k]; s_im[k] = cim[k]; }
  __syncthreads();
  constexpr int U = 4;
  double accr = 0.0, acci = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < dim; i0 += U * stride) {
    c128 a[U], b[U];
    uint64_t idx[U];
    double wr[U], wi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = i0 + (uint64_t)u * stride;
      const bool live = idx[u] < dim;
      a[u] = live ? amp[idx[u]] : make_double2(0.0, 0.0);
      b[u] = live && xmask ? amp[idx[u] ^ xmask] : a[u];
      idx[u] |= offset;
      wr[u] = wi[u] = 0.0;
    }
    for (int k = 0; k < nt; ++k) {
      const uint64_t m = s_zy[k];
      const double cr = s_re[k], ci = HAS_IM ? s_im[k] : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool neg = __popcll(idx[u] & m) & 1;
        wr[u] += neg ? -cr : cr;
        if (HAS_IM) wi[u] += neg ? -ci : ci;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {                      // (an amplitude past the end contributes 0 * w)
      const double pr = b[u].x * a[u].x + b[u].y * a[u].y, pi = b[u].x * a[u].y - b[u].y * a[u].x;   // conj(b) * a
      accr += pr * wr[u] - pi * wi[u];
      acci += pr * wi[u] + pi * wr[u];
    }
  }
  block_sum2(accr, acci);
  if (threadIdx.x == 0) { partial[blockIdx.x] = accr; partial[kMaxPartials + blockIdx.x] = acci; }
}

// The same sum with TABLES instead of a loop over the terms (up to 56 terms per launch, states of at least 2^11
// amplitudes).  The sign of term k at index i is the parity of i & zy_k = parity(lo & zy_k) xor parity(hi & zy_k) with
// lo = i mod 2^11.  L[lo] holds the nt parities of lo as one bit mask (built once per block), H the parities of the
// chunk's high bits (once per chunk of 2^11 amplitudes), so X = H ^ L[lo] has bit k set where term k counts negative;
// and sum_k +-c_k is read off seven terms at a time from T[g][7-bit pattern].  Per amplitude: one 64-bit and up to
// eight 8-byte shared-memory reads and as many additions, against nine integer instructions and an addition PER TERM
// in k_expect -- whose 42-term QAOA energy at 28 qubits was bound by the integer pipe at ~4 ms, six reads of the state.
constexpr int kExpLoBits = 11, kExpGroupBits = 7, kExpMaxTerms = 56, kExpGroups = kExpMaxTerms / kExpGroupBits;

template <bool HAS_IM>
__global__ void __launch_bounds__(kThreads) k_expect_tab(const c128* __restrict__ amp, uint64_t dim, uint64_t offset,
                                                         uint64_t xmask, int nt, const uint64_t* __restrict__ zy,
                                                         const double* __restrict__ cre, const double* __restrict__ cim,
                                                         double* __restrict__ partial) {
  __shared__ uint64_t L[1 << kExpLoBits];
  __shared__ double Tr[kExpGroups][1 << kExpGroupBits];
  __shared__ double Ti[HAS_IM ? kExpGroups : 1][1 << kExpGroupBits];
  __shared__ uint64_t s_zy[kExpMaxTerms];
  __shared__ double s_re[kExpMaxTerms], s_im[kExpMaxTerms];
  for (int k = threadIdx.x; k < kExpMaxTerms; k += blockDim.x) {
    s_zy[k] = k < nt ? zy[k] : 0ull;
    s_re[k] = k < nt ? cre[k] : 0.0;
    s_im[k] = k < nt ? cim[k] : 0.0;
  }
  __syncthreads();
  for (uint32_t lo = threadIdx.x; lo < (1u << kExpLoBits); lo += blockDim.x) {
    uint64_t x = 0;
    for (int k = 0; k < nt; ++k) x |= (uint64_t)(__popcll((uint64_t)lo & s_zy[k]) & 1) << k;
    L[lo] = x;
  }
  for (uint32_t e = threadIdx.x; e < (uint32_t)(kExpGroups << kExpGroupBits); e += blockDim.x) {
    const uint32_t g = e >> kExpGroupBits, pat = e & ((1u << kExpGroupBits) - 1u);
    double tr = 0.0, ti = 0.0;
    for (int j = 0; j < kExpGroupBits; ++j) {
      const int k = (int)g * kExpGroupBits + j;
      const bool neg = (pat >> j) & 1u;
      tr += neg ? -s_re[k] : s_re[k];
      ti += neg ? -s_im[k] : s_im[k];
    }
    Tr[g][pat] = tr;
    if (HAS_IM) Ti[g][pat] = ti;
  }
  __syncthreads();
  const int ng = (nt + kExpGroupBits - 1) / kExpGroupBits;
  const uint64_t lo_mask = (1ull << kExpLoBits) - 1ull;
  double accr = 0.0, acci = 0.0;
  const uint64_t n_chunks = dim >> kExpLoBits;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t base = c << kExpLoBits, hi = (base | offset) & ~lo_mask;
    uint64_t H = 0;
    for (int k = 0; k < nt; ++k) H |= (uint64_t)(__popcll(hi & s_zy[k]) & 1) << k;
    for (uint32_t lo = threadIdx.x; lo < (1u << kExpLoBits); lo += blockDim.x) {
      const uint64_t i = base | lo;
      const c128 a = amp[i];
      const c128 b = xmask ? amp[i ^ xmask] : a;
      const uint64_t X = H ^ L[lo];
      double wr = 0.0, wi = 0.0;
      for (int g = 0; g < ng; ++g) {
        const uint32_t pat = (uint32_t)(X >> (g * kExpGroupBits)) & ((1u << kExpGroupBits) - 1u);
        wr += Tr[g][pat];
        if (HAS_IM) wi += Ti[g][pat];
      }
      const double pr = b.x * a.x + b.y * a.y, pi = b.x * a.y - b.y * a.x;   // conj(b) * a
      accr += pr * wr - pi * wi;
      acci += pr * wi + pi * wr;
    }
  }
  block_sum2(accr, acci);
  if (threadIdx.x == 0) { partial[blockIdx.x] = accr; partial[kMaxPartials + blockIdx.x] = acci; }
}

// out[0] = sum partial[0..n), out[1] = sum partial[kMaxPartials .. +n) in a fixed order.
__global__ void __launch_bounds__(kThreads) k_finish(const double* __restrict__ partial, int n, int two,
                                                     double* __restrict__ out) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    a += partial[i];
    if (two) b += partial[kMaxPartials + i];
  }
  block_sum2(a, b);
  if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

// ---- measurement collapse -------------------------------------------------------------------------

// mode 0: keep the `outcome` half scaled by s, zero the other half.
// mode 1 (reset): result always ends in the bit=0 half.
__global__ void __launch_bounds__(kThreads) k_collapse(c128* __restrict__ amp, uint64_t count, int target,
                                                       int outcome, double s, int mode) {
  const uint64_t tbit = 1ull << target;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const c128 zero = make_double2(0.0, 0.0);
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t i0 = insert_zero(k, target), i1 = i0 | tbit;
    if (outcome == 0) {
      const c128 a = amp[i0];
      amp[i0] = make_double2(a.x * s, a.y * s);
      amp[i1] = zero;
    } else {
      const c128 a = amp[i1];
      const c128 r = make_double2(a.x * s, a.y * s);
      if (mode == 1) { amp[i0] = r; amp[i1] = zero; }
      else           { amp[i0] = zero; amp[i1] = r; }
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_first_above(const c128* __restrict__ amp, uint64_t dim, double eps2,
                                                          unsigned long long* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long best = ~0ull;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
    if (i >= best) break;
    const c128 a = amp[i];
    if (a.x * a.x + a.y * a.y > eps2) { best = i; break; }
  }
  if (best != ~0ull) atomicMin(out, best);
}

// The same search in *wire* order for a state whose wires sit on permuted index bits (sharded states never undo
// their layout): key(i) = sum of bit p of (index_offset | i) << wire_of_bit[p], found through five 256-entry tables;
// the smallest key wins.  A full scan: a smaller physical index does not mean a smaller key.
struct WireOfBit { int8_t w[40]; };
__global__ void __launch_bounds__(kThreads) k_first_above_wire(const c128* __restrict__ amp, uint64_t dim, uint64_t offset,
                                                               double eps2, WireOfBit wob, int n_bits,
                                                               unsigned long long* __restrict__ out) {
  __shared__ unsigned long long tab[5][256];
  for (int idx = threadIdx.x; idx < 5 * 256; idx += blockDim.x) {
    const int c = idx >> 8, v = idx & 255;
    unsigned long long k = 0;
    for (int b = 0; b < 8; ++b)
      if (((v >> b) & 1) && 8 * c + b < n_bits) k |= 1ull << wob.w[8 * c + b];
    tab[c][v] = k;
  }
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long best = ~0ull;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
    const c128 a = amp[i];
    if (a.x * a.x + a.y * a.y > eps2) {
      const uint64_t f = offset | i;
      const unsigned long long k = tab[0][f & 255] | tab[1][(f >> 8) & 255] | tab[2][(f >> 16) & 255] |
                                   tab[3][(f >> 24) & 255] | tab[4][(f >> 32) & 255];
      best = k < best ? k : best;
    }
  }
  if (best != ~0ull) atomicMin(out, best);
}

// ---- batched sampler ----------------------------------------------------------------------------------
// Conditional marginal for one measured qubit: for every pending shot group g (amplitudes whose already
// measured qubits equal that group's outcomes) accumulate ||psi[.., bit=0]||^2 and ||psi[..]||^2.
// The generic implementation walks shots x amplitudes through the program-level API (see b200_sample).

}  // namespace b200

using namespace b200;

// =====================================================================================================
// C ABI
// =====================================================================================================

extern "C" {

const char* b200_last_error(void) { return g_error.c_str(); }

int b200_version(void) { return 104; }   // 104: b200_marginal_low; 103: b200_apply_program_ex; 102: table variants 17 entries apart (tile.cu)

int b200_device_count(int* out) {
  B200_REQUIRE(out != nullptr, "b200_device_count: null out");
  B200_CUDA(cudaGetDeviceCount(out));
  return 0;
}

static int state_init_common(b200_state* st) {
  B200_CUDA(cudaSetDevice(st->device));
  B200_CUDA(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
  B200_CUDA(cudaMalloc(&st->d_partial, sizeof(double) * 2 * kMaxPartials));
  B200_CUDA(cudaMalloc(&st->d_result, sizeof(double) * 8));
  B200_CUDA(cudaMallocHost(&st->h_result, sizeof(double) * 8));
  B200_CUDA(cudaEventCreate(&st->prog.start));
  B200_CUDA(cudaEventCreate(&st->prog.stop));
  return 0;
}

int b200_state_create(int device, int n_qubits, b200_state** out) {
  B200_REQUIRE(out != nullptr, "b200_state_create: null out");
  B200_REQUIRE(n_qubits >= 0 && n_qubits <= 40, "b200_state_create: n_qubits out of range [0, 40]");
  b200_state* st = new b200_state();
  st->device = device;
  st->n = n_qubits;
  st->dim = 1ull << n_qubits;
  st->owns = true;
  if (state_init_common(st)) { delete st; return 1; }
  cudaError_t e = cudaMalloc(&st->amp, sizeof(c128) * st->dim);
  if (e != cudaSuccess) {
    cuda_ok(e, "cudaMalloc(state)", __FILE__, __LINE__);
    st->amp = nullptr;
    b200_state_destroy(st);
    return 1;
  }
  *out = st;
  return b200_state_set_basis(st, 0);
}

int b200_state_wrap(int device, int n_qubits, void* device_amps, b200_state** out) {
  B200_REQUIRE(out != nullptr && device_amps != nullptr, "b200_state_wrap: null argument");
  B200_REQUIRE(n_qubits >= 0 && n_qubits <= 40, "b200_state_wrap: n_qubits out of range [0, 40]");
  b200_state* st = new b200_state();
  st->device = device;
  st->n = n_qubits;
  st->dim = 1ull << n_qubits;
  st->owns = false;
  st->amp = static_cast<c128*>(device_amps);
  if (state_init_common(st)) { delete st; return 1; }
  *out = st;
  return 0;
}

int b200_state_destroy(b200_state* st) {
  if (!st) return 0;
  cudaSetDevice(st->device);
  if (st->stream) cudaStreamSynchronize(st->stream);
  if (st->owns && st->amp) cudaFree(st->amp);
  if (st->d_partial) cudaFree(st->d_partial);
  if (st->d_samp) cudaFree(st->d_samp);
  if (st->d_result) cudaFree(st->d_result);
  if (st->h_result) cudaFreeHost(st->h_result);
  if (st->d_words) cudaFree(st->d_words);
  if (st->d_reals) cudaFree(st->d_reals);
  if (st->prog.start) cudaEventDestroy(st->prog.start);
  if (st->prog.stop) cudaEventDestroy(st->prog.stop);
  for (cudaEvent_t e : st->op_events) cudaEventDestroy(e);
  for (cudaEvent_t e : st->user_events) if (e) cudaEventDestroy(e);
  if (st->stream) cudaStreamDestroy(st->stream);
  delete st;
  return 0;
}

int b200_state_set_index_offset(b200_state* st, uint64_t offset) {
  B200_REQUIRE(st != nullptr, "b200_state_set_index_offset: null state");
  B200_REQUIRE((offset & (st->dim - 1)) == 0, "b200_state_set_index_offset: offset overlaps the state's own index bits");
  st->index_offset = offset;
  return 0;
}

int b200_state_n_qubits(const b200_state* st) { return st ? st->n : -1; }
void* b200_state_device_ptr(const b200_state* st) { return st ? (void*)st->amp : nullptr; }

int b200_sync(b200_state* st) {
  B200_REQUIRE(st != nullptr, "b200_sync: null state");
  B200_CUDA(cudaSetDevice(st->device));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_state_set_basis(b200_state* st, uint64_t index) {
  B200_REQUIRE(st != nullptr, "b200_state_set_basis: null state");
  B200_REQUIRE(index < st->dim, "b200_state_set_basis: index out of range");
  B200_CUDA(cudaSetDevice(st->device));
  B200_CUDA(cudaMemsetAsync(st->amp, 0, sizeof(c128) * st->dim, st->stream));
  k_set_one<<<1, 1, 0, st->stream>>>(st->amp, index);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_state_upload(b200_state* st, const double* host, uint64_t offset, uint64_t count) {
  B200_REQUIRE(st != nullptr && host != nullptr, "b200_state_upload: null argument");
  B200_REQUIRE(offset <= st->dim && count <= st->dim - offset, "b200_state_upload: range exceeds the state");
  B200_CUDA(cudaSetDevice(st->device));
  B200_CUDA(cudaMemcpyAsync(st->amp + offset, host, sizeof(c128) * count, cudaMemcpyHostToDevice, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_state_download(const b200_state* st, double* host, uint64_t offset, uint64_t count) {
  B200_REQUIRE(st != nullptr && host != nullptr, "b200_state_download: null argument");
  B200_REQUIRE(offset <= st->dim && count <= st->dim - offset, "b200_state_download: range exceeds the state");
  B200_CUDA(cudaSetDevice(st->device));
  B200_CUDA(cudaMemcpyAsync(host, st->amp + offset, sizeof(c128) * count, cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_state_copy(b200_state* dst, const b200_state* src) {
  B200_REQUIRE(dst != nullptr && src != nullptr, "b200_state_copy: null argument");
  B200_REQUIRE(dst->n == src->n, "b200_state_copy: qubit counts differ");
  B200_CUDA(cudaSetDevice(dst->device));
  B200_CUDA(cudaStreamSynchronize(src->stream));
  B200_CUDA(cudaMemcpyAsync(dst->amp, src->amp, sizeof(c128) * dst->dim, cudaMemcpyDeviceToDevice, dst->stream));
  B200_CUDA(cudaStreamSynchronize(dst->stream));
  return 0;
}

// ---- program execution ----------------------------------------------------------------------------

static int ensure_capacity(b200_state* st, size_t n_words, size_t n_reals) {
  if (n_words > st->cap_words) {
    if (st->d_words) B200_CUDA(cudaFree(st->d_words));
    st->d_words = nullptr;
    st->cap_words = std::max<size_t>(n_words * 2, 1024);
    B200_CUDA(cudaMalloc(&st->d_words, sizeof(uint64_t) * st->cap_words));
  }
  if (n_reals > st->cap_reals) {
    if (st->d_reals) B200_CUDA(cudaFree(st->d_reals));
    st->d_reals = nullptr;
    st->cap_reals = std::max<size_t>(n_reals * 2, 1024);
    B200_CUDA(cudaMalloc(&st->d_reals, sizeof(double) * st->cap_reals));
  }
  return 0;
}

// Sorted zero-insertion positions for the bits of `mask` (plus nothing else).
static bool positions_of(uint64_t mask, BitPositions* bp) {
  bp->n = 0;
  for (int b = 0; b < 64; ++b)
    if (mask >> b & 1) {
      if (bp->n == 8) return false;
      bp->pos[bp->n++] = b;
    }
  return true;
}

static int run_simple_op(b200_state* st, const b200_op& op, const double* reals, size_t n_reals) {
  const int n = st->n;
  const uint64_t full = st->dim - 1;
  char msg[160];
  if (op.kind != B200_OP_SCALE) {
    if (op.target < 0 || op.target >= n) {
      snprintf(msg, sizeof msg, "op target %d out of range for %d qubits", op.target, n);
      set_error(msg);
      return 1;
    }
    if ((op.cmask & ~full) || (op.cmask >> op.target & 1)) {
      set_error("op control mask out of range or overlapping the target");
      return 1;
    }
  }
  const uint64_t tbit = 1ull << (op.kind == B200_OP_SCALE ? 0 : op.target);
  BitPositions bp;
  switch (op.kind) {
    case B200_OP_MAT:
    case B200_OP_X: {
      if (!positions_of(op.cmask | tbit, &bp)) { set_error("too many control qubits (max 7)"); return 1; }
      const uint64_t count = st->dim >> bp.n;
      if (op.kind == B200_OP_X) {
        k_x<<<grid_for(count), kThreads, 0, st->stream>>>(st->amp, count, bp, op.cmask, tbit);
      } else {
        B200_REQUIRE(op.roff + 8 <= n_reals, "MAT payload out of range");
        M2 m;
        memcpy(m.v, reals + op.roff, sizeof m.v);
        k_mat<<<grid_for(count), kThreads, 0, st->stream>>>(st->amp, count, bp, op.cmask, tbit, m);
      }
      break;
    }
    case B200_OP_DIAG: {
      B200_REQUIRE(op.roff + 4 <= n_reals, "DIAG payload out of range");
      const double* d = reals + op.roff;
      if (d[0] == 1.0 && d[1] == 0.0) {
        if (!positions_of(op.cmask | tbit, &bp)) { set_error("too many control qubits (max 7)"); return 1; }
        const uint64_t count = st->dim >> bp.n;
        k_phase<<<grid_for(count), kThreads, 0, st->stream>>>(st->amp, count, bp, op.cmask | tbit,
                                                             make_double2(d[2], d[3]));
      } else {
        if (!positions_of(op.cmask, &bp)) { set_error("too many control qubits (max 8)"); return 1; }
        const uint64_t count = st->dim >> bp.n;
        D2 dd;
        memcpy(dd.v, d, sizeof dd.v);
        k_diag2<<<grid_for(count), kThreads, 0, st->stream>>>(st->amp, count, bp, op.cmask, tbit, dd);
      }
      break;
    }
    case B200_OP_SWAP: {
      B200_REQUIRE(op.aux >= 0 && op.aux < n && op.aux != op.target, "SWAP: bad second qubit");
      const uint64_t abit = tbit, bbit = 1ull << op.aux;
      positions_of(abit | bbit, &bp);
      const uint64_t count = st->dim >> 2;
      k_swapbits<<<grid_for(count), kThreads, 0, st->stream>>>(st->amp, count, bp, abit, bbit);
      break;
    }
    case B200_OP_SCALE: {
      B200_REQUIRE(op.roff + 2 <= n_reals, "SCALE payload out of range");
      k_scale<<<grid_for(st->dim), kThreads, 0, st->stream>>>(st->amp, st->dim,
                                                            make_double2(reals[op.roff], reals[op.roff + 1]));
      break;
    }
    default:
      snprintf(msg, sizeof msg, "unknown op kind %d", op.kind);
      set_error(msg);
      return 1;
  }
  st->launches++;
  B200_CUDA(cudaGetLastError());
  return 0;
}

int b200_apply_program(b200_state* st, const b200_op* ops, int n_ops, const uint64_t* words, size_t n_words,
                       const double* reals, size_t n_reals) {
  return b200_apply_program_ex(st, ops, n_ops, words, n_words, reals, n_reals, 0);
}

// Every op takes what it needs from the host arrays at launch time (a TILE pass travels as kernel parameters, the
// single-gate kernels take their matrices by value): nothing is uploaded, so a program may be handed over in pieces
// (B200_APPLY_ASYNC: return without waiting; B200_APPLY_CONTINUE: a further piece of the program whose first piece
// started the timer) while the host is still planning the rest.
int b200_apply_program_ex(b200_state* st, const b200_op* ops, int n_ops, const uint64_t* words, size_t n_words,
                          const double* reals, size_t n_reals, int flags) {
  B200_REQUIRE(st != nullptr, "b200_apply_program: null state");
  B200_REQUIRE(n_ops == 0 || ops != nullptr, "b200_apply_program: null ops");
  B200_CUDA(cudaSetDevice(st->device));
  if (st->profiling) {
    while (st->op_events.size() < (size_t)2 * n_ops) {
      cudaEvent_t e;
      B200_CUDA(cudaEventCreate(&e));
      st->op_events.push_back(e);
    }
    st->op_kinds.assign(n_ops, 0);
  }
  if (!(flags & B200_APPLY_CONTINUE)) B200_CUDA(cudaEventRecord(st->prog.start, st->stream));
  uint64_t pending_init = kNoInit;
  for (int i = 0; i < n_ops; ++i) {
    const b200_op& op = ops[i];
    if (st->profiling) { B200_CUDA(cudaEventRecord(st->op_events[2 * i], st->stream)); st->op_kinds[i] = op.kind; }
    if (op.kind == B200_OP_INIT) {
      const uint64_t lo = op.cmask & (st->dim - 1);
      B200_REQUIRE(st->n >= 63 || (op.cmask >> st->n) < (1ull << 24), "INIT: basis index out of range");
      if (i + 1 < n_ops && ops[i + 1].kind == B200_OP_TILE) {
        pending_init = op.cmask;                           // folded into the next pass
      } else {
        B200_CUDA(cudaMemsetAsync(st->amp, 0, sizeof(c128) * st->dim, st->stream));
        if ((op.cmask & ~(st->dim - 1)) == st->index_offset) {
          k_set_one<<<1, 1, 0, st->stream>>>(st->amp, lo);
          st->launches++;
          B200_CUDA(cudaGetLastError());
        }
      }
    } else if (op.kind == B200_OP_TILE) {
      B200_REQUIRE(op.woff + (uint64_t)op.wlen <= n_words, "TILE descriptor out of range");
      if (launch_tile_pass(st, words + op.woff, op.wlen, reals, n_reals, pending_init))
        return 1;
      pending_init = kNoInit;
    } else if (op.kind == B200_OP_PERMUTE) {
      B200_REQUIRE(op.wlen >= 0 && op.woff + (uint64_t)op.wlen <= n_words, "PERMUTE descriptor out of range");
      if (launch_permute(st, words + op.woff, op.wlen)) return 1;
    } else {
      if (run_simple_op(st, op, reals, n_reals)) return 1;
    }
    if (st->profiling) B200_CUDA(cudaEventRecord(st->op_events[2 * i + 1], st->stream));
  }
  B200_CUDA(cudaEventRecord(st->prog.stop, st->stream));
  st->prog.valid = true;
  if (!(flags & B200_APPLY_ASYNC)) B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_last_program_ms(b200_state* st, double* ms) {
  B200_REQUIRE(st != nullptr && ms != nullptr, "b200_last_program_ms: null argument");
  B200_REQUIRE(st->prog.valid, "b200_last_program_ms: no program has run");
  B200_CUDA(cudaSetDevice(st->device));
  float f = 0.f;
  B200_CUDA(cudaEventSynchronize(st->prog.stop));
  B200_CUDA(cudaEventElapsedTime(&f, st->prog.start, st->prog.stop));
  *ms = f;
  return 0;
}

int b200_set_profiling(b200_state* st, int on) {
  B200_REQUIRE(st != nullptr, "b200_set_profiling: null state");
  st->profiling = on != 0;
  if (!on) st->op_kinds.clear();
  return 0;
}

int b200_last_op_times(b200_state* st, float* ms, int32_t* kinds, int cap, int* n) {
  B200_REQUIRE(st != nullptr && n != nullptr, "b200_last_op_times: null argument");
  B200_CUDA(cudaSetDevice(st->device));
  const int have = (int)st->op_kinds.size();
  *n = have;
  for (int i = 0; i < have && i < cap; ++i) {
    float f = 0.f;
    B200_CUDA(cudaEventElapsedTime(&f, st->op_events[2 * i], st->op_events[2 * i + 1]));
    if (ms) ms[i] = f;
    if (kinds) kinds[i] = st->op_kinds[i];
  }
  return 0;
}

int b200_launch_count(const b200_state* st, uint64_t* out) {
  B200_REQUIRE(st != nullptr && out != nullptr, "b200_launch_count: null argument");
  *out = st->launches;
  return 0;
}

// ---- reductions / measurement -------------------------------------------------------------------------

static int reduce_blocks(uint64_t count) {
  int g = grid_for(count);
  return std::min(g, kSMs * 8);
}

static int finish_to_host(b200_state* st, int blocks, int two) {
  k_finish<<<1, kThreads, 0, st->stream>>>(st->d_partial, blocks, two, st->d_result);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(st->h_result, st->d_result, sizeof(double) * 2, cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_prob_zero(b200_state* st, int target, double* p0) {
  B200_REQUIRE(st != nullptr && p0 != nullptr, "b200_prob_zero: null argument");
  B200_REQUIRE(target >= 0 && target < st->n, "b200_prob_zero: target out of range");
  B200_CUDA(cudaSetDevice(st->device));
  const uint64_t count = st->dim >> 1;
  const int blocks = reduce_blocks(count);
  k_prob<<<blocks, kThreads, 0, st->stream>>>(st->amp, count, target, st->d_partial);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  if (finish_to_host(st, blocks, 0)) return 1;
  *p0 = st->h_result[0];
  return 0;
}

int b200_norm2(b200_state* st, double* out) {
  B200_REQUIRE(st != nullptr && out != nullptr, "b200_norm2: null argument");
  B200_CUDA(cudaSetDevice(st->device));
  const int blocks = reduce_blocks(st->dim);
  k_prob<<<blocks, kThreads, 0, st->stream>>>(st->amp, st->dim, -1, st->d_partial);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  if (finish_to_host(st, blocks, 0)) return 1;
  *out = st->h_result[0];
  return 0;
}

int b200_collapse(b200_state* st, int target, int outcome, double p, int mode) {
  B200_REQUIRE(st != nullptr, "b200_collapse: null state");
  B200_REQUIRE(target >= 0 && target < st->n, "b200_collapse: target out of range");
  B200_REQUIRE((outcome == 0 || outcome == 1) && (mode == 0 || mode == 1), "b200_collapse: bad outcome/mode");
  B200_CUDA(cudaSetDevice(st->device));
  const uint64_t count = st->dim >> 1;
  // the reference divides by math.sqrt(p) (numpy_backend.py:459,464); multiply by the reciprocal would
  // round differently, so the kernel multiplies by 1/sqrt(p) computed once in double -- within 1 ulp.
  const double s = 1.0 / sqrt(p);
  k_collapse<<<grid_for(count), kThreads, 0, st->stream>>>(st->amp, count, target, outcome, s, mode);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_scale(b200_state* st, double re, double im) {
  B200_REQUIRE(st != nullptr, "b200_scale: null state");
  B200_CUDA(cudaSetDevice(st->device));
  k_scale<<<grid_for(st->dim), kThreads, 0, st->stream>>>(st->amp, st->dim, make_double2(re, im));
  st->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(st->stream));
  return 0;
}

int b200_first_above(b200_state* st, double eps, uint64_t* index, double* re, double* im) {
  B200_REQUIRE(st != nullptr && index != nullptr && re != nullptr && im != nullptr, "b200_first_above: null argument");
  B200_CUDA(cudaSetDevice(st->device));
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(st->d_result + 4);
  B200_CUDA(cudaMemsetAsync(d_out, 0xff, sizeof(unsigned long long), st->stream));
  k_first_above<<<grid_for(st->dim), kThreads, 0, st->stream>>>(st->amp, st->dim, eps * eps, d_out);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  unsigned long long* h_out = reinterpret_cast<unsigned long long*>(st->h_result + 4);
  B200_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  *index = *h_out;
  *re = *im = 0.0;
  if (*h_out != ~0ull) {
    B200_CUDA(cudaMemcpyAsync(st->h_result, st->amp + *h_out, sizeof(c128), cudaMemcpyDeviceToHost, st->stream));
    B200_CUDA(cudaStreamSynchronize(st->stream));
    *re = st->h_result[0];
    *im = st->h_result[1];
  }
  return 0;
}

int b200_first_above_wire(b200_state* st, double eps, const int32_t* wire_of_bit, int n_bits, uint64_t* key) {
  B200_REQUIRE(st != nullptr && wire_of_bit != nullptr && key != nullptr, "b200_first_above_wire: null argument");
  B200_REQUIRE(n_bits >= st->n && n_bits <= 40, "b200_first_above_wire: n_bits must cover the shard and stay below 41");
  WireOfBit wob;
  uint64_t seen = 0;
  for (int b = 0; b < n_bits; ++b) {
    B200_REQUIRE(wire_of_bit[b] >= 0 && wire_of_bit[b] < n_bits && !((seen >> wire_of_bit[b]) & 1),
                 "b200_first_above_wire: wire_of_bit is not a permutation");
    seen |= 1ull << wire_of_bit[b];
    wob.w[b] = (int8_t)wire_of_bit[b];
  }
  B200_CUDA(cudaSetDevice(st->device));
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(st->d_result + 4);
  B200_CUDA(cudaMemsetAsync(d_out, 0xff, sizeof(unsigned long long), st->stream));
  k_first_above_wire<<<reduce_blocks(st->dim), kThreads, 0, st->stream>>>(st->amp, st->dim, st->index_offset, eps * eps, wob,
                                                                         n_bits, d_out);
  st->launches++;
  B200_CUDA(cudaGetLastError());
  unsigned long long* h_out = reinterpret_cast<unsigned long long*>(st->h_result + 4);
  B200_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st->stream));
  B200_CUDA(cudaStreamSynchronize(st->stream));
  *key = *h_out;
  return 0;
}

int b200_expect_group(b200_state* st, uint64_t xmask, int n_terms, const uint64_t* zymask, const double* cre,
                      const double* cim, double* out_re, double* out_im) {
  B200_REQUIRE(st != nullptr && out_re != nullptr && out_im != nullptr, "b200_expect_group: null argument");
  B200_REQUIRE(n_terms >= 0 && (n_terms == 0 || (zymask && cre && cim)), "b200_expect_group: bad terms");
  B200_REQUIRE((xmask & ~(st->dim - 1)) == 0, "b200_expect_group: xmask out of range");
  B200_CUDA(cudaSetDevice(st->device));
  *out_re = *out_im = 0.0;
  const bool tables = st->n >= kExpLoBits && getenv("B200_EXPECT_LOOP") == nullptr;
  const int kChunk = tables ? kExpMaxTerms : 1024;
  for (int base = 0; base < n_terms; base += kChunk) {
    const int nt = std::min(kChunk, n_terms - base);
    if (ensure_capacity(st, nt, 2 * (size_t)nt)) return 1;
    B200_CUDA(cudaStreamSynchronize(st->stream));
    B200_CUDA(cudaMemcpyAsync(st->d_words, zymask + base, sizeof(uint64_t) * nt, cudaMemcpyHostToDevice, st->stream));
    B200_CUDA(cudaMemcpyAsync(st->d_reals, cre + base, sizeof(double) * nt, cudaMemcpyHostToDevice, st->stream));
    B200_CUDA(cudaMemcpyAsync(st->d_reals + nt, cim + base, sizeof(double) * nt, cudaMemcpyHostToDevice, st->stream));
    const int blocks = reduce_blocks(st->dim);
    const size_t smem = (size_t)nt * (sizeof(uint64_t) + 2 * sizeof(double));
    bool has_im = false;
    for (int k = 0; k < nt; ++k) has_im |= cim[base + k] != 0.0;
    if (tables && has_im)
      k_expect_tab<true><<<blocks, kThreads, 0, st->stream>>>(st->amp, st->dim, st->index_offset, xmask, nt, st->d_words,
                                                             st->d_reals, st->d_reals + nt, st->d_partial);
    else if (tables)
      k_expect_tab<false><<<blocks, kThreads, 0, st->stream>>>(st->amp, st->dim, st->index_offset, xmask, nt, st->d_words,
                                                              st->d_reals, st->d_reals + nt, st->d_partial);
    else if (has_im)
      k_expect<true><<<blocks, kThreads, smem, st->stream>>>(st->amp, st->dim, st->index_offset, xmask, nt, st->d_words,
                                                            st->d_reals, st->d_reals + nt, st->d_partial);
    else
      k_expect<false><<<blocks, kThreads, smem, st->stream>>>(st->amp, st->dim, st->index_offset, xmask, nt, st->d_words,
                                                             st->d_reals, st->d_reals + nt, st->d_partial);
    st->launches++;
    B200_CUDA(cudaGetLastError());
    if (finish_to_host(st, blocks, 1)) return 1;
    *out_re += st->h_result[0];
    *out_im += st->h_result[1];
  }
  return 0;
}

int b200_event_record(b200_state* st, int slot) {
  B200_REQUIRE(st != nullptr, "b200_event_record: null state");
  B200_REQUIRE(slot >= 0 && slot < 16, "b200_event_record: slot out of range [0, 16)");
  B200_CUDA(cudaSetDevice(st->device));
  if (!st->user_events[slot]) B200_CUDA(cudaEventCreate(&st->user_events[slot]));
  B200_CUDA(cudaEventRecord(st->user_events[slot], st->stream));
  return 0;
}

int b200_event_elapsed_ms(b200_state* st, int slot_a, int slot_b, double* ms) {
  B200_REQUIRE(st != nullptr && ms != nullptr, "b200_event_elapsed_ms: null argument");
  B200_REQUIRE(slot_a >= 0 && slot_a < 16 && slot_b >= 0 && slot_b < 16, "b200_event_elapsed_ms: slot out of range");
  B200_REQUIRE(st->user_events[slot_a] && st->user_events[slot_b], "b200_event_elapsed_ms: slot never recorded");
  B200_CUDA(cudaSetDevice(st->device));
  B200_CUDA(cudaEventSynchronize(st->user_events[slot_b]));
  float f = 0.f;
  B200_CUDA(cudaEventElapsedTime(&f, st->user_events[slot_a], st->user_events[slot_b]));
  *ms = f;
  return 0;
}

int b200_host_alloc(size_t bytes, void** out) {
  B200_REQUIRE(out != nullptr, "b200_host_alloc: null out");
  B200_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}

int b200_host_free(void* p) {
  if (p) B200_CUDA(cudaFreeHost(p));
  return 0;
}

}  // extern "C"
