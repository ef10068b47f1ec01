// Shared declarations for the b200sv translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/b200sv.h"

namespace b200 {

typedef double2 c128;   // (re, im), 16 bytes: one LDG.128 / STG.128 per amplitude

struct Timing {
  cudaEvent_t start = nullptr, stop = nullptr;
  bool valid = false;
};

}  // namespace b200

struct b200_state {
  int device = 0;
  int n = 0;                       // qubits held in this buffer (a shard when used multi-GPU)
  uint64_t dim = 1;                // 2^n
  uint64_t index_offset = 0;       // index bits above n (the shard's rank bits), seen by controls and phases
  b200::c128* amp = nullptr;       // device amplitudes
  bool owns = true;
  cudaStream_t stream = nullptr;
  // scratch for reductions: per-block partials + result slots (device), mirrored pinned host slots
  double* d_partial = nullptr;     // kMaxPartials * 2 doubles
  double* d_result = nullptr;      // 8 doubles
  double* h_result = nullptr;      // pinned, 8 doubles
  // device copy of the current program payloads (grown on demand)
  uint64_t* d_words = nullptr; size_t cap_words = 0;
  double* d_reals = nullptr;   size_t cap_reals = 0;
  b200::Timing prog;
  bool profiling = false;
  std::vector<cudaEvent_t> op_events;      // 2 per op while profiling
  std::vector<int32_t> op_kinds;
  uint64_t launches = 0;
  cudaEvent_t user_events[16] = {};
  // grow-only scratch of the sampler's reductions (sample.cu): group bits, per-block partials, per-group sums
  void* d_samp = nullptr;      size_t cap_samp = 0;
};

namespace b200 {

constexpr int kSMs = 148;
constexpr int kMaxPartials = 4096;

void set_error(const std::string& msg);
bool cuda_ok(cudaError_t e, const char* what, const char* file, int line);

#define B200_CUDA(call)                                                     \
  do {                                                                      \
    if (!b200::cuda_ok((call), #call, __FILE__, __LINE__)) return 1;        \
  } while (0)

#define B200_REQUIRE(cond, msg)                                             \
  do {                                                                      \
    if (!(cond)) { b200::set_error(msg); return 1; }                        \
  } while (0)

// ---- complex helpers --------------------------------------------------------------------------
__host__ __device__ __forceinline__ c128 cmul(c128 a, c128 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ c128 cfma(c128 a, c128 b, c128 acc) {   // acc + a*b
  return make_double2(acc.x + a.x * b.x - a.y * b.y, acc.y + a.x * b.y + a.y * b.x);
}

// Insert a zero bit at position p of k (bits >= p move up by one).
__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t k, int p) {
  uint64_t lo = k & ((1ull << p) - 1ull);
  return ((k >> p) << (p + 1)) | lo;
}

struct BitPositions {    // ascending positions at which zero bits are inserted
  int n;
  int pos[8];
};

__host__ __device__ __forceinline__ uint64_t spread(uint64_t k, const BitPositions& bp) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < bp.n) k = insert_zero(k, bp.pos[j]);
  return k;
}

// tile.cu
// init_index != kNoInit: the pass starts from the basis state |init_index> instead of reading the buffer
constexpr uint64_t kNoInit = ~0ull;
int launch_tile_pass(b200_state* st, const uint64_t* h_desc, int wlen, const double* h_reals, size_t n_reals,
                     uint64_t init_index);

// permute.cu
int launch_permute(b200_state* st, const uint64_t* pairs, int n_pairs);

}  // namespace b200
