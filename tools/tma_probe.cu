// tma_probe.cu -- can TMA feed a tile pass at HBM speed?  (measurement tool, not part of the product path)
//
// The tile pass of csrc/tile.cu moves 2^12-amplitude tiles (64 KB) whose rows of 2^c amplitudes are contiguous in HBM
// and otherwise scattered over the index space.  This probe runs the data movement of such a pass alone: one
// persistent CTA per SM, a ring of 3 tile buffers in shared memory, one producer warp, and two groups of 256 threads
// that take alternate tiles, make `rounds` round trips shared memory -> 16 registers -> shared memory with `fp`
// dependent-free DFMAs per amplitude in each, and hand the tile back with a TMA store or with 16-byte stores from the
// registers of the last round.
//
// Version 1 of this probe used one cp.async.bulk per 2^c-amplitude row (profiles/r02_bulk_row_probe.json): a B200 SM
// retires one such copy per ~31 cycles whatever its size (256 B rows: 2.5 TB/s, 512 B rows: 4.6 TB/s for load + store),
// so row-granular bulk copies cannot carry 256-byte rows.  This version describes the state to the TMA unit as a
// rank-5 tensor of doubles whose dimensions are the runs of tile bits (each followed by the non-tile bits above it):
// one cp.async.bulk.tensor per tile and direction, SWIZZLE_128B (16-byte chunk ^= 128-byte line, i.e. amplitude slot
// l ^ ((l >> 3) & 7)) for conflict-free quarter-warp accesses.  It checks the data (every amplitude must come back
// with `rounds` added to its real part) and prints GB/s.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
//   tools/tma_probe [n_qubits=30]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s\", \"line\": %d}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int A = 12, TILE = 1 << A, STAGES = 3, GROUP = 256, NG = 2;

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(sptr(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load5(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                          int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(sptr(smem_dst)), "l"(tm), "r"(sptr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store5(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3,
                                           int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(tm), "r"(sptr(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(GROUP) : "memory"); }

__host__ __device__ inline uint32_t swz(uint32_t l) { return l ^ ((l >> 3) & 7u); }

// ---- the tile as a tensor box ----------------------------------------------------------------------------------------
// Field i of the tensor covers index bits [s_i, s_(i+1)): its low `box` bits are tile bits, the rest is fixed per tile.
// Field 0 is (re/im, bits 0..2) = 16 doubles = one 128-byte swizzle line.  More than 5 fields: the top runs of tile bits
// are enumerated by separate copies (`n_sub` per tile, each landing on a contiguous part of the shared-memory tile).
struct TileMap {
  int rank;                 // fields in use (<= 5)
  int shift[5];             // index bit where field i starts (field 0: -1)
  int sub_bits;             // top local bits enumerated by separate copies
  int sub_pos[12];          // their index positions
  uint32_t box_doubles;     // doubles per copy
};

static bool build_tile_map(int n, const int* tile, int a, TileMap& tm, cuuint64_t* gdim, cuuint64_t* gstride,
                           cuuint32_t* box) {
  // chunks of <= 8 contiguous tile bits, starting at bit 3
  int starts[16], lens[16], nch = 0;
  for (int i = 3; i < a;) {
    int j = i + 1;
    while (j < a && tile[j] == tile[j - 1] + 1 && j - i < 8) ++j;
    starts[nch] = tile[i];
    lens[nch++] = j - i;
    i = j;
  }
  if (a < 3 || tile[0] != 0 || tile[1] != 1 || tile[2] != 2) return false;
  const int keep = nch < 4 ? nch : 4;
  tm.rank = 1 + keep;
  tm.sub_bits = 0;
  for (int k = keep; k < nch; ++k)
    for (int b = 0; b < lens[k]; ++b) tm.sub_pos[tm.sub_bits++] = starts[k] + b;
  tm.shift[0] = -1;
  box[0] = 16;
  for (int k = 0; k < keep; ++k) { tm.shift[1 + k] = starts[k]; box[1 + k] = 1u << lens[k]; }
  for (int i = 0; i < tm.rank; ++i) {
    const int lo = tm.shift[i] + 1, hi = (i + 1 < tm.rank ? tm.shift[i + 1] : n) + 1;      // in double-index bits
    gdim[i] = 1ull << (hi - lo);
    if (i > 0) gstride[i - 1] = 8ull << lo;
  }
  for (int i = tm.rank; i < 5; ++i) { gdim[i] = 1; box[i] = 1; gstride[i - 1] = 16ull << n; tm.shift[i] = n; }
  tm.box_doubles = 1;
  for (int i = 0; i < 5; ++i) tm.box_doubles *= box[i];
  return true;
}

struct Params {
  int rounds, fp, direct, n, dbg;
  TileMap map;
  int tile[A];              // index positions of the tile bits
  int regs[4][4];           // per round: local bits held in registers
};

struct Smem {
  uint64_t full[STAGES][2], empty[STAGES];
  uint32_t base[4][GROUP];     // per round, per thread: swizzled slot of its thread bits
  alignas(16) uint32_t delta[4][4];        // per round, per register bit: slot XOR mask
  uint64_t gthr[GROUP];        // direct store: where the thread bits of the last round sit in the index space
  uint64_t gbit[4];
  uint64_t spread[3][1024];    // tile number -> index bits outside the tile, 10 bits per table
};

__global__ void __launch_bounds__(32 + NG * GROUP, 1)
k_probe(double2* __restrict__ amp, uint64_t n_tiles, const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw + STAGES * TILE * 16);
  double2* const stage0 = reinterpret_cast<double2*>(raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s][0], 1); mbar_init(&S.full[s][1], 1); mbar_init(&S.empty[s], GROUP); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = tid; idx < 3 * 1024; idx += blockDim.x) {
    uint64_t x = (uint64_t)(idx & 1023) << (10 * (idx >> 10));
    for (int i = 0; i < A; ++i) { const int pos = p.tile[i]; x = ((x >> pos) << (pos + 1)) | (x & ((1ull << pos) - 1)); }
    S.spread[idx >> 10][idx & 1023] = x;
  }
  if (tid >= 32) {
    const int t = (tid - 32) % GROUP;
    for (int rd = 0; rd < 4; ++rd) {
      // thread bits: the local bits that are not register bits; the first three with distinct bank classes
      // (swz(1 << b) & 7 = 1, 2, 4: bits 0/3, 1/4, 2/5)
      int thr[8], nthr = 0;
      bool used[A] = {};
      for (int k = 0; k < 4; ++k) used[p.regs[rd][k]] = true;
      for (uint32_t want = 1; want <= 4; want <<= 1)
        for (int b = 0; b < A; ++b)
          if (!used[b] && (swz(1u << b) & 7u) == want) { thr[nthr++] = b; used[b] = true; break; }
      for (int b = 0; b < A; ++b)
        if (!used[b]) { thr[nthr++] = b; used[b] = true; }
      uint32_t l = 0;
      for (int b = 0; b < 8; ++b)
        if ((t >> b) & 1) l |= 1u << thr[b];
      S.base[rd][t] = swz(l) | (l << 16);
      if (t < 4) S.delta[rd][t] = swz(1u << p.regs[rd][t]);
      if (rd == (p.rounds + 3) % 4) {
        uint64_t go = 0;
        for (int b = 0; b < A; ++b)
          if ((l >> b) & 1) go |= 1ull << p.tile[b];
        S.gthr[t] = go;
        if (t < 4) S.gbit[t] = 1ull << p.tile[p.regs[rd][t]];
      }
    }
  }
  __syncthreads();
  const uint64_t first = blockIdx.x, step = gridDim.x;
  auto tile_base = [&](uint64_t t) -> uint64_t {
    return S.spread[0][t & 1023] | S.spread[1][(t >> 10) & 1023] | S.spread[2][(t >> 20) & 1023];
  };
  // coordinates of the box of the tile at `base` (sub-copy `sub`)
  auto coords = [&](uint64_t base, int sub, int (&c)[5]) {
    for (int k = 0; k < p.map.sub_bits; ++k)
      if ((sub >> k) & 1) base |= 1ull << p.map.sub_pos[k];
    c[0] = (int)((base & ((1ull << p.map.shift[1]) - 1)) << 1);
#pragma unroll
    for (int i = 1; i < 5; ++i) {
      const int lo = p.map.shift[i], hi = i + 1 < 5 ? p.map.shift[i + 1] : p.n;
      c[i] = lo >= p.n ? 0 : (int)((base >> lo) & ((1ull << (hi - lo)) - 1));
    }
  };
  const int n_sub = 1 << p.map.sub_bits;
  const uint32_t sub_amps = p.map.box_doubles / 2;

  if (tid < 32) {                                   // ---- producer warp
    uint32_t j = 0;
    for (uint64_t tile = first; tile < n_tiles; tile += step, ++j) {
      const int s = j % STAGES;
      const uint32_t ph = (j / STAGES) & 1u;
      mbar_wait(&S.empty[s], ph ^ 1u);
      uint64_t* const fb = &S.full[s][(j / STAGES) & 1u];
      if (p.dbg & 2) { if (tid == 0) mbar_arrive(fb); continue; }
      if (tid == 0) mbar_expect_tx(fb, TILE * 16);
      __syncwarp();
      const uint64_t base = tile_base(tile);
      for (int sub = tid; sub < n_sub; sub += 32) {
        int c[5];
        coords(base, sub, c);
        tma_load5(stage0 + (size_t)s * TILE + (size_t)sub * sub_amps, &tmap, fb, c[0], c[1], c[2], c[3], c[4]);
      }
    }
    return;
  }
  // ---- compute groups
  const int g = (tid - 32) / GROUP, t = (tid - 32) % GROUP;
  uint32_t j = 0;
  for (uint64_t tile = first; tile < n_tiles; tile += step, ++j) {
    if ((int)(j % NG) != g) continue;
    const int s = j % STAGES;
    double2* st = stage0 + (size_t)s * TILE;
    const uint64_t base = tile_base(tile);
    // Parity waits cannot tell phase u from phase u + 2, and are wrong when the barrier is a phase BEHIND the waiter.
    // (1) The threads of a group stay within one tile of each other (group barrier).  (2) Stage s is used alternately
    // by the two groups, so with one `full` barrier per stage a group can get there while the OTHER group's previous use
    // of the stage is still loading (loads complete out of order on a busy GPU): the parity wait then returns at once
    // (the first version of this probe died with 'unspecified launch failure' whenever the consumers were faster than
    // the loads).  Two `full` barriers per stage, taken alternately: each is only ever waited on by ONE group, which
    // consumed its previous phase itself.
    group_sync(g);
    mbar_wait(&S.full[s][(j / STAGES) & 1u], (j / (2 * STAGES)) & 1u);
    double2 v[16];
    bool stored = false;
    for (int rd = 0; rd < p.rounds; ++rd) {
      const uint32_t b0 = S.base[rd & 3][t] & 0xFFFFu, lthr = S.base[rd & 3][t] >> 16;
      uint32_t lreg[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lreg[k] = 1u << p.regs[rd & 3][k];
      const uint4 d = *reinterpret_cast<const uint4*>(S.delta[rd & 3]);
      const uint32_t dl[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t o = b0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((i >> k) & 1) o ^= dl[k];
        v[i] = st[o];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {               // += the amplitude's local index: checks the slot mapping
        uint32_t l = lthr;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((i >> k) & 1) l |= lreg[k];
        v[i].x += (double)l;
      }
      for (int f = 0; f < p.fp; f += 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i].x = fma(v[i].x, 0.999999, 1e-9); v[i].y = fma(v[i].y, 0.999999, 1e-9); }
      }
      if (p.direct && rd == p.rounds - 1) {
        mbar_arrive(&S.empty[s]);                   // (after the values were used: the loads have landed)
        double2* go = amp + base + S.gthr[t];
        const uint64_t gb[4] = {S.gbit[0], S.gbit[1], S.gbit[2], S.gbit[3]};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint64_t o = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((i >> k) & 1) o |= gb[k];
          __stcg(go + o, v[i]);
        }
        stored = true;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint32_t o = b0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((i >> k) & 1) o ^= dl[k];
          st[o] = v[i];
        }
        if (rd < p.rounds - 1) group_sync(g);
      }
    }
    if (!stored) {
      if (p.rounds > 0) { fence_async(); group_sync(g); }
      if (t < n_sub && !(p.dbg & 1)) {
        int c[5];
        coords(base, t, c);
        tma_store5(&tmap, st + (size_t)t * sub_amps, c[0], c[1], c[2], c[3], c[4]);
        bulk_commit();
        bulk_wait_read();
      }
      mbar_arrive(&S.empty[s]);                     // (256 arrivals: the storing threads arrive after their reads)
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void k_fill(double2* amp, uint64_t dim) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * blockDim.x)
    amp[i] = make_double2((double)(i & 0xFFFFFF), (double)(i >> 24));
}
struct TileBits { int pos[A]; };
__global__ void k_verify(const double2* amp, uint64_t dim, double add, TileBits tb, unsigned long long* bad) {
  unsigned long long b = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 a = amp[i];
    uint32_t l = 0;
    for (int k = 0; k < A; ++k) l |= (uint32_t)((i >> tb.pos[k]) & 1ull) << k;
    if (a.x != (double)(i & 0xFFFFFF) + add * (double)l || a.y != (double)(i >> 24)) ++b;
  }
  if (b) atomicAdd(bad, b);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 30;
  const int only = argc > 2 ? atoi(argv[2]) : -1;              // debugging: run one case ...
  const int grid_arg = argc > 3 ? atoi(argv[3]) : 0;           // ... on this many CTAs ...
  const uint64_t tiles_arg = argc > 4 ? (uint64_t)atoll(argv[4]) : 0;   // ... over this many tiles
  const int dbg = argc > 5 ? atoi(argv[5]) : 0;                // 1: no stores, 2: no loads
  const uint64_t dim = 1ull << n, n_tiles = tiles_arg ? tiles_arg : dim >> A;
  double2* amp;
  CHECK(cudaMalloc(&amp, dim * sizeof(double2)));
  unsigned long long* d_bad;
  CHECK(cudaMalloc(&d_bad, 8));
  int sms = 0;
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("{\"error\": \"no cuTensorMapEncodeTiled\"}\n"); return 1; }
  const size_t smem = (size_t)STAGES * TILE * 16 + sizeof(Smem);
  CHECK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  // tile shapes: index positions of the 12 tile bits
  struct Shape { const char* name; int tile[A]; };
  Shape shapes[] = {
      {"rows of 16 (bits 0-3) + top 8 bits", {0, 1, 2, 3, n - 8, n - 7, n - 6, n - 5, n - 4, n - 3, n - 2, n - 1}},
      {"rows of 32 (bits 0-4) + top 7 bits", {0, 1, 2, 3, 4, n - 7, n - 6, n - 5, n - 4, n - 3, n - 2, n - 1}},
      {"contiguous (bits 0-11)", {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}},
      {"rows of 16 + 4 runs of 2 bits", {0, 1, 2, 3, 8, 9, 13, 14, 18, 19, n - 2, n - 1}},
      {"rows of 16 + 8 isolated bits (6 fields: 32 copies per tile)", {0, 1, 2, 3, 6, 9, 12, 15, 18, 21, 24, 27}},
  };
  struct Case { int shape, rounds, fp, direct; CUtensorMapL2promotion l2; };
  const Case cases[] = {
      {0, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE},  {0, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {1, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {2, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {3, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {4, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {0, 1, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {0, 1, 0, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {0, 3, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {0, 3, 0, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {4, 3, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {4, 3, 0, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {1, 3, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {0, 3, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {0, 3, 20, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {0, 4, 24, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {0, 4, 24, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {0, 3, 40, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}, {0, 4, 40, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
  };
  printf("{\"n\": %d, \"sms\": %d, \"smem\": %zu, \"cases\": [\n", n, sms, smem);
  bool firstc = true;
  int case_no = -1;
  for (const Case& cs : cases) {
    ++case_no;
    if (only >= 0 && case_no != only) continue;
    Params p{};
    p.rounds = cs.rounds; p.fp = cs.fp; p.direct = cs.direct; p.n = n; p.dbg = dbg;
    for (int i = 0; i < A; ++i) p.tile[i] = shapes[cs.shape].tile[i];
    // register bits per round: no round holds both bits of a bank class (0/3, 1/4, 2/5)
    const int regs[4][4] = {{0, 1, 2, 6}, {3, 4, 5, 7}, {8, 9, 10, 11}, {0, 4, 8, 9}};
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 4; ++k) p.regs[r][k] = regs[r][k];
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    if (!build_tile_map(n, p.tile, A, p.map, gdim, gstride, box)) { printf("{\"error\": \"tile map\"}\n"); return 1; }
    CUtensorMap tmap;
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, amp, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, cs.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("{\"error\": \"cuTensorMapEncodeTiled %d\", \"gdim\": [%llu,%llu,%llu,%llu,%llu], \"box\": [%u,%u,%u,%u,%u]}\n", (int)r,
             (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2], (unsigned long long)gdim[3],
             (unsigned long long)gdim[4], box[0], box[1], box[2], box[3], box[4]);
      return 1;
    }
    k_fill<<<sms * 8, 256>>>(amp, dim);
    CHECK(cudaMemset(d_bad, 0, 8));
    float best = 1e30f, sum = 0;
    const int reps = 5;
    for (int it = 0; it < 2 + reps; ++it) {
      CHECK(cudaEventRecord(e0));
      k_probe<<<grid_arg ? grid_arg : sms, 32 + NG * GROUP, smem>>>(amp, n_tiles, tmap, p);
      CHECK(cudaEventRecord(e1));
      CHECK(cudaEventSynchronize(e1));
      CHECK(cudaGetLastError());
      float ms;
      CHECK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    long long bad = -1;
    if (cs.fp == 0 && !tiles_arg) {
      TileBits tb;
      for (int i = 0; i < A; ++i) tb.pos[i] = p.tile[i];
      k_verify<<<sms * 8, 256>>>(amp, dim, (double)(cs.rounds * (2 + reps)), tb, d_bad);
      unsigned long long hb = 0;
      CHECK(cudaMemcpy(&hb, d_bad, 8, cudaMemcpyDeviceToHost));
      bad = (long long)hb;
    }
    const double gb = 32.0 * (double)dim / 1e9;
    printf("%s {\"tile\": \"%s\", \"copies_per_tile\": %d, \"l2_promotion\": %d, \"rounds\": %d, \"dfma_per_amp_round\": %d, "
           "\"direct_out\": %d, \"ms_avg\": %.3f, \"ms_best\": %.3f, \"GBps\": %.1f, \"wrong_amplitudes\": %lld}",
           firstc ? "" : ",\n", shapes[cs.shape].name, 1 << p.map.sub_bits, (int)cs.l2, cs.rounds, cs.fp, cs.direct,
           sum / reps, best, gb / (sum / reps) * 1e3, bad);
    firstc = false;
    fflush(stdout);
  }
  printf("\n]}\n");
  return 0;
}
