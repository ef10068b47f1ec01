// tma_min.cu -- smallest possible checks of the TMA paths used by tools/tma_probe.cu (debugging aid, not product code).
//   tools/tma_min <test> <bytes>
//     test 0: one cp.async.bulk global -> shared of <bytes>, read back by threads
//     test 1: one rank-5 cp.async.bulk.tensor load of a 64 KB box (SWIZZLE_128B), read back by threads
//     test 2: test 1 + tensor store to a second buffer
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("error: %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_bulk(const double* src, double* dst, uint32_t bytes) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sptr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sptr(raw)), "l"(src), "r"(bytes), "r"(sptr(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(sptr(&bar)), "r"(0) : "memory");
  const double* s = reinterpret_cast<const double*>(raw);
  for (uint32_t i = threadIdx.x; i < bytes / 8; i += blockDim.x) dst[i] = s[i];
}

__global__ void k_tensor(const __grid_constant__ CUtensorMap tmap, double* dst, int do_store,
                         const __grid_constant__ CUtensorMap tmap_out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sptr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(&bar)), "r"(65536) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(sptr(raw)), "l"(&tmap), "r"(sptr(&bar)), "r"(0), "r"(2), "r"(0), "r"(0), "r"(0) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(sptr(&bar)), "r"(0) : "memory");
  const double* s = reinterpret_cast<const double*>(raw);
  for (uint32_t i = threadIdx.x; i < 8192; i += blockDim.x) dst[i] = s[i];
  if (do_store) {
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                   ::"l"(&tmap_out), "r"(sptr(raw)), "r"(0), "r"(2), "r"(0), "r"(0), "r"(0) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 0;
  const uint32_t bytes = argc > 2 ? (uint32_t)atoi(argv[2]) : 512;
  const int n = 20;                                  // 2^20 amplitudes = 2^21 doubles
  const size_t nd = 2ull << n;
  double *src, *dst, *out2;
  CHECK(cudaMalloc(&src, nd * 8));
  CHECK(cudaMalloc(&dst, 65536));
  CHECK(cudaMalloc(&out2, nd * 8));
  CHECK(cudaMemset(out2, 0, nd * 8));
  std::vector<double> h(nd);
  for (size_t i = 0; i < nd; ++i) h[i] = (double)i;
  CHECK(cudaMemcpy(src, h.data(), nd * 8, cudaMemcpyHostToDevice));
  std::vector<double> got(8192);
  if (test == 0) {
    CHECK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k_bulk<<<1, 128, 65536>>>(src, dst, bytes);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(got.data(), dst, bytes, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (uint32_t i = 0; i < bytes / 8; ++i) bad += got[i] != (double)i;
    printf("test 0 bytes %u: ok, %zu wrong\n", bytes, bad);
    return 0;
  }
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  // tile bits {0,1,2,3} + top 8 bits of n = 20: fields (re/im,0..2) | bit 3 + bits 4..11 | bits 12..19
  cuuint64_t gdim[5] = {16, 1ull << 9, 256, 1, 1};
  cuuint64_t gstride[4] = {128, 8ull << 13, 16ull << n, 16ull << n};
  cuuint32_t box[5] = {16, 2, 256, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap tm, tm2;
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      bytes == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = encode(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, out2, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       bytes == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d\n", (int)r, (int)r2);
  if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return 1;
  CHECK(cudaFuncSetAttribute(k_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  k_tensor<<<1, 128, 65536>>>(tm, dst, test == 2, tm2);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(got.data(), dst, 65536, cudaMemcpyDeviceToHost));
  // expected: local index l (12 bits) -> index i = (l & 15) | 1 << 4 (coordinate 2 of field 1 = bit 4 set... field 1 starts
  // at bit 3, so coordinate 2 means index bit 4) | (l >> 4) << 12; smem slot = l ^ ((l >> 3) & 7) with the swizzle
  size_t bad = 0;
  for (uint32_t l = 0; l < 4096; ++l) {
    const uint64_t i = (l & 15u) | (1u << 4) | ((uint64_t)(l >> 4) << 12);
    const uint32_t slot = bytes == 1 ? l : (l ^ ((l >> 3) & 7u));
    bad += got[2 * slot] != (double)(2 * i) || got[2 * slot + 1] != (double)(2 * i + 1);
  }
  printf("test %d: load ok, %zu wrong slots (first doubles: %.0f %.0f %.0f %.0f | %.0f %.0f)\n", test, bad, got[0], got[1],
         got[2], got[3], got[16], got[17]);
  if (test == 2) {
    CHECK(cudaMemcpy(h.data(), out2, nd * 8, cudaMemcpyDeviceToHost));
    size_t badst = 0, nonzero = 0;
    for (size_t d = 0; d < nd; ++d) {
      if (h[d] != 0.0) { ++nonzero; badst += h[d] != (double)d; }
    }
    printf("store: %zu doubles written, %zu wrong\n", nonzero, badst);
  }
  return 0;
}
