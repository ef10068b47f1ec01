// fp64_probe.cu -- how fast are the two FP64 paths of a B200 SM?  (measurement tool, not part of the product path)
//
// BASELINE.json's north star allows a dense k <= 5 unitary applied with FP64 tensor-core MMA (DMMA) "only if it
// beats the CUDA-core path".  A dense 2^k x 2^k complex mat-vec spends 8 * 2^k real flops per amplitude where the
// gate-by-gate register path spends ~16 per 2x2 gate (8 for real matrices), so DMMA has to deliver more than
// 2^k / (2 * gates) times the DFMA rate to win.  This probe measures both rates on the device it runs on.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_probe tools/fp64_probe.cu && tools/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 16 DFMA + `MIX` x 16 integer ops per iteration, all independent: do integer instructions issue in the shadow of
// the half-rate FP64 pipe, or does every DFMA hold the dispatch port for its two cycles?
template <int MIX>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b, unsigned k) {
  double c[16];
  unsigned u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i] = threadIdx.x * 1e-9 + i; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      c[i] = fma(c[i], a, b);
#pragma unroll
      for (int m = 0; m < MIX; ++m) u[i] = (u[i] ^ k) + (u[i] >> 3);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one dependent DFMA chain, one warp: cycles per DFMA = the FP64 pipeline latency seen by a lone warp
__global__ void k_latency(double* out, long long* cycles, int iters, double a, double b) {
  double c = threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c = fma(c, a, b);
  }
  const long long t1 = clock64();
  out[threadIdx.x] = c;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = i; }
  const double av = a + (threadIdx.x & 3) * 1e-12, bv = b + (threadIdx.x >> 2) * 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(av), "d"(bv));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* d;
  CHECK(cudaMalloc(&d, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  float ms_fma = 0, ms_mma = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CHECK(cudaEventRecord(e0));
    k_dfma<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms_fma, e0, e1));
    CHECK(cudaEventRecord(e0));
    k_dmma<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms_mma, e0, e1));
  }
  float ms_mix1 = 0, ms_mix3 = 0;
  CHECK(cudaEventRecord(e0));
  k_mix<1><<<blocks, threads>>>(d, iters, 0.999999, 1e-9, 12345u);
  CHECK(cudaEventRecord(e1));
  CHECK(cudaEventSynchronize(e1));
  CHECK(cudaEventElapsedTime(&ms_mix1, e0, e1));
  CHECK(cudaEventRecord(e0));
  k_mix<3><<<blocks, threads>>>(d, iters, 0.999999, 1e-9, 12345u);
  CHECK(cudaEventRecord(e1));
  CHECK(cudaEventSynchronize(e1));
  CHECK(cudaEventElapsedTime(&ms_mix3, e0, e1));
  long long* d_cyc;
  CHECK(cudaMalloc(&d_cyc, sizeof(long long)));
  k_latency<<<1, 32>>>(d, d_cyc, 4096, 0.999999, 1e-9);
  long long h_cyc = 0;
  CHECK(cudaMemcpy(&h_cyc, d_cyc, sizeof h_cyc, cudaMemcpyDeviceToHost));
  const double dfma_latency = (double)h_cyc / (4096.0 * 16);
  CHECK(cudaGetLastError());
  const double flops_fma = 2.0 * 16 * iters * (double)blocks * threads;
  const double flops_mma = 512.0 * 8 * iters * (double)blocks * (threads / 32);
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f, "
         "\"dfma_plus_3_int_ops_each_ms\": %.3f, \"dfma_plus_9_int_ops_each_ms\": %.3f, "
         "\"dfma_dependent_latency_cycles\": %.2f}\n",
         prop.name, prop.multiProcessorCount, flops_fma / ms_fma / 1e9, flops_mma / ms_mma / 1e9, ms_fma, ms_mma, ms_mix1,
         ms_mix3, dfma_latency);
  return 0;
}
