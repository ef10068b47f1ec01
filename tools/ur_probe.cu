// ur_probe.cu -- what keeps ptxas from using the uniform datapath (LDCU + UR operands of DFMA / DMUL) for a record
// loop whose records and coefficients are kernel parameters.  Compile-only probe (no GPU needed):
//
//   for v in 0 1 2 3 4; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DVARIANT=$v -cubin -o /tmp/u.cubin \
//     tools/ur_probe.cu && cuobjdump -sass /tmp/u.cubin | grep -c 'DFMA.*UR\|DMUL.*UR'; done
//
//   VARIANT 0  no barrier in the tile loop, group number a constant            -> 96 of 96 FP64 instructions take a UR operand
//   VARIANT 1  group number through __reduce_max_sync (REDUX writes a UR)      -> 96
//   VARIANT 2  + `asm volatile("bar.sync ...")` inside the tile loop           -> 0   (also: __syncthreads, __syncwarp,
//   VARIANT 3  both                                                            -> 0    __shfl_*_sync, an EMPTY asm volatile)
//   VARIANT 4  the barrier as a plain `asm` with a "memory" clobber and a dummy result that is consumed -> 96
//
// Further blockers found while converting csrc/tile.cu (each alone turns the whole loop nest back into per-thread code):
// a shared-memory address that mixes a uniform loop counter with the thread index, a shift or XOR of per-thread data by a
// uniform amount, a `switch` (or >= 4 equality tests of one value: the compiler rebuilds the switch) lowered to a jump
// table, a few instructions predicated by a uniform condition, 16-bit parameter loads.  tile.cu keeps everything the
// per-thread side needs in a second stream in shared memory (PassConst::vs) for that reason.
#include <cstdint>
struct RC { uint16_t n_gates, first; uint16_t srb[4]; uint32_t pad; uint64_t o_lo, o_hi; };
struct PP { uint32_t n_rounds; uint32_t pad[7]; RC rh[6]; uint4 rec[161]; double pay[2560]; };
#if VARIANT == 4
__device__ __forceinline__ uint32_t gsync(int g) { uint32_t d; asm("{\n\tbar.sync %1, %2;\n\tmov.u32 %0, 0;\n\t}" : "=r"(d) : "r"(1 + g), "n"(256) : "memory"); return d; }
#else
__device__ __forceinline__ uint32_t gsync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(256) : "memory"); return 0; }
#endif
#ifndef VARIANT
#define VARIANT 0
#endif
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ PP P, double2* __restrict__ out, uint64_t n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
#if (VARIANT & 1) || VARIANT == 4
  const uint32_t g = __reduce_max_sync(0xFFFFFFFFu, tid / 256);
#else
  const uint32_t g = 0;
#endif
  const uint32_t t = tid % 256;
  double2 v[16];
  for (uint32_t j = g;; j += 2) {
    const uint64_t tile = blockIdx.x + (uint64_t)j * gridDim.x;
    if (tile >= n_tiles) break;
    unsigned char* smb = smem + (j % 3) * 65536;
#if VARIANT & 6
    smb += gsync(g);
#endif
    for (int rd = 0; rd < (int)P.n_rounds; ++rd) {
      const int ng = P.rh[rd].n_gates, first = P.rh[rd].first;
      const uint32_t s0 = (uint32_t)P.rh[rd].srb[0] << 4;
      for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<double2*>(smb + ((t * 16 + i) * 16 ^ s0));
      for (int gi = first; gi < first + ng; ++gi) {
        const uint4 q = P.rec[gi];
        const double* pay = P.pay + (q.y & 0xFFFF) * 2;
        if (q.x & 4u) {
          const double z = pay[0], x = pay[1];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            asm("fma.rn.f64 %0, %4, %2, %0;\n\tfma.rn.f64 %1, %4, %3, %1;\n\t"
                "fma.rn.f64 %2, %5, %0, %2;\n\tfma.rn.f64 %3, %5, %1, %3;"
                : "+d"(v[i].x), "+d"(v[i].y), "+d"(v[i + 1].x), "+d"(v[i + 1].y) : "d"(z), "d"(x));
          }
        }
        if (q.x & 0x100u) {
          const double2* tab = reinterpret_cast<const double2*>(P.pay + (q.y >> 16) * 2);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const double2 f = tab[i];
            const double nfy = -f.y;
            asm("{\n\t.reg .f64 t0, t1;\n\t"
                "mul.f64 t0, %1, %4;\n\tmul.f64 t1, %1, %2;\n\tfma.rn.f64 %1, %0, %3, t1;\n\tfma.rn.f64 %0, %0, %2, t0;\n\t}"
                : "+d"(v[i].x), "+d"(v[i].y) : "d"(f.x), "d"(f.y), "d"(nfy));
          }
        }
      }
      for (int i = 0; i < 16; ++i) *reinterpret_cast<double2*>(smb + ((t * 16 + i) * 16 ^ s0)) = v[i];
#if VARIANT & 6
      smb += gsync(g);
#endif
    }
  }
}
