#include <stdlib.h>
// mma_rate.cu - microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) as a function of N, of the
// shared-memory layout of the operands (no-swizzle K-major with a plane pitch = the "shifted window" layout of
// kernels_tc.cu, no-swizzle MN-major = the wgrad layout, 128-byte swizzle K-major = the TMA layout) and of how many CTAs
// share the SM.  Data is garbage (timing only).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct Args {
  int N, iters, layoutA, layoutB;   // layout: 0 no-swizzle K-major (LBO = pitch, SBO = 128), 1 no-swizzle MN-major, 2 SW128 K-major
  unsigned pitchA, pitchB;
  int rotate;                       // advance the A start address by 16 B each MMA (tap shifts)
  long long* out;
};

__device__ __forceinline__ unsigned long long mk(unsigned addr, unsigned lbo, unsigned sbo, unsigned layout) {
  unsigned long long d = 0;
  d |= (unsigned long long)((addr >> 4) & 0x3FFF);
  d |= (unsigned long long)((lbo >> 4) & 0x3FFF) << 16;
  d |= (unsigned long long)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= (unsigned long long)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128) k(Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ unsigned tbase;
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<unsigned*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tm = tbase;
  if (tid == 0) {
    const unsigned A = smem_u32(smem), B = A + 40960;
    unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(a.N >> 3) << 17) | (8u << 24);
    if (a.layoutA == 1) idesc |= 1u << 15;
    if (a.layoutB == 1) idesc |= 1u << 16;
    unsigned long long da, db;
    if (a.layoutA == 0) da = mk(A, a.pitchA, 128, 0);
    else if (a.layoutA == 1) da = mk(A, 128, a.pitchA, 0);
    else da = mk(A, 16, 1024, 2);
    if (a.layoutB == 0) db = mk(B, a.pitchB, 128, 0);
    else if (a.layoutB == 1) db = mk(B, 128, a.pitchB, 0);
    else db = mk(B, 16, 1024, 2);
    long long t0 = clock64();
    for (int i = 0; i < a.iters; ++i) {
      unsigned long long dai = da + (a.rotate ? (unsigned long long)(i & 15) : 0ull);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(dai), "l"(db), "r"(idesc), "r"(i)
          : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    unsigned ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) { a.out[0] = t1 - t0; a.out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[3] = {"K-major/noswz", "MN-major/noswz", "K-major/sw128"};
  const int iters = 2048;
  // does the stride between the two K core matrices of A (LBO) matter?  (tc2 uses a 128-byte-multiple plane pitch)
  for (unsigned lbo : {4096u, 4112u, 4160u, 3840u, 3856u, 2048u + 1024u})
    for (int N : {32, 128}) {
      Args a{N, iters, 0, 0, lbo, (unsigned)(N * 16), 1, out};
      k<<<148, 128, 96 * 1024, 0>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("LBO_A=%u N=%3d rot=1 : issue %.1f cyc/mma, complete %.1f cyc/mma\n", lbo, N, (double)out[0] / iters, (double)out[1] / iters);
    }
  for (unsigned lbb : {512u, 528u, 2048u, 2064u})
    for (int N : {32, 128}) {
      Args a{N, iters, 0, 0, 4112u, lbb, 1, out};
      if (lbb < (unsigned)N * 16) continue;
      k<<<148, 128, 96 * 1024, 0>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("LBO_B=%u N=%3d rot=1 : issue %.1f cyc/mma, complete %.1f cyc/mma\n", lbb, N, (double)out[0] / iters, (double)out[1] / iters);
    }
  // MN-major operands (the weight-gradient layout): stride between the 8-channel planes (SBO of A / B)
  for (unsigned pa : {2048u, 2064u, 3840u, 3856u})
    for (unsigned pb : {2048u, 2064u})
      for (int N : {32, 128})
        for (int ctas = 1; ctas <= 2; ++ctas) {
          Args a{N, iters, 1, 1, pa, pb, 1, out};
          k<<<148 * ctas, 128, 96 * 1024, 0>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("MN-major pitchA=%u pitchB=%u N=%3d ctas/SM=%d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", pa, pb, N, ctas,
                 (double)out[0] / iters, (double)out[1] / iters);
        }
  if (getenv("MMA_RATE_SHORT")) return 0;
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int la = 0; la < 3; ++la)
      for (int lb = 0; lb < 3; ++lb) {
        if ((la == 1) != (lb == 1)) continue;
        for (int N : {16, 32, 64, 128, 256}) {
          for (int rot = 0; rot < 2; ++rot) {
            if (rot && la != 0) continue;
            Args a{N, iters, la, lb, (unsigned)(la == 1 ? 2064 : 4112), (unsigned)(lb == 1 ? 1040 : N * 16), rot, out};
            k<<<148 * ctas, 128, 96 * 1024, 0>>>(a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("ctas/SM=%d A=%-14s B=%-14s N=%3d rot=%d : issue %.1f cyc/mma, complete %.1f cyc/mma  (floor %d)\n", ctas,
                   names[la], names[lb], N, rot, (double)out[0] / iters, (double)out[1] / iters, 128 * N / 256);
          }
        }
      }
  return 0;
}
