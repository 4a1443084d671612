// tc_ptx.cuh - PTX wrappers shared by the tcgen05 kernels (mbarrier, bulk copies, TMEM, UMMA descriptors).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace tcptx {

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (reported CUDA error), never a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
#ifdef SVAE_MBAR_POLL
  for (unsigned it = 0; !mbar_test_wait(bar, parity); ++it) {
    if (it > (1u << 28)) __trap();
  }
#else
  for (unsigned it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 26)) __trap();
#endif
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, unsigned cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-column variant: v[16..31] are zeroed.  Accumulators whose width is an odd multiple of 16 must not be read with the
// 32-column load: the over-read leaves the CTA's TMEM allocation (and faults when the allocation sits at the top of TMEM,
// which happens as soon as CTAs of different kernels share an SM).
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[32]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
#pragma unroll
  for (int i = 16; i < 32; ++i) v[i] = 0.f;
}
// up to 32 columns starting at taddr, `avail` = columns left in the accumulator (multiple of 16)
__device__ __forceinline__ void tmem_ld_upto32(unsigned taddr, float (&v)[32], int avail) {
  if (avail >= 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
}

// One lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler knows the guarded region runs on a single
// lane and keeps warp-uniform operands (descriptors, TMEM addresses) in uniform registers instead of wrapping every
// tcgen05.mma in a lane-election loop.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp-uniform broadcast of lane 0's value (also tells the compiler the result is uniform)
__device__ __forceinline__ unsigned uniform_u32(unsigned v) { return __shfl_sync(0xffffffffu, v, 0); }

// Shared-memory matrix descriptor, canonical K-major layout without swizzle (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4 ; [16,30) leading byte offset >> 4 (between the two 8-element K groups of one MMA) ;
//   [32,46) stride byte offset >> 4 (between 8-row groups) ; [46,48) version = 1 (sm_100) ; [61,64) layout type 0.
__device__ __forceinline__ unsigned long long make_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at [7,10) /
// [10,13), a/b K-major (0) at 15 / 16, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr unsigned make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// warp-level transpose-reduction: on return lane l holds the sum over the 32 lanes of v[l]
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = up ? v[i] : v[i + off];
      float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&f)[8]) {
  __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
  uint4 r;
  r.x = *reinterpret_cast<unsigned*>(&b0); r.y = *reinterpret_cast<unsigned*>(&b1);
  r.z = *reinterpret_cast<unsigned*>(&b2); r.w = *reinterpret_cast<unsigned*>(&b3);
  return r;
}

}  // namespace tcptx

// ---- TMA tensor loads (cp.async.bulk.tensor, UTMALDG in SASS) -----------------------------------------------------------
#include <cuda.h>
namespace tcptx {
// 2-D tiled load: box (dim0 = c0.., dim1 = c1..) of the tensor described by `tmap` -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* tmap, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(tmap)) : "memory");
}
}  // namespace tcptx
