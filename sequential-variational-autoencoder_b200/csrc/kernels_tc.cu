// kernels_tc.cu - tcgen05 / TMEM "shifted-window" implicit-GEMM kernels for the 4x4 convolutions and transposed
// convolutions of the chain (bf16 operands, fp32 accumulation in tensor memory).
//
// Idea (B200-first, not a cuDNN-style im2col): for a 4x4 window every tap reads the SAME input pixels, only shifted.
// The kernel walks the batch in a zero-padded linear pixel space q = (n*Hp + r)*Wp + c whose padding columns / rows are
// shared between neighbouring rows / images (wrap-around), so that every tap is a constant offset "q + shift".  A CTA
//   1. loads the halo [q0 - lo, q0 + 128 + hi) of its 128-row tile ONCE (fp32 NHWC -> bf16, zero fill), into a planar
//      shared-memory layout  [8-channel chunk][pixel][16 B]  - exactly the canonical no-swizzle K-major UMMA layout with
//      SBO = 128 B, LBO = plane pitch - in which "shift by s pixels" is "start address + 16*s";
//   2. streams the pre-packed bf16 weights of each (channel chunk, tap) with one cp.async.bulk (TMA 1-D bulk copy) per
//      stage through an mbarrier ring;
//   3. issues tcgen05.mma (M=128, N=Cout, K=16) per tap and 16-channel slice from ONE thread, accumulating in TMEM;
//   4. drains TMEM with tcgen05.ld, writes fp32 NHWC (channel-window aware, optional accumulate) and reduces the
//      per-channel batch-norm statistics (sum, sum of squares) in the epilogue.
// Stride-2 convs read four parity planes; stride-2 transposed convs are four output phases = four accumulators.
//
// Reference ops replaced: Conv2D / Conv2DBackpropInput as launched for conv2d_bn_lrelu, conv2d_t_bn(_relu) and their
// input gradients (abstract_network.py:18,37,56; sequential_vae.py:1273).
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int MAX_STAGES = 4;
constexpr int W2_STAGES_MAX = 6;   // weight ring of the TMA-fed kernel: streaming layers are bound by the bytes one CTA keeps in flight

struct TcParams {
  const float* in; int in_ld, in_coff, cin_valid, in_vec;
  const __nv_bfloat16* wp;
  float* out; int out_ld, out_coff, n_valid, out_vec;
  double* stats;
  unsigned long long* dbg;   // optional per-CTA phase timestamps (globaltimer ns): [cta][8]
  int B, Hin, Win, Hout, Wout;
  int Cin_p, N_p;   // channel counts padded to multiples of 16 (zero weights / zero activations in the padding)
  int mode;         // 0: stride-1 window (or 1x1 = fully connected), 1: stride-2 gather (4 parity planes), 2: stride-2 phases
  int Hp, Wp, Hv, Wv;
  int lo, HL, HLpad;
  int KC, NC, JC;   // channels per chunk, number of chunks, 8-channel groups per chunk
  int nplanes, nacc, ntaps;
  int tps;          // taps per weight stage (one bulk copy brings tps consecutive taps of one channel chunk)
  int accumulate;
  long long Q;      // padded positions
  int stages, a_bufs;
  unsigned a_buf_bytes, b_stage_max, tmem_cols;
  signed char acc[16], plane[16];
  int shift[16];
};

using namespace tcptx;
}
extern void* g_tc_debug_buffer;
namespace {


struct SmemHeader {
  unsigned long long full_b[MAX_STAGES], empty_b[MAX_STAGES], a_ready[2], a_free[2], acc_done;
  unsigned tmem_base;
  unsigned pad;
  float s_sum[4][128], s_sq[4][128];   // per-warp partial column statistics of one accumulator
  volatile unsigned long long ts[16];  // debug timestamps
  unsigned tap_a_lo[16];               // per tap: low word of the A descriptor (start address of the shifted window >> 4)
  unsigned tap_dtmem[16];              // per tap: TMEM address of its accumulator
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
#ifdef SVAE_DBG_CLOCK64
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
#else
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#endif
  return t;
}
// timestamps go to shared memory while the kernel runs (a global store ahead of fence.proxy.async would stall it) and are
// flushed to global once at the very end
// (executed by the whole warp, lane 0's store is predicated: an `if (tid == 0)` wrapper would split lane 0 from its warp
//  and serialise the two groups through everything that follows)
#define DBG_MARK(slot) do { if (P.dbg != nullptr) { const unsigned long long t_ = gtimer(); if ((threadIdx.x & 31) == 0) hdr->ts[slot] = t_; } } while (0)

// 8 consecutive channels [ch0, ch0+8) of one pixel -> 8 bf16; channels >= valid are zero (channel padding)
__device__ __forceinline__ uint4 load8(const float* __restrict__ px, int ch0, int valid, int vec) {
  float f[8];
  if (vec) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(px + ch0));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(px + ch0) + 1);
    f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (ch0 + e < valid) ? __ldg(px + ch0 + e) : 0.f;
  }
  return pack8_bf16(f);
}

// Raw (fp32) 8-channel group: loads are issued first for a whole batch of items so that several are in flight per thread.
struct Raw8 { float4 a, b; };
__device__ __forceinline__ Raw8 load8_raw(const float* __restrict__ px, int ch0, int valid, int vec) {
  Raw8 r;
  if (vec) {
    r.a = __ldg(reinterpret_cast<const float4*>(px + ch0));
    r.b = __ldg(reinterpret_cast<const float4*>(px + ch0) + 1);
  } else {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (ch0 + e < valid) ? __ldg(px + ch0 + e) : 0.f;
    r.a = make_float4(f[0], f[1], f[2], f[3]); r.b = make_float4(f[4], f[5], f[6], f[7]);
  }
  return r;
}
__device__ __forceinline__ uint4 pack_raw(const Raw8& r) {
  const float f[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
  return pack8_bf16(f);
}

// Stage one halo (or plain 128-row tile when lo = hi = 0) of an NHWC fp32 tensor into the planar bf16 layout
//   dst[(plane*J + j) * pitch + i*16]   for i in [0, HL), j in [0, J), plane in [0, nplanes)
// q = q_first + i walks the padded pixel space (Hp x Wp per image, valid window Hv x Wv, `sm` = parity-plane stride).
// 32-bit index math (Q < 2^31 is checked on the host) and 4 items per thread in flight.
__device__ __forceinline__ void stage_halo(unsigned char* dst, unsigned pitch, const float* __restrict__ src, int ld, int coff,
                                           int ch_base, int ch_valid, int vec, int HL, int J, int nplanes, int sm,
                                           int q_first, int Q, int Hp, int Wp, int Hv, int Wv, int Hsrc, int Wsrc, int tid,
                                           int nthreads) {
  const int per_plane = HL * J;
  const int items = per_plane * nplanes;
  // Branch-free body: every lane computes a (clamped, always legal) address and a validity flag, ALL loads of the batch
  // are issued before the first use, invalid items are zeroed by a select.  Divergent "valid / zero" paths would make
  // the two lane groups of a warp run the loop one after the other.
  constexpr int U = 8;
  for (int base = tid; base < items; base += nthreads * U) {
    Raw8 raw[U];
    unsigned off[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = min(base + u * nthreads, items - 1);
      const int pl = it / per_plane;
      const int rem = it - pl * per_plane;
      const int i = rem / J, j = rem - i * J;
      off[u] = (unsigned)(pl * J + j) * pitch + (unsigned)i * 16u;
      const int q = q_first + i;
      const int qc = min(max(q, 0), Q - 1);
      const int t = qc / Wp;
      const int cc = qc - t * Wp;
      const int n = t / Hp;
      const int r = t - n * Hp;
      ok[u] = (q >= 0) & (q < Q) & (r < Hv) & (cc < Wv);
      const int ih = min(r, Hv - 1) * sm + (pl >> 1), iw = min(cc, Wv - 1) * sm + (pl & 1);
      const float* px = src + ((size_t)(n * Hsrc + ih) * Wsrc + iw) * ld + coff;
      raw[u] = load8_raw(px, ch_base + j * 8, ch_valid, vec);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthreads < items) {
        uint4 v = pack_raw(raw[u]);
        if (!ok[u]) v = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dst + off[u]) = v;
      }
    }
  }
}

// 6 warps: 0-3 halo producers then epilogue, 4 weight loader (TMA bulk), 5 MMA issuer + TMEM allocator.
// blockIdx.x = 128-row tile of the padded pixel space, blockIdx.y = 128-column tile of the output channels.
__global__ void __launch_bounds__(192, 3) tc_conv_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem_raw);
  unsigned char* b_smem = smem_raw + ((sizeof(SmemHeader) + 127) & ~127u);
  unsigned char* a_smem = b_smem + (size_t)P.stages * P.b_stage_max;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q0 = (long long)blockIdx.x * TILE_M;
  const int n0 = blockIdx.y * 128;
  if (warp == 0) DBG_MARK(0);
  const int Nt = min(128, P.N_p - n0);               // this CTA's MMA N (multiple of 16)
  const unsigned b_tap_bytes = (unsigned)(P.KC * Nt * 2);
  const unsigned b_stage_bytes = b_tap_bytes * (unsigned)P.tps;

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(smem_u32(&hdr->full_b[s]), 1); mbar_init(smem_u32(&hdr->empty_b[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&hdr->a_ready[i]), 128); mbar_init(smem_u32(&hdr->a_free[i]), 1); }
    mbar_init(smem_u32(&hdr->acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(smem_u32(&hdr->tmem_base), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = hdr->tmem_base;
  const unsigned LBO_A = (unsigned)P.HLpad * 16u;   // bytes between consecutive 8-channel planes of the halo
  if (warp == 0) DBG_MARK(1);

  if (warp < 4) {
    // ================= halo producers: fp32 NHWC global -> bf16 planar smem (once per channel chunk) =================
    const int sm = P.mode == 1 ? 2 : 1;
    for (int c = 0; c < P.NC; ++c) {
      const int buf = c % P.a_bufs;
      if (c >= P.a_bufs) mbar_wait(smem_u32(&hdr->a_free[buf]), ((c / P.a_bufs) - 1) & 1);
      unsigned char* abuf = a_smem + (size_t)buf * P.a_buf_bytes;
      stage_halo(abuf, LBO_A, P.in, P.in_ld, P.in_coff, c * P.KC, P.cin_valid, P.in_vec, P.HL, P.JC, P.nplanes, sm,
                 (int)q0 - P.lo, (int)P.Q, P.Hp, P.Wp, P.Hv, P.Wv, P.Hin, P.Win, tid, 128);
      if (warp == 0) DBG_MARK(14);
#ifndef SVAE_NO_PROXY_FENCE
      fence_proxy_async();                           // generic-proxy smem writes -> visible to the tensor core (async proxy)
#endif
#ifdef SVAE_NAMED_BAR
      asm volatile("bar.arrive 2, 160;" ::: "memory");   // producers (128) arrive, the MMA warp (32) syncs
#else
      mbar_arrive(smem_u32(&hdr->a_ready[buf]));
#endif
    }
    if (warp == 0) DBG_MARK(2);
    DBG_MARK(8 + warp);

    // ================= epilogue: TMEM -> registers -> global (+ batch-norm statistics) =================
    mbar_wait(smem_u32(&hdr->acc_done), 0);
    tc_fence_after();
    if (warp == 0) DBG_MARK(3);
    const int m = warp * 32 + lane;                 // accumulator row == TMEM lane
    const long long q = q0 + m;
    bool valid = q < P.Q;
    int n = 0, r = 0, cc = 0;
    if (valid) {
      cc = (int)(q % P.Wp);
      const long long t = q / P.Wp;
      r = (int)(t % P.Hp);
      n = (int)(t / P.Hp);
      valid = r < P.Hv && cc < P.Wv;
    }
    for (int a = 0; a < P.nacc; ++a) {
      int oh = r, ow = cc;
      if (P.mode == 2) { oh = 2 * r + (a >> 1); ow = 2 * cc + (a & 1); }
      float* orow = P.out + (((size_t)n * P.Hout + oh) * P.Wout + ow) * P.out_ld + P.out_coff + n0;
      for (int nn = 0; nn < Nt; nn += 32) {
        float v[32];
        tmem_ld_upto32(tmem_base + ((unsigned)(warp * 32) << 16) + (unsigned)(a * Nt + nn), v, Nt - nn);
        const int ncols = max(0, min(min(32, Nt - nn), P.n_valid - (n0 + nn)));   // real (unpadded) columns of this group
        if (valid) {
          if (P.out_vec) {
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              if (k < ncols) {
                float4 o = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                float4* dst = reinterpret_cast<float4*>(orow + nn + k);
                if (P.accumulate) { float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                *dst = o;
                v[k] = o.x; v[k + 1] = o.y; v[k + 2] = o.z; v[k + 3] = o.w;
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              if (k < ncols) {
                float o = v[k];
                if (P.accumulate) o += orow[nn + k];
                orow[nn + k] = o;
                v[k] = o;
              }
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0.f;
        }
        if (P.stats != nullptr) {
          float sq[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) { if (k >= ncols) v[k] = 0.f; sq[k] = v[k] * v[k]; }
          const float cs = warp_colsum32(v, lane);
          const float cq = warp_colsum32(sq, lane);
          // accumulate this warp's column totals over accumulators (phases share channels)
          if (a == 0) { hdr->s_sum[warp][nn + lane] = cs; hdr->s_sq[warp][nn + lane] = cq; }
          else { hdr->s_sum[warp][nn + lane] += cs; hdr->s_sq[warp][nn + lane] += cq; }
        }
      }
    }
    tc_fence_before();
    if (warp == 0) DBG_MARK(4);
    if (P.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      for (int col = tid; col < Nt; col += 128) {
        if (n0 + col < P.n_valid) {
          const float s = hdr->s_sum[0][col] + hdr->s_sum[1][col] + hdr->s_sum[2][col] + hdr->s_sum[3][col];
          const float s2 = hdr->s_sq[0][col] + hdr->s_sq[1][col] + hdr->s_sq[2][col] + hdr->s_sq[3][col];
          atomicAdd(&P.stats[n0 + col], (double)s);
          atomicAdd(&P.stats[P.n_valid + n0 + col], (double)s2);
        }
      }
    }
  } else if (warp == 4) {
    // ================= weight loader: one bulk copy per (channel chunk, tap) through the stage ring =================
    // (the whole warp walks the ring and waits converged; one lane issues - a lone lane spinning while its 31
    //  siblings sit in the final bar.sync is scheduled poorly)
    int stage = 0; unsigned phase = 0;
    // packed weights: [n tile][chunk][tap][8-channel group][n][8]; every full tile holds 128*Cin_p*ntaps elements
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(P.wp) + (size_t)blockIdx.y * 128 * P.Cin_p * P.ntaps * 2;
    for (int c = 0; c < P.NC; ++c)
      for (int s = 0; s < P.ntaps; s += P.tps) {
        mbar_wait(smem_u32(&hdr->empty_b[stage]), phase ^ 1);
        if (lane == 0) {
          mbar_expect_tx(smem_u32(&hdr->full_b[stage]), b_stage_bytes);
          bulk_g2s(smem_u32(b_smem + (size_t)stage * P.b_stage_max), wsrc + ((size_t)c * P.ntaps + s) * b_tap_bytes,
                   b_stage_bytes, smem_u32(&hdr->full_b[stage]));
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
  } else {
    // ================= MMA issuer (warp converged on the waits, one elected lane issues) =================
    // The issue loop must be a handful of instructions per MMA (a single thread feeds the tensor core): everything
    // that depends only on the tap is computed once, by 16 lanes in parallel, while the producers are still loading.
    const unsigned idesc = make_idesc(TILE_M, Nt);
    const unsigned LBO_B = (unsigned)Nt * 16u;
    if (lane < P.ntaps) {
      hdr->tap_a_lo[lane] = ((unsigned)(P.plane[lane] * P.JC) * LBO_A + (unsigned)(P.lo + P.shift[lane]) * 16u) >> 4;
      hdr->tap_dtmem[lane] = tmem_base + (unsigned)(P.acc[lane] * Nt);
    }
    __syncwarp();
    // descriptor high words are tap-independent: LBO | SBO | version
    const unsigned long long a_hi = ((unsigned long long)((128u >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((LBO_A >> 4) & 0x3FFF) << 16);
    const unsigned long long b_hi = ((unsigned long long)((128u >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((LBO_B >> 4) & 0x3FFF) << 16);
    const unsigned a_kstep = (2u * LBO_A) >> 4, b_kstep = (2u * LBO_B) >> 4;   // one 16-channel slice further
    const int nk = P.KC / 16;
    int stage = 0; unsigned phase = 0;
    unsigned started = 0;                           // bit a set once accumulator a has been written
    DBG_MARK(12);
    for (int c = 0; c < P.NC; ++c) {
      const int buf = c % P.a_bufs;
      mbar_wait(smem_u32(&hdr->a_ready[buf]), (c / P.a_bufs) & 1);
      if (c == 0) DBG_MARK(6);
      const unsigned abase4 = smem_u32(a_smem + (size_t)buf * P.a_buf_bytes) >> 4;
      for (int s0 = 0; s0 < P.ntaps; s0 += P.tps) {
        mbar_wait(smem_u32(&hdr->full_b[stage]), phase);
        tc_fence_after();
        if (lane == 0) {
          unsigned b_lo = smem_u32(b_smem + (size_t)stage * P.b_stage_max) >> 4;
          for (int tl = 0; tl < P.tps; ++tl) {
            const int s = s0 + tl;
            const unsigned d_tmem = hdr->tap_dtmem[s];
            const unsigned a = (unsigned)P.acc[s];
            unsigned a_lo = abase4 + hdr->tap_a_lo[s];
            unsigned bk = b_lo;
            unsigned acc_flag = (started >> a) & 1u;
#pragma unroll 4
            for (int kk = 0; kk < nk; ++kk) {
              umma_bf16(d_tmem, a_hi | (unsigned long long)(a_lo & 0x3FFF), b_hi | (unsigned long long)(bk & 0x3FFF), idesc,
                        acc_flag);
              acc_flag = 1u;
              a_lo += a_kstep;
              bk += b_kstep;
            }
            started |= 1u << a;
            b_lo += b_tap_bytes >> 4;
          }
          umma_commit(smem_u32(&hdr->empty_b[stage]));   // frees the weight stage once these MMAs have read it
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) umma_commit(smem_u32(&hdr->a_free[buf]));   // halo buffer reusable
      __syncwarp();
    }
    if (lane == 0) umma_commit(smem_u32(&hdr->acc_done));
    __syncwarp();
    DBG_MARK(7);
  }
  __syncthreads();
  if (warp == 0) DBG_MARK(5);
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
  if (P.dbg != nullptr && tid < 16) P.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + tid] = hdr->ts[tid];
}

static inline int round16(int v) { return (v + 15) / 16 * 16; }
static inline int pick_kc(int cin_p) { return cin_p % 128 == 0 ? 128 : cin_p % 64 == 0 ? 64 : cin_p % 32 == 0 ? 32 : 16; }

// weights -> bf16 [n tile][chunk c][tap s][8-channel group j][n][8]  (canonical K-major, no swizzle: LBO = Nt*16, SBO = 128);
// zero in the channel padding
__global__ void tc_pack_kernel(Geom g, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int KC, int Cin_p,
                               int N_p, int TW) {
  const int JC = KC / 8;
  const int ntaps = g.KH * g.KW;
  const long long per_tile = (long long)TW * Cin_p * ntaps;
  const long long total = (long long)N_p * Cin_p * ntaps;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int tile = (int)(e / per_tile);
    long long t = e - tile * per_tile;
    const int Nt = min(TW, N_p - tile * TW);
    const int k8 = (int)(t % 8); t /= 8;
    const int n = (int)(t % Nt); t /= Nt;
    const int j = (int)(t % JC); t /= JC;
    const int s = (int)(t % ntaps); t /= ntaps;
    const int c = (int)t;
    const int ci = c * KC + j * 8 + k8, co = tile * TW + n;
    float v = 0.f;
    if (ci < g.Cin && co < g.Cout)
      v = g.w_out_major == 0 ? w[((size_t)s * g.Cin + ci) * g.Cout + co] : w[((size_t)s * g.Cout + co) * g.Cin + ci];
    out[e] = __float2bfloat16_rn(v);
  }
}

__device__ __forceinline__ void pack_one(const Geom& g, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                         int KC, int Cin_p, int N_p, int TW, long long e, const float* __restrict__ w2 = nullptr,
                                         int n1 = 0) {
  const int JC = KC / 8;
  const int ntaps = g.KH * g.KW;
  const long long per_tile = (long long)TW * Cin_p * ntaps;
  const int tile = (int)(e / per_tile);
  long long t = e - tile * per_tile;
  const int Nt = min(TW, N_p - tile * TW);
  const int k8 = (int)(t % 8); t /= 8;
  const int n = (int)(t % Nt); t /= Nt;
  const int j = (int)(t % JC); t /= JC;
  const int s = (int)(t % ntaps); t /= ntaps;
  const int c = (int)t;
  const int ci = c * KC + j * 8 + k8, co = tile * TW + n;
  float v = 0.f;
  if (ci < g.Cin && co < g.Cout) {
    if (w2 == nullptr) {
      v = g.w_out_major == 0 ? w[((size_t)s * g.Cin + ci) * g.Cout + co] : w[((size_t)s * g.Cout + co) * g.Cin + ci];
    } else if (g.w_out_major == 1) {   // forward form: [tap][co][ci], the two tensors split the co range
      v = co < n1 ? w[((size_t)s * n1 + co) * g.Cin + ci] : w2[((size_t)s * (g.Cout - n1) + (co - n1)) * g.Cin + ci];
    } else {                           // input-gradient form: [tap][ci][co], the two tensors split the ci range
      v = ci < n1 ? w[((size_t)s * n1 + ci) * g.Cout + co] : w2[((size_t)s * (g.Cin - n1) + (ci - n1)) * g.Cout + co];
    }
  }
  out[e] = __float2bfloat16_rn(v);
}

// every tensor-core layer's operand copy in one launch: blockIdx.y = table entry
__global__ void tc_pack_batched_kernel(const TcPackEntry* __restrict__ tab) {
  const TcPackEntry E = tab[blockIdx.y];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E.total; e += (long long)gridDim.x * blockDim.x)
    pack_one(E.g, E.w, reinterpret_cast<__nv_bfloat16*>(E.out), E.KC, E.Cin_p, E.N_p, E.TW, e, E.w2, E.n1);
}

bool build_params(const Geom& g, TcParams& P) {
  memset(&P, 0, sizeof P);
  P.B = g.B; P.Hin = g.Hin; P.Win = g.Win; P.Hout = g.Hout; P.Wout = g.Wout;
  P.cin_valid = g.Cin; P.n_valid = g.Cout;
  P.Cin_p = round16(g.Cin); P.N_p = round16(g.Cout);
  P.accumulate = g.accumulate;
  P.KC = pick_kc(P.Cin_p); P.NC = P.Cin_p / P.KC; P.JC = P.KC / 8;
  int hi;
  if (g.KH == 1) {            // fully connected = 1x1 window over a 1x1 "image"
    if (g.KW != 1 || g.Hin != 1 || g.Win != 1 || g.Hout != 1 || g.Wout != 1) return false;
    P.mode = 0; P.nplanes = 1; P.nacc = 1; P.ntaps = 1;
    P.Hp = P.Wp = P.Hv = P.Wv = 1;
    P.acc[0] = 0; P.plane[0] = 0; P.shift[0] = 0;
    P.lo = 0; hi = 0;
  } else if (g.stride == 1) {
    P.mode = 0; P.nplanes = 1; P.nacc = 1; P.ntaps = 16;
    P.Hp = g.Hin + 2; P.Wp = g.Win + 2; P.Hv = g.Hin; P.Wv = g.Win;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        const int s = kh * 4 + kw;
        const int dh = g.mode == 0 ? kh - 1 : 1 - kh, dw = g.mode == 0 ? kw - 1 : 1 - kw;
        P.acc[s] = 0; P.plane[s] = 0; P.shift[s] = dh * P.Wp + dw;
      }
    P.lo = g.mode == 0 ? P.Wp + 1 : 2 * P.Wp + 2;
    hi = g.mode == 0 ? 2 * P.Wp + 2 : P.Wp + 1;
  } else if (g.mode == 0) {   // stride-2 conv: parity planes of the input, output grid
    P.mode = 1; P.nplanes = 4; P.nacc = 1; P.ntaps = 16;
    P.Hp = g.Hout + 1; P.Wp = g.Wout + 1; P.Hv = g.Hout; P.Wv = g.Wout;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        const int s = kh * 4 + kw;
        const int dh = kh == 0 ? -1 : kh == 3 ? 1 : 0, dw = kw == 0 ? -1 : kw == 3 ? 1 : 0;
        P.acc[s] = 0; P.plane[s] = (signed char)((((kh + 1) & 1) << 1) | ((kw + 1) & 1)); P.shift[s] = dh * P.Wp + dw;
      }
    P.lo = P.Wp + 1; hi = P.Wp + 1;
  } else {                    // stride-2 transposed conv: input grid, four output phases
    P.mode = 2; P.nplanes = 1; P.nacc = 4; P.ntaps = 16;
    P.Hp = g.Hin + 1; P.Wp = g.Win + 1; P.Hv = g.Hin; P.Wv = g.Win;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        const int s = kh * 4 + kw;
        const int ph = (kh + 1) & 1, pw = (kw + 1) & 1;
        const int dh = kh == 3 ? -1 : kh == 0 ? 1 : 0, dw = kw == 3 ? -1 : kw == 0 ? 1 : 0;
        P.acc[s] = (signed char)(ph * 2 + pw); P.plane[s] = 0; P.shift[s] = dh * P.Wp + dw;
      }
    P.lo = P.Wp + 1; hi = P.Wp + 1;
    if (P.N_p > 128) return false;                 // four accumulators of one N tile
  }
  P.HL = TILE_M + P.lo + hi;
  P.HLpad = P.HL | 1;                               // odd plane pitch (in 16-byte units): conflict-free producer stores
  P.Q = (long long)g.B * P.Hp * P.Wp;
  if (P.Q + TILE_M + P.HL >= (1ll << 31) || (long long)g.B * g.Hin * g.Win >= (1ll << 31)) return false;   // 32-bit pixel indices
  P.a_buf_bytes = (unsigned)(P.nplanes * P.JC * P.HLpad * 16);
  P.a_bufs = P.NC > 1 ? 2 : 1;
  const int nt_max = P.N_p < 128 ? P.N_p : 128;
  // Weight stages: one bulk copy brings `tps` consecutive taps (<= 32 KB), so that small layers get all their weights in
  // one or two copies instead of 16 latency-bound round trips; <= 64 KB of weight stages keeps >= 2 CTAs per SM.
  const unsigned tap_bytes = (unsigned)(P.KC * nt_max * 2);
  P.tps = 1;
  while (P.tps * 2 <= P.ntaps && (unsigned)(P.tps * 2) * tap_bytes <= 32 * 1024) P.tps *= 2;
  P.b_stage_max = tap_bytes * (unsigned)P.tps;
  {
    const int loads = P.NC * (P.ntaps / P.tps);
    int st = (int)((64 * 1024) / P.b_stage_max);
    if (st < 2) st = 2;
    if (st > MAX_STAGES) st = MAX_STAGES;
    if (st > loads) st = loads;
    P.stages = st;
    const size_t fixed = ((sizeof(SmemHeader) + 127) & ~(size_t)127) + (size_t)P.a_bufs * P.a_buf_bytes + 128;
    while (P.stages > 1 && fixed + (size_t)P.stages * P.b_stage_max > 227 * 1024) --P.stages;
    if (fixed + (size_t)P.stages * P.b_stage_max > 227 * 1024) return false;
  }
  unsigned cols = (unsigned)(P.nacc * nt_max), t = 32;
  while (t < cols) t <<= 1;
  P.tmem_cols = t;
  return t <= 512;
}

size_t smem_bytes(const TcParams& P) {
  return ((sizeof(SmemHeader) + 127) & ~(size_t)127) + (size_t)P.stages * P.b_stage_max + (size_t)P.a_bufs * P.a_buf_bytes + 128;
}

// =====================================================================================================================
// Weight gradient:  dW[tap][a][b] = sum_q X[q + shift(tap)][a] * dY[q][b]   (conv-gather geometry, q = padded positions)
// Both operands live in the same planar shared-memory layout as above, now consumed as MN-major matrices (the reduction
// dimension K is the pixel index): D[a, b] += A[a x 16 px] * B[16 px x b] with M = 128 input channels (rows beyond the
// real channel count read neighbouring planes and are ignored), N = output-channel block, one TMEM accumulator per tap
// of the CTA's tap group.  A CTA walks its share of the 128-pixel tiles, accumulating in TMEM the whole time, and
// flushes once with red.global.add.f32.
struct TwParams {
  const float* x; int x_ld, x_coff, x_vec;
  const float* dy; int dy_ld, dy_coff, dy_vec;
  float* dw; int dw_vec;
  float* dw2; int a_split;  // rows (X channels) >= a_split go to dw2 [16][Ca - a_split][Cb] (two parameter tensors, one contraction)
  int B, Hx, Wx, Ca, Hy, Wy, Cb;   // Ca, Cb: real channel counts (dw is [16][Ca][Cb]); padded counts are mblocks*CaB, nblocks*N
  int mode;                 // 0 stride 1 (one plane), 1 stride 2 (four parity planes of X)
  int Hp, Wp, Hv, Wv, lo, HL, HLpad, YLpad;
  int CaB, JA, N, JN;       // channels of X per CTA (<=128), CaB/8, channels of dY per CTA (<=128), N/8
  int nplanes, taps_per_cta, ngroups, mblocks, nblocks, bufs;
  long long Q, tiles;
  unsigned x_buf_bytes, y_buf_bytes, tmem_cols;
  signed char plane[16];
  int shift[16];
};

struct SmemHeaderW {
  unsigned long long ready[2], free_[2], acc_done;
  unsigned tmem_base, pad;
};

__host__ __device__ constexpr unsigned make_idesc_mn(int M, int N) {   // both operands MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__global__ void __launch_bounds__(160) tc_wgrad_kernel(const __grid_constant__ TwParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemHeaderW* hdr = reinterpret_cast<SmemHeaderW*>(smem_raw);
  unsigned char* bufs = smem_raw + 128;
  const unsigned stage_bytes = P.x_buf_bytes + P.y_buf_bytes;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uniform_u32((unsigned)(tid >> 5));
  int by = blockIdx.y;
  const int nb = by % P.nblocks; by /= P.nblocks;
  const int mb = by % P.mblocks; by /= P.mblocks;
  const int grp = by;
  const int a0 = mb * P.CaB, b0 = nb * P.N;
  const long long my_tiles = (P.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles blockIdx.x, +gridDim.x, ...

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&hdr->ready[i]), 128); mbar_init(smem_u32(&hdr->free_[i]), 1); }
    mbar_init(smem_u32(&hdr->acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&hdr->tmem_base), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = uniform_u32(hdr->tmem_base);
  const unsigned PITCH_X = (unsigned)P.HLpad * 16u, PITCH_Y = (unsigned)P.YLpad * 16u;

  if (warp < 4) {
    const int sm = P.mode == 1 ? 2 : 1;
    for (long long it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it % P.bufs);
      if (it >= P.bufs) mbar_wait(smem_u32(&hdr->free_[buf]), (unsigned)((it / P.bufs) - 1) & 1u);
      unsigned char* xb = bufs + (size_t)buf * stage_bytes;
      unsigned char* yb = xb + P.x_buf_bytes;
      const long long q0 = ((long long)blockIdx.x + it * gridDim.x) * TILE_M;
      stage_halo(xb, PITCH_X, P.x, P.x_ld, P.x_coff, a0, P.Ca, P.x_vec, P.HL, P.JA, P.nplanes, sm, (int)q0 - P.lo, (int)P.Q,
                 P.Hp, P.Wp, P.Hv, P.Wv, P.Hx, P.Wx, tid, 128);
      stage_halo(yb, PITCH_Y, P.dy, P.dy_ld, P.dy_coff, b0, P.Cb, P.dy_vec, TILE_M, P.JN, 1, 1, (int)q0, (int)P.Q, P.Hp, P.Wp,
                 P.Hv, P.Wv, P.Hy, P.Wy, tid, 128);
      fence_proxy_async();
      mbar_arrive(smem_u32(&hdr->ready[buf]));
    }
    // ---- epilogue: D row = X channel (TMEM lane), columns = dY channels; one accumulator per tap of the group ----
    mbar_wait(smem_u32(&hdr->acc_done), 0);
    tc_fence_after();
    const int a = warp * 32 + lane;
    for (int tl = 0; tl < P.taps_per_cta; ++tl) {
      const int tap = grp * P.taps_per_cta + tl;
      for (int n0 = 0; n0 < P.N; n0 += 32) {
        float v[32];
        tmem_ld_upto32(tmem_base + ((unsigned)(warp * 32) << 16) + (unsigned)(tl * P.N + n0), v, P.N - n0);
        if (a0 + a < P.Ca && my_tiles > 0) {
          float* dst = P.dw + ((size_t)tap * P.Ca + a0 + a) * P.Cb + b0 + n0;
          const int ncols = max(0, min(min(32, P.N - n0), P.Cb - (b0 + n0)));
          if (P.dw_vec) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              if (k < ncols) atomicAdd(reinterpret_cast<float4*>(dst + k), make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < ncols) atomicAdd(dst + k, v[k]);
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // issue loop on uniform operands (see tc2_conv_kernel): descriptor high words are loop invariants, the low words
    // advance by 16 (= 256 bytes = 16 pixels) per MMA
    const unsigned idesc = make_idesc_mn(TILE_M, P.N);
    // MN-major canonical layout: 8 pixels (K) at 16 B, next K group at LBO = 128 B, next 8 channels at SBO = pitch
    const unsigned long long a_hi = ((unsigned long long)((PITCH_X >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((128u >> 4) & 0x3FFF) << 16);
    const unsigned long long b_hi = ((unsigned long long)((PITCH_Y >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((128u >> 4) & 0x3FFF) << 16);
    const bool leader = elect_one();
    const int my_tiles_i = (int)my_tiles;
    unsigned first = 1u;
    for (int it = 0; it < my_tiles_i; ++it) {
      const int buf = it % P.bufs;
      mbar_wait(smem_u32(&hdr->ready[buf]), (unsigned)(it / P.bufs) & 1u);
      tc_fence_after();
      if (leader) {
        const unsigned xb4 = smem_u32(bufs + (size_t)buf * stage_bytes) >> 4;
        const unsigned yb4 = xb4 + (P.x_buf_bytes >> 4);
        for (int tl = 0; tl < P.taps_per_cta; ++tl) {
          const int tap = grp * P.taps_per_cta + tl;
          const unsigned d_tmem = tmem_base + (unsigned)(tl * P.N);
          unsigned a_lo = xb4 + (unsigned)(P.plane[tap] * P.JA) * (PITCH_X >> 4) + (unsigned)(P.lo + P.shift[tap]);
          unsigned b_lo = yb4;
          unsigned acc_flag = first ^ 1u;
#pragma unroll
          for (int k = 0; k < TILE_M / 16; ++k) {
            umma_bf16(d_tmem, a_hi | (unsigned long long)(a_lo & 0x3FFF), b_hi | (unsigned long long)(b_lo & 0x3FFF), idesc, acc_flag);
            acc_flag = 1u;
            a_lo += 16u;
            b_lo += 16u;
          }
        }
        umma_commit(smem_u32(&hdr->free_[buf]));
      }
      first = 0u;
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&hdr->acc_done));
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

bool build_wparams(const Geom& g, TwParams& P, int sm_count, int tmem_cap = 512) {
  // g: conv-gather geometry (mode 0) with X = [B,Hin,Win,Cin], dY = [B,Hout,Wout,Cout]
  memset(&P, 0, sizeof P);
  if (g.mode != 0 || g.KH != 4 || g.KW != 4 || g.pad != 1) return false;
  const int Ca_p = round16(g.Cin), Cb_p = round16(g.Cout);
  if ((Ca_p > 128 && Ca_p % 128) || (Cb_p > 128 && Cb_p % 128)) return false;
  P.B = g.B; P.Hx = g.Hin; P.Wx = g.Win; P.Ca = g.Cin; P.Hy = g.Hout; P.Wy = g.Wout; P.Cb = g.Cout;
  P.CaB = Ca_p < 128 ? Ca_p : 128; P.JA = P.CaB / 8; P.mblocks = Ca_p / P.CaB;
  P.N = Cb_p < 128 ? Cb_p : 128; P.JN = P.N / 8; P.nblocks = Cb_p / P.N;
  int hi;
  if (g.stride == 1) {
    P.mode = 0; P.nplanes = 1;
    P.Hp = g.Hin + 2; P.Wp = g.Win + 2; P.Hv = g.Hin; P.Wv = g.Win;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) { P.plane[kh * 4 + kw] = 0; P.shift[kh * 4 + kw] = (kh - 1) * P.Wp + (kw - 1); }
    P.lo = P.Wp + 1; hi = 2 * P.Wp + 2;
  } else if (g.stride == 2) {
    P.mode = 1; P.nplanes = 4;
    P.Hp = g.Hout + 1; P.Wp = g.Wout + 1; P.Hv = g.Hout; P.Wv = g.Wout;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        const int dh = kh == 0 ? -1 : kh == 3 ? 1 : 0, dw = kw == 0 ? -1 : kw == 3 ? 1 : 0;
        P.plane[kh * 4 + kw] = (signed char)((((kh + 1) & 1) << 1) | ((kw + 1) & 1));
        P.shift[kh * 4 + kw] = dh * P.Wp + dw;
      }
    P.lo = P.Wp + 1; hi = P.Wp + 1;
  } else {
    return false;
  }
  P.HL = TILE_M + P.lo + hi;
  P.HLpad = P.HL | 1;
  P.YLpad = TILE_M | 1;
  P.Q = (long long)g.B * P.Hp * P.Wp;
  if (P.Q + TILE_M + P.HL >= (1ll << 31) || (long long)g.B * g.Hin * g.Win >= (1ll << 31)) return false;
  P.tiles = (P.Q + TILE_M - 1) / TILE_M;
  P.taps_per_cta = tmem_cap / P.N; if (P.taps_per_cta > 16) P.taps_per_cta = 16;
  if (P.taps_per_cta < 1) return false;
  while (16 % P.taps_per_cta) --P.taps_per_cta;     // the tap groups must tile the 16 taps exactly (N = 48, 80, ...)
  P.ngroups = 16 / P.taps_per_cta;
  unsigned cols = (unsigned)(P.taps_per_cta * P.N), t = 32;
  while (t < cols) t <<= 1;
  P.tmem_cols = t;
  // X buffer: every MMA reads 16 consecutive 8-channel planes starting at its parity plane's first one
  P.x_buf_bytes = (unsigned)(((P.nplanes - 1) * P.JA + 16) * P.HLpad * 16);
  P.y_buf_bytes = (unsigned)(P.JN * P.YLpad * 16);
  P.bufs = 2;
  if (128 + 2 * (size_t)(P.x_buf_bytes + P.y_buf_bytes) > 227 * 1024) P.bufs = 1;
  if (128 + (size_t)P.bufs * (P.x_buf_bytes + P.y_buf_bytes) > 227 * 1024) return false;
  (void)sm_count;
  return true;
}


}  // namespace


// =====================================================================================================================
// TMA-fed persistent variant of the shifted-window kernel ("tc2").
// Activations arrive as bf16 copies in a zero-padded NHWC layout written by the producing elementwise kernel
// (BfAct, common.cuh): the halo of a tile is then a contiguous pixel range of every 8-channel plane and one elected thread
// stages it with cp.async.bulk.tensor (2-D box = 8 channels x <= 256 pixels, out-of-range pixels zero-filled by the
// TMA unit) - no SIMT staging, no index math, no fp32 re-read.  The CTA is persistent: it walks tiles blockIdx.x,
// +gridDim.x, ...; the packed weights of layers up to 144 KB stay resident in shared memory for all of them; the
// accumulator is double-buffered in TMEM so that the epilogue of tile i overlaps the MMAs of tile i+1.
// Warps: 0 halo producer (TMA), 1 weight producer (bulk copies), 2 MMA issuer + TMEM owner, 3-6 epilogue.
namespace {

constexpr int A_STAGES_MAX = 4;
constexpr unsigned W_RESIDENT_MAX = 144 * 1024;
constexpr int ACC_BUFS_MAX = 16;   // accumulator sets of a CTA (MODE 2 keeps one per tile: 16 x 32 columns = all of TMEM)

struct Tc2Params {
  TcParams t;               // geometry (mode, padded space, taps, chunking) as for the SIMT-staged kernel
  int tiles;
  int nsplit, boxp;         // TMA boxes per 8-channel plane, pixels per box
  unsigned a_pitch;         // bytes between planes in shared memory
  unsigned a_bytes;         // one halo (all planes of one channel chunk)
  int a_stages;
  int ntw;                  // output channels per CTA (blockIdx.z walks the sub-tiles of a 128-column group; 128: whole group)
  int wtw;                  // tile width the weights were packed with (128, or = ntw: the CTA's weights are then contiguous)
  int w_resident;           // 1: all weights of the CTA's N tile live in shared memory
  unsigned w_bytes_ntile;   // packed bytes of one full N tile (128 columns)
  unsigned w_region;        // shared-memory bytes reserved for weights (resident copy or ring)
  int w_stages;             // ring mode: stages of b_stage_max bytes
  int acc_bufs;             // TMEM accumulator sets (2 = epilogue overlaps the next tile's MMAs)
  unsigned acc_cols;        // columns per set (nacc * Nt rounded for the allocation)
  unsigned tmem_cols;
  const __nv_bfloat16* a_src;   // bf16 planar activation copy, pointing at (group chan0/8, pixel 0 of plane 0)
  long long plane_rows;     // rows between parity planes
  long long group_rows;     // rows between 8-channel groups
  BnBwdFuse fz;             // fz.y != nullptr: batch-norm backward pass 1 of the consuming block fused into the epilogue
  BnFwdFuse ff;             // MODE 2: batch-norm forward of THIS block fused (all tiles resident in TMEM, grid barrier, pass 2)
  unsigned grid_ctas;       // MODE 2: CTAs of the launch (grid-barrier target)
  // per-tap issue table (host-built): A start offset inside a halo stage in 16-byte units (parity plane + lo + shift), and
  // accumulator index | 0x80 when the tap is the first one that writes its accumulator
  unsigned tap_a[16];
  unsigned char tap_f[16];
};

struct SmemHeader2 {
  unsigned long long a_full[A_STAGES_MAX], a_empty[A_STAGES_MAX], w_full[W2_STAGES_MAX], w_empty[W2_STAGES_MAX];
  unsigned long long acc_full[ACC_BUFS_MAX], acc_empty[2];
  unsigned tmem_base, pad;
  float s_sum[4][128], s_sq[4][128];
  // fused batch-norm backward: per channel of this CTA  f_mean = shift (-mean*rstd), f_rstd = scale, f_beta (read as float4)
  alignas(16) float f_mean[128];
  alignas(16) float f_rstd[128];
  alignas(16) float f_beta[128];
  unsigned tap_a_lo[16], tap_dcol[16];
  volatile unsigned long long ts[16];   // debug timestamps (svae_debug_set_buffer)
};
#define DBG2(slot) do { if (P.dbg != nullptr) { const unsigned long long t_ = gtimer(); if ((threadIdx.x & 31) == 0) hdr->ts[slot] = t_; } } while (0)

// grid-wide barrier of the MODE 2 kernel: every CTA of the launch is resident (the host sizes the grid to the SMs' capacity and
// the only other barrier user, if any, is an earlier launch on the same stream whose CTAs are all resident already), so the
// wait is bounded by the slowest CTA's pass 1.  A protocol bug becomes a trap (reported CUDA error), never a hung GPU.
__device__ __forceinline__ void grid_barrier_arrive_wait(unsigned* counter, unsigned target) {
  __threadfence();
  atomicAdd(counter, 1u);
  unsigned seen = 0;
  for (unsigned it = 0;; ++it) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    if (seen >= target) break;
    __nanosleep(64);
    if (it > (1u << 24)) __trap();
  }
}

// MODE 0: plain contraction (+ forward statistics).  MODE 1: the epilogue carries batch-norm backward pass 1 of the consuming
// block (BnBwdFuse).  MODE 2: batch-norm forward of this block fused (BnFwdFuse).  Separate instantiations so that the plain
// kernel keeps its register budget.
template <int MODE>
__device__ __forceinline__ void tc2_conv_body(const Tc2Params& PP, const int zsub) {
  constexpr bool FUSED = MODE == 1;
  constexpr bool BNF = MODE == 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const TcParams& P = PP.t;
  SmemHeader2* hdr = reinterpret_cast<SmemHeader2*>(smem_raw);
  unsigned char* w_smem = smem_raw + ((sizeof(SmemHeader2) + 127) & ~127u);
  // Output-channel tiling: blockIdx.y = 128-column tile of the packed weights, blockIdx.z = sub-tile of PP.ntw columns
  // inside it.  Layers with few pixel tiles (4x4 / 8x8 maps, B = 100) are bound by streaming their weights into one CTA
  // per tile; splitting the channels spreads that stream (and the MMAs) over up to 4x as many SMs.
  const int n0 = blockIdx.y * 128 + zsub * PP.ntw;
  const int Nt = min(PP.ntw, min(128, P.N_p - (int)blockIdx.y * 128) - zsub * PP.ntw);
  const int wtile = n0 / PP.wtw;                                  // packed tile holding this CTA's columns
  const int tile_w = min(PP.wtw, P.N_p - wtile * PP.wtw);         // its width
  const bool w_sub = Nt != tile_w;     // sub-tile: one bulk copy per (chunk, tap, 8-channel group) row of Nt*16 bytes
  const unsigned b_tap_bytes = (unsigned)(P.KC * Nt * 2);
  const unsigned b_stage_bytes = b_tap_bytes * (unsigned)P.tps;
  unsigned char* a_smem = w_smem + PP.w_region;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uniform_u32((unsigned)(tid >> 5));   // warp-uniform for the compiler: role branches stay convergent
  const int my_tiles = (PP.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp == 0) DBG2(0);

  if (tid == 0) {
    for (int s = 0; s < A_STAGES_MAX; ++s) { mbar_init(smem_u32(&hdr->a_full[s]), 1); mbar_init(smem_u32(&hdr->a_empty[s]), 1); }
    for (int s = 0; s < W2_STAGES_MAX; ++s) { mbar_init(smem_u32(&hdr->w_full[s]), 1); mbar_init(smem_u32(&hdr->w_empty[s]), 1); }
    for (int b = 0; b < ACC_BUFS_MAX; ++b) mbar_init(smem_u32(&hdr->acc_full[b]), 1);
    for (int b = 0; b < 2; ++b) mbar_init(smem_u32(&hdr->acc_empty[b]), 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&hdr->tmem_base), PP.tmem_cols);
  if (warp >= 3) {
    for (int i = tid - 96; i < 4 * 128; i += 128) { (&hdr->s_sum[0][0])[i] = 0.f; (&hdr->s_sq[0][0])[i] = 0.f; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = uniform_u32(hdr->tmem_base);
  if (warp == 0) DBG2(1);
  // Resident weights are fetched BEFORE the grid dependency is resolved: the packed operand copies are written once at the
  // start of the step (>= 2 kernels back, see launch_k), so under programmatic dependent launch this load - like the barrier
  // and TMEM setup above - overlaps the tail of the predecessor kernel.
  const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(P.wp) + (size_t)wtile * PP.wtw * P.Cin_p * P.ntaps * 2 +
                              (size_t)(n0 - wtile * PP.wtw) * 16;
  const unsigned src_row = (unsigned)tile_w * 16u, dst_row = (unsigned)Nt * 16u;   // one (chunk, tap, group) row: [n][8] bf16
  if (warp == 1 && PP.w_resident && my_tiles > 0) {
    const unsigned total = (unsigned)(P.NC * P.ntaps) * b_tap_bytes;
    const unsigned bar = smem_u32(&hdr->w_full[0]);
    if (lane == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (!w_sub) {
      if (lane == 0)
        for (unsigned off = 0; off < total; off += 32768u)
          bulk_g2s(smem_u32(w_smem + off), wsrc + off, min(32768u, total - off), bar);
    } else {
      const int rows = P.NC * P.ntaps * P.JC;
      for (int r = lane; r < rows; r += 32)
        bulk_g2s(smem_u32(w_smem) + (unsigned)r * dst_row, wsrc + (size_t)r * src_row, dst_row, bar);
    }
  }
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ================= halo producer: one bulk copy (TMA 1-D) per (parity plane, 8-channel group) =================
    int it = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int q0 = ((int)blockIdx.x + ti * (int)gridDim.x) * TILE_M;
      for (int c = 0; c < P.NC; ++c, ++it) {
        const int s = it % PP.a_stages;
        if (it >= PP.a_stages) mbar_wait(smem_u32(&hdr->a_empty[s]), (unsigned)((it / PP.a_stages) - 1) & 1u);
        if (lane == 0) {
          const unsigned bar = smem_u32(&hdr->a_full[s]);
          mbar_expect_tx(bar, PP.a_bytes);
          const unsigned abase = smem_u32(a_smem + (size_t)s * PP.a_bytes);
          for (int pl = 0; pl < P.nplanes; ++pl)
            for (int j = 0; j < P.JC; ++j)
              bulk_g2s(abase + (unsigned)(pl * P.JC + j) * PP.a_pitch,
                       PP.a_src + ((long long)(c * P.JC + j) * PP.group_rows + pl * PP.plane_rows + (q0 - P.lo)) * 8,
                       PP.a_pitch, bar);
        }
        __syncwarp();
        if (it == 0) DBG2(2);
      }
    }
  } else if (warp == 1) {
    // ================= weight producer =================
    if (!PP.w_resident) {
      int stage = 0; unsigned phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti)
        for (int c = 0; c < P.NC; ++c)
          for (int s = 0; s < P.ntaps; s += P.tps) {
            mbar_wait(smem_u32(&hdr->w_empty[stage]), phase ^ 1);
            if (lane == 0) mbar_expect_tx(smem_u32(&hdr->w_full[stage]), b_stage_bytes);
            __syncwarp();
            if (!w_sub) {
              if (lane == 0)
                bulk_g2s(smem_u32(w_smem + (size_t)stage * P.b_stage_max), wsrc + ((size_t)c * P.ntaps + s) * b_tap_bytes,
                         b_stage_bytes, smem_u32(&hdr->w_full[stage]));
            } else {
              const int rows = P.tps * P.JC;
              const size_t row0 = ((size_t)c * P.ntaps + s) * P.JC;
              for (int r = lane; r < rows; r += 32)
                bulk_g2s(smem_u32(w_smem + (size_t)stage * P.b_stage_max) + (unsigned)r * dst_row, wsrc + (row0 + r) * src_row, dst_row,
                         smem_u32(&hdr->w_full[stage]));
            }
            __syncwarp();
            if (++stage == PP.w_stages) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 2) {
    // ================= MMA issuer =================
    // Every operand of the issue loop is derived from kernel parameters, loop counters and the (uniform) shared-memory
    // window, so that descriptors live in uniform registers: the loop is a handful of uniform ALU ops per tcgen05.mma.
    const unsigned idesc = make_idesc(TILE_M, Nt);
    const unsigned LBO_A = PP.a_pitch, LBO_B = (unsigned)Nt * 16u;
    const unsigned long long a_hi = ((unsigned long long)((128u >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((LBO_A >> 4) & 0x3FFF) << 16);
    const unsigned long long b_hi = ((unsigned long long)((128u >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((LBO_B >> 4) & 0x3FFF) << 16);
    const unsigned a_kstep = (2u * LBO_A) >> 4, b_kstep = (2u * LBO_B) >> 4;
    const int nk = P.KC / 16;
    const bool leader = elect_one();
    int stage = 0; unsigned phase = 0;
    int it = 0;
    if (PP.w_resident && my_tiles > 0) mbar_wait(smem_u32(&hdr->w_full[0]), 0);
    DBG2(3);
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int b = BNF ? ti : ti % PP.acc_bufs;     // MODE 2: every tile keeps its own accumulator set until pass 2
      if (!BNF && ti >= PP.acc_bufs) mbar_wait(smem_u32(&hdr->acc_empty[b]), (unsigned)((ti / PP.acc_bufs) - 1) & 1u);
      tc_fence_after();
      const unsigned acc_base = tmem_base + (unsigned)b * PP.acc_cols;
      for (int c = 0; c < P.NC; ++c, ++it) {
        const int s = it % PP.a_stages;
        mbar_wait(smem_u32(&hdr->a_full[s]), (unsigned)(it / PP.a_stages) & 1u);
        tc_fence_after();
        if (it == 0) DBG2(4);
        const unsigned abase4 = smem_u32(a_smem + (size_t)s * PP.a_bytes) >> 4;
        for (int s0 = 0; s0 < P.ntaps; s0 += P.tps) {
          unsigned b_lo;
          if (PP.w_resident) {
            b_lo = smem_u32(w_smem + ((size_t)c * P.ntaps + s0) * b_tap_bytes) >> 4;
          } else {
            mbar_wait(smem_u32(&hdr->w_full[stage]), phase);
            tc_fence_after();
            b_lo = smem_u32(w_smem + (size_t)stage * P.b_stage_max) >> 4;
          }
          if (leader) {
            // The per-tap operands come from a host-built table in the parameter space; the NEXT tap's entry is fetched
            // before the current tap's MMAs are issued, so the constant-load latency hides behind them (fetched on demand
            // it made the issue loop, not the tensor pipe, the pace setter: ~118 instead of ~60 cycles per MMA).
            unsigned na = PP.tap_a[s0], nf = PP.tap_f[s0];
            for (int tl = 0; tl < P.tps; ++tl) {
              const unsigned ca = na, cf = nf;
              const int nx = min(s0 + tl + 1, 15);
              na = PP.tap_a[nx]; nf = PP.tap_f[nx];
              const unsigned d_tmem = acc_base + (cf & 0x7Fu) * (unsigned)Nt;
              unsigned a_lo = abase4 + ca;
              unsigned bk = b_lo;
              unsigned acc_flag = (c > 0 || (cf & 0x80u) == 0u) ? 1u : 0u;
              for (int kk = 0; kk < nk; ++kk) {
                umma_bf16(d_tmem, a_hi | (unsigned long long)(a_lo & 0x3FFF), b_hi | (unsigned long long)(bk & 0x3FFF), idesc,
                          acc_flag);
                acc_flag = 1u;
                a_lo += a_kstep;
                bk += b_kstep;
              }
              b_lo += b_tap_bytes >> 4;
            }
            if (!PP.w_resident) umma_commit(smem_u32(&hdr->w_empty[stage]));
          }
          __syncwarp();
          if (!PP.w_resident) { if (++stage == PP.w_stages) { stage = 0; phase ^= 1; } }
        }
        if (leader) umma_commit(smem_u32(&hdr->a_empty[s]));
        __syncwarp();
      }
      if (leader) umma_commit(smem_u32(&hdr->acc_full[b]));
      __syncwarp();
      if (ti == 0) DBG2(5);
      if (ti == my_tiles - 1) DBG2(8);
    }
  } else {
    // ================= epilogue warps 3..6: TMEM lane quarter = warp % 4 =================
    const int quarter = warp & 3;
    const int et = tid - 96;                          // 0..127 inside the epilogue group
    const BnBwdFuse& fz = PP.fz;
    constexpr bool fused = FUSED;
    if constexpr (FUSED) {
      // xhat = y*scale + shift; channels outside the batch norm get scale = shift = 0 and beta = 1: xhat = 0, act' = 1, i.e.
      // they pass through the same arithmetic unchanged and add nothing to the second sum - no per-element branches
      for (int col = et; col < 128; col += 128) {
        const int c = n0 + col;
        float scale = 0.f, shift = 0.f, bt = 1.f;
        if (col < Nt && c < fz.C) {
          const double m = fz.stats[c] / (double)fz.rows;
          double var = fz.stats[fz.C + c] / (double)fz.rows - m * m;
          if (var < 0.0) var = 0.0;
          const float mean = (float)m;
          scale = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
          shift = -mean * scale;
          bt = fz.beta[c];
        }
        hdr->f_mean[col] = shift; hdr->f_rstd[col] = scale; hdr->f_beta[col] = bt;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int b = BNF ? ti : ti % PP.acc_bufs;
      mbar_wait(smem_u32(&hdr->acc_full[b]), BNF ? 0u : ((unsigned)(ti / PP.acc_bufs) & 1u));
      tc_fence_after();
      if (ti == 0 && warp == 3) DBG2(6);
      const int q0 = ((int)blockIdx.x + ti * (int)gridDim.x) * TILE_M;
      const int q = q0 + quarter * 32 + lane;
      bool valid = q < (int)P.Q;
      int n = 0, r = 0, cc = 0;
      if (valid) {
        const int t = q / P.Wp;
        cc = q - t * P.Wp;
        n = t / P.Hp;
        r = t - n * P.Hp;
        valid = r < P.Hv && cc < P.Wv;
      }
      const unsigned acc_base = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)b * PP.acc_cols;
      for (int a = 0; a < P.nacc; ++a) {
        int oh = r, ow = cc;
        if (P.mode == 2) { oh = 2 * r + (a >> 1); ow = 2 * cc + (a & 1); }
        float* orow = P.out + (((size_t)n * P.Hout + oh) * P.Wout + ow) * P.out_ld + P.out_coff + n0;
        for (int nn = 0; nn < Nt; nn += 32) {
          float v[32];
          tmem_ld_upto32(acc_base + (unsigned)(a * Nt + nn), v, Nt - nn);
          const int ncols = max(0, min(min(32, Nt - nn), P.n_valid - (n0 + nn)));
          if constexpr (FUSED) {
            // ---- da -> g = da * act'(xhat + beta + residual) for the consuming block's channels; sums of g and g*xhat ----
            // Phase A issues every global load of the 32-column group (y, residual, previous value) before the first use, so
            // that they are all in flight together; phase B computes and stores.  (Interleaved, each load would wait behind
            // the preceding store: the compiler cannot prove that the output does not alias y.)
            const size_t pix = ((size_t)n * P.Hout + oh) * P.Wout + ow;
            const int cbase = n0 + nn;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* yp = reinterpret_cast<const float4*>(fz.y + pix * fz.C + cbase);
            const float4* rp = reinterpret_cast<const float4*>(fz.res + pix * fz.res_ld + fz.res_coff + cbase);
            float4* op = reinterpret_cast<float4*>(orow + nn);
            const int nbn = valid ? min(ncols, fz.C - cbase) : 0;          // columns of this group inside the batch norm
            const int nout = valid ? ncols : 0;
            // one auxiliary operand per element: the residual (enters the activation derivative) OR the previous value of
            // the output (accumulate) - the host never asks for both at once
            const bool has_res = fz.res != nullptr, acc = P.accumulate != 0;
            const float4* xp = has_res ? rp : op;
            const int naux = has_res ? nbn : (acc ? nout : 0);
            float4 yv[8], av[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              yv[j] = 4 * j < nbn ? __ldg(yp + j) : z4;
              av[j] = 4 * j < naux ? xp[j] : z4;
            }
            if (!valid) {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = 0.f;
            }
            const float neg = fz.act == ACT_LRELU ? SVAE_LRELU_SLOPE : fz.act == ACT_RELU ? 0.f : 1.f;
            const float rsel = has_res ? 1.f : 0.f, osel = has_res ? 0.f : 1.f;
            float gy[32];     // g*y: sum g*xhat = scale * sum g*y + shift * sum g (applied per column at the end)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = 4 * j;
              const float4 sc = *reinterpret_cast<const float4*>(&hdr->f_rstd[nn + k]);
              const float4 sh = *reinterpret_cast<const float4*>(&hdr->f_mean[nn + k]);
              const float4 bt = *reinterpret_cast<const float4*>(&hdr->f_beta[nn + k]);
              const float yy[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w}, aa[4] = {av[j].x, av[j].y, av[j].z, av[j].w};
              const float ss[4] = {sc.x, sc.y, sc.z, sc.w}, hh[4] = {sh.x, sh.y, sh.z, sh.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float pre = fmaf(yy[e], ss[e], hh[e]) + bb[e] + rsel * aa[e];
                const float gval = fmaf(osel, aa[e], v[k + e]) * (pre > 0.f ? 1.f : neg);
                v[k + e] = gval;
                gy[k + e] = gval * yy[e];
              }
              if (k < nout) {
                const float4 o = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                op[j] = o;
                if (fz.dres != nullptr && k < nbn) {
                  float4* dr = reinterpret_cast<float4*>(fz.dres + pix * fz.C + cbase) + j;
                  float4 d = o;
                  if (fz.dres_acc) { const float4 q4 = *dr; d.x += q4.x; d.y += q4.y; d.z += q4.z; d.w += q4.w; }
                  *dr = d;
                }
              }
            }
            const float cs = warp_colsum32(v, lane);
            const float cq = warp_colsum32(gy, lane);
            hdr->s_sum[quarter][nn + lane] += cs;
            hdr->s_sq[quarter][nn + lane] += cq;
            continue;
          }
          if (valid) {
            if (P.out_vec) {
#pragma unroll
              for (int k = 0; k < 32; k += 4) {
                if (k < ncols) {
                  float4 o = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
                  float4* dst = reinterpret_cast<float4*>(orow + nn + k);
                  if (P.accumulate) { float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                  *dst = o;
                  v[k] = o.x; v[k + 1] = o.y; v[k + 2] = o.z; v[k + 3] = o.w;
                }
              }
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                if (k < ncols) {
                  float o = v[k];
                  if (P.accumulate) o += orow[nn + k];
                  orow[nn + k] = o;
                  v[k] = o;
                }
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.f;
          }
          if (P.stats != nullptr) {
            float sq[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) { if (k >= ncols) v[k] = 0.f; sq[k] = v[k] * v[k]; }
            const float cs = warp_colsum32(v, lane);
            const float cq = warp_colsum32(sq, lane);
            hdr->s_sum[quarter][nn + lane] += cs;      // this warp's slot: accumulated over phases and over the CTA's tiles
            hdr->s_sq[quarter][nn + lane] += cq;
          }
        }
      }
      tc_fence_before();
      if (!BNF) mbar_arrive(smem_u32(&hdr->acc_empty[b]));
      if (ti == 0 && warp == 3) DBG2(7);
      if (ti == my_tiles - 1 && warp == 3) DBG2(9);
    }
    if (fused) {
      if (my_tiles > 0) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < Nt; col += 128) {
        if (n0 + col < fz.C) {
          const float s = hdr->s_sum[0][col] + hdr->s_sum[1][col] + hdr->s_sum[2][col] + hdr->s_sum[3][col];
          const float s2 = hdr->s_sq[0][col] + hdr->s_sq[1][col] + hdr->s_sq[2][col] + hdr->s_sq[3][col];
          atomicAdd(&fz.S[n0 + col], (double)s);
          atomicAdd(&fz.S[fz.C + n0 + col], (double)hdr->f_rstd[col] * (double)s2 + (double)hdr->f_mean[col] * (double)s);   // sum g*xhat
        }
      }
      }
    } else if (P.stats != nullptr && my_tiles > 0) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < Nt; col += 128) {
        if (n0 + col < P.n_valid) {
          const float s = hdr->s_sum[0][col] + hdr->s_sum[1][col] + hdr->s_sum[2][col] + hdr->s_sum[3][col];
          const float s2 = hdr->s_sq[0][col] + hdr->s_sq[1][col] + hdr->s_sq[2][col] + hdr->s_sq[3][col];
          atomicAdd(&P.stats[n0 + col], (double)s);
          atomicAdd(&P.stats[P.n_valid + n0 + col], (double)s2);
        }
      }
    }
    if constexpr (BNF) {
      // ---- grid-wide barrier: the channel statistics are final once every CTA of the launch has added its share ----
      const BnFwdFuse& ff = PP.ff;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) grid_barrier_arrive_wait(ff.counter, PP.grid_ctas);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < 128; col += 128) {
        const int c = n0 + col;
        float scale = 0.f, shift = 0.f;
        if (col < Nt && c < P.n_valid) {
          const double m = __ldcg(&P.stats[c]) / (double)ff.rows;
          double var = __ldcg(&P.stats[P.n_valid + c]) / (double)ff.rows - m * m;
          if (var < 0.0) var = 0.0;
          const float mean = (float)m;
          scale = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
          shift = ff.beta[c] - mean * scale;
        }
        hdr->f_rstd[col] = scale; hdr->f_mean[col] = shift;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- pass 2: accumulators -> normalise (+ shortcut) -> activation -> fp32 channel window and / or bf16 planar copy ----
      const float neg = ff.act == ACT_LRELU ? SVAE_LRELU_SLOPE : ff.act == ACT_RELU ? 0.f : 1.f;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int q0 = ((int)blockIdx.x + ti * (int)gridDim.x) * TILE_M;
        const int q = q0 + quarter * 32 + lane;
        bool valid = q < (int)P.Q;
        int n = 0, r = 0, cc = 0;
        if (valid) {
          const int t = q / P.Wp;
          cc = q - t * P.Wp;
          n = t / P.Hp;
          r = t - n * P.Hp;
          valid = r < P.Hv && cc < P.Wv;
        }
        const unsigned acc_base = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)ti * PP.acc_cols;
        for (int a = 0; a < P.nacc; ++a) {
          int oh = r, ow = cc;
          if (P.mode == 2) { oh = 2 * r + (a >> 1); ow = 2 * cc + (a & 1); }
          const size_t pix = ((size_t)n * P.Hout + oh) * P.Wout + ow;
          for (int nn = 0; nn < Nt; nn += 32) {
            float v[32];
            tmem_ld_upto32(acc_base + (unsigned)(a * Nt + nn), v, Nt - nn);      // warp-collective: outside the `valid` branch
            const int ncols = max(0, min(min(32, Nt - nn), P.n_valid - (n0 + nn)));
            if (!valid) continue;
            const int cbase = n0 + nn;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              if (k < ncols) {
                float o[8];
                const float4 sc0 = *reinterpret_cast<const float4*>(&hdr->f_rstd[nn + k]);
                const float4 sc1 = *reinterpret_cast<const float4*>(&hdr->f_rstd[nn + k + 4]);
                const float4 sh0 = *reinterpret_cast<const float4*>(&hdr->f_mean[nn + k]);
                const float4 sh1 = *reinterpret_cast<const float4*>(&hdr->f_mean[nn + k + 4]);
                const float ss[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                const float hh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(v[k + e], ss[e], hh[e]);
                if (ff.res != nullptr) {
                  const float4* rp = reinterpret_cast<const float4*>(ff.res + pix * ff.res_ld + ff.res_coff + cbase + k);
                  const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                  o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w; o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = o[e] > 0.f ? o[e] : o[e] * neg;
                if (ff.out != nullptr) {
                  float4* op = reinterpret_cast<float4*>(ff.out + pix * ff.out_ld + ff.out_coff + cbase + k);
                  op[0] = make_float4(o[0], o[1], o[2], o[3]);
                  op[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
                if (ff.bf.a.p != nullptr)
                  *reinterpret_cast<uint4*>(ff.bf.a.p + bf_index(ff.bf.a, n, oh, ow, ff.bf.coff + cbase + k)) = pack8_bf16(o);
              }
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 0) DBG2(10);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, PP.tmem_cols);
  }
  if (P.dbg != nullptr) {
    __syncthreads();
    if (tid < 16) P.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + tid] = tid == 15 ? (unsigned long long)my_tiles : hdr->ts[tid];
  }
}

template <int MODE>
__global__ void __launch_bounds__(224, MODE != 0 ? 2 : 1) tc2_conv_kernel(const __grid_constant__ Tc2Params PP) {
  tc2_conv_body<MODE>(PP, (int)blockIdx.z);
}
// batched variant (LaunchCtx::multi): blockIdx.z = item * nsub + channel sub-tile; the items' parameter blocks travel by value
__global__ void __launch_bounds__(224, 1) tc2_conv_multi_kernel(const __grid_constant__ MultiArgs<Tc2Params> A, int nsub) {
  tc2_conv_body<0>(A.v[blockIdx.z / nsub], (int)(blockIdx.z % nsub));
}
static void tc2_conv_multi_launch(const void* host_args, int items, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  const int nsub = (int)grid.z;
  multi_launch_chunks<Tc2Params>(tc2_conv_multi_kernel, host_args, items, grid, block, smem, st, (unsigned)nsub, nsub);
}

constexpr size_t TC2_SMEM_CAP = 227 * 1024;   // shared-memory budget of the TMA-fed conv kernel
bool build_params2(const Geom& g, Tc2Params& PP, int ntw = 128) {
  memset(&PP, 0, sizeof PP);
  TcParams& P = PP.t;
  if (g.KH != 4) return false;
  if (!build_params(g, P)) return false;
  PP.ntw = ntw;
  if (ntw < 128) {
    // sub-tiles must tile every 128-column packed tile exactly
    if (ntw % 16 || ntw < 16 || (P.N_p >= 128 ? (P.N_p % 128 || 128 % ntw) : (P.N_p % ntw))) return false;
    const unsigned tap_bytes = (unsigned)(P.KC * ntw * 2);     // weight stages as build_params sizes them, for the narrower tile
    P.tps = 1;
    while (P.tps * 2 <= P.ntaps && (unsigned)(P.tps * 2) * tap_bytes <= 32 * 1024) P.tps *= 2;
    P.b_stage_max = tap_bytes * (unsigned)P.tps;
  }
  PP.tiles = (int)((P.Q + TILE_M - 1) / TILE_M);
  PP.nsplit = 1;
  PP.boxp = (P.HL + 7) & ~7;            // pixels per plane copy (16 B each); keeps every plane 128-byte aligned
  PP.a_pitch = (unsigned)PP.boxp * 16u;
  PP.a_bytes = (unsigned)(P.nplanes * P.JC) * PP.a_pitch;
  const int nt_max = min(P.N_p < 128 ? P.N_p : 128, ntw);
  PP.w_bytes_ntile = (unsigned)(P.NC * P.ntaps) * (unsigned)(P.KC * 128 * 2);
  const unsigned w_need = (unsigned)(P.NC * P.ntaps) * (unsigned)(P.KC * nt_max * 2);
  PP.w_resident = w_need <= W_RESIDENT_MAX ? 1 : 0;
  const size_t hdr = (sizeof(SmemHeader2) + 127) & ~(size_t)127;
  size_t w_region;
  if (PP.w_resident) {
    w_region = PP.w_bytes_ntile;        // region sized for a full tile; only w_need bytes are filled
    if (nt_max < 128) w_region = w_need;
  } else {
    static const int ws_max = getenv("SVAE_W_STAGES") ? atoi(getenv("SVAE_W_STAGES")) : W2_STAGES_MAX;
    PP.w_stages = ws_max < 2 ? 2 : (ws_max > W2_STAGES_MAX ? W2_STAGES_MAX : ws_max);
    while (PP.w_stages > 2 && hdr + (size_t)PP.w_stages * P.b_stage_max + 2 * (size_t)PP.a_bytes + 256 > TC2_SMEM_CAP) --PP.w_stages;
    w_region = (size_t)PP.w_stages * P.b_stage_max;
  }
  w_region = (w_region + 127) & ~(size_t)127;
  PP.w_region = (unsigned)w_region;
  size_t left = TC2_SMEM_CAP - hdr - w_region - 256;
  if (hdr + w_region + 256 > TC2_SMEM_CAP || left < PP.a_bytes) return false;
  PP.a_stages = (int)(left / PP.a_bytes);
  if (PP.a_stages > A_STAGES_MAX) PP.a_stages = A_STAGES_MAX;
  const int want = P.NC > 1 ? 3 : 2;   // keep the footprint modest: two CTAs per SM help the short layers
  if (PP.a_stages > want) PP.a_stages = want;
  if (PP.a_stages < 1) return false;
  // Narrow layers (N <= 64) are paced by the per-CTA MMA issue latency (59 cycles per MMA alone, ~40 with two CTAs on the SM,
  // scripts/mma_rate.cu): when dropping the halo prefetch depth lets two CTAs share an SM, do it - the second CTA covers the
  // halo latency that the shallower ring exposes.
  {
    static const bool two = !(getenv("SVAE_CONV_2CTA") && getenv("SVAE_CONV_2CTA")[0] == '0');
    const size_t cap = 113 * 1024;
    if (two && nt_max <= 64 && PP.w_resident && hdr + w_region + 256 + (size_t)PP.a_stages * PP.a_bytes > cap &&
        hdr + w_region + 256 + (size_t)PP.a_bytes <= cap) {
      while (PP.a_stages > 1 && hdr + w_region + 256 + (size_t)PP.a_stages * PP.a_bytes > cap) --PP.a_stages;
    }
  }
  unsigned cols = (unsigned)(P.nacc * nt_max), t = 32;
  while (t < cols) t <<= 1;
  PP.acc_cols = t;
  PP.acc_bufs = 2 * t <= 512 ? 2 : 1;
  PP.tmem_cols = t * (unsigned)PP.acc_bufs;
  {
    const unsigned plane4 = ((unsigned)P.JC * PP.a_pitch) >> 4;     // one parity plane further (in 16-byte units)
    unsigned seen = 0;
    for (int tp = 0; tp < 16; ++tp) {
      const int tq = tp < P.ntaps ? tp : 0;
      PP.tap_a[tp] = (unsigned)P.plane[tq] * plane4 + (unsigned)(P.lo + P.shift[tq]);
      const unsigned a = (unsigned)P.acc[tq];
      PP.tap_f[tp] = (unsigned char)(a | (((seen >> a) & 1u) ? 0u : 0x80u));
      seen |= 1u << a;
    }
  }
  return PP.tmem_cols <= 512;
}

size_t smem_bytes2(const Tc2Params& PP) {
  const TcParams& P = PP.t;
  const size_t hdr = (sizeof(SmemHeader2) + 127) & ~(size_t)127;
  return hdr + PP.w_region + (size_t)PP.a_stages * PP.a_bytes + 256;
}

// ---- bf16 padded activation copies ---------------------------------------------------------------------------------
__global__ void bf_fill_kernel(BfAct d, const float* __restrict__ src, int ld, int coff, int C) {
  // one thread per (row of a group, 8-channel group): zeros in the slack and the padding
  const int G = d.Cpad / 8;
  const int nplanes = d.kind == 2 ? 4 : 1;
  const int Hv = d.kind == 2 ? d.H / 2 : d.H, Wv = d.kind == 2 ? d.W / 2 : d.W;
  const long long Q = (long long)d.B * d.Hp * d.Wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.group_rows * G; i += (long long)gridDim.x * blockDim.x) {
    const int g8 = (int)(i / d.group_rows);
    const long long row = i - (long long)g8 * d.group_rows;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long rr = row - d.front;
    if (rr >= 0 && rr < nplanes * d.plane_rows) {
      const int pl = (int)(rr / d.plane_rows);
      const long long q = rr - (long long)pl * d.plane_rows;
      if (q < Q) {
        const int cc = (int)(q % d.Wp);
        const long long t = q / d.Wp;
        const int r = (int)(t % d.Hp), n = (int)(t / d.Hp);
        if (r < Hv && cc < Wv) {
          const int h = d.kind == 2 ? 2 * r + (pl >> 1) : r, w = d.kind == 2 ? 2 * cc + (pl & 1) : cc;
          const float* px = src + (((size_t)n * d.H + h) * d.W + w) * ld + coff;
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (g8 * 8 + e < C) ? px[g8 * 8 + e] : 0.f;
        }
      }
    }
    *reinterpret_cast<uint4*>(d.p + (size_t)i * 8) = pack8_bf16(f);
  }
}

}  // namespace

BfAct bf_act_describe(int kind, int B, int H, int W, int C) {
  BfAct d{};
  d.kind = kind; d.B = B; d.H = H; d.W = W;
  d.Cpad = (C + 15) / 16 * 16;           // the kernels chunk channels by 16: the padding channels are real zero planes
  if (kind == 0) { d.Hp = H + 2; d.Wp = W + 2; }
  else if (kind == 1) { d.Hp = H + 1; d.Wp = W + 1; }
  else { d.Hp = H / 2 + 1; d.Wp = W / 2 + 1; }
  const long long Q = (long long)B * d.Hp * d.Wp;
  const int reach = 2 * d.Wp + 2;         // largest |tap shift| of any geometry that reads this layout
  d.front = (reach + 7) & ~7;
  d.plane_rows = kind == 2 ? ((Q + reach + 7) & ~7LL) : Q;
  const long long tail = 128 + 2 * reach + 24;   // the tap-stacked weight gradient reads copies shifted by up to one padded row + 3 pixels
  d.group_rows = (d.front + (kind == 2 ? 4 : 1) * d.plane_rows + tail + 7) & ~7LL;
  d.p = nullptr;
  return d;
}
size_t bf_act_bytes(const BfAct& d) { return (size_t)d.group_rows * (d.Cpad / 8) * 16; }

int bf_act_fill(const LaunchCtx& lc, const BfAct& d, View src, int C) {
  const long long items = d.group_rows * (d.Cpad / 8);
  long long blocks = (items + 255) / 256;
  if (blocks > lc.sm_count * 16) blocks = lc.sm_count * 16;
  ProfScope ps(lc, KC_MISC, 0.0, 6.0 * (double)items * 8);
  bf_fill_kernel<<<(unsigned)blocks, 256, 0, lc.stream>>>(d, src.p, src.ld, src.coff, C);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

bool tc2_supported(const Geom& g) {
  Tc2Params PP;
  Geom gg = g; if (gg.B < 1) gg.B = 1;
  if (!(g.KH == 4 && g.KW == 4 && g.pad == 1 && (g.stride == 1 || g.stride == 2))) return false;
  return build_params2(gg, PP) && smem_bytes2(PP) <= 227 * 1024;
}

// layout kind the tc2 kernel wants for the INPUT of geometry g
int tc2_input_kind(const Geom& g) { return g.stride == 1 ? 0 : (g.mode == 0 ? 2 : 1); }

// Output channels per CTA for geometry g (B set): few pixel tiles (small feature maps) -> split the output channels over more
// CTAs (see the kernel's header comment); 128 = no split.  The plan packs the weights with this tile width so that every
// CTA's weights are one contiguous run (one large bulk copy per stage instead of 512-byte rows: 65 vs 18 GB/s per CTA).
int tc2_pick_ntw(const Geom& g, int sm_count) {
  Tc2Params PP;
  if (!build_params2(g, PP)) return 128;
  const int tiles128 = (PP.t.N_p + 127) / 128;
  const int ncap = PP.t.N_p < 128 ? PP.t.N_p : 128;
  static const int force = getenv("SVAE_NTW") ? atoi(getenv("SVAE_NTW")) : 0;
  int best = 128;
  for (int split = 2; split <= 4; split *= 2) {
    const int w = ncap / split;
    if (w < 32 || ncap % split || (long long)PP.tiles * tiles128 * split > (long long)sm_count) break;   // one wave, one CTA per SM
    best = w;
  }
  if (force) best = force;
  if (best < 128) {
    Tc2Params Q2;
    if (!(build_params2(g, Q2, best) && smem_bytes2(Q2) <= 227 * 1024)) best = 128;
  }
  return best;
}

bool tc2_fuse_supported(const Geom& g, View out, int C) {
  return C > 0 && C % 8 == 0 && C <= g.Cout && g.Cout % 4 == 0 && out.ld % 4 == 0 && out.coff % 4 == 0 &&
         (((uintptr_t)out.p & 15) == 0) && tc2_supported(g);
}

// Launch shape of geometry g: CTAs along x, CTAs per SM, and - for the all-resident (MODE 2) variant - accumulator sets per
// CTA and the tensor-memory allocation that holds them.  Returns false when the all-resident variant does not fit.
namespace {
struct Tc2Launch { int ctas, per_sm, ntiles, nsub, tpc; unsigned tmem_cols; };
bool tc2_launch_shape(const Tc2Params& PP, int sm_count, bool all_resident, Tc2Launch& L) {
  const TcParams& P = PP.t;
  const size_t smem = smem_bytes2(PP);
  L.nsub = ((P.N_p < 128 ? P.N_p : 128) + PP.ntw - 1) / PP.ntw;
  L.ntiles = (P.N_p + 127) / 128;
  int per_sm = (int)((227 * 1024) / smem);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  if (!all_resident) {
    if (PP.tmem_cols * (unsigned)per_sm > 512) per_sm = 1;
    int ctas = sm_count * per_sm / (L.ntiles * L.nsub);
    if (ctas < 1) ctas = 1;
    if (ctas > PP.tiles) ctas = PP.tiles;
    L.ctas = ctas; L.per_sm = per_sm; L.tpc = 0; L.tmem_cols = PP.tmem_cols;
    return true;
  }
  for (; per_sm >= 1; --per_sm) {
    int ctas = sm_count * per_sm / (L.ntiles * L.nsub);
    if (ctas < 1) return false;            // more channel tiles than the machine holds at once: no grid barrier possible
    if (ctas > PP.tiles) ctas = PP.tiles;
    const int tpc = (PP.tiles + ctas - 1) / ctas;
    unsigned cols = 32;
    while (cols < (unsigned)tpc * PP.acc_cols) cols <<= 1;
    if (tpc <= ACC_BUFS_MAX && cols * (unsigned)per_sm <= 512) {
      L.ctas = ctas; L.per_sm = per_sm; L.tpc = tpc; L.tmem_cols = cols;
      return true;
    }
  }
  return false;
}
}  // namespace

bool tc2_bnf_supported(const Geom& g, int w_tile_width, int sm_count, View out, int C) {
  // SVAE_BNF=1 enables the single-kernel conv + batch-norm forward.  Measured on B200 (CelebA-64, B = 100, T = 8): 97 fewer
  // launches per step but 12.39 instead of 11.94 ms/step - the all-resident accumulators take most of an SM's tensor memory, so
  // the next kernel's programmatic-dependent-launch prologue can no longer overlap, and 4 epilogue warps per CTA do the work a
  // standalone elementwise kernel spreads over the whole SM.  It also shares the hazard of every grid barrier outside a
  // cooperative launch: a co-resident CTA of another stream that blocks in tcgen05.alloc on columns held by a spinning CTA while
  // occupying the shared memory its sibling needs closes a cycle (seen at B = 256).  Off by default; kept for the record.
  static const bool enabled = getenv("SVAE_BNF") && getenv("SVAE_BNF")[0] == '1';
  if (!enabled || C <= 0 || C % 8 != 0 || C != g.Cout) return false;
  if (out.ld % 4 || out.coff % 4 || (((uintptr_t)out.p) & 15)) return false;
  Tc2Params PP;
  if (!build_params2(g, PP)) return false;
  const int best = w_tile_width < 128 ? w_tile_width : tc2_pick_ntw(g, sm_count);
  if (best < 128) {
    Tc2Params Q2;
    if (build_params2(g, Q2, best) && smem_bytes2(Q2) <= 227 * 1024) PP = Q2;
    else if (w_tile_width < 128) return false;
  }
  if (smem_bytes2(PP) > 227 * 1024) return false;
  Tc2Launch L;
  return tc2_launch_shape(PP, sm_count, true, L);
}

int tc2_gather_gemm(const LaunchCtx& lc, const Geom& g, const BfAct& in, int chan0, const void* w_packed, View out,
                    double* stats, const BnBwdFuse* fuse, int w_tile_width, const BnFwdFuse* fwd_fuse) {
  Tc2Params PP;
  if (!build_params2(g, PP)) { svae_global_error() = "tc2: unsupported geometry"; return -1; }
  {
    // weights packed with a narrower tile: that IS the CTA width; packed with 128: the heuristic may still split, the CTA
    // then gathers its columns row by row
    const int best = w_tile_width < 128 ? w_tile_width : tc2_pick_ntw(g, lc.sm_count);
    if (best < 128) {
      Tc2Params Q2;
      if (build_params2(g, Q2, best) && smem_bytes2(Q2) <= 227 * 1024) PP = Q2;
      else if (w_tile_width < 128) { svae_global_error() = "tc2: weights packed for a tile width this launch cannot use"; return -1; }
    }
    PP.wtw = w_tile_width < 128 ? w_tile_width : 128;
  }
  TcParams& P = PP.t;
  if (in.kind != tc2_input_kind(g) || in.Hp != P.Hp || in.Wp != P.Wp || (chan0 & 7)) {
    svae_global_error() = "tc2: activation copy is not in the layout this geometry reads";
    return -1;
  }
  P.wp = reinterpret_cast<const __nv_bfloat16*>(w_packed);
  P.out = out.p; P.out_ld = out.ld; P.out_coff = out.coff;
  P.out_vec = (out.ld % 4 == 0) && (out.coff % 4 == 0) && (((uintptr_t)out.p & 15) == 0) && (g.Cout % 4 == 0);
  P.stats = stats;
  if (fuse != nullptr) {
    if (stats != nullptr || !P.out_vec || fuse->C % 8 || fuse->C > g.Cout || (fuse->res != nullptr && g.accumulate)) { svae_global_error() = "tc2: fused batch-norm backward needs a vectorised output and no forward statistics"; return -1; }
    PP.fz = *fuse;
  }
  if (chan0 + P.Cin_p > in.Cpad || P.lo > in.front) { svae_global_error() = "tc2: channel window / slack of the activation copy too small"; return -1; }
  P.dbg = reinterpret_cast<unsigned long long*>(g_tc_debug_buffer);
  PP.a_src = in.p + ((long long)(chan0 / 8) * in.group_rows + in.front) * 8;
  PP.plane_rows = in.plane_rows;
  PP.group_rows = in.group_rows;
  const size_t smem = smem_bytes2(PP);
  static bool configured = false;
  static int occ[3] = {1, 1, 1};
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc2_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(tc2_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(tc2_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(tc2_conv_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM_CAP));
    configured = true;
  }
  Tc2Launch L;
  const bool bnf = fwd_fuse != nullptr;
  if (bnf && (fuse != nullptr || stats == nullptr || g.accumulate)) { svae_global_error() = "tc2: fused batch-norm forward needs forward statistics and a plain store"; return -1; }
  if (!tc2_launch_shape(PP, lc.sm_count, bnf, L)) { svae_global_error() = "tc2: accumulator tiles do not fit tensor memory for the fused batch-norm forward"; return -1; }
  if (bnf) {
    // the grid barrier needs every CTA resident at once: check what the device really grants this kernel at this footprint
    int fit = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, tc2_conv_kernel<2>, 224, smem));
    (void)occ;
    if (fit < L.per_sm) {
      if (fit < 1) { svae_global_error() = "tc2: fused batch-norm forward kernel does not fit an SM"; return -1; }
      Tc2Params Q1 = PP;       // retry with one CTA per SM
      const int ctas = lc.sm_count / (L.ntiles * L.nsub);
      if (ctas < 1) { svae_global_error() = "tc2: fused batch-norm forward grid exceeds the machine"; return -1; }
      L.ctas = ctas > PP.tiles ? PP.tiles : ctas; L.per_sm = 1;
      L.tpc = (PP.tiles + L.ctas - 1) / L.ctas;
      unsigned cols = 32;
      while (cols < (unsigned)L.tpc * PP.acc_cols) cols <<= 1;
      if (L.tpc > ACC_BUFS_MAX || cols > 512) { svae_global_error() = "tc2: accumulator tiles do not fit tensor memory (one CTA per SM)"; return -1; }
      L.tmem_cols = cols;
      (void)Q1;
    }
    PP.acc_bufs = L.tpc;
    PP.tmem_cols = L.tmem_cols;
    PP.ff = *fwd_fuse;
    PP.grid_ctas = (unsigned)(L.ctas * L.ntiles * L.nsub);
  }
  const double pix = (double)g.B * (g.Hin * g.Win < g.Hout * g.Wout ? g.Hin * g.Win : g.Hout * g.Wout);
  const double out_elems = (double)g.B * g.Hout * g.Wout * g.Cout;
  ProfScope ps(lc, KC_GEMM_TC, 2.0 * pix * g.KH * g.KW * g.Cin * g.Cout,
               2.0 * (double)g.B * g.Hin * g.Win * g.Cin + 4.0 * out_elems + 2.0 * g.KH * g.KW * g.Cin * g.Cout +
                   (bnf ? (fwd_fuse->res ? 4.0 : 0.0) * out_elems + (fwd_fuse->out ? 4.0 : 0.0) * out_elems +
                              (fwd_fuse->bf.a.p ? 2.0 : 0.0) * out_elems : 0.0), &g);
  const dim3 grid((unsigned)L.ctas, (unsigned)L.ntiles, (unsigned)L.nsub);
  if (lc.multi != nullptr) {
    if (bnf || fuse != nullptr) { svae_global_error() = "tc2: fused epilogues are not available in batched launches"; return -1; }
    if (lc.multi->add(&tc2_conv_multi_launch, &PP, sizeof PP, grid, dim3(224), smem) != 0) { svae_global_error() = lc.multi->err; return -1; }
    return 0;
  }
  if (bnf) CUDA_TRY(launch_k(lc, tc2_conv_kernel<2>, grid, dim3(224), smem, PP));
  else if (fuse != nullptr) CUDA_TRY(launch_k(lc, tc2_conv_kernel<1>, grid, dim3(224), smem, PP));
  else CUDA_TRY(launch_k(lc, tc2_conv_kernel<0>, grid, dim3(224), smem, PP));
  return 0;
}


// =====================================================================================================================
// TMA-fed weight gradient ("tc2w"): the same contraction as tc_wgrad_kernel, with both operands arriving as bf16 planar
// copies (BfAct).  X (shifted-window side) and dY live in the SAME padded pixel space, so a tile is
//   X : nplanes * JA bulk copies of HL pixels x 16 B        dY : JN bulk copies of 128 pixels x 16 B
// issued by one thread into a 3..4-deep ring; padding pixels of dY are zero, which is what masks the halo.
// Warps: 0 producer, 1 MMA issuer + TMEM owner, 2-5 epilogue (once, at the end: red.global.add of the CTA's partial dW).
namespace {

constexpr int W_STAGES_MAX = 4;

// Tap stacking.  The MMA is always M = 128 rows (X channels) x N (dY channels) x K = 16 pixels, and costs the same ~59 cycles
// whatever M holds: with CaB = 32 X channels three quarters of every instruction were padding.  Instead S = 128 / CaB taps are
// STACKED along M: the shared-memory X buffer holds `ncopy` plane blocks (JA 8-channel planes each) whose sources are
// pre-shifted so that ONE start-address offset serves every block of a tap set -
//   stride 1: S copies of the halo shifted by 0..3 pixels (and by one padded row for S = 8); a set = the kw (and kh pair) taps
//             of its rows, reached by moving the start address by (kh-1)*Wp - 1 pixels;
//   stride 2: the four parity planes, plane (ph,pw) shifted by its first tap's (dh,dw): tap (i,j) of EVERY plane is then
//             i*Wp + j pixels further, so one MMA covers the four taps that share (i,j) - without a single extra load.
// 16 / S instructions per 16 pixels instead of 16, 16 / S accumulators of N columns instead of 16.
struct Tw2Params {
  TwParams t;
  const __nv_bfloat16* x_src;   // X copy at (group 0, pixel 0 of plane 0)
  const __nv_bfloat16* y_src;   // dY copy at (group 0, pixel 0)
  long long x_plane_rows, x_group_rows, y_group_rows;
  unsigned x_pitch, y_pitch;    // bytes between 8-channel planes in shared memory
  unsigned x_bytes, y_bytes;    // per stage
  unsigned stage_bytes;
  int stages;
  int S, nsets, sets_per_cta, ngroups, ncopy;
  int copy_plane[8], copy_off[8];            // source parity plane and pixel offset of every plane block
  int set_block[16], set_off[16];            // per tap set: first plane block, start offset in pixels (added to lo)
  signed char set_tap[16][8];                // per tap set: tap of row block i (i < S)
  unsigned tmem_cols;
};

struct SmemHeaderW2 {
  unsigned long long full[W_STAGES_MAX], empty[W_STAGES_MAX], acc_done;
  unsigned tmem_base, pad;
};

__device__ __forceinline__ void tc2_wgrad_body(const Tw2Params& PP) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const TwParams& P = PP.t;
  SmemHeaderW2* hdr = reinterpret_cast<SmemHeaderW2*>(smem_raw);
  unsigned char* bufs = smem_raw + 128;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uniform_u32((unsigned)(tid >> 5));
  int by = blockIdx.y;
  const int nb = by % P.nblocks; by /= P.nblocks;
  const int mb = by % P.mblocks; by /= P.mblocks;
  const int grp = by;
  const int a0 = mb * P.CaB, b0 = nb * P.N;
  const int my_tiles = (int)((P.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < W_STAGES_MAX; ++i) { mbar_init(smem_u32(&hdr->full[i]), 1); mbar_init(smem_u32(&hdr->empty[i]), 1); }
    mbar_init(smem_u32(&hdr->acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&hdr->tmem_base), PP.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = uniform_u32(hdr->tmem_base);

  if (warp == 0) {
    // ---- producer ----
    const __nv_bfloat16* xs = PP.x_src + (long long)(a0 / 8) * PP.x_group_rows * 8;
    const __nv_bfloat16* ys = PP.y_src + (long long)(b0 / 8) * PP.y_group_rows * 8;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % PP.stages;
      if (it >= PP.stages) mbar_wait(smem_u32(&hdr->empty[s]), (unsigned)((it / PP.stages) - 1) & 1u);
      if (lane == 0) {
        const long long q0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TILE_M;
        const unsigned bar = smem_u32(&hdr->full[s]);
        mbar_expect_tx(bar, PP.x_bytes + PP.y_bytes);
        const unsigned xb = smem_u32(bufs + (size_t)s * PP.stage_bytes);
        const unsigned yb = xb + PP.x_bytes;
        for (int c = 0; c < PP.ncopy; ++c)
          for (int j = 0; j < P.JA; ++j)
            bulk_g2s(xb + (unsigned)(c * P.JA + j) * PP.x_pitch,
                     xs + ((long long)j * PP.x_group_rows + PP.copy_plane[c] * PP.x_plane_rows + (q0 - P.lo + PP.copy_off[c])) * 8,
                     PP.x_pitch, bar);
        for (int j = 0; j < P.JN; ++j)
          bulk_g2s(yb + (unsigned)j * PP.y_pitch, ys + ((long long)j * PP.y_group_rows + q0) * 8, PP.y_pitch, bar);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer: D[a, b] (+)= X[a x 16 px] * dY[16 px x b], one accumulator per tap of the CTA's tap group ----
    const unsigned idesc = make_idesc_mn(TILE_M, P.N);
    const unsigned long long a_hi = ((unsigned long long)((PP.x_pitch >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((128u >> 4) & 0x3FFF) << 16);
    const unsigned long long b_hi = ((unsigned long long)((PP.y_pitch >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                                    ((unsigned long long)((128u >> 4) & 0x3FFF) << 16);
    const bool leader = elect_one();
    unsigned first = 1u;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % PP.stages;
      mbar_wait(smem_u32(&hdr->full[s]), (unsigned)(it / PP.stages) & 1u);
      tc_fence_after();
      if (leader) {
        const unsigned xb4 = smem_u32(bufs + (size_t)s * PP.stage_bytes) >> 4;
        const unsigned yb4 = xb4 + (PP.x_bytes >> 4);
        for (int tl = 0; tl < PP.sets_per_cta; ++tl) {
          const int set = grp * PP.sets_per_cta + tl;
          const unsigned d_tmem = tmem_base + (unsigned)(tl * P.N);
          unsigned a_lo = xb4 + (unsigned)(PP.set_block[set] * P.JA) * (PP.x_pitch >> 4) + (unsigned)(P.lo + PP.set_off[set]);
          unsigned b_lo = yb4;
          unsigned acc_flag = first ^ 1u;
#pragma unroll
          for (int k = 0; k < TILE_M / 16; ++k) {
            umma_bf16(d_tmem, a_hi | (unsigned long long)(a_lo & 0x3FFF), b_hi | (unsigned long long)(b_lo & 0x3FFF), idesc, acc_flag);
            acc_flag = 1u;
            a_lo += 16u;
            b_lo += 16u;
          }
        }
        umma_commit(smem_u32(&hdr->empty[s]));
      }
      first = 0u;
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&hdr->acc_done));
    __syncwarp();
  } else {
    // ---- epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1): row = X channel, columns = dY channels ----
    mbar_wait(smem_u32(&hdr->acc_done), 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;           // accumulator row = (row block i, X channel a)
    const int iblk = m / P.CaB;
    const int a = m - iblk * P.CaB;
    for (int tl = 0; tl < PP.sets_per_cta; ++tl) {
      const int set = grp * PP.sets_per_cta + tl;
      const int tap = iblk < PP.S ? PP.set_tap[set][iblk] : 0;
      for (int n0 = 0; n0 < P.N; n0 += 32) {
        float v[32];
        tmem_ld_upto32(tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(tl * P.N + n0), v, P.N - n0);
        if (iblk < PP.S && a0 + a < P.Ca && my_tiles > 0) {
          const int aa = a0 + a;
          float* dst = P.dw2 == nullptr ? P.dw + ((size_t)tap * P.Ca + aa) * P.Cb + b0 + n0
                       : aa < P.a_split ? P.dw + ((size_t)tap * P.a_split + aa) * P.Cb + b0 + n0
                                        : P.dw2 + ((size_t)tap * (P.Ca - P.a_split) + (aa - P.a_split)) * P.Cb + b0 + n0;
          const int ncols = max(0, min(min(32, P.N - n0), P.Cb - (b0 + n0)));
          if (P.dw_vec) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              if (k < ncols) atomicAdd(reinterpret_cast<float4*>(dst + k), make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < ncols) atomicAdd(dst + k, v[k]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, PP.tmem_cols);
  }
}

__global__ void __launch_bounds__(192, 1) tc2_wgrad_kernel(const __grid_constant__ Tw2Params PP) { tc2_wgrad_body(PP); }
// batched variant (LaunchCtx::multi): blockIdx.z = item
__global__ void __launch_bounds__(192, 1) tc2_wgrad_multi_kernel(const __grid_constant__ MultiArgs<Tw2Params> A) { tc2_wgrad_body(A.v[blockIdx.z]); }
static void tc2_wgrad_multi_launch(const void* host_args, int items, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  multi_launch_chunks<Tw2Params>(tc2_wgrad_multi_kernel, host_args, items, grid, block, smem, st, 1u);
}

size_t smem_bytes_w2(const Tw2Params& PP);

// two_per_sm: half the tensor memory (half the taps per CTA) and at most ~110 KB of shared memory, so that two CTAs share an
// SM and interleave their MMAs - narrow layers (N <= 64) are paced by the per-CTA issue latency (59 cycles per MMA alone,
// 40 with two CTAs, scripts/mma_rate.cu), not by the tensor pipe.
bool build_wparams2(const Geom& g, Tw2Params& PP, bool two_per_sm = false) {
  memset(&PP, 0, sizeof PP);
  TwParams& P = PP.t;
  if (!build_wparams(g, P, 148, two_per_sm ? 256 : 512)) return false;
  // ---- tap stacking tables (see Tw2Params)
  {
    static const bool stack = !(getenv("SVAE_WGRAD_STACK") && getenv("SVAE_WGRAD_STACK")[0] == '0');
    int S = stack ? 128 / P.CaB : 1;
    if (P.mode == 1 && S > 4) S = 4;
    if (S > 8) S = 8;
    PP.S = S; PP.nsets = 16 / S;
    const int Wp = P.Wp;
    if (P.mode == 0) {
      PP.ncopy = S;
      for (int c = 0; c < S; ++c) { PP.copy_plane[c] = 0; PP.copy_off[c] = (c >> 2) * Wp + (c & 3); }
      if (S == 1) PP.copy_off[0] = 0;
      if (S == 2) { PP.copy_off[0] = 0; PP.copy_off[1] = 1; }
      for (int st = 0; st < PP.nsets; ++st) {
        PP.set_block[st] = 0;
        if (S == 1) { PP.set_off[st] = P.shift[st]; PP.set_tap[st][0] = (signed char)st; }
        else if (S == 2) { const int kh = st >> 1, kp = st & 1; PP.set_off[st] = (kh - 1) * Wp + 2 * kp - 1;
                           for (int i = 0; i < 2; ++i) PP.set_tap[st][i] = (signed char)(kh * 4 + 2 * kp + i); }
        else if (S == 4) { PP.set_off[st] = (st - 1) * Wp - 1; for (int i = 0; i < 4; ++i) PP.set_tap[st][i] = (signed char)(st * 4 + i); }
        else { PP.set_off[st] = (2 * st - 1) * Wp - 1; for (int i = 0; i < 8; ++i) PP.set_tap[st][i] = (signed char)((2 * st + (i >> 2)) * 4 + (i & 3)); }
      }
    } else {
      // parity plane p = (ph << 1) | pw holds taps kh in {1,3} (ph = 0: dh = 0,+1) or {0,2} (ph = 1: dh = -1,0), same for kw
      PP.ncopy = 4;
      for (int pl = 0; pl < 4; ++pl) {
        const int ph = pl >> 1, pw = pl & 1;
        PP.copy_plane[pl] = pl;
        PP.copy_off[pl] = S > 1 ? (ph ? -Wp : 0) + (pw ? -1 : 0) : 0;
      }
      auto tap_of = [](int pl, int i, int j) {
        const int ph = pl >> 1, pw = pl & 1;
        const int kh = ph == 0 ? (i == 0 ? 1 : 3) : (i == 0 ? 0 : 2);
        const int kw = pw == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 0 : 2);
        return kh * 4 + kw;
      };
      if (S == 1) {
        for (int st = 0; st < 16; ++st) { PP.set_block[st] = P.plane[st]; PP.set_off[st] = P.shift[st]; PP.set_tap[st][0] = (signed char)st; }
      } else {
        // sets ordered (first plane block, i, j); S planes per set
        int st = 0;
        for (int pb = 0; pb < 4; pb += S)
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j, ++st) {
              PP.set_block[st] = pb; PP.set_off[st] = i * Wp + j;
              for (int k = 0; k < S; ++k) PP.set_tap[st][k] = (signed char)tap_of(pb + k, i, j);
            }
      }
    }
    int cap = two_per_sm ? 256 : 512;
    PP.sets_per_cta = cap / P.N; if (PP.sets_per_cta > PP.nsets) PP.sets_per_cta = PP.nsets;
    { static const int spc = getenv("SVAE_WGRAD_SPC") ? atoi(getenv("SVAE_WGRAD_SPC")) : 0;   // experiment: cap on tap sets per CTA
      if (spc > 0 && PP.sets_per_cta > spc) PP.sets_per_cta = spc; }
    if (PP.sets_per_cta < 1) return false;
    while (PP.nsets % PP.sets_per_cta) --PP.sets_per_cta;
    PP.ngroups = PP.nsets / PP.sets_per_cta;
    unsigned cols = (unsigned)(PP.sets_per_cta * P.N), t = 32;
    while (t < cols) t <<= 1;
    PP.tmem_cols = t;
  }
  if (two_per_sm && (P.N > 64 || PP.tmem_cols > 256)) return false;
  const unsigned hl = (unsigned)((P.HL + 7) & ~7);
  PP.x_pitch = hl * 16u;
  PP.y_pitch = 128u * 16u;
  PP.x_bytes = (unsigned)(PP.ncopy * P.JA) * PP.x_pitch;
  PP.y_bytes = (unsigned)P.JN * PP.y_pitch;
  PP.stage_bytes = PP.x_bytes + PP.y_bytes;
  // The A descriptor always spans 16 planes (M = 128 rows): planes beyond the JA real ones are never copied, they only have
  // to be readable, i.e. inside the CTA's shared-memory window (their rows are ignored by the epilogue).
  const size_t reach = 128 + (size_t)(W_STAGES_MAX - 1) * 0;   // header
  (void)reach;
  int st = W_STAGES_MAX;
  if (two_per_sm) {
    const size_t cap = 110 * 1024;
    PP.stages = st;
    while (PP.stages > 2 && smem_bytes_w2(PP) > cap) --PP.stages;
    return smem_bytes_w2(PP) <= cap;
  }
  while (st > 1 && 128 + (size_t)st * PP.stage_bytes > 200 * 1024) --st;
  if (128 + (size_t)st * PP.stage_bytes > 200 * 1024) return false;
  PP.stages = st;
  return true;
}

size_t smem_bytes_w2(const Tw2Params& PP) {
  const TwParams& P = PP.t;
  // last stage's X buffer must see 16 readable planes from its highest parity-plane base
  const size_t last_x = 128 + (size_t)(PP.stages - 1) * PP.stage_bytes;
  int max_block = 0;
  for (int st = 0; st < PP.nsets; ++st) if (PP.set_block[st] > max_block) max_block = PP.set_block[st];
  const size_t need_read = last_x + (size_t)(max_block * P.JA + 16) * PP.x_pitch;
  size_t sz = 128 + (size_t)PP.stages * PP.stage_bytes;
  if (need_read > sz) sz = need_read;
  return sz + 128;
}

}  // namespace

bool tc2_wgrad_supported(const Geom& g) {
  Tw2Params PP;
  Geom gg = g; if (gg.B < 1) gg.B = 1;
  return build_wparams2(gg, PP) && smem_bytes_w2(PP) <= 227 * 1024;
}

// g: conv-gather geometry (X = conv input side in its tc2 input layout, dY = conv output side in the layout of the SAME
// padded pixel space: kind 0 for stride 1, kind 1 for stride 2)
int tc2_wgrad(const LaunchCtx& lc, const Geom& g, const BfAct& x, const BfAct& dy, float* dw, float* dw2, int a_split) {
  Tw2Params PP;
  if (!build_wparams2(g, PP)) { svae_global_error() = "tc2 wgrad: unsupported geometry"; return -1; }
  int per_sm = 1;
  {
    static const int mode = getenv("SVAE_WGRAD_2CTA") ? atoi(getenv("SVAE_WGRAD_2CTA")) : 0;   // measured neutral (4.94 vs 5.07 ms of wgrad per step): off
    Tw2Params Q2;
    if (mode && build_wparams2(g, Q2, true)) { PP = Q2; per_sm = 2; }
  }
  TwParams& P = PP.t;
  const int ykind = g.stride == 1 ? 0 : 1;
  if (x.kind != tc2_input_kind(g) || x.Hp != P.Hp || x.Wp != P.Wp || dy.kind != ykind || dy.Hp != P.Hp || dy.Wp != P.Wp ||
      P.lo + (P.mode == 1 && PP.S > 1 ? P.Wp + 1 : 0) > x.front || P.mblocks * P.CaB > x.Cpad || P.nblocks * P.N > dy.Cpad) {
    svae_global_error() = "tc2 wgrad: operand copies are not in the layouts this geometry reads";
    return -1;
  }
  PP.x_src = x.p + (long long)x.front * 8;
  PP.y_src = dy.p + (long long)dy.front * 8;
  PP.x_plane_rows = x.plane_rows; PP.x_group_rows = x.group_rows; PP.y_group_rows = dy.group_rows;
  P.dw = dw; P.dw2 = dw2; P.a_split = a_split;
  P.dw_vec = (g.Cout % 4 == 0) && (((uintptr_t)dw & 15) == 0) && (((uintptr_t)dw2 & 15) == 0);
  const size_t smem = smem_bytes_w2(PP);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc2_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(tc2_wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM_CAP));
    configured = true;
  }
  // How the work is cut: `sets_per_cta` tap sets per CTA (the rest of the sets go to other CTAs of blockIdx.y, which re-read
  // the operands) x `splits` CTAs along the pixels (each flushes its partial dW with vector red.global.add).  Measured on
  // B200 (scripts/wgrad_ab.sh, profiles/r2_wgrad_split.md): the flush sustains ~800 B/cycle over the whole chip whatever the
  // number of CTAs, one CTA pulls ~34 B/cycle of operands from L2, an MMA issues every 59 (N <= 64) or 64 cycles.  With every
  // SM holding all sets of a 16..128-channel layer the flush was 2-6x the MMA time; the model picks the cheapest cut.
  int best_sets = PP.sets_per_cta;
  long long best_splits = 1;
  {
    static const bool model = !(getenv("SVAE_WGRAD_MODEL") && getenv("SVAE_WGRAD_MODEL")[0] == '0');
    double best = 1e30;
    const int sm = lc.sm_count * per_sm;
    for (int sets = PP.sets_per_cta; sets >= 1; --sets) {
      if (PP.nsets % sets) continue;
      const int gy_ = (PP.nsets / sets) * P.mblocks * P.nblocks;
      long long smax = ((long long)sm + gy_ - 1) / gy_;
      if (smax > P.tiles) smax = P.tiles;
      if (smax < 1) smax = 1;
      for (long long sp = smax; sp >= 1; sp = (sp > 8 ? sp * 7 / 8 : sp - 1)) {
        const double tpc = (double)((P.tiles + sp - 1) / sp);
        const double load = tpc * PP.stage_bytes / 34.0;
        const double mma = tpc * sets * 8.0 * (P.N > 64 ? 64.0 : 59.0);
        const double acc_bytes = (double)sets * 128.0 * P.N * 4.0;
        const double flush = std::max(acc_bytes / 32.0, (double)gy_ * sp * acc_bytes / 800.0);
        const double cost = std::max(load, mma) + flush + 1500.0;
        if (cost < best) { best = cost; best_sets = sets; best_splits = sp; }
        if (!model) break;   // SVAE_WGRAD_MODEL=0: all sets on every CTA, as many pixel splits as SMs (round-1 behaviour)
      }
      if (!model) break;
    }
  }
  if (best_sets != PP.sets_per_cta) {
    PP.sets_per_cta = best_sets;
    PP.ngroups = PP.nsets / best_sets;
    unsigned cols = (unsigned)(best_sets * P.N), t = 32;
    while (t < cols) t <<= 1;
    PP.tmem_cols = t;
  }
  const int gy = PP.ngroups * P.mblocks * P.nblocks;
  long long splits = best_splits;
  { static const int ms = getenv("SVAE_WGRAD_MAXSPLIT") ? atoi(getenv("SVAE_WGRAD_MAXSPLIT")) : 0;   // experiment: cap on the pixel splits
    if (ms > 0 && splits > ms) splits = ms; }
  if (splits < 1) splits = 1;
  const double pix = (double)g.B * (g.Hin * g.Win < g.Hout * g.Wout ? g.Hin * g.Win : g.Hout * g.Wout);
  ProfScope ps(lc, KC_WGRAD_TC, 2.0 * pix * 16 * g.Cin * g.Cout,
               2.0 * ((double)g.B * g.Hin * g.Win * g.Cin + (double)g.B * g.Hout * g.Wout * g.Cout) + 4.0 * 16.0 * g.Cin * g.Cout, &g);
  if (lc.multi != nullptr) {
    if (lc.multi->add(&tc2_wgrad_multi_launch, &PP, sizeof PP, dim3((unsigned)splits, (unsigned)gy), dim3(192), smem) != 0) {
      svae_global_error() = lc.multi->err;
      return -1;
    }
    return 0;
  }
  tc2_wgrad_kernel<<<dim3((unsigned)splits, (unsigned)gy), 192, smem, lc.stream>>>(PP);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

void* g_tc_debug_buffer = nullptr;   // set through svae_debug_set_buffer (scripts/diag_phases.py only)

bool tc_supported(const Geom& g) {
  const bool conv = g.KH == 4 && g.KW == 4 && g.pad == 1 && (g.stride == 1 || g.stride == 2);
  const bool fc = g.KH == 1 && g.KW == 1 && g.Hin == 1 && g.Win == 1;
  if (!conv && !fc) return false;
  if (g.Cin < 1 || g.Cout < 1) return false;
  if (fc) return tc_fc_supported(g.Cin, g.Cout);   // kernels_fc.cu; skinny projections / heads stay on the fp32 row-dot kernels
  TcParams P;
  Geom gg = g;
  if (gg.B < 1) gg.B = 1;
  if (!build_params(gg, P)) return false;
  return smem_bytes(P) <= 227 * 1024;
}

size_t tc_packed_bytes(const Geom& g) { return (size_t)g.KH * g.KW * round16(g.Cin) * round16(g.Cout) * 2; }

int tc_pack_weights(const LaunchCtx& lc, const Geom& g, const float* w, void* w_packed, int tile_width) {
  const int Cin_p = round16(g.Cin), N_p = round16(g.Cout);
  const long long total = (long long)g.KH * g.KW * Cin_p * N_p;
  int blocks = (int)((total + 255) / 256);
  if (blocks > lc.sm_count * 8) blocks = lc.sm_count * 8;
  ProfScope ps(lc, KC_PACK, 0.0, 6.0 * total);
  tc_pack_kernel<<<blocks, 256, 0, lc.stream>>>(g, w, reinterpret_cast<__nv_bfloat16*>(w_packed), pick_kc(Cin_p), Cin_p, N_p,
                                                tile_width);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

TcPackEntry tc_pack_entry(const Geom& g, const float* w, void* w_packed, int tile_width, const float* w2, int n1) {
  TcPackEntry e;
  e.g = g; e.w = w; e.out = w_packed; e.TW = tile_width; e.w2 = w2; e.n1 = n1;
  e.Cin_p = round16(g.Cin); e.N_p = round16(g.Cout); e.KC = pick_kc(e.Cin_p);
  e.total = (long long)g.KH * g.KW * e.Cin_p * e.N_p;
  return e;
}

int tc_pack_batched(const LaunchCtx& lc, const void* dev_entries, int n, double total_elems) {
  if (n <= 0) return 0;
  ProfScope ps(lc, KC_PACK, 0.0, 6.0 * total_elems);
  // blocks per entry: enough in total to fill the machine a few times over, whether the table holds every layer of the model
  // (one launch after a full Adam step) or one chain step's layers (bucketed update)
  // (entries differ 1000x in size: blocks beyond an entry's element count exit at once, the large entries need the threads)
  int bx = (64 * lc.sm_count + n - 1) / n;
  if (bx < (lc.sm_count >= 100 ? 32 : 4)) bx = lc.sm_count >= 100 ? 32 : 4;   // a reduced SM budget (bucketed update beside the chain): few blocks
  if (bx > 256) bx = 256;
  tc_pack_batched_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, lc.stream>>>(reinterpret_cast<const TcPackEntry*>(dev_entries));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int tc_gather_gemm(const LaunchCtx& lc, const Geom& g, View in, const void* w_packed, View out, double* stats) {
  if (g.KH == 1) { svae_global_error() = "tcgen05 conv kernel: fully-connected layers run on tc_fc"; return -1; }
  TcParams P;
  if (!build_params(g, P)) { svae_global_error() = "tcgen05: unsupported geometry"; return -1; }
  P.in = in.p; P.in_ld = in.ld; P.in_coff = in.coff;
  P.in_vec = (in.ld % 4 == 0) && (in.coff % 4 == 0) && (((uintptr_t)in.p & 15) == 0) && (g.Cin % 8 == 0);
  P.wp = reinterpret_cast<const __nv_bfloat16*>(w_packed);
  P.out = out.p; P.out_ld = out.ld; P.out_coff = out.coff;
  P.out_vec = (out.ld % 4 == 0) && (out.coff % 4 == 0) && (((uintptr_t)out.p & 15) == 0) && (g.Cout % 4 == 0);
  P.stats = stats;
  P.dbg = reinterpret_cast<unsigned long long*>(g_tc_debug_buffer);
  const size_t smem = smem_bytes(P);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const long long tiles = (P.Q + TILE_M - 1) / TILE_M;
  const int ntiles = (P.N_p + 127) / 128;
  const double pix = (double)g.B * (g.Hin * g.Win < g.Hout * g.Wout ? g.Hin * g.Win : g.Hout * g.Wout);
  ProfScope ps(lc, KC_GEMM_TC, 2.0 * pix * g.KH * g.KW * g.Cin * g.Cout,
               4.0 * ((double)g.B * g.Hin * g.Win * g.Cin + (double)g.B * g.Hout * g.Wout * g.Cout) +
                   2.0 * g.KH * g.KW * g.Cin * g.Cout, &g);
  tc_conv_kernel<<<dim3((unsigned)tiles, (unsigned)ntiles), 192, smem, lc.stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static Geom wgrad_conv_geom(const Geom& fwd) {
  // conv-gather geometry whose "input" side is X and "output" side is dY (for a transposed conv the roles swap)
  Geom g = fwd;
  if (fwd.mode == 1) {
    g.Hin = fwd.Hout; g.Win = fwd.Wout; g.Cin = fwd.Cout;
    g.Hout = fwd.Hin; g.Wout = fwd.Win; g.Cout = fwd.Cin;
    g.mode = 0;
  }
  return g;
}

bool tc_wgrad_supported(const Geom& fwd) {
  if (fwd.KH == 1 && fwd.KW == 1 && fwd.Hin == 1 && fwd.Win == 1) return tc_fc_supported(fwd.Cin, fwd.Cout);
  TwParams P;
  Geom g = wgrad_conv_geom(fwd);
  if (g.B < 1) g.B = 1;
  return build_wparams(g, P, 148);
}

int tc_wgrad(const LaunchCtx& lc, const Geom& g, View x, View dy, float* dw) {
  if (g.KH == 1)   // fully connected: dW[K,N] = X[B,K]^T . dY[B,N]
    return tc_fc(lc, 2, x.p + x.coff, x.ld, dy.p + dy.coff, dy.ld, dw, g.Cout, g.B, g.Cin, g.Cout, 0);
  TwParams P;
  if (!build_wparams(g, P, lc.sm_count)) { svae_global_error() = "tcgen05 wgrad: unsupported geometry"; return -1; }
  P.x = x.p; P.x_ld = x.ld; P.x_coff = x.coff;
  P.x_vec = (x.ld % 4 == 0) && (x.coff % 4 == 0) && (((uintptr_t)x.p & 15) == 0) && (g.Cin % 8 == 0);
  P.dy = dy.p; P.dy_ld = dy.ld; P.dy_coff = dy.coff;
  P.dy_vec = (dy.ld % 4 == 0) && (dy.coff % 4 == 0) && (((uintptr_t)dy.p & 15) == 0) && (g.Cout % 8 == 0);
  P.dw = dw;
  P.dw_vec = (g.Cout % 4 == 0) && (((uintptr_t)dw & 15) == 0);
  const size_t smem = 128 + (size_t)P.bufs * (P.x_buf_bytes + P.y_buf_bytes);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int gy = P.ngroups * P.mblocks * P.nblocks;
  long long splits = (lc.sm_count + gy - 1) / gy;    // about one CTA per SM: every extra split costs a full flush of atomics
  if (splits > P.tiles) splits = P.tiles;
  if (splits < 1) splits = 1;
  const double pix = (double)g.B * (g.Hin * g.Win < g.Hout * g.Wout ? g.Hin * g.Win : g.Hout * g.Wout);
  ProfScope ps(lc, KC_WGRAD_TC, 2.0 * pix * 16 * g.Cin * g.Cout,
               4.0 * ((double)g.B * g.Hin * g.Win * g.Cin + (double)g.B * g.Hout * g.Wout * g.Cout + 16.0 * g.Cin * g.Cout), &g);
  tc_wgrad_kernel<<<dim3((unsigned)splits, (unsigned)gy), 160, smem, lc.stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

