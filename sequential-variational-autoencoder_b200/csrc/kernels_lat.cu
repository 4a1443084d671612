// kernels_lat.cu - the latent projections P_i = fc_bn_lrelu(z_i) of split_latent (sequential_vae.py:1796-1806) as ONE
// kernel per direction.
//
// The "contraction" has K = latent_dims[i] inputs (2-3 in the MNIST / CelebA nets, 20-30 in the LSUN one) and up to
// S*S*F = 32768 outputs; the batch norm is 2-D (per output feature over the batch, abstract_network.py:64-71).  One thread
// owns one output feature for the whole batch, so the batch statistics, the backward sums, the weight gradient column and
// the beta gradient never leave its registers: no statistics pass, no atomics on the weights, no intermediate dy tensor.
//   forward : y = z.W (kept for the backward), batch statistics, lrelu(bn(y)) -> fp32 concat slot + bf16 planar copy
//   backward: g = da*lrelu', S1 = sum g, S2 = sum g*xhat, dy = rstd*(g - S1/B - xhat*S2/B),
//             dW[:,f] = z^T dy, dbeta[f] = S1, dz[b,:] += dy[b,f] W[:,f] (warp transpose-reduction -> shared -> global)
// HBM-bound: forward writes 4+4+2 bytes per element, backward reads 8.
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int LAT_RY = 8;                 // row phases per block
constexpr int LAT_THREADS = 32 * LAT_RY;

__device__ __forceinline__ float lat_act(float v, int act) {
  if (act == ACT_LRELU) return fmaxf(fminf(v * SVAE_LRELU_SLOPE, 0.f), v);
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}
__device__ __forceinline__ float lat_act_grad(float pre, int act) {
  if (act == ACT_LRELU) return pre > 0.f ? 1.f : SVAE_LRELU_SLOPE;
  if (act == ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  return 1.f;
}

// Block = 32 features (x, one warp per row phase: coalesced 128-byte rows) x LAT_RY row phases (y): a thread owns feature
// f for the rows b = y, y + LAT_RY, ...; per-feature sums are combined across the row phases through shared memory.
template <int NM>
__device__ __forceinline__ void lat_fwd_fused_kernel_body(const float* __restrict__ z, int z_ld, int z_coff, const float* __restrict__ w,
                     const float* __restrict__ beta, int B, int KZ, int N, int act, float* __restrict__ y,
                     double* __restrict__ stats, FeatView out, BfDst bf) {
  extern __shared__ float s_z[];   // [B][NM]
  __shared__ double s_s[LAT_RY][32], s_q[LAT_RY][32];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < B * NM; i += LAT_THREADS) {
    const int b = i / NM, k = i - b * NM;
    s_z[i] = k < KZ ? z[(size_t)b * z_ld + z_coff + k] : 0.f;
  }
  __syncthreads();
  const int f = blockIdx.x * 32 + threadIdx.x;
  const bool live = f < N;
  float wk[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) wk[k] = (live && k < KZ) ? __ldg(w + (size_t)k * N + f) : 0.f;
  double s = 0.0, q = 0.0;
  if (live)
    for (int b = threadIdx.y; b < B; b += LAT_RY) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < NM; ++k) v = fmaf(s_z[b * NM + k], wk[k], v);
      if (y != nullptr) y[(size_t)b * N + f] = v;
      s += (double)v;
      q += (double)v * (double)v;
    }
  s_s[threadIdx.y][threadIdx.x] = s;
  s_q[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (!live) return;
  s = 0.0; q = 0.0;
#pragma unroll
  for (int i = 0; i < LAT_RY; ++i) { s += s_s[i][threadIdx.x]; q += s_q[i][threadIdx.x]; }   // same order in every row phase
  if (threadIdx.y == 0 && stats != nullptr) {
    stats[f] = s;            // the same (sum, sum of squares) the statistics-fused contraction epilogues leave
    stats[N + f] = q;
  }
  const double m = s / (double)B;
  double var = q / (double)B - m * m;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m;
  const float rstd = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
  const float sh = beta[f] - mean * rstd;
  // feature -> (pixel, channel) inside the fp32 concat slot and inside the bf16 planar copy
  const int pix = f / out.inner, c = f - pix * out.inner;
  const size_t ooff = (size_t)pix * out.ld + out.coff + c;
  const size_t ostride = (size_t)out.ppr * out.ld;
  const bool has_bf = bf.a.p != nullptr;
  const int binner = bf.inner ? bf.inner : out.inner, bppr = bf.inner ? bf.ppr : out.ppr;
  const int bpix = f / binner, bc = f - bpix * binner;
  const int HW = has_bf ? bf.a.H * bf.a.W : 1, Wd = has_bf ? bf.a.W : 1;
  for (int b = threadIdx.y; b < B; b += LAT_RY) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < NM; ++k) v = fmaf(s_z[b * NM + k], wk[k], v);
    v = lat_act(fmaf(v, rstd, sh), act);
    if (out.p != nullptr) out.p[(size_t)b * ostride + ooff] = v;
    if (has_bf) {
      const int64_t p = (int64_t)b * bppr + bpix;
      const int n = (int)(p / HW);
      const int hw = (int)(p - (int64_t)n * HW);
      const int hh = hw / Wd;
      bf.a.p[bf_index(bf.a, n, hh, hw - hh * Wd, bf.coff + bc)] = __float2bfloat16_rn(v);
    }
  }
}
template <int NM>
__global__ void __launch_bounds__(LAT_THREADS)
lat_fwd_fused_kernel(const float* __restrict__ z, int z_ld, int z_coff, const float* __restrict__ w,
                     const float* __restrict__ beta, int B, int KZ, int N, int act, float* __restrict__ y,
                     double* __restrict__ stats, FeatView out, BfDst bf) { lat_fwd_fused_kernel_body<NM>(z, z_ld, z_coff, w, beta, B, KZ, N, act, y, stats, out, bf); }

// Forward for narrow latent groups (K <= 8): y = z.W is a rank-K map of the batch, so its batch statistics follow from the
// K x K second moments of z -  mean_y[f] = m.W[:,f],  var_y[f] = W[:,f]^T Cov(z) W[:,f]  - and no pass over the batch is needed
// to normalise: every block first reduces the moments of z (B x K values, L2-resident), then the kernel is a pure map over
// (row chunk, feature tile) with K FMAs per element.  Rows split freely over blockIdx.y (generation runs B = 4096).
template <int NM>
__device__ __forceinline__ void lat_fwd_moment_kernel_body(const float* __restrict__ z, int z_ld, int z_coff, const float* __restrict__ w,
                      const float* __restrict__ beta, int B, int KZ, int N, int act, int rows_per_block,
                      float* __restrict__ y, double* __restrict__ stats, FeatView out, BfDst bf,
                      const double* __restrict__ mom_in, double* __restrict__ mom_out) {
  constexpr int NP = NM * (NM + 1) / 2;
  __shared__ double s_part[LAT_RY][NM + NP];
  __shared__ double s_mom[NM + NP];           // sum z_k, then sum z_k z_k' (k <= k')
  extern __shared__ float s_z[];              // [rows_per_block][NM] rows of this block
  const int tid = threadIdx.y * 32 + threadIdx.x;
  double acc[NM + NP];
#pragma unroll
  for (int i = 0; i < NM + NP; ++i) acc[i] = 0.0;
  // mom_in: the moments were reduced once by a one-block launch of this kernel (mom_out != nullptr) - large batches
  if (mom_in == nullptr)
  for (int b = tid; b < B; b += LAT_THREADS) {
    float zk[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) zk[k] = k < KZ ? __ldg(z + (size_t)b * z_ld + z_coff + k) : 0.f;
    int q = NM;
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      acc[k] += (double)zk[k];
#pragma unroll
      for (int k2 = k; k2 < NM; ++k2) acc[q++] += (double)zk[k] * (double)zk[k2];
    }
  }
#pragma unroll
  for (int i = 0; i < NM + NP; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_part[threadIdx.y][i] = v;
  }
  const int r0 = blockIdx.y * rows_per_block, r1 = min(B, r0 + rows_per_block);
  for (int i = tid; i < (r1 - r0) * NM; i += LAT_THREADS) {
    const int b = i / NM, k = i - b * NM;
    s_z[i] = k < KZ ? z[(size_t)(r0 + b) * z_ld + z_coff + k] : 0.f;
  }
  __syncthreads();
  if (tid < NM + NP) {
    double v = 0.0;
    if (mom_in != nullptr) v = mom_in[tid];
    else {
#pragma unroll
      for (int i = 0; i < LAT_RY; ++i) v += s_part[i][tid];
    }
    s_mom[tid] = v;
    if (mom_out != nullptr) mom_out[tid] = v;
  }
  if (mom_out != nullptr) return;              // moments-only launch
  __syncthreads();
  const int f = blockIdx.x * 32 + threadIdx.x;
  if (f >= N) return;
  float wk[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) wk[k] = k < KZ ? __ldg(w + (size_t)k * N + f) : 0.f;
  // batch statistics of y[:, f] from the moments of z
  double sy = 0.0, syy = 0.0;
  {
    int q = NM;
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      sy += s_mom[k] * (double)wk[k];
#pragma unroll
      for (int k2 = k; k2 < NM; ++k2) {
        const double t = s_mom[q++] * (double)wk[k] * (double)wk[k2];
        syy += k2 == k ? t : 2.0 * t;
      }
    }
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && stats != nullptr) { stats[f] = sy; stats[N + f] = syy; }
  const double m = sy / (double)B;
  double var = syy / (double)B - m * m;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m;
  const float rstd = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
  const float sh = beta[f] - mean * rstd;
  const int pix = f / out.inner, c = f - pix * out.inner;
  const size_t ooff = (size_t)pix * out.ld + out.coff + c;
  const size_t ostride = (size_t)out.ppr * out.ld;
  const bool has_bf = bf.a.p != nullptr;
  const int binner = bf.inner ? bf.inner : out.inner, bppr = bf.inner ? bf.ppr : out.ppr;
  const int bpix = f / binner, bc = f - bpix * binner;
  const int HW = has_bf ? bf.a.H * bf.a.W : 1, Wd = has_bf ? bf.a.W : 1;
  // one image per row (ppr == H*W): the element's address is linear in the image index - no index math in the loop
  const bool lin = has_bf && bppr == HW;
  size_t bf_base = 0, bf_step = 0;
  if (lin) {
    const int hh = bpix / Wd;
    bf_base = bf_index(bf.a, 0, hh, bpix - hh * Wd, bf.coff + bc);
    bf_step = bf_index(bf.a, 1, hh, bpix - hh * Wd, bf.coff + bc) - bf_base;
  }
  for (int b = r0 + threadIdx.y; b < r1; b += LAT_RY) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < NM; ++k) v = fmaf(s_z[(b - r0) * NM + k], wk[k], v);
    if (y != nullptr) y[(size_t)b * N + f] = v;
    v = lat_act(fmaf(v, rstd, sh), act);
    if (out.p != nullptr) out.p[(size_t)b * ostride + ooff] = v;
    if (lin) {
      bf.a.p[bf_base + (size_t)b * bf_step] = __float2bfloat16_rn(v);
    } else if (has_bf) {
      const int64_t p = (int64_t)b * bppr + bpix;
      const int n = (int)(p / HW);
      const int hw = (int)(p - (int64_t)n * HW);
      const int hh = hw / Wd;
      bf.a.p[bf_index(bf.a, n, hh, hw - hh * Wd, bf.coff + bc)] = __float2bfloat16_rn(v);
    }
  }
}
template <int NM>
__global__ void __launch_bounds__(LAT_THREADS)
lat_fwd_moment_kernel(const float* __restrict__ z, int z_ld, int z_coff, const float* __restrict__ w,
                      const float* __restrict__ beta, int B, int KZ, int N, int act, int rows_per_block,
                      float* __restrict__ y, double* __restrict__ stats, FeatView out, BfDst bf,
                      const double* __restrict__ mom_in, double* __restrict__ mom_out) { lat_fwd_moment_kernel_body<NM>(z, z_ld, z_coff, w, beta, B, KZ, N, act, rows_per_block, y, stats, out, bf, mom_in, mom_out); }

// RB rows of NM products are reduced across the warp (= 32 features of one row phase) at a time: 32 values per
// transpose-reduction.  A block walks feature tiles blockIdx.x, +gridDim.x, ... and keeps its dz partial sums in shared
// memory across them, so the global atomics on dz number gridDim.x per element, not N/32.
template <int NM>
__device__ __forceinline__ void lat_bwd_fused_kernel_body(FeatView da, const float* __restrict__ y, const double* __restrict__ stats,
                     const float* __restrict__ beta, const float* __restrict__ z, int z_ld, int z_coff,
                     const float* __restrict__ w, int B, int KZ, int N, int act, float* __restrict__ dw,
                     float* __restrict__ dbeta, float* __restrict__ dz, int dz_ld, int dz_coff) {
  constexpr int RB = 32 / NM;
  extern __shared__ float s_mem[];   // [B][NM] z, then [B][NM] dz accumulators
  __shared__ float s_a[LAT_RY][32], s_b[LAT_RY][32];
  __shared__ float s_w[LAT_RY][NM][33];
  float* s_z = s_mem;
  float* s_dz = s_mem + (size_t)B * NM;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < B * NM; i += LAT_THREADS) {
    const int b = i / NM, k = i - b * NM;
    s_z[i] = k < KZ ? z[(size_t)b * z_ld + z_coff + k] : 0.f;
    s_dz[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x;
  const size_t dstride = (size_t)da.ppr * da.ld;
  const int tiles = (N + 31) / 32;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int f = tile * 32 + lane;
    const bool live = f < N;        // dead lanes keep running: they take part in the warp reductions with zeros
    float mean = 0.f, rstd = 0.f, bt = 0.f;
    size_t doff = 0;
    float wk[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) wk[k] = 0.f;
    if (live) {
      const double m = stats[f] / (double)B;
      double var = stats[N + f] / (double)B - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
      bt = beta[f];
      const int pix = f / da.inner;
      doff = (size_t)pix * da.ld + da.coff + (f - pix * da.inner);
#pragma unroll
      for (int k = 0; k < NM; ++k) wk[k] = k < KZ ? __ldg(w + (size_t)k * N + f) : 0.f;
    }
    float S1 = 0.f, S2 = 0.f;
    if (live) {
#pragma unroll 4
      for (int b = threadIdx.y; b < B; b += LAT_RY) {
        const float xh = (__ldg(y + (size_t)b * N + f) - mean) * rstd;
        const float g = __ldg(da.p + (size_t)b * dstride + doff) * lat_act_grad(xh + bt, act);
        S1 += g;
        S2 += g * xh;
      }
    }
    s_a[threadIdx.y][lane] = S1;
    s_b[threadIdx.y][lane] = S2;
    __syncthreads();
    S1 = 0.f; S2 = 0.f;
#pragma unroll
    for (int i = 0; i < LAT_RY; ++i) { S1 += s_a[i][lane]; S2 += s_b[i][lane]; }
    const float m1 = S1 / (float)B, m2 = S2 / (float)B;
    float aw[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) aw[k] = 0.f;
    for (int b0 = threadIdx.y; b0 < B; b0 += LAT_RY * RB) {
      float v[32];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int b = b0 + r * LAT_RY;
        float d = 0.f;
        if (live && b < B) {
          const float xh = (__ldg(y + (size_t)b * N + f) - mean) * rstd;
          const float g = __ldg(da.p + (size_t)b * dstride + doff) * lat_act_grad(xh + bt, act);
          d = rstd * (g - m1 - xh * m2);
        }
#pragma unroll
        for (int k = 0; k < NM; ++k) {
          if (b < B) aw[k] = fmaf(s_z[b * NM + k], d, aw[k]);
          v[r * NM + k] = d * wk[k];
        }
      }
      const float tot = tcptx::warp_colsum32(v, lane);    // lane l: sum over the warp's features of value l = (row l/NM, k l%NM)
      const int b = b0 + (lane / NM) * LAT_RY;
      if (b < B) s_dz[b * NM + (lane % NM)] += tot;       // row b belongs to this warp alone (b % LAT_RY == threadIdx.y)
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) s_w[threadIdx.y][k][lane] = aw[k];
    __syncthreads();
    if (live && threadIdx.y == 0) {
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        if (k < KZ) {
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < LAT_RY; ++i) t += s_w[i][k][lane];
          dw[(size_t)k * N + f] += t;                      // this block owns columns f of the weight gradient
        }
      }
      if (dbeta != nullptr) dbeta[f] = S1;
    }
    __syncthreads();                                       // s_a / s_b / s_w are reused by the next tile
  }
  for (int i = tid; i < B * NM; i += LAT_THREADS) {
    const int b = i / NM, k = i - b * NM;
    if (k < KZ) atomicAdd(dz + (size_t)b * dz_ld + dz_coff + k, s_dz[i]);
  }
}
template <int NM>
__global__ void __launch_bounds__(LAT_THREADS)
lat_bwd_fused_kernel(FeatView da, const float* __restrict__ y, const double* __restrict__ stats,
                     const float* __restrict__ beta, const float* __restrict__ z, int z_ld, int z_coff,
                     const float* __restrict__ w, int B, int KZ, int N, int act, float* __restrict__ dw,
                     float* __restrict__ dbeta, float* __restrict__ dz, int dz_ld, int dz_coff) { lat_bwd_fused_kernel_body<NM>(da, y, stats, beta, z, z_ld, z_coff, w, B, KZ, N, act, dw, dbeta, dz, dz_ld, dz_coff); }

// ---- 2-D batch norm of a fully-connected block (fc_bn_lrelu, abstract_network.py:64-71: enc.fc, dec.fc) ------------------
// rows = batch (<= a few hundred), feats = 384 .. 6144.  The same ownership as above - 32 features x LAT_RY row phases per
// block - makes the statistics AND the normalisation one kernel (forward) and both backward passes one kernel.
__global__ void __launch_bounds__(LAT_THREADS)
bn2d_fwd_kernel(const float* __restrict__ y, const float* __restrict__ beta, int B, int N, int act,
                double* __restrict__ stats, FeatView out, BfDst bf) {
  __shared__ double s_s[LAT_RY][32], s_q[LAT_RY][32];
  pdl_wait();
  pdl_trigger();
  const int f = blockIdx.x * 32 + threadIdx.x;
  const bool live = f < N;
  double s = 0.0, q = 0.0;
  if (live) {
#pragma unroll 4
    for (int b = threadIdx.y; b < B; b += LAT_RY) {
      const float v = __ldg(y + (size_t)b * N + f);
      s += (double)v;
      q += (double)v * (double)v;
    }
  }
  s_s[threadIdx.y][threadIdx.x] = s;
  s_q[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (!live) return;
  s = 0.0; q = 0.0;
#pragma unroll
  for (int i = 0; i < LAT_RY; ++i) { s += s_s[i][threadIdx.x]; q += s_q[i][threadIdx.x]; }
  if (threadIdx.y == 0) { stats[f] = s; stats[N + f] = q; }
  const double m = s / (double)B;
  double var = q / (double)B - m * m;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m;
  const float rstd = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
  const float sh = beta[f] - mean * rstd;
  const int pix = f / out.inner, c = f - pix * out.inner;
  const size_t ooff = (size_t)pix * out.ld + out.coff + c;
  const size_t ostride = (size_t)out.ppr * out.ld;
  const bool has_bf = bf.a.p != nullptr;
  const int binner = bf.inner ? bf.inner : out.inner, bppr = bf.inner ? bf.ppr : out.ppr;
  const int bpix = f / binner, bc = f - bpix * binner;
  const int HW = has_bf ? bf.a.H * bf.a.W : 1, Wd = has_bf ? bf.a.W : 1;
  for (int b = threadIdx.y; b < B; b += LAT_RY) {
    const float v = lat_act(fmaf(__ldg(y + (size_t)b * N + f), rstd, sh), act);
    if (out.p != nullptr) out.p[(size_t)b * ostride + ooff] = v;
    if (has_bf) {
      const int64_t p = (int64_t)b * bppr + bpix;
      const int n = (int)(p / HW);
      const int hw = (int)(p - (int64_t)n * HW);
      const int hh = hw / Wd;
      bf.a.p[bf_index(bf.a, n, hh, hw - hh * Wd, bf.coff + bc)] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void __launch_bounds__(LAT_THREADS)
bn2d_bwd_kernel(FeatView da, const float* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ beta,
                int B, int N, int act, float* __restrict__ dy, float* __restrict__ dbeta) {
  __shared__ float s_a[LAT_RY][32], s_b[LAT_RY][32];
  pdl_wait();
  pdl_trigger();
  const int f = blockIdx.x * 32 + threadIdx.x;
  const bool live = f < N;
  float mean = 0.f, rstd = 0.f, bt = 0.f;
  size_t doff = 0;
  const size_t dstride = (size_t)da.ppr * da.ld;
  if (live) {
    const double m = stats[f] / (double)B;
    double var = stats[N + f] / (double)B - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)SVAE_BN_EPS));
    bt = beta[f];
    const int pix = f / da.inner;
    doff = (size_t)pix * da.ld + da.coff + (f - pix * da.inner);
  }
  float S1 = 0.f, S2 = 0.f;
  if (live) {
#pragma unroll 4
    for (int b = threadIdx.y; b < B; b += LAT_RY) {
      const float xh = (__ldg(y + (size_t)b * N + f) - mean) * rstd;
      const float g = __ldg(da.p + (size_t)b * dstride + doff) * lat_act_grad(xh + bt, act);
      S1 += g;
      S2 += g * xh;
    }
  }
  s_a[threadIdx.y][threadIdx.x] = S1;
  s_b[threadIdx.y][threadIdx.x] = S2;
  __syncthreads();
  if (!live) return;
  S1 = 0.f; S2 = 0.f;
#pragma unroll
  for (int i = 0; i < LAT_RY; ++i) { S1 += s_a[i][threadIdx.x]; S2 += s_b[i][threadIdx.x]; }
  const float m1 = S1 / (float)B, m2 = S2 / (float)B;
  if (threadIdx.y == 0 && dbeta != nullptr) dbeta[f] = S1;
#pragma unroll 4
  for (int b = threadIdx.y; b < B; b += LAT_RY) {
    const float xh = (__ldg(y + (size_t)b * N + f) - mean) * rstd;
    const float g = __ldg(da.p + (size_t)b * dstride + doff) * lat_act_grad(xh + bt, act);
    dy[(size_t)b * N + f] = rstd * (g - m1 - xh * m2);
  }
}

int width_class(int n) { return n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

}  // namespace

bool lat_fused_supported(int B, int KZ) { return KZ >= 1 && KZ <= 32 && B >= 1 && (size_t)B * width_class(KZ) * 8 <= 160 * 1024; }

template <typename K>
static int allow_smem(K kernel, size_t bytes) {   // opt in to > 48 KB of static + dynamic shared memory (large batch x wide latent group)
  if (bytes > 8 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

#define LAT_DISPATCH(KZV, ...)                                  \
  switch (width_class(KZV)) {                                   \
    case 4: { constexpr int NM = 4; __VA_ARGS__ } break;        \
    case 8: { constexpr int NM = 8; __VA_ARGS__ } break;        \
    case 16: { constexpr int NM = 16; __VA_ARGS__ } break;      \
    default: { constexpr int NM = 32; __VA_ARGS__ } break;      \
  }

int lat_fwd_fused(const LaunchCtx& lc, View z, const float* w, const float* beta, int B, int KZ, int N, int act, float* y,
                  double* stats, FeatView out, BfDst bf, double* mom_scratch) {
  if (!(KZ >= 1 && KZ <= 8 && B >= 1) && !lat_fused_supported(B, KZ)) { svae_global_error() = "lat_fwd_fused: unsupported batch / latent width"; return -1; }
  Geom tg{}; tg.B = B; tg.Cin = KZ; tg.Cout = N;
  ProfScope ps(lc, KC_SKINNY, 4.0 * B * KZ * (double)N, (double)B * N * (4.0 + (out.p ? 4.0 : 0.0) + (bf.a.p ? 2.0 : 0.0)), &tg);
  const unsigned blocks = (unsigned)((N + 31) / 32);
  if (KZ <= 8) {   // statistics from the moments of z: rows split over blockIdx.y, no pass over the batch
    const int rpb = B <= 128 ? B : 128;
    const dim3 grid(blocks, (unsigned)((B + rpb - 1) / rpb));
    // large batches: the moments of z are reduced once by a one-block launch instead of by every block
    const double* mom = (B > 512 && mom_scratch != nullptr && lc.multi == nullptr) ? mom_scratch : nullptr;
    if (lc.multi != nullptr) {   // batched launch: recorded, every block reduces the moments of z itself
      if (KZ <= 4)
        return MULTI_RECORD(lat_fwd_moment_kernel_body<4>, LAT_THREADS, lc, grid, dim3(32, LAT_RY), (size_t)rpb * 4 * sizeof(float), z.p, z.ld,
                            z.coff, w, beta, B, KZ, N, act, rpb, y, stats, out, bf, (const double*)nullptr, (double*)nullptr);
      return MULTI_RECORD(lat_fwd_moment_kernel_body<8>, LAT_THREADS, lc, grid, dim3(32, LAT_RY), (size_t)rpb * 8 * sizeof(float), z.p, z.ld,
                          z.coff, w, beta, B, KZ, N, act, rpb, y, stats, out, bf, (const double*)nullptr, (double*)nullptr);
    }
    if (KZ <= 4) {
      if (mom) lat_fwd_moment_kernel<4><<<dim3(1, 1), dim3(32, LAT_RY), 0, lc.stream>>>(
                   z.p, z.ld, z.coff, w, beta, B, KZ, N, act, 0, nullptr, nullptr, FeatView{}, BfDst{}, nullptr, mom_scratch);
      lat_fwd_moment_kernel<4><<<grid, dim3(32, LAT_RY), (size_t)rpb * 4 * sizeof(float), lc.stream>>>(
          z.p, z.ld, z.coff, w, beta, B, KZ, N, act, rpb, y, stats, out, bf, mom, nullptr);
    } else {
      if (mom) lat_fwd_moment_kernel<8><<<dim3(1, 1), dim3(32, LAT_RY), 0, lc.stream>>>(
                   z.p, z.ld, z.coff, w, beta, B, KZ, N, act, 0, nullptr, nullptr, FeatView{}, BfDst{}, nullptr, mom_scratch);
      lat_fwd_moment_kernel<8><<<grid, dim3(32, LAT_RY), (size_t)rpb * 8 * sizeof(float), lc.stream>>>(
          z.p, z.ld, z.coff, w, beta, B, KZ, N, act, rpb, y, stats, out, bf, mom, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  LAT_DISPATCH(KZ, {
    const size_t smem = (size_t)B * NM * sizeof(float);
    if (lc.multi != nullptr) {
      if (smem > 40 * 1024) { svae_global_error() = "lat_fwd_fused: batch too large for a batched launch"; return -1; }
      return MULTI_RECORD(lat_fwd_fused_kernel_body<NM>, LAT_THREADS, lc, dim3(blocks), dim3(32, LAT_RY), smem, z.p, z.ld, z.coff, w, beta, B, KZ,
                          N, act, y, stats, out, bf);
    }
    if (allow_smem(lat_fwd_fused_kernel<NM>, smem) != 0) return -3;
    lat_fwd_fused_kernel<NM><<<blocks, dim3(32, LAT_RY), smem, lc.stream>>>(z.p, z.ld, z.coff, w, beta, B, KZ, N, act, y, stats, out, bf);
  });
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lat_bwd_fused(const LaunchCtx& lc, FeatView da, const float* y, const double* stats, const float* beta, View z,
                  const float* w, int B, int KZ, int N, int act, float* dw, float* dbeta, View dz) {
  if (!lat_fused_supported(B, KZ)) { svae_global_error() = "lat_bwd_fused: unsupported batch / latent width"; return -1; }
  Geom tg{}; tg.B = B; tg.Cin = KZ; tg.Cout = N;
  ProfScope ps(lc, KC_SKINNY, 8.0 * B * KZ * (double)N, 8.0 * B * (double)N, &tg);
  unsigned blocks = (unsigned)((N + 31) / 32);
  static const unsigned per_sm = getenv("SVAE_LAT_BLOCKS_PER_SM") ? (unsigned)atoi(getenv("SVAE_LAT_BLOCKS_PER_SM")) : 4u;
  if (blocks > per_sm * (unsigned)lc.sm_count) blocks = per_sm * (unsigned)lc.sm_count;
  LAT_DISPATCH(KZ, {
    const size_t smem = 2 * (size_t)B * NM * sizeof(float);
    if (lc.multi != nullptr) {
      if (smem > 40 * 1024) { svae_global_error() = "lat_bwd_fused: batch too large for a batched launch"; return -1; }
      return MULTI_RECORD(lat_bwd_fused_kernel_body<NM>, LAT_THREADS, lc, dim3(blocks), dim3(32, LAT_RY), smem, da, y, stats, beta, z.p, z.ld,
                          z.coff, w, B, KZ, N, act, dw, dbeta, dz.p, dz.ld, dz.coff);
    }
    if (allow_smem(lat_bwd_fused_kernel<NM>, smem) != 0) return -3;
    lat_bwd_fused_kernel<NM><<<blocks, dim3(32, LAT_RY), smem, lc.stream>>>(da, y, stats, beta, z.p, z.ld, z.coff, w, B, KZ, N, act, dw,
                                                                        dbeta, dz.p, dz.ld, dz.coff);
  });
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// statistics + normalisation + activation of a [B, N] fully-connected output in one launch (stats are left for the backward)
int bn2d_fwd(const LaunchCtx& lc, const float* y, const float* beta, int B, int N, int act, double* stats, FeatView out, BfDst bf) {
  Geom tg{}; tg.B = B; tg.Cout = N;
  ProfScope ps(lc, KC_BN_FWD, 7.0 * B * (double)N, (double)B * N * (4.0 + (out.p ? 4.0 : 0.0) + (bf.a.p ? 2.0 : 0.0)), &tg);
  CUDA_TRY(launch_k(lc, bn2d_fwd_kernel, dim3((unsigned)((N + 31) / 32)), dim3(32, LAT_RY), 0, y, beta, B, N, act, stats, out, bf));
  return 0;
}

// both passes of the batch-norm backward of a [B, N] fully-connected output in one launch
int bn2d_bwd(const LaunchCtx& lc, FeatView da, const float* y, const double* stats, const float* beta, int B, int N, int act,
             float* dy, float* dbeta) {
  Geom tg{}; tg.B = B; tg.Cout = N;
  ProfScope ps(lc, KC_BN_BWD_APPLY, 13.0 * B * (double)N, 12.0 * B * (double)N, &tg);
  CUDA_TRY(launch_k(lc, bn2d_bwd_kernel, dim3((unsigned)((N + 31) / 32)), dim3(32, LAT_RY), 0, da, y, stats, beta, B, N, act, dy, dbeta));
  return 0;
}
