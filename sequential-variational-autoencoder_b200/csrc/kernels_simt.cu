// kernels_simt.cu - fp32 SIMT contractions: the strict-parity kernel family (SVAE_OPERAND_FP32) and the home of the
// shapes that are not tensor-core shaped (Cin = 1/3 first layers, N = 3/1 output deconvs, skinny heads / latent
// projections).  One gather-form implicit GEMM covers conv fwd, transposed-conv fwd, both dgrads and the fully
// connected layers; one wgrad kernel covers every weight gradient (see Geom in common.cuh).
//
// Reference ops replaced: tf.contrib.layers.convolution2d / convolution2d_transpose / fully_connected data paths and
// their autodiff gradients (abstract_network.py:18,37,56,65; sequential_vae.py:1273).
#include "common.cuh"

namespace {

constexpr int BK = 16;

__device__ __forceinline__ bool gather_coord(const Geom& g, int o, int k, int in_size, int& i) {
  if (g.mode == 0) {
    i = o * g.stride - g.pad + k;
    return i >= 0 && i < in_size;
  }
  int t = o + g.pad - k;
  if (t < 0) return false;
  i = t / g.stride;
  return (t - i * g.stride) == 0 && i < in_size;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gather_gemm_kernel(Geom g, View in, const float* __restrict__ w, View out, double* __restrict__ stats) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_PER = (BM * BK) / NT;  // A elements per thread per chunk
  constexpr int W_PER = (BN * BK + NT - 1) / NT;
  static_assert((BM * BK) % NT == 0, "tile/threads mismatch");
  __shared__ float As[BK][BM + 1];
  __shared__ float Ws[BK][BN + 1];
  __shared__ float s_sum[BN], s_sq[BN];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t M = (int64_t)g.B * g.Hout * g.Wout;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // rows this thread stages into As: kk = tid % BK, m_local = tid / BK + i * (NT / BK)
  const int a_kk = tid % BK;
  int rb[A_PER], roh[A_PER], row_[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    int64_t m = m0 + tid / BK + i * (NT / BK);
    if (m < M) {
      int ow = (int)(m % g.Wout);
      int64_t r = m / g.Wout;
      roh[i] = (int)(r % g.Hout);
      rb[i] = (int)(r / g.Hout);
      row_[i] = ow;
    } else {
      rb[i] = -1; roh[i] = 0; row_[i] = 0;
    }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int taps = g.KH * g.KW;
  for (int tap = 0; tap < taps; ++tap) {
    const int kh = tap / g.KW, kw = tap % g.KW;
    size_t abase[A_PER];
    bool aval[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int ih = 0, iw = 0;
      bool v = rb[i] >= 0 && gather_coord(g, roh[i], kh, g.Hin, ih) && gather_coord(g, row_[i], kw, g.Win, iw);
      aval[i] = v;
      abase[i] = v ? (((size_t)rb[i] * g.Hin + ih) * g.Win + iw) * in.ld + in.coff : 0;
    }
    for (int c0 = 0; c0 < g.Cin; c0 += BK) {
      // stage A
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int ci = c0 + a_kk;
        float v = (aval[i] && ci < g.Cin) ? __ldg(in.p + abase[i] + ci) : 0.f;
        As[a_kk][tid / BK + i * (NT / BK)] = v;
      }
      // stage W
#pragma unroll
      for (int i = 0; i < W_PER; ++i) {
        int e = tid + i * NT;
        if (e < BN * BK) {
          int kk, n;
          if (g.w_out_major == 0) { n = e % BN; kk = e / BN; } else { kk = e % BK; n = e / BK; }
          int ci = c0 + kk, co = n0 + n;
          float v = 0.f;
          if (ci < g.Cin && co < g.Cout)
            v = g.w_out_major == 0 ? __ldg(w + ((size_t)tap * g.Cin + ci) * g.Cout + co)
                                   : __ldg(w + ((size_t)tap * g.Cout + co) * g.Cin + ci);
          Ws[kk][n] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Ws[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue: store (+ accumulate) and per-channel batch-norm statistics of the *stored* value
  if (stats != nullptr) {
    for (int i = tid; i < BN; i += NT) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
    __syncthreads();
  }
  float csum[TN], csq[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { csum[j] = 0.f; csq[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
    float* orow = out.p + (size_t)m * out.ld + out.coff;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = n0 + tx * TN + j;
      if (co >= g.Cout) continue;
      float v = acc[i][j];
      if (g.accumulate) v += orow[co];
      orow[co] = v;
      csum[j] += v;
      csq[j] += v * v;
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      atomicAdd(&s_sum[tx * TN + j], csum[j]);
      atomicAdd(&s_sq[tx * TN + j], csq[j]);
    }
    __syncthreads();
    for (int i = tid; i < BN; i += NT) {
      int co = n0 + i;
      if (co < g.Cout) {
        atomicAdd(&stats[co], (double)s_sum[i]);
        atomicAdd(&stats[g.Cout + co], (double)s_sq[i]);
      }
    }
  }
}

// dW(tap,a,b) += sum over a slice of rows
template <int BA, int BB, int TA, int TB>
__global__ void __launch_bounds__((BA / TA) * (BB / TB))
wgrad_kernel(Geom g, View x, View dy, float* __restrict__ dw, int tilesA, int tilesB, int64_t rows_per_split) {
  constexpr int NT = (BA / TA) * (BB / TB);
  __shared__ float Xs[BK][BA + 1];
  __shared__ float Ys[BK][BB + 1];
  const int tid = threadIdx.x;
  const int tx = tid % (BB / TB), ty = tid / (BB / TB);
  int bid = blockIdx.x;
  const int tb = bid % tilesB; bid /= tilesB;
  const int ta = bid % tilesA; bid /= tilesA;
  const int tap = bid;
  const int kh = tap / g.KW, kw = tap % g.KW;
  const int a0 = ta * BA, b0 = tb * BB;
  const int Ca = g.Cin, Cb = g.Cout;
  const int64_t M = (int64_t)g.B * g.Hout * g.Wout;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(M, r_begin + rows_per_split);

  float acc[TA][TB];
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    // stage X (gathered) : element e -> row kk = e / BA, col a = e % BA (channel-contiguous => coalesced)
    for (int e = tid; e < BK * BA; e += NT) {
      int kk = e / BA, a = e % BA;
      int64_t m = r0 + kk;
      float v = 0.f;
      if (m < r_end && a0 + a < Ca) {
        int ow = (int)(m % g.Wout);
        int64_t r = m / g.Wout;
        int oh = (int)(r % g.Hout);
        int b = (int)(r / g.Hout);
        int ih, iw;
        if (gather_coord(g, oh, kh, g.Hin, ih) && gather_coord(g, ow, kw, g.Win, iw))
          v = __ldg(x.p + (((size_t)b * g.Hin + ih) * g.Win + iw) * x.ld + x.coff + a0 + a);
      }
      Xs[kk][a] = v;
    }
    for (int e = tid; e < BK * BB; e += NT) {
      int kk = e / BB, b = e % BB;
      int64_t m = r0 + kk;
      float v = 0.f;
      if (m < r_end && b0 + b < Cb) v = __ldg(dy.p + (size_t)m * dy.ld + dy.coff + b0 + b);
      Ys[kk][b] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TA], b[TB];
#pragma unroll
      for (int i = 0; i < TA; ++i) a[i] = Xs[kk][ty * TA + i];
#pragma unroll
      for (int j = 0; j < TB; ++j) b[j] = Ys[kk][tx * TB + j];
#pragma unroll
      for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TA; ++i) {
    int a = a0 + ty * TA + i;
    if (a >= Ca) continue;
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      int b = b0 + tx * TB + j;
      if (b >= Cb) continue;
      float* dst = dw + ((size_t)tap * Ca + a) * Cb + b;
      if (gridDim.y == 1) *dst += acc[i][j]; else atomicAdd(dst, acc[i][j]);
    }
  }
}

// ---- skinny fully-connected kernels ------------------------------------------------------------------------------
// Recognition heads (layers.fully_connected, sequential_vae.py:1592-1609): [B, K] x [K, n] with K = 8192..32768 and
// n = latent_dims[l] <= 32 - GEMV-like, bound by reading the activation map once.  All heads that read the same feature
// map run in ONE launch; K (forward) / the batch (weight gradient) is split across blocks and combined with atomics.
constexpr int SK_NMAX = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mu_pre / sd_pre [B, Z] (pre-zeroed) += flat[b, kslice] . W_h[kslice, :]  (+ bias from the first K slice)
// NM: compile-time bound on the widths in this launch (the latent groups are 2-3 wide in the MNIST / CelebA nets, 20-30
// in the LSUN one): the per-element inner loops are NM long, not SK_NMAX.
template <int NM>
__device__ __forceinline__ void heads_fwd_kernel_body(const float* __restrict__ flat, int K, HeadSet hs, float* __restrict__ mu_pre, float* __restrict__ sd_pre,
                 int Z) {
  __shared__ float red[8][SK_NMAX];
  const int b = blockIdx.x;
  const int kper = (K + gridDim.y - 1) / gridDim.y;
  const int k0 = blockIdx.y * kper, k1 = min(K, k0 + kper);
  const float* arow = flat + (size_t)b * K;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int h = 0; h < hs.nheads; ++h) {
    const int n = hs.n[h];
    const float* __restrict__ w = hs.w[h];
    float acc[NM];
#pragma unroll
    for (int i = 0; i < NM; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
      const float av = __ldg(arow + k);
      const float* wr = w + (size_t)k * n;
#pragma unroll
      for (int i = 0; i < NM; ++i)
        if (i < n) acc[i] = fmaf(av, __ldg(wr + i), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      if (i < n) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[wid][i] = v;
      }
    }
    __syncthreads();
    if (threadIdx.x < n) {
      float v = blockIdx.y == 0 ? hs.b[h][threadIdx.x] : 0.f;
      for (int wi = 0; wi < 8; ++wi) v += red[wi][threadIdx.x];
      float* dst = (hs.is_sd[h] ? sd_pre : mu_pre) + (size_t)b * Z + hs.col[h] + threadIdx.x;
      atomicAdd(dst, v);
    }
    __syncthreads();
  }
}
template <int NM>
__global__ void __launch_bounds__(256)
heads_fwd_kernel(const float* __restrict__ flat, int K, HeadSet hs, float* __restrict__ mu_pre, float* __restrict__ sd_pre,
                 int Z) { heads_fwd_kernel_body<NM>(flat, K, hs, mu_pre, sd_pre, Z); }

// d_flat[b, k] = sum_h sum_n d_h[b, col_h + n] * W_h[k, n]   (overwrite)
__device__ __forceinline__ void heads_dgrad_kernel_body(HeadSet hs, const float* __restrict__ dmu, const float* __restrict__ dsd, int Z, int K,
                   float* __restrict__ d_flat) {
  __shared__ float s_d[4 * SK_NMAX];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < hs.nheads * SK_NMAX; i += blockDim.x) {
    const int h = i / SK_NMAX, n = i % SK_NMAX;
    s_d[i] = n < hs.n[h] ? (hs.is_sd[h] ? dsd : dmu)[(size_t)b * Z + hs.col[h] + n] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float v = 0.f;
  for (int h = 0; h < hs.nheads; ++h) {
    const int n = hs.n[h];
    const float* wr = hs.w[h] + (size_t)k * n;
    for (int i = 0; i < n; ++i) v = fmaf(s_d[h * SK_NMAX + i], __ldg(wr + i), v);
  }
  d_flat[(size_t)b * K + k] = v;
}
__global__ void __launch_bounds__(256)
heads_dgrad_kernel(HeadSet hs, const float* __restrict__ dmu, const float* __restrict__ dsd, int Z, int K,
                   float* __restrict__ d_flat) { heads_dgrad_kernel_body(hs, dmu, dsd, Z, K, d_flat); }

// gw_h[k, n] += sum_{b in slice} flat[b, k] * d_h[b, col_h + n] ; gb_h[n] += sum_b d_h[b, col_h + n]   (grads pre-zeroed)
template <int NM>
__device__ __forceinline__ void heads_wgrad_kernel_body(const float* __restrict__ flat, HeadSet hs, const float* __restrict__ dmu, const float* __restrict__ dsd,
                   int B, int Z, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int bper = (B + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * bper, r1 = min(B, r0 + bper);
  for (int h = 0; h < hs.nheads; ++h) {
    const int n = hs.n[h];
    const float* __restrict__ d = (hs.is_sd[h] ? dsd : dmu) + hs.col[h];
    if (k < K) {
      float acc[NM];
#pragma unroll
      for (int i = 0; i < NM; ++i) acc[i] = 0.f;
      for (int b = r0; b < r1; ++b) {
        const float av = __ldg(flat + (size_t)b * K + k);
        const float* dr = d + (size_t)b * Z;
#pragma unroll
        for (int i = 0; i < NM; ++i)
          if (i < n) acc[i] = fmaf(av, __ldg(dr + i), acc[i]);
      }
#pragma unroll
      for (int i = 0; i < NM; ++i)
        if (i < n) atomicAdd(hs.gw[h] + (size_t)k * n + i, acc[i]);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < n) {
      float sum = 0.f;
      for (int b = 0; b < B; ++b) sum += d[(size_t)b * Z + threadIdx.x];
      hs.gb[h][threadIdx.x] = sum;
    }
  }
}
template <int NM>
__global__ void __launch_bounds__(256)
heads_wgrad_kernel(const float* __restrict__ flat, HeadSet hs, const float* __restrict__ dmu, const float* __restrict__ dsd,
                   int B, int Z, int K) { heads_wgrad_kernel_body<NM>(flat, hs, dmu, dsd, B, Z, K); }

// ---- tiled head kernels (v2) ------------------------------------------------------------------------------------------
// The kernels above walk ONE batch row per block (GEMV): every block re-reads the whole weight slice, with one load per
// FMA.  That is fine for the 2-3 wide latent groups of the MNIST / CelebA nets but not for the LSUN net (groups of 20-30,
// up to 120 columns per feature map, B = 256): measured 286 / 259 / 142 us per launch against ~7 us of HBM time.  The v2
// kernels treat the heads of one feature map as what they are - a [B, K] x [K, ntot] GEMM with a huge K - and tile it:
// all columns of all heads side by side (dense, column c = sum of the widths before head h + i), operands staged through
// shared memory (forward) or held in registers (input / weight gradients), >= 4 FMAs per shared-memory load.
constexpr int HV_BM = 64;    // batch rows per tile
constexpr int HV_KC = 32;    // K chunk staged per iteration (forward)

struct HeadCols {            // dense column map of a HeadSet, built per block in shared memory
  int head[128], idx[128];
};

__device__ __forceinline__ int heads_ntot(const HeadSet& hs) {
  int t = 0;
  for (int h = 0; h < hs.nheads; ++h) t += hs.n[h];
  return t;
}

// forward: grid (row tiles, K slices); thread (tx, ty) owns rows ty*4..+4 and columns tx*TN..+TN of the tile
template <int NT>
__device__ __forceinline__ void heads_fwd_tiled_kernel_body(const float* __restrict__ flat, int B, int K, int kslice, HeadSet hs, float* __restrict__ mu_pre,
                       float* __restrict__ sd_pre, int Z) {
  constexpr int TN = NT / 16;
  __shared__ __align__(16) float s_a[HV_KC][HV_BM + 4];
  __shared__ __align__(16) float s_w[HV_KC][NT];
  __shared__ int s_dst[NT];        // destination offset inside a [Z] row, bit 30 set: sd_pre
  __shared__ float s_bias[NT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * HV_BM;
  const int k0 = blockIdx.y * kslice, k1 = min(K, k0 + kslice);
  const int ntot = heads_ntot(hs);
  for (int c = tid; c < NT; c += 256) {
    int h = 0, base = 0;
    while (h < hs.nheads && c >= base + hs.n[h]) { base += hs.n[h]; ++h; }
    if (h < hs.nheads) {
      s_dst[c] = (hs.col[h] + (c - base)) | (hs.is_sd[h] ? (1 << 30) : 0);
      s_bias[c] = hs.b[h][c - base];
    } else {
      s_dst[c] = -1;
      s_bias[c] = 0.f;
    }
  }
  for (int i = tid; i < HV_KC * NT; i += 256) (&s_w[0][0])[i] = 0.f;   // pad columns stay zero
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  __syncthreads();
  // Software pipeline: the next chunk's global loads are issued into registers before the current chunk is multiplied, so
  // the L2 / HBM latency of a chunk overlaps the FMAs of the previous one instead of adding to them.
  constexpr int WREG = (HV_KC * NT + 255) / 256;     // weight elements per thread and chunk (dense columns: <= this)
  float ra[HV_BM / 8], rw[WREG];
  const int akk = tid & 31, ar0 = tid >> 5;
  auto fetch = [&](int kc) {
    const int kn = min(HV_KC, k1 - kc);
#pragma unroll
    for (int j = 0; j < HV_BM / 8; ++j) {
      const int b = row0 + ar0 + 8 * j;
      ra[j] = (b < B && akk < kn) ? __ldg(flat + (size_t)b * K + kc + akk) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < WREG; ++j) {
      // element e of the dense [HV_KC][ntot] chunk: row kk, dense column c -> head h, local column
      const int e = tid + 256 * j;
      const int kk = e / ntot, c = e - kk * ntot;
      float v = 0.f;
      if (kk < kn) {
        int h = 0, base = 0;
        while (c >= base + hs.n[h]) { base += hs.n[h]; ++h; }
        v = __ldg(hs.w[h] + (size_t)(kc + kk) * hs.n[h] + (c - base));
      }
      rw[j] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < HV_BM / 8; ++j) s_a[akk][ar0 + 8 * j] = ra[j];
#pragma unroll
    for (int j = 0; j < WREG; ++j) {
      const int e = tid + 256 * j;
      const int kk = e / ntot, c = e - kk * ntot;
      if (kk < HV_KC) s_w[kk][c] = rw[j];
    }
  };
  if (k0 < k1) fetch(k0);
  for (int kc = k0; kc < k1; kc += HV_KC) {
    stash();
    __syncthreads();
    if (kc + HV_KC < k1) fetch(kc + HV_KC);
#pragma unroll 8
    for (int kk = 0; kk < HV_KC; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&s_a[kk][ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float w[TN];
      if constexpr (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&s_w[kk][tx * TN + j]);
          w[j] = w4.x; w[j + 1] = w4.y; w[j + 2] = w4.z; w[j + 3] = w4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) w[j] = s_w[kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = row0 + ty * 4 + i;
    if (b >= B) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = tx * TN + j;
      if (c >= ntot) continue;
      const int d = s_dst[c];
      float* dst = ((d >> 30) & 1 ? sd_pre : mu_pre) + (size_t)b * Z + (d & 0x3fffffff);
      atomicAdd(dst, acc[i][j] + (blockIdx.y == 0 ? s_bias[c] : 0.f));
    }
  }
}
template <int NT>
__global__ void __launch_bounds__(256)
heads_fwd_tiled_kernel(const float* __restrict__ flat, int B, int K, int kslice, HeadSet hs, float* __restrict__ mu_pre,
                       float* __restrict__ sd_pre, int Z) { heads_fwd_tiled_kernel_body<NT>(flat, B, K, kslice, hs, mu_pre, sd_pre, Z); }

// input gradient: d_flat[b, k] = sum_c d[b, c] * W[k, c]  (overwrite).  Thread = one k with its NT weights in registers;
// the d rows of the tile sit in shared memory and are read as broadcast float4 (4 FMAs per shared load).
template <int NT>
__device__ __forceinline__ void heads_dgrad_tiled_kernel_body(HeadSet hs, const float* __restrict__ dmu, const float* __restrict__ dsd, int B, int Z, int K,
                         float* __restrict__ d_flat) {
  __shared__ __align__(16) float s_d[HV_BM][NT];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.y * HV_BM;
  const int rows = min(HV_BM, B - row0);
  for (int i = tid; i < HV_BM * NT; i += 256) {
    const int r = i / NT, c = i - r * NT;
    int h = 0, base = 0;
    while (h < hs.nheads && c >= base + hs.n[h]) { base += hs.n[h]; ++h; }
    float v = 0.f;
    if (h < hs.nheads && r < rows) v = (hs.is_sd[h] ? dsd : dmu)[(size_t)(row0 + r) * Z + hs.col[h] + (c - base)];
    s_d[r][c] = v;
  }
  const int k = blockIdx.x * 256 + tid;
  float w[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) w[c] = 0.f;
  if (k < K) {
    int base = 0;
    for (int h = 0; h < hs.nheads; ++h) {
      const int n = hs.n[h];
      const float* __restrict__ wr = hs.w[h] + (size_t)k * n;
#pragma unroll
      for (int c = 0; c < NT; ++c) {
        const int i = c - base;
        if (i >= 0 && i < n) w[c] = __ldg(wr + i);
      }
      base += n;
    }
  }
  __syncthreads();
  if (k >= K) return;
  for (int r = 0; r < rows; ++r) {
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < NT; c += 4) {
      const float4 d4 = *reinterpret_cast<const float4*>(&s_d[r][c]);
      v = fmaf(d4.x, w[c], v); v = fmaf(d4.y, w[c + 1], v); v = fmaf(d4.z, w[c + 2], v); v = fmaf(d4.w, w[c + 3], v);
    }
    d_flat[(size_t)(row0 + r) * K + k] = v;
  }
}
template <int NT>
__global__ void __launch_bounds__(256)
heads_dgrad_tiled_kernel(HeadSet hs, const float* __restrict__ dmu, const float* __restrict__ dsd, int B, int Z, int K,
                         float* __restrict__ d_flat) { heads_dgrad_tiled_kernel_body<NT>(hs, dmu, dsd, B, Z, K, d_flat); }

// weight gradient: gw[k, c] = sum_b flat[b, k] * d[b, c].  Thread = one k with CT accumulators in registers for the CT dense
// columns of its column group (blockIdx.y); the d rows sit in shared memory (broadcast float4).  Every (k, c) has exactly one
// owner, so there are no atomics (the GEMV kernel splits the batch and pays K * ntot atomics per slice - 6-8 M per launch
// for the LSUN heads); machine filling comes from 128-thread blocks and the column groups instead.
// gb[c] = sum_b d[b, c] from the blocks of row blockIdx.x == 0.
template <int CT>
__device__ __forceinline__ void heads_wgrad_tiled_kernel_body(const float* __restrict__ flat, HeadSet hs, const float* __restrict__ dmu,
                         const float* __restrict__ dsd, int B, int Z, int K) {
  __shared__ __align__(16) float s_d[HV_BM][CT];
  __shared__ int s_src[CT];      // source offset inside a [Z] row, bit 30: dsd, -1: pad column
  const int tid = threadIdx.x;
  const int k = blockIdx.x * 128 + tid;
  const int c0 = blockIdx.y * CT;
  const int ntot = heads_ntot(hs);
  for (int c = tid; c < CT; c += 128) {
    const int cc = c0 + c;
    int h = 0, base = 0;
    while (h < hs.nheads && cc >= base + hs.n[h]) { base += hs.n[h]; ++h; }
    s_src[c] = (h < hs.nheads && cc < ntot) ? ((hs.col[h] + (cc - base)) | (hs.is_sd[h] ? (1 << 30) : 0)) : -1;
  }
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = 0.f;
  for (int rb = 0; rb < B; rb += HV_BM) {
    const int rows = min(HV_BM, B - rb);
    __syncthreads();
    for (int i = tid; i < HV_BM * CT; i += 128) {
      const int r = i / CT, c = i - r * CT;
      const int src = s_src[c];
      float v = 0.f;
      if (src >= 0 && r < rows) v = ((src >> 30) & 1 ? dsd : dmu)[(size_t)(rb + r) * Z + (src & 0x3fffffff)];
      s_d[r][c] = v;
    }
    __syncthreads();
    if (k < K) {
      const float* __restrict__ ap = flat + (size_t)rb * K + k;
      for (int r = 0; r < rows; r += 4) {
        float a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = r + u < rows ? __ldg(ap + (size_t)(r + u) * K) : 0.f;   // 4 loads in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r + u < rows) {
#pragma unroll
            for (int c = 0; c < CT; c += 4) {
              const float4 d4 = *reinterpret_cast<const float4*>(&s_d[r + u][c]);
              acc[c] = fmaf(a[u], d4.x, acc[c]); acc[c + 1] = fmaf(a[u], d4.y, acc[c + 1]);
              acc[c + 2] = fmaf(a[u], d4.z, acc[c + 2]); acc[c + 3] = fmaf(a[u], d4.w, acc[c + 3]);
            }
          }
        }
      }
    }
  }
  if (k < K) {
    int base = 0;
    for (int h = 0; h < hs.nheads; ++h) {
      const int n = hs.n[h];
      float* __restrict__ g = hs.gw[h] + (size_t)k * n;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const int i = c0 + c - base;
        if (i >= 0 && i < n) g[i] += acc[c];      // pre-zeroed, single owner
      }
      base += n;
    }
  }
  if (blockIdx.x == 0 && tid < CT) {
    const int src = s_src[tid];
    if (src >= 0) {
      const float* __restrict__ d = ((src >> 30) & 1 ? dsd : dmu) + (src & 0x3fffffff);
      float sum = 0.f;
      for (int b = 0; b < B; ++b) sum += d[(size_t)b * Z];
      const int cc = c0 + tid;
      int h = 0, base = 0;
      while (cc >= base + hs.n[h]) { base += hs.n[h]; ++h; }
      hs.gb[h][cc - base] = sum;
    }
  }
}
template <int CT>
__global__ void __launch_bounds__(128)
heads_wgrad_tiled_kernel(const float* __restrict__ flat, HeadSet hs, const float* __restrict__ dmu,
                         const float* __restrict__ dsd, int B, int Z, int K) { heads_wgrad_tiled_kernel_body<CT>(flat, hs, dmu, dsd, B, Z, K); }

// Latent projections (split_latent, sequential_vae.py:1801-1806): [B, kz<=32] x [kz, N] with N up to 32768.
// dW[k, f] += sum_{b in slice} z[b, k] * dy[b, f]      (grid: f tiles x batch slices, grads pre-zeroed)
template <int NM>
__global__ void __launch_bounds__(256)
lat_wgrad_kernel(View z, const float* __restrict__ dy, int B, int KZ, int N, float* __restrict__ dw) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  const int bper = (B + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * bper, r1 = min(B, r0 + bper);
  float acc[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) acc[k] = 0.f;
  for (int b = r0; b < r1; ++b) {
    const float dv = __ldg(dy + (size_t)b * N + f);
    const float* zr = z.p + (size_t)b * z.ld + z.coff;
#pragma unroll
    for (int k = 0; k < NM; ++k)
      if (k < KZ) acc[k] = fmaf(__ldg(zr + k), dv, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < NM; ++k)
    if (k < KZ) atomicAdd(dw + (size_t)k * N + f, acc[k]);
}

// dz[b, k] (pre-zeroed window) += sum_{f in slice} dy[b, f] * W[k, f]
template <int NM>
__global__ void __launch_bounds__(256)
lat_dz_kernel(const float* __restrict__ dy, const float* __restrict__ w, int KZ, int N, View dz) {
  __shared__ float red[8][SK_NMAX];
  const int b = blockIdx.x;
  const int fper = (N + gridDim.y - 1) / gridDim.y;
  const int f0 = blockIdx.y * fper, f1 = min(N, f0 + fper);
  float acc[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
    const float dv = __ldg(dy + (size_t)b * N + f);
#pragma unroll
    for (int k = 0; k < NM; ++k)
      if (k < KZ) acc[k] = fmaf(dv, __ldg(w + (size_t)k * N + f), acc[k]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NM; ++k) {
    if (k < KZ) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) red[wid][k] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < KZ) {
    float v = 0.f;
    for (int wi = 0; wi < 8; ++wi) v += red[wi][threadIdx.x];
    atomicAdd(dz.p + (size_t)b * dz.ld + dz.coff + threadIdx.x, v);
  }
}

}  // namespace

int simt_gather_gemm(const LaunchCtx& lc, const Geom& g, View in, const float* w, View out, double* stats) {
  const int64_t M = (int64_t)g.B * g.Hout * g.Wout;
  if (M <= 0) return 0;
  const double pix = (double)g.B * (g.Hin * g.Win < g.Hout * g.Wout ? g.Hin * g.Win : g.Hout * g.Wout);
  ProfScope ps(lc, KC_GEMM_SIMT, 2.0 * pix * g.KH * g.KW * g.Cin * g.Cout,
               4.0 * ((double)g.B * g.Hin * g.Win * g.Cin + (double)M * g.Cout + (double)g.KH * g.KW * g.Cin * g.Cout), &g);
  if (g.Cout > 32) {
    dim3 grid((unsigned)((M + 63) / 64), (g.Cout + 63) / 64);
    gather_gemm_kernel<64, 64, 4, 4><<<grid, 256, 0, lc.stream>>>(g, in, w, out, stats);
  } else if (g.Cout > 8) {
    dim3 grid((unsigned)((M + 127) / 128), (g.Cout + 31) / 32);
    gather_gemm_kernel<128, 32, 4, 4><<<grid, 256, 0, lc.stream>>>(g, in, w, out, stats);
  } else {
    dim3 grid((unsigned)((M + 255) / 256), 1);
    gather_gemm_kernel<256, 8, 4, 2><<<grid, 256, 0, lc.stream>>>(g, in, w, out, stats);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int simt_wgrad(const LaunchCtx& lc, const Geom& g, View x, View dy, float* dw) {
  const int64_t M = (int64_t)g.B * g.Hout * g.Wout;
  const int taps = g.KH * g.KW;
  const int tilesA = (g.Cin + 63) / 64, tilesB = (g.Cout + 63) / 64;
  const int64_t base = (int64_t)taps * tilesA * tilesB;
  int64_t want = (4LL * lc.sm_count + base - 1) / base;
  int64_t max_split = (M + 4 * BK - 1) / (4 * BK);
  int64_t ksplit = want < 1 ? 1 : (want > max_split ? max_split : want);
  if (ksplit < 1) ksplit = 1;
  int64_t rows_per_split = ((M + ksplit - 1) / ksplit + BK - 1) / BK * BK;
  ksplit = (M + rows_per_split - 1) / rows_per_split;
  dim3 grid((unsigned)base, (unsigned)ksplit);
  const double pix = (double)g.B * (g.Hin * g.Win < g.Hout * g.Wout ? g.Hin * g.Win : g.Hout * g.Wout);
  ProfScope ps(lc, KC_WGRAD_SIMT, 2.0 * pix * taps * g.Cin * g.Cout,
               4.0 * ((double)g.B * g.Hin * g.Win * g.Cin + (double)M * g.Cout + (double)taps * g.Cin * g.Cout), &g);
  wgrad_kernel<64, 64, 4, 4><<<grid, 256, 0, lc.stream>>>(g, x, dy, dw, tilesA, tilesB, rows_per_split);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// width class of a launch: the smallest of 4 / 8 / 16 / 32 that holds every head (or the latent group)
static int width_class(int nmax) { return nmax <= 4 ? 4 : nmax <= 8 ? 8 : nmax <= 16 ? 16 : 32; }
static int heads_nmax(const HeadSet& hs) {
  int m = 0;
  for (int h = 0; h < hs.nheads; ++h) m = hs.n[h] > m ? hs.n[h] : m;
  return m;
}
#define SK_DISPATCH(NMAXV, CALL)                         \
  switch (width_class(NMAXV)) {                          \
    case 4: { constexpr int NM = 4; CALL; } break;       \
    case 8: { constexpr int NM = 8; CALL; } break;       \
    case 16: { constexpr int NM = 16; CALL; } break;     \
    default: { constexpr int NM = 32; CALL; } break;     \
  }

// Tiled head kernels (v2): the default for every feature map whose heads are <= 128 columns wide in total; SVAE_HEADS_TILED=0
// selects the per-row GEMV kernels (A/B: CelebA B=100 12.04 -> 11.88 ms/step, LSUN B=256 T=16 110.5 -> 54.7 ms/step with the
// tiled ones).  NT = padded column count of the launch.
static bool heads_use_tiled(int ntot) {
  static const int mode = [] { const char* e = getenv("SVAE_HEADS_TILED"); return e ? atoi(e) : -1; }();
  if (ntot > 128) return false;
  if (mode >= 0) return mode != 0;
  return true;
}
#define HV_DISPATCH(NTOTV, CALL)                                   \
  do {                                                             \
    const int _n = (NTOTV);                                        \
    if (_n <= 16) { constexpr int NT = 16; CALL; }                 \
    else if (_n <= 32) { constexpr int NT = 32; CALL; }            \
    else if (_n <= 64) { constexpr int NT = 64; CALL; }            \
    else { constexpr int NT = 128; CALL; }                         \
  } while (0)

static int split_for(int rows_or_blocks, int sm_count, int max_split) {
  int s = (2 * sm_count + rows_or_blocks - 1) / rows_or_blocks;
  if (s < 1) s = 1;
  if (s > max_split) s = max_split;
  return s;
}

int heads_fwd(const LaunchCtx& lc, const float* flat, int B, int K, const HeadSet& hs, float* mu_pre, float* sd_pre, int Z) {
  int ntot = 0;
  for (int h = 0; h < hs.nheads; ++h) ntot += hs.n[h];
  // K slices per image: enough blocks (~8 per SM) that the L2 latency of the two short load streams is hidden; every slice
  // still spans >= 1024 inputs so that the per-slice reduction + atomics stay small
  int ks = (8 * lc.sm_count + B - 1) / B;
  const int ks_max = (K + 1023) / 1024;
  if (ks > ks_max) ks = ks_max;
  if (ks < 1) ks = 1;
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * K * ntot, 4.0 * ((double)B * K + (double)K * ntot));
  if (heads_use_tiled(ntot)) {
    // row tiles x K slices: ~2 blocks per SM, every slice a multiple of the staged chunk and >= 256 inputs
    const int rt = (B + HV_BM - 1) / HV_BM;
    int kslices = (2 * lc.sm_count + rt - 1) / rt;
    const int ks_cap = (K + 255) / 256;
    if (kslices > ks_cap) kslices = ks_cap;
    if (kslices < 1) kslices = 1;
    const int kslice = ((K + kslices - 1) / kslices + HV_KC - 1) / HV_KC * HV_KC;
    kslices = (K + kslice - 1) / kslice;
    if (lc.multi != nullptr) {
      int r = 0;
      HV_DISPATCH(ntot, (r = MULTI_RECORD(heads_fwd_tiled_kernel_body<NT>, 256, lc, dim3(rt, kslices), dim3(256), 0, flat, B, K, kslice, hs,
                                          mu_pre, sd_pre, Z)));
      return r;
    }
    HV_DISPATCH(ntot, (heads_fwd_tiled_kernel<NT><<<dim3(rt, kslices), 256, 0, lc.stream>>>(flat, B, K, kslice, hs, mu_pre,
                                                                                            sd_pre, Z)));
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  if (lc.multi != nullptr) {
    int r = 0;
    SK_DISPATCH(heads_nmax(hs), (r = MULTI_RECORD(heads_fwd_kernel_body<NM>, 256, lc, dim3(B, ks), dim3(256), 0, flat, K, hs, mu_pre, sd_pre, Z)));
    return r;
  }
  SK_DISPATCH(heads_nmax(hs), (heads_fwd_kernel<NM><<<dim3(B, ks), 256, 0, lc.stream>>>(flat, K, hs, mu_pre, sd_pre, Z)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int heads_dgrad(const LaunchCtx& lc, const HeadSet& hs, const float* dmu, const float* dsd, int B, int Z, int K,
                float* d_flat) {
  int ntot = 0;
  for (int h = 0; h < hs.nheads; ++h) ntot += hs.n[h];
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * K * ntot, 4.0 * ((double)B * K + (double)K * ntot));
  if (heads_use_tiled(ntot)) {
    if (lc.multi != nullptr) {
      int r = 0;
      HV_DISPATCH(ntot, (r = MULTI_RECORD(heads_dgrad_tiled_kernel_body<NT>, 256, lc, dim3((K + 255) / 256, (B + HV_BM - 1) / HV_BM),
                                          dim3(256), 0, hs, dmu, dsd, B, Z, K, d_flat)));
      return r;
    }
    HV_DISPATCH(ntot, (heads_dgrad_tiled_kernel<NT><<<dim3((K + 255) / 256, (B + HV_BM - 1) / HV_BM), 256, 0, lc.stream>>>(
                          hs, dmu, dsd, B, Z, K, d_flat)));
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  if (lc.multi != nullptr) return MULTI_RECORD(heads_dgrad_kernel_body, 256, lc, dim3((K + 255) / 256, B), dim3(256), 0, hs, dmu, dsd, Z, K, d_flat);
  heads_dgrad_kernel<<<dim3((K + 255) / 256, B), 256, 0, lc.stream>>>(hs, dmu, dsd, Z, K, d_flat);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int heads_wgrad(const LaunchCtx& lc, const float* flat, const HeadSet& hs, const float* dmu, const float* dsd, int B, int Z,
                int K) {
  int ntot = 0;
  for (int h = 0; h < hs.nheads; ++h) ntot += hs.n[h];
  const int kb = (K + 255) / 256;
  const int bs = split_for(kb, lc.sm_count, (B + 7) / 8);
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * K * ntot, 4.0 * ((double)B * K + (double)K * ntot));
  if (heads_use_tiled(ntot)) {
    // 128 k per block x column groups of 16 or 32: narrower groups when the k blocks alone do not fill the machine
    const int kb128 = (K + 127) / 128;
    if (kb128 * ((ntot + 31) / 32) >= 2 * lc.sm_count) {
      if (lc.multi != nullptr) return MULTI_RECORD(heads_wgrad_tiled_kernel_body<32>, 128, lc, dim3(kb128, (ntot + 31) / 32), dim3(128), 0, flat, hs, dmu, dsd, B, Z, K);
      heads_wgrad_tiled_kernel<32><<<dim3(kb128, (ntot + 31) / 32), 128, 0, lc.stream>>>(flat, hs, dmu, dsd, B, Z, K);
    } else {
      if (lc.multi != nullptr) return MULTI_RECORD(heads_wgrad_tiled_kernel_body<16>, 128, lc, dim3(kb128, (ntot + 15) / 16), dim3(128), 0, flat, hs, dmu, dsd, B, Z, K);
      heads_wgrad_tiled_kernel<16><<<dim3(kb128, (ntot + 15) / 16), 128, 0, lc.stream>>>(flat, hs, dmu, dsd, B, Z, K);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  if (lc.multi != nullptr) {
    int r = 0;
    SK_DISPATCH(heads_nmax(hs), (r = MULTI_RECORD(heads_wgrad_kernel_body<NM>, 256, lc, dim3(kb, bs), dim3(256), 0, flat, hs, dmu, dsd, B, Z, K)));
    return r;
  }
  SK_DISPATCH(heads_nmax(hs), (heads_wgrad_kernel<NM><<<dim3(kb, bs), 256, 0, lc.stream>>>(flat, hs, dmu, dsd, B, Z, K)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lat_wgrad(const LaunchCtx& lc, View z, const float* dy, int B, int KZ, int N, float* dw) {
  const int fb = (N + 255) / 256;
  const int bs = split_for(fb, lc.sm_count, (B + 7) / 8);
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * KZ * N, 4.0 * ((double)B * N + (double)KZ * N));
  SK_DISPATCH(KZ, (lat_wgrad_kernel<NM><<<dim3(fb, bs), 256, 0, lc.stream>>>(z, dy, B, KZ, N, dw)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int lat_dz(const LaunchCtx& lc, const float* dy, const float* w, int B, int KZ, int N, View dz) {
  const int fs = split_for(B, lc.sm_count, (N + 2047) / 2048);
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * KZ * N, 4.0 * ((double)B * N + (double)KZ * N));
  SK_DISPATCH(KZ, (lat_dz_kernel<NM><<<dim3(B, fs), 256, 0, lc.stream>>>(dy, w, KZ, N, dz)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
