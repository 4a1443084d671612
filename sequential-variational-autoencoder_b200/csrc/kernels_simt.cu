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

// ---- skinny fully-connected kernels ---------------------------------------------------------------------------
constexpr int SK_NMAX = 32;

// out[b, n] = bias[n] + sum_k a[b,k] * W(k,n) ; one block per (row b, group of SK_NMAX outputs)
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(View a, int K, const float* __restrict__ w, int w_n_major, const float* __restrict__ bias, View out,
                  int N) {
  const int b = blockIdx.x;
  const int n0 = blockIdx.y * SK_NMAX;
  const int nn = min(SK_NMAX, N - n0);
  float acc[SK_NMAX];
#pragma unroll
  for (int n = 0; n < SK_NMAX; ++n) acc[n] = 0.f;
  const float* arow = a.p + (size_t)b * a.ld + a.coff;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float av = __ldg(arow + k);
    if (w_n_major) {
#pragma unroll
      for (int n = 0; n < SK_NMAX; ++n)
        if (n < nn) acc[n] = fmaf(av, __ldg(w + (size_t)(n0 + n) * K + k), acc[n]);
    } else {
      const float* wr = w + (size_t)k * N + n0;
#pragma unroll
      for (int n = 0; n < SK_NMAX; ++n)
        if (n < nn) acc[n] = fmaf(av, __ldg(wr + n), acc[n]);
    }
  }
  __shared__ float red[8][SK_NMAX];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < SK_NMAX; ++n) {
    float v = acc[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][n] = v;
  }
  __syncthreads();
  if (threadIdx.x < nn) {
    float v = bias ? bias[n0 + threadIdx.x] : 0.f;
    for (int wi = 0; wi < 8; ++wi) v += red[wi][threadIdx.x];
    out.p[(size_t)b * out.ld + out.coff + n0 + threadIdx.x] = v;
  }
}

// din[b,k] (+)= sum_n dout[b,n] * w[k*N + n]      (w is [K,N])
__global__ void __launch_bounds__(256)
skinny_dgrad_kernel(View dout, int B, int N, const float* __restrict__ w, int K, View din, int accumulate) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ float s_d[];
  for (int n = threadIdx.x; n < N; n += blockDim.x) s_d[n] = dout.p[(size_t)b * dout.ld + dout.coff + n];
  __syncthreads();
  if (k >= K) return;
  const float* wr = w + (size_t)k * N;
  float v = 0.f;
  for (int n = 0; n < N; ++n) v = fmaf(s_d[n], __ldg(wr + n), v);
  float* dst = din.p + (size_t)b * din.ld + din.coff + k;
  if (accumulate) v += *dst;
  *dst = v;
}

// w_n_major == 0: dw[k*N + n] = sum_b a[b,k] * dout[b,n]   (thread per k; heads: K large, N small)
// w_n_major == 1: dw[n*K + k] = same value, K small, N large (latent projections [lat, feats]): thread per n
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(View a, View dout, int B, int K, int N, float* __restrict__ dw, float* __restrict__ dbias,
                    int thread_over_n) {
  if (!thread_over_n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) {
      for (int n0 = 0; n0 < N; n0 += SK_NMAX) {
        float acc[SK_NMAX];
#pragma unroll
        for (int n = 0; n < SK_NMAX; ++n) acc[n] = 0.f;
        const int nn = min(SK_NMAX, N - n0);
        for (int b = 0; b < B; ++b) {
          float av = __ldg(a.p + (size_t)b * a.ld + a.coff + k);
          const float* dr = dout.p + (size_t)b * dout.ld + dout.coff + n0;
#pragma unroll
          for (int n = 0; n < SK_NMAX; ++n)
            if (n < nn) acc[n] = fmaf(av, __ldg(dr + n), acc[n]);
        }
#pragma unroll
        for (int n = 0; n < SK_NMAX; ++n)
          if (n < nn) dw[(size_t)k * N + n0 + n] = acc[n];
      }
    }
    if (dbias != nullptr && blockIdx.x == 0) {
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dout.p[(size_t)b * dout.ld + dout.coff + n];
        dbias[n] = s;
      }
    }
  } else {
    // thread per output feature n; K (<= SK_NMAX) latent inputs; dw laid out [K, N] (reference [in,out])
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc[SK_NMAX];
#pragma unroll
    for (int k = 0; k < SK_NMAX; ++k) acc[k] = 0.f;
    for (int b = 0; b < B; ++b) {
      float dv = __ldg(dout.p + (size_t)b * dout.ld + dout.coff + n);
      const float* ar = a.p + (size_t)b * a.ld + a.coff;
#pragma unroll
      for (int k = 0; k < SK_NMAX; ++k)
        if (k < K) acc[k] = fmaf(__ldg(ar + k), dv, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < SK_NMAX; ++k)
      if (k < K) dw[(size_t)k * N + n] = acc[k];
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

int skinny_fwd(const LaunchCtx& lc, View a, int B, int K, const float* w, int w_n_major, const float* bias, View out,
               int N) {
  dim3 grid(B, (N + SK_NMAX - 1) / SK_NMAX);
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * K * N, 4.0 * ((double)B * K + (double)K * N + (double)B * N));
  skinny_fwd_kernel<<<grid, 256, 0, lc.stream>>>(a, K, w, w_n_major, bias, out, N);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int skinny_dgrad(const LaunchCtx& lc, View dout, int B, int N, const float* w, int K, View din, int accumulate) {
  dim3 grid((K + 255) / 256, B);
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * K * N, 4.0 * ((double)B * K + (double)K * N + (double)B * N));
  skinny_dgrad_kernel<<<grid, 256, N * sizeof(float), lc.stream>>>(dout, B, N, w, K, din, accumulate);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int skinny_wgrad(const LaunchCtx& lc, View a, View dout, int B, int K, int N, float* dw, float* dbias,
                 int thread_over_n) {
  int span = thread_over_n ? N : K;
  ProfScope ps(lc, KC_SKINNY, 2.0 * B * K * N, 4.0 * ((double)B * K + (double)K * N + (double)B * N));
  skinny_wgrad_kernel<<<(span + 255) / 256, 256, 0, lc.stream>>>(a, dout, B, K, N, dw, dbias, thread_over_n);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
