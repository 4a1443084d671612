// kernels_fc.cu - tcgen05 kernel for the fully-connected layers of the chain (fc_bn_lrelu: enc.fc 2048->384 and
// dec.fc {512,896}->6144, abstract_network.py:64-71; call sites sequential_vae.py:1704,1775) and their gradients.
//
// At batch 100 these are weight-streaming problems (dec.fc: 5.5 M weights for 100 rows): the fp32 master weights are
// read ONCE, straight from the parameter arena, converted to bf16 on the way into shared memory - no packed operand
// copy, no repack after Adam.  One kernel covers the three contractions; they differ only in which matrix dimension is
// contiguous in memory, i.e. in the "major-ness" of the UMMA operands:
//   forward  Y[m,n]  = sum_k X[m,k]  W[k,n]     A = X  (K-major)   B = W  (MN-major)   reduction = k
//   dgrad    dX[m,k] = sum_n dY[m,n] W[k,n]     A = dY (K-major)   B = W  (K-major)    reduction = n
//   wgrad    dW[k,n] = sum_m X[m,k]  dY[m,n]    A = X  (MN-major)  B = dY (MN-major)   reduction = m
// The reduction is cut into 64-element chunks that four producer warps stage (fp32 -> bf16, zero fill) into the canonical
// no-swizzle layouts  [8-element group of the memory-contiguous dim][index of the other dim][16 B]  through a 3-deep
// mbarrier ring; one elected thread issues tcgen05.mma (M = 128, N = 128, K = 16) into a TMEM accumulator; the same four
// warps drain it with tcgen05.ld.  Forward / dgrad split the reduction across blockIdx.z (red.global.add into a zeroed
// output) so that >= 148 CTAs stream the weights; wgrad tiles the [K, N] output and needs no split.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int FC_RC = 64;       // reduction elements per stage
constexpr int FC_STAGES = 3;
constexpr unsigned FC_PITCH_K = 129 * 16;   // K-major operand: 8 planes (reduction groups) x 128 rows x 16 B, odd pitch
constexpr unsigned FC_PITCH_MN = 65 * 16;   // MN-major operand: 16 planes (MN groups) x 64 reduction rows x 16 B
constexpr unsigned FC_OP_BYTES = 16 * FC_PITCH_MN > 8 * FC_PITCH_K ? 16 * FC_PITCH_MN : 8 * FC_PITCH_K;

struct FcOperand {
  const float* p;
  int ld;
  int mn_major;    // 1: memory rows are reduction indices, columns are this operand's M/N index
  int mn_extent;   // valid extent of the M/N index
  int vec;         // rows 16-byte aligned (ld % 4 == 0, base aligned)
};

struct FcParams {
  FcOperand a, b;
  float* out; int out_ld;
  int R;            // reduction length
  int chunks_per_split;
  int atomic;       // 1: red.add into a zeroed / accumulating output, 0: plain store
};

struct FcSmem {
  unsigned long long ready[FC_STAGES], free_[FC_STAGES], acc_done;
  unsigned tmem_base, pad;
};

// Stage one 64-element reduction chunk [r0, r0+64) of an operand tile (128 M/N indices starting at mn0).
__device__ __forceinline__ void fc_stage(unsigned char* dst, const FcOperand& op, int mn0, int r0, int R, int tid) {
  // K-major: outer = M/N index (128), inner = reduction (64 -> 8 groups);  MN-major: outer = reduction (64), inner = M/N (128 -> 16 groups)
  const int groups = op.mn_major ? 16 : 8;
  const int outer = op.mn_major ? FC_RC : 128;
  const unsigned pitch = op.mn_major ? FC_PITCH_MN : FC_PITCH_K;
  const int o_base = op.mn_major ? r0 : mn0, i_base = op.mn_major ? mn0 : r0;
  const int o_lim = op.mn_major ? R : op.mn_extent, i_lim = op.mn_major ? op.mn_extent : R;
#ifndef SVAE_FC_U
#define SVAE_FC_U 8
#endif
  constexpr int U = SVAE_FC_U;   // loads in flight per thread and pass (the stage is 1024 16-byte items: one pass at U = 8)
  for (int base = tid; base < outer * groups; base += 128 * U) {
    float4 v0[U], v1[U];
    unsigned off[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = base + u * 128;            // outer*groups = 1024: never out of range
      const int o = it / groups, g = it - o * groups;
      off[u] = (unsigned)g * pitch + (unsigned)o * 16u;
      const int row = o_base + o, col = i_base + g * 8;
      v0[u] = make_float4(0.f, 0.f, 0.f, 0.f); v1[u] = v0[u];
      if (row < o_lim && col < i_lim) {
        const float* src = op.p + (size_t)row * op.ld + col;
        if (op.vec && col + 8 <= i_lim) {
          v0[u] = __ldg(reinterpret_cast<const float4*>(src));
          v1[u] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (col + e < i_lim) ? __ldg(src + e) : 0.f;
          v0[u] = make_float4(f[0], f[1], f[2], f[3]); v1[u] = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float f[8] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w, v1[u].x, v1[u].y, v1[u].z, v1[u].w};
      *reinterpret_cast<uint4*>(dst + off[u]) = pack8_bf16(f);
    }
  }
}

// grid: x = B-side (output column) tile, y = A-side (output row) tile, z = reduction split.  160 threads:
// warps 0-3 producers then epilogue, warp 4 MMA issuer + TMEM owner.
__global__ void __launch_bounds__(160) tc_fc_kernel(const __grid_constant__ FcParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FcSmem* hdr = reinterpret_cast<FcSmem*>(smem_raw);
  unsigned char* bufs = smem_raw + 128;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uniform_u32((unsigned)(tid >> 5));
  const int col0 = blockIdx.x * 128, row0 = blockIdx.y * 128;
  const int nchunks_total = (P.R + FC_RC - 1) / FC_RC;
  const int c_begin = blockIdx.z * P.chunks_per_split;
  const int c_end = min(nchunks_total, c_begin + P.chunks_per_split);
  const int my_chunks = max(0, c_end - c_begin);

  if (tid == 0) {
    for (int i = 0; i < FC_STAGES; ++i) { mbar_init(smem_u32(&hdr->ready[i]), 128); mbar_init(smem_u32(&hdr->free_[i]), 1); }
    mbar_init(smem_u32(&hdr->acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&hdr->tmem_base), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = uniform_u32(hdr->tmem_base);

  if (warp < 4) {
    for (int it = 0; it < my_chunks; ++it) {
      const int buf = it % FC_STAGES;
      if (it >= FC_STAGES) mbar_wait(smem_u32(&hdr->free_[buf]), (unsigned)((it / FC_STAGES) - 1) & 1u);
      unsigned char* ab = bufs + (size_t)buf * 2 * FC_OP_BYTES;
      const int r0 = (c_begin + it) * FC_RC;
      fc_stage(ab, P.a, row0, r0, P.R, tid);
      fc_stage(ab + FC_OP_BYTES, P.b, col0, r0, P.R, tid);
      fence_proxy_async();
      mbar_arrive(smem_u32(&hdr->ready[buf]));
    }
    // ---- epilogue: accumulator row = output row (TMEM lane), 128 columns ----
    mbar_wait(smem_u32(&hdr->acc_done), 0);
    tc_fence_after();
    const int row = row0 + warp * 32 + lane;
    const bool row_ok = row < P.a.mn_extent && my_chunks > 0;
    const bool vec_ok = (P.out_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
    for (int n0 = 0; n0 < 128; n0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((unsigned)(warp * 32) << 16) + (unsigned)n0, v);
      if (!row_ok) continue;
      const int ncols = max(0, min(32, P.b.mn_extent - (col0 + n0)));
      float* dst = P.out + (size_t)row * P.out_ld + col0 + n0;
      if (vec_ok && ncols == 32) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 o = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
          if (P.atomic) atomicAdd(reinterpret_cast<float4*>(dst + k), o);
          else *reinterpret_cast<float4*>(dst + k) = o;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < ncols) { if (P.atomic) atomicAdd(dst + k, v[k]); else dst[k] = v[k]; }
      }
    }
    tc_fence_before();
  } else {
    // instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = 128, major bits 15 (A) / 16 (B)
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(P.a.mn_major != 0) << 15) |
                           ((unsigned)(P.b.mn_major != 0) << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const bool leader = elect_one();
    for (int it = 0; it < my_chunks; ++it) {
      const int buf = it % FC_STAGES;
      mbar_wait(smem_u32(&hdr->ready[buf]), (unsigned)(it / FC_STAGES) & 1u);
      tc_fence_after();
      if (leader) {
        const unsigned ab = smem_u32(bufs + (size_t)buf * 2 * FC_OP_BYTES);
        const unsigned bb = ab + FC_OP_BYTES;
#pragma unroll
        for (int s = 0; s < FC_RC / 16; ++s) {
          // K-major: the two 8-element reduction groups of one MMA are two planes (LBO = plane pitch), 8-row groups 128 B apart
          // MN-major: 16 reduction rows = 256 B inside a plane (LBO = 128 B between the 8-row halves), planes = 8 M/N indices (SBO)
          const unsigned long long adesc = P.a.mn_major ? make_desc(ab + (unsigned)s * 256u, 128u, FC_PITCH_MN)
                                                        : make_desc(ab + (unsigned)(2 * s) * FC_PITCH_K, FC_PITCH_K, 128u);
          const unsigned long long bdesc = P.b.mn_major ? make_desc(bb + (unsigned)s * 256u, 128u, FC_PITCH_MN)
                                                        : make_desc(bb + (unsigned)(2 * s) * FC_PITCH_K, FC_PITCH_K, 128u);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&hdr->free_[buf]));
      }
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&hdr->acc_done));
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

__global__ void zero_rows_kernel(float* __restrict__ p, int ld, int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[(i / cols) * ld + (i % cols)] = 0.f;
}

}  // namespace

bool tc_fc_supported(int K, int N) { return K >= 64 && N >= 64; }

// mode 0: out[M,N] = x[M,K] . w[K,N] ; mode 1: out[M,K] = x[M,N] . w[K,N]^T ; mode 2: out[K,N] = x[M,K]^T . w[M,N]
// (x = first operand with leading dimension ldx, w = second operand with ldw).  accumulate: add into `out`.
int tc_fc(const LaunchCtx& lc, int mode, const float* x, int ldx, const float* w, int ldw, float* out, int ldo, int M, int K,
          int N, int accumulate) {
  FcParams P;
  memset(&P, 0, sizeof P);
  auto vec = [](const float* p, int ld) { return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? 1 : 0; };
  int rows, cols;
  if (mode == 0) {
    P.a = FcOperand{x, ldx, 0, M, vec(x, ldx)}; P.b = FcOperand{w, ldw, 1, N, vec(w, ldw)}; P.R = K; rows = M; cols = N;
  } else if (mode == 1) {
    P.a = FcOperand{x, ldx, 0, M, vec(x, ldx)}; P.b = FcOperand{w, ldw, 0, K, vec(w, ldw)}; P.R = N; rows = M; cols = K;
  } else {
    P.a = FcOperand{x, ldx, 1, K, vec(x, ldx)}; P.b = FcOperand{w, ldw, 1, N, vec(w, ldw)}; P.R = M; rows = K; cols = N;
  }
  P.out = out; P.out_ld = ldo;
  const int ct = (cols + 127) / 128, rt = (rows + 127) / 128;
  const int nchunks = (P.R + FC_RC - 1) / FC_RC;
  int split = ct * rt <= lc.sm_count ? lc.sm_count / (ct * rt) : 1;     // at most one CTA per SM: a second, nearly empty wave doubles the launch
  if (split > nchunks) split = nchunks;
  if (split < 1) split = 1;
  P.chunks_per_split = (nchunks + split - 1) / split;
  split = (nchunks + P.chunks_per_split - 1) / P.chunks_per_split;
  P.atomic = (split > 1 || accumulate) ? 1 : 0;
  if (split > 1 && !accumulate) {
    if (ldo == cols) {
      CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows * cols, lc.stream));
      if (lc.pdl_state) *lc.pdl_state = 0;   // a memset node now ends the stream: the next kernel takes a full dependency
    } else {
      zero_rows_kernel<<<lc.sm_count, 256, 0, lc.stream>>>(out, ldo, rows, cols);
      CUDA_TRY(cudaGetLastError());
    }
  }
  static bool configured = false;
  const size_t smem = 128 + (size_t)FC_STAGES * 2 * FC_OP_BYTES;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  Geom g{}; g.B = M; g.Hin = g.Win = g.Hout = g.Wout = 1; g.Cin = K; g.Cout = N; g.KH = g.KW = 1; g.stride = 1; g.mode = mode;
  ProfScope ps(lc, mode == 2 ? KC_WGRAD_TC : KC_GEMM_TC, 2.0 * M * K * N, 4.0 * ((double)K * N + (double)M * K + (double)M * N), &g);
  tc_fc_kernel<<<dim3((unsigned)ct, (unsigned)rt, (unsigned)split), 160, smem, lc.stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
