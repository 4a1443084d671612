// common.cuh - shared declarations for libsvae (B200 / sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include <string>
#include <vector>

#define SVAE_BN_EPS 1e-3f     // tf.contrib.layers.batch_norm default epsilon (abstract_network.py:22)
#define SVAE_LRELU_SLOPE 0.1f // lrelu rate (abstract_network.py:8)

enum Act { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2 };

// A dense NHWC tensor seen through a channel window: element (row r, channel c) lives at p[r*ld + coff + c].
// This is how channel concats (sequential_vae.py:1716,1834) are expressed without copies: producers write, and
// consumers read, their channel window of the shared buffer.
struct View {
  float* p;
  int ld;
  int coff;
};
static inline View mkview(float* p, int ld, int coff = 0) { return View{p, ld, coff}; }

// Addressing of one feature of a batch-norm'd tensor.  4-D BN (per channel over N,H,W): rows = pixels, feats = C,
// inner = C, ppr = 1.  2-D BN over a projected latent plane (fc_bn_lrelu reshaped to [B,S,S,F], sequential_vae.py:
// 1803-1804): rows = B, feats = S*S*F, inner = F, ppr = S*S, so that feature f = (pixel f/F, channel f%F) is found
// inside a (possibly concatenated) NHWC buffer.
struct FeatView {
  float* p;
  int ld;    // channels of the underlying buffer
  int coff;  // first channel of the window
  int inner; // channels in the window
  int ppr;   // pixels per row
};
__host__ __device__ static inline size_t fv_addr(const FeatView& v, int64_t row, int f) {
  int pix = f / v.inner;
  int c = f - pix * v.inner;
  return ((size_t)row * v.ppr + pix) * v.ld + v.coff + c;
}

// Geometry of one contraction (conv / transposed conv / fully connected) in gather form:
//   out[b,oh,ow,co] = sum_{kh,kw,ci} in[b,ih,iw,ci] * W(kh,kw,ci,co)
//   mode 0 (conv gather):    ih = oh*stride - pad + kh
//   mode 1 (deconv gather):  ih = (oh + pad - kh)/stride when divisible (adjoint of mode 0)
//   w_out_major 0: W(tap,ci,co) = w[(tap*Cin + ci)*Cout + co]   1: w[(tap*Cout + co)*Cin + ci]
struct Geom {
  int B, Hin, Win, Cin, Hout, Wout, Cout;
  int KH, KW, stride, pad, mode, w_out_major, accumulate;
};

// ---- per-kernel-class event profiler (bench.py's live roofline numbers) -------------------------------------------
enum KClass {
  KC_GEMM_SIMT = 0, KC_WGRAD_SIMT, KC_SKINNY, KC_BN_FWD, KC_BN_BWD_REDUCE, KC_BN_BWD_APPLY, KC_OUT_MIX, KC_REPARAM,
  KC_ADAM, KC_MISC, KC_GEMM_TC, KC_WGRAD_TC, KC_PACK, KC_COUNT
};
static const char* const kKClassNames[KC_COUNT] = {
    "gather_gemm_simt", "wgrad_simt", "skinny_fc", "bn_act_fwd", "bn_bwd_reduce", "bn_bwd_apply", "out_mix",
    "reparam_kl", "adam_clip", "misc", "gather_gemm_tcgen05", "wgrad_tcgen05", "pack_weights"};

struct Profiler {
  bool enabled = false;
  struct Rec { cudaEvent_t a, b; int kc; int geo[8]; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t pool_used = 0;
  int64_t launches[KC_COUNT] = {0};
  double flops[KC_COUNT] = {0}, bytes[KC_COUNT] = {0}, ms[KC_COUNT] = {0};
  cudaEvent_t get() {
    if (pool_used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[pool_used++];
  }
};

struct MultiRec;
struct LaunchCtx {
  cudaStream_t stream;
  int64_t* launches;
  int sm_count;
  Profiler* prof;
  // Batched ("multi") launches: while set, the kernel wrappers that support it RECORD their launch (kernel, grid, argument
  // block) instead of launching; multi_flush (model.cu) then issues ONE launch per recorded slot over all recorded items, the item
  // index being blockIdx.z.  See MultiRec below.
  MultiRec* multi = nullptr;
  // Programmatic dependent launch (chain stream only, see launch_k): *pdl_state == 1 when the node enqueued last on this
  // stream is a kernel; nullptr: never use it.  ProfScope sets it after every launch, non-kernel operations clear it.
  int* pdl_state = nullptr;
};

// RAII: brackets one kernel launch with events when profiling is on, and counts the launch either way.
struct ProfScope {
  const LaunchCtx& lc;
  cudaEvent_t b = nullptr;
  int kc;
  ProfScope(const LaunchCtx& lc_, int kc_, double flops, double bytes, const Geom* g = nullptr) : lc(lc_), kc(kc_) {
    if (lc.multi != nullptr) { multi_note(lc.multi, kc_, flops, bytes); return; }   // recorded, not launched: counted at the flush
    ++*lc.launches;
    Profiler* p = lc.prof;
    if (p && p->enabled) {
      cudaEvent_t a = p->get();
      b = p->get();
      cudaEventRecord(a, lc.stream);
      Profiler::Rec r{a, b, kc, {0, 0, 0, 0, 0, 0, 0, 0}};
      if (g) { r.geo[0] = g->B; r.geo[1] = g->Hin; r.geo[2] = g->Cin; r.geo[3] = g->Hout; r.geo[4] = g->Cout;
               r.geo[5] = g->KH; r.geo[6] = g->stride; r.geo[7] = g->mode; }
      p->recs.push_back(r);
      p->launches[kc]++; p->flops[kc] += flops; p->bytes[kc] += bytes;
    }
  }
  ~ProfScope() {
    if (lc.multi != nullptr) return;
    if (b) cudaEventRecord(b, lc.stream);
    if (lc.pdl_state) *lc.pdl_state = 1;
  }
  static void multi_note(MultiRec* m, int kc, double flops, double bytes);
};

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// The chain (encoder -> decoder -> ... of step t, then t+1; the reverse in the backward) is ~600 short dependent kernels.
// A kernel launched through launch_k on the chain stream right after another kernel carries the programmatic-stream-
// serialization attribute: its CTAs are scheduled as soon as every CTA of the predecessor has passed pdl_trigger() (or
// exited), run their prologue (barrier init, TMEM allocation, weight loads, index math) and block in pdl_wait() until the
// predecessor has completed and flushed.  Rules every kernel launched this way follows: (1) pdl_wait() before the first
// access to anything another kernel of the step writes, and before its own first global write; (2) pdl_trigger() only
// after pdl_wait(), so that when a kernel's prologue runs, everything two or more kernels back is complete.
// Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(const LaunchCtx& lc, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                   Args&&... args) {
  if (lc.multi != nullptr) return cudaErrorNotSupported;   // a sequence is being recorded and this wrapper cannot record
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = lc.stream;
  cudaLaunchAttribute at[1];
  if (lc.pdl_state != nullptr && *lc.pdl_state == 1) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      svae_set_cuda_error(_e, #expr, __FILE__, __LINE__);                                     \
      return -3;                                                                              \
    }                                                                                         \
  } while (0)

void svae_set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
std::string& svae_global_error();

// cooperative launch (all CTAs co-resident: the kernel may use a grid-wide barrier); never combined with the PDL attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_coop(const LaunchCtx& lc, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = lc.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- batched ("multi") launches ---------------------------------------------------------------------------------------
// The T recognition nets of a step (sequential_vae.py:1022: q(z_t | x) depends on x only) and their latent projections are T
// independent copies of the same layer sequence on different buffers and weights.  Instead of T x ~47 small launches, the
// sequence is RECORDED once per net - every wrapper appends its kernel's argument block to the slot it occupies in the
// sequence - and issued as ONE launch per slot whose grid is the single-net grid times T along z: the kernel picks the
// argument block of item blockIdx.z from its parameter space and runs the unchanged single-net body.  Buffers, layouts,
// weights and numerics are exactly those of the per-net launches (every parity test applies unchanged); only the launch
// count and the machine fill (T x more CTAs per launch) change.
struct MultiSlot {
  void (*launch)(const void* host_args, int items, dim3 grid, dim3 block, size_t smem, cudaStream_t st) = nullptr;
  dim3 grid, block;
  size_t smem = 0, psize = 0;
  int kc = KC_MISC;
  double flops = 0, bytes = 0;
  std::vector<unsigned char> host;   // items * psize argument bytes
  int count = 0;
  int lane = 0;                      // 1: weight gradient (may be issued on a second stream, see multi_flush)
};
struct MultiRec {
  std::vector<MultiSlot> slots;
  int cursor = 0, items = 0, lane = 0;
  int pend_kc = KC_MISC; double pend_flops = 0, pend_bytes = 0;
  bool pending = false;   // a wrapper announced a launch (ProfScope) that has not been recorded yet
  std::string err;
  void begin_item() { cursor = 0; ++items; }
  // returns 0, or -1 when the sequence of this item differs from the first item's (not batchable)
  int add(void (*launch)(const void*, int, dim3, dim3, size_t, cudaStream_t), const void* args, size_t psize, dim3 grid, dim3 block,
          size_t smem) {
    if (items == 1) {
      MultiSlot sl;
      sl.launch = launch; sl.grid = grid; sl.block = block; sl.smem = smem; sl.psize = psize; sl.kc = pend_kc; sl.lane = lane;
      slots.push_back(sl);
    }
    if (cursor >= (int)slots.size()) { err = "batched launch: items record different kernel sequences"; return -1; }
    MultiSlot& sl = slots[cursor++];
    if (sl.launch != launch || sl.psize != psize || sl.grid.x != grid.x || sl.grid.y != grid.y || sl.grid.z != grid.z ||
        sl.block.x != block.x || sl.block.y != block.y || sl.block.z != block.z || sl.smem != smem) {
      err = "batched launch: items record different kernels / grids for the same slot";
      return -1;
    }
    sl.host.insert(sl.host.end(), static_cast<const unsigned char*>(args), static_cast<const unsigned char*>(args) + psize);
    sl.count += 1;
    sl.flops += pend_flops; sl.bytes += pend_bytes;
    pend_flops = pend_bytes = 0;
    pending = false;
    return 0;
  }
};
inline void ProfScope::multi_note(MultiRec* m, int kc, double flops, double bytes) {
  // a wrapper without batched-launch support would have LAUNCHED its kernel while the sequence is being recorded: caught here
  // (by the next wrapper) or at the flush
  if (m->pending) m->err = "batched launch: a kernel wrapper without batched-launch support ran inside a recorded sequence";
  m->pend_kc = kc; m->pend_flops = flops; m->pend_bytes = bytes; m->pending = true;
}

// Generic trampoline for kernels whose body is a __device__ function of plain arguments: the argument tuples of up to
// MULTI_MAX items travel BY VALUE in the kernel parameter space (constant bank, like ordinary kernel arguments: no device-side
// argument buffer, nothing to upload, capturable as is); item blockIdx.z runs the unchanged single-item body (it may use
// blockIdx.x / blockIdx.y, static and dynamic shared memory, __syncthreads ...).  More than MULTI_MAX items: several launches.
#include <cuda/std/tuple>
constexpr int MULTI_MAX = 8;
template <typename T>
struct MultiArgs { T v[MULTI_MAX]; };
template <auto Body, int MAXT, typename... A>
__global__ void __launch_bounds__(MAXT) svae_multi_kernel(const __grid_constant__ MultiArgs<cuda::std::tuple<A...>> args) {
  const cuda::std::tuple<A...>& t = args.v[blockIdx.z];
  cuda::std::apply([](const A&... a) { Body(a...); }, t);
}
// launcher shared by every batched kernel: K = the __global__ function taking MultiArgs<T> (+ extra trailing arguments)
template <typename T, typename K, typename... Extra>
static inline void multi_launch_chunks(K kernel, const void* host_args, int items, dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       unsigned z_per_item, Extra... extra) {
  static_assert(sizeof(MultiArgs<T>) <= 16 * 1024, "argument blocks of one batched launch must fit the kernel parameter space");
  for (int i0 = 0; i0 < items; i0 += MULTI_MAX) {
    const int n = items - i0 < MULTI_MAX ? items - i0 : MULTI_MAX;
    MultiArgs<T> a;
    memset(static_cast<void*>(&a), 0, sizeof a);
    memcpy(static_cast<void*>(a.v), static_cast<const unsigned char*>(host_args) + (size_t)i0 * sizeof(T), (size_t)n * sizeof(T));
    grid.z = z_per_item * (unsigned)n;
    kernel<<<grid, block, smem, st>>>(a, extra...);
  }
}
template <auto Body, int MAXT, typename... A>
struct MultiLaunch {
  static void launch(const void* host_args, int items, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    multi_launch_chunks<cuda::std::tuple<A...>>(svae_multi_kernel<Body, MAXT, A...>, host_args, items, grid, block, smem, st, 1u);
  }
};
template <auto Body, int MAXT, typename... KA, typename... Args>
static inline int multi_record(void (*)(KA...), const LaunchCtx& lc, dim3 grid, dim3 block, size_t smem, Args&&... args) {
  static_assert(sizeof...(KA) == sizeof...(Args), "argument count of the batched launch differs from the kernel body's");
  alignas(16) unsigned char buf[sizeof(cuda::std::tuple<KA...>)];
  memset(buf, 0, sizeof buf);
  new (buf) cuda::std::tuple<KA...>(static_cast<KA>(args)...);
  if (grid.z != 1) { svae_global_error() = "batched launch: the kernel already uses grid.z"; return -1; }
  if (lc.multi->add(&MultiLaunch<Body, MAXT, KA...>::launch, buf, sizeof buf, grid, block, smem) != 0) {
    svae_global_error() = lc.multi->err;
    return -1;
  }
  return 0;
}
// MULTI_RECORD(body, max threads, lc, grid, block, smem, args...): record one launch of `body` (see LaunchCtx::multi)
#define MULTI_RECORD(body, maxt, lc, grid, block, smem, ...) multi_record<body, maxt>(body, lc, grid, block, smem, __VA_ARGS__)

// ---- bf16 activation copies for the TMA-fed kernels ------------------------------------------------------------------
// A [B,H,W,C] tensor stored as bf16, PLANAR by 8-channel group, over the zero-padded linear pixel space of its consumer:
//     element (n,h,w,c)  ->  p[ ((c/8) * group_rows + front + plane*plane_rows + q) * 8 + c%8 ]
//   kind 0: q = (n*(H+2) + h)*(W+2) + w          (stride-1 windows)
//   kind 1: q = (n*(H+1) + h)*(W+1) + w          (stride-2 transposed-conv input)
//   kind 2: plane = (h&1)*2 + (w&1), q = (n*(H/2+1) + h/2)*(W/2+1) + w/2   (stride-2 conv input: four parity planes)
// so that the halo of a 128-pixel tile is, per 8-channel group, ONE contiguous run of 16-byte pixels: a single
// cp.async.bulk (TMA 1-D) brings it into exactly the canonical no-swizzle UMMA layout.  `front` zero rows before the
// first pixel, a zero gap after every plane and a zero tail make every halo an in-bounds read; padding pixels are zero
// and stay zero (producers only ever write valid pixels).
struct BfAct {
  __nv_bfloat16* p;
  int kind, B, H, W, Cpad, Hp, Wp;
  int front;              // zero rows before pixel 0 of the first plane
  long long plane_rows;   // rows between parity planes (kind 2), = pixels + gap
  long long group_rows;   // rows of one 8-channel group (front + planes + tail)
};
__host__ __device__ static inline size_t bf_index(const BfAct& d, int n, int h, int w, int c) {
  long long q;
  int pl = 0;
  if (d.kind == 2) { q = ((long long)n * d.Hp + (h >> 1)) * d.Wp + (w >> 1); pl = (h & 1) * 2 + (w & 1); }
  else q = ((long long)n * d.Hp + h) * d.Wp + w;
  return (size_t)(((long long)(c >> 3) * d.group_rows + d.front + pl * d.plane_rows + q) * 8 + (c & 7));
}
// Where an elementwise producer writes the bf16 copy of (a channel window of) its output; p == nullptr: no copy.
struct BfDst {
  BfAct a;
  int coff;        // first channel of the producer's window inside the copy (multiple of 8)
  int inner, ppr;  // feature -> (pixel, channel) mapping of a 2-D batch norm whose output is a [.., S, S, inner] map
                   // (0, 0: take the mapping of the fp32 output view)
};
// Batch-norm backward pass 1 fused into the epilogue of the input-gradient kernel that PRODUCES da (= dL/d(activated
// output) of the block described here): the first C output channels of that kernel are turned into
// g = da * act'(xhat + beta + residual) before they are stored, sum g and sum g*xhat are reduced into S, and the shortcut
// gradient is written (dres = or += g).  The block's backward is then one pass (bn_bwd_apply reading g where da was).
struct BnBwdFuse {
  const float* y;          // the block's pre-BN contraction output [rows, C]
  const double* stats;     // its forward statistics [2C]
  const float* beta;       // [C]
  const float* res;        // residual added before the activation (4-D view) or nullptr
  int res_ld, res_coff;
  float* dres;             // shortcut gradient [rows, C] or nullptr
  int dres_acc;
  double* S;               // [2C] backward sums (pre-zeroed)
  int C, act;
  long long rows;
};
// Batch-norm FORWARD fused into the contraction that produces the block's pre-BN tensor (conv2d_bn_lrelu / conv2d_t_bn[_relu]
// as ONE kernel, abstract_network.py:17-24,36-61): every accumulator tile of the CTA stays in tensor memory, pass 1 writes the
// pre-BN tensor (kept for the backward) and reduces the channel statistics, a grid-wide barrier makes them final, pass 2 reads
// the accumulators again, normalises, adds the ladder shortcut, applies the activation and writes the activated output in
// the layouts its consumers read (fp32 channel window and / or the bf16 planar copy of the next TMA-fed contraction).
struct BnFwdFuse {
  const float* beta;                 // [C]
  const float* res;                  // tensor added before the activation (4-D view) or nullptr
  int res_ld, res_coff;
  float* out;                        // fp32 activated output (channel window of a concat buffer) or nullptr
  int out_ld, out_coff;
  BfDst bf;                          // bf16 planar copy of the activated output (bf.a.p == nullptr: none)
  unsigned* counter;                 // grid-barrier counter, zeroed by the caller before the launch
  int act;
  long long rows;                    // B*Hout*Wout (batch-norm population)
};
BfAct bf_act_describe(int kind, int B, int H, int W, int C);
size_t bf_act_bytes(const BfAct& d);
int bf_act_fill(const LaunchCtx& lc, const BfAct& d, View src, int C);   // fp32 NHWC window -> padded bf16 copy (incl. zeros)
bool tc2_supported(const Geom& g);
int tc2_input_kind(const Geom& g);
// w_tile_width: tile width the weights were packed with (tc_pack_entry / tc_pack_weights): 128 or tc2_pick_ntw(g)
int tc2_gather_gemm(const LaunchCtx& lc, const Geom& g, const BfAct& in, int chan0, const void* w_packed, View out,
                    double* stats, const BnBwdFuse* fuse = nullptr, int w_tile_width = 128, const BnFwdFuse* fwd_fuse = nullptr);
// can the contraction g (B set, weights packed with w_tile_width) keep all of its accumulator tiles in tensor memory, i.e. run
// with a BnFwdFuse?  C = output channels
bool tc2_bnf_supported(const Geom& g, int w_tile_width, int sm_count, View out, int C);
int tc2_pick_ntw(const Geom& g, int sm_count);
bool tc2_fuse_supported(const Geom& g, View out, int C);   // can this launch carry a BnBwdFuse over its first C output channels?
bool tc2_wgrad_supported(const Geom& g);   // g: conv-gather geometry (see tc_wgrad)
// dw2 != nullptr: rows (X channels) [a_split, Ca) of the gradient belong to a second parameter tensor [16][Ca - a_split][Cb]
int tc2_wgrad(const LaunchCtx& lc, const Geom& g, const BfAct& x, const BfAct& dy, float* dw, float* dw2 = nullptr, int a_split = 0);
// ---- SIMT fp32 contractions (kernels_simt.cu) ------------------------------------------------------------------
int simt_gather_gemm(const LaunchCtx& lc, const Geom& g, View in, const float* w, View out, double* stats);
// dW(tap,a,b) = sum_rows X(gathered at tap, channel a) * dY(row, channel b); layout w[(tap*Ca + a)*Cb + b].
// Geometry: X is [B,Hin,Win,Ca] (g.Cin = Ca), dY is [B,Hout,Wout,Cb] (g.Cout = Cb), conv-gather (mode 0) indexing.
int simt_wgrad(const LaunchCtx& lc, const Geom& g, View x, View dy, float* dw);
// skinny fully-connected helpers (recognition heads, latent projections)
struct HeadSet {   // all heads that read one flattened feature map (sequential_vae.py:1592-1609)
  int nheads;
  const float* w[4]; const float* b[4];
  float* gw[4]; float* gb[4];
  int n[4], col[4], is_sd[4];
};
int heads_fwd(const LaunchCtx& lc, const float* flat, int B, int K, const HeadSet& hs, float* mu_pre, float* sd_pre, int Z);
int heads_dgrad(const LaunchCtx& lc, const HeadSet& hs, const float* dmu, const float* dsd, int B, int Z, int K, float* d_flat);
int heads_wgrad(const LaunchCtx& lc, const float* flat, const HeadSet& hs, const float* dmu, const float* dsd, int B, int Z, int K);
int lat_wgrad(const LaunchCtx& lc, View z, const float* dy, int B, int KZ, int N, float* dw);
int lat_dz(const LaunchCtx& lc, const float* dy, const float* w, int B, int KZ, int N, View dz);

// ---- fused latent projections (kernels_lat.cu): fc (K <= 32) + 2-D batch norm + activation in one kernel per direction --
bool lat_fused_supported(int B, int KZ);
// y[B,N] = z.W ; stats = (sum, sumsq) per feature ; act(bn(y)) -> out (fp32 concat slot, may be NULL) and bf (may be NULL)
// mom_scratch: 64 doubles of scratch (second moments of z, reduced once for batches > 512), may be NULL
int lat_fwd_fused(const LaunchCtx& lc, View z, const float* w, const float* beta, int B, int KZ, int N, int act, float* y,
                  double* stats, FeatView out, BfDst bf, double* mom_scratch = nullptr);
// dw[KZ,N] += z^T dy ; dbeta = sum g ; dz[B, window] += dy.W^T  (dy never materialised)
int lat_bwd_fused(const LaunchCtx& lc, FeatView da, const float* y, const double* stats, const float* beta, View z,
                  const float* w, int B, int KZ, int N, int act, float* dw, float* dbeta, View dz);

// 2-D batch norm (per feature over the batch) of a fully-connected block, one launch per direction
int bn2d_fwd(const LaunchCtx& lc, const float* y, const float* beta, int B, int N, int act, double* stats, FeatView out, BfDst bf);
int bn2d_bwd(const LaunchCtx& lc, FeatView da, const float* y, const double* stats, const float* beta, int B, int N, int act,
             float* dy, float* dbeta);

// ---- elementwise / reductions (kernels_elem.cu) ----------------------------------------------------------------
int col_stats(const LaunchCtx& lc, const float* y, int64_t rows, int C, double* stats);
// out.p may be NULL when only the bf16 copy is wanted; bf.a.p may be NULL
int bn_act_fwd(const LaunchCtx& lc, const float* y, const double* stats, const float* beta, int64_t rows, int feats,
               int act, FeatView residual, FeatView out, BfDst bf = BfDst{});
// pass 1: dyhat = da * act'(bn(y)+res) ; S += (sum dyhat, sum dyhat*xhat) ; optional dres = dyhat (or += when acc)
int bn_bwd_reduce(const LaunchCtx& lc, FeatView da, const float* y, const double* stats, const float* beta,
                  int64_t rows, int feats, int act, FeatView residual, float* dyhat, double* S, float* dres,
                  int dres_accumulate);
// pass 2: dy = rstd * (dyhat - S1/rows - xhat*S2/rows) in place ; dbeta = S1
int bn_bwd_apply(const LaunchCtx& lc, float* dyhat, const float* y, const double* stats, const double* S, int64_t rows,
                 int feats, float* dbeta, BfDst bf = BfDst{});
// Both passes in ONE cooperative kernel (4-D batch norm, C % 8 == 0): pass 1 reduces the two sums, a grid-wide barrier, pass 2
// recomputes g from da / y (L2-resident) and writes dy - g is never materialised, dy fp32 only when `dy` != nullptr.
// S must have 2C + 1 doubles, zeroed: the last one is the barrier counter.  Returns 1 when the layout is not supported
// (the caller then runs bn_bwd_reduce + bn_bwd_apply).
int bn_bwd_fused(const LaunchCtx& lc, FeatView da, const float* y, const double* stats, const float* beta, int64_t rows,
                 int feats, int act, FeatView residual, float* dy, double* S, float* dres, int dres_accumulate, float* dbeta,
                 BfDst bf);
// the same with g read through a 4-D channel window (where a fused input-gradient epilogue left it) and the fp32 dy
// optional (dy == nullptr: only the bf16 copy is written)
int bn_bwd_apply_from(const LaunchCtx& lc, View g, float* dy, const float* y, const double* stats, const double* S, int64_t rows,
                      int feats, float* dbeta, BfDst bf);
struct OutMixParams {
  int64_t pixels;  // B*H*W
  int C;
  int has_gate;
  float lo, hi, minr, maxr;
  int gxld;        // channel stride of the dL/dx_t buffers gx_in / gx_prev (0: = C)
};
// x_t = mix(sigmoid(u+bias)) ; accumulates sum (x_t - tgt)^2 into *recon_sum
int out_mix_fwd(const LaunchCtx& lc, const OutMixParams& p, const float* u, const float* b_out, const float* b_gate,
                const float* xprev, const float* tgt, float* xt, double* recon_sum, BfDst xt_bf = BfDst{});
// g = gx_in (or 0) + coef*(x_t - tgt) ; du, gx_prev, bias grads
int out_mix_bwd(const LaunchCtx& lc, const OutMixParams& p, const float* u, const float* b_out, const float* b_gate,
                const float* xprev, const float* tgt, const float* xt, const float* gx_in, float coef, float* du,
                float* gx_prev, float* db_out, float* db_gate, BfDst du_out_bf = BfDst{}, BfDst du_gate_bf = BfDst{},
                int gate_in_out_bf = 0);   // 1: the gate's gradient goes to channel C of du_out_bf (merged output + gate contraction)
struct ReparamParams {
  int B, Z;
  float clip, prior;
};
// Per-iteration scalars live in device memory so that a captured CUDA graph of the train step can be replayed with new
// values (learning-rate schedule + Adam bias correction, KL warm-up coefficient, Philox key / counter).
struct SvaeDyn {
  float lr_t;                    // lr * sqrt(1 - b2^t) / (1 - b1^t)          sequential_vae.py:1267,1356
  float reg;                     // reg_coeff                                 :1357
  unsigned long long seed;       // Philox key of the in-kernel eps
  unsigned long long iteration;  // Philox counter base = (iteration*T + t) * max_batch*Z
};
// eps == NULL: Philox N(0,1) keyed by dyn->seed at counter (dyn->iteration*T + t)*stride + i
int reparam_fwd(const LaunchCtx& lc, const ReparamParams& p, const float* mu_pre, const float* sd_pre, const float* eps,
                const SvaeDyn* dyn, int T, int t, uint64_t stride, float* eps_store, float* mu, float* sd, float* z,
                double* kl_sum);
// KL gradient coefficient = dyn->reg * kl_scale
int reparam_bwd(const LaunchCtx& lc, const ReparamParams& p, const float* dz, const float* mu_pre, const float* mu,
                const float* sd, const float* eps, const SvaeDyn* dyn, float kl_scale, float* dmu_pre, float* dsd_pre);
// dyn != NULL: the step size is read from dyn->lr_t (lr_t ignored)
int adam_update(const LaunchCtx& lc, float* p, const float* g, float* m, float* v, int64_t n, const SvaeDyn* dyn, float lr_t,
                float beta1, float beta2, float eps, float clip, float grad_scale);
// A run of tied gradient slices (homogeneous chains): `members` slices of `n` floats (multiple of 4) at element offsets off[]
// of the gradient arena; tie_reduce leaves their element-wise sum (members added in ascending order) in every slice.
struct TieRun {
  int64_t n;
  int members;
  int64_t off[64];   // SVAE_MAX_STEPS (static_assert in model.cu)
};
int tie_reduce(const LaunchCtx& lc, float* grad_arena, const TieRun& run);
int apply_noise(const LaunchCtx& lc, const float* x, float* out, int64_t n, float pepper_prob, float salt_prob, float scale,
                float lo, float hi, uint64_t seed, float* draws /* [3,n] keep, salt, gaussian; may be nullptr */);
int fill_normal(const LaunchCtx& lc, float* dst, int64_t n, uint64_t seed, uint64_t counter_base);
// xs = xt + (dyn ? dyn->reg : reg_fixed) * sigma * noise  (+ bf16 planar copy); noise == NULL: Philox keyed (dyn->seed ^ key_xor)
// at counter (dyn ? dyn->iteration * t_stride : 0) + base_fixed + element
int chain_noise(const LaunchCtx& lc, const float* xt, const float* noise, const SvaeDyn* dyn, float reg_fixed, float sigma,
                uint64_t key_xor, uint64_t base_fixed, uint64_t t_stride, float* xs, int64_t pixels, int C, BfDst xt_bf);
int axpy_inplace(const LaunchCtx& lc, float* dst, const float* src, int64_t n);  // dst += src
// debug probes (parity tests only): dense fp32 copies of a bf16 planar copy / of a feature view
int probe_bf_unpack(const LaunchCtx& lc, const BfAct& a, int coff, int C, int B, float* dst);   // -> [B,H,W,C]
int probe_fv_gather(const LaunchCtx& lc, const FeatView& v, int64_t rows, int feats, float* dst);   // -> [rows,feats]

// ---- tcgen05 contractions (kernels_tc.cu) ----------------------------------------------------------------------
struct TcPlan;  // opaque per-layer packed-weight + schedule state
bool tc_supported(const Geom& g);
bool tc_wgrad_supported(const Geom& fwd);
// g: conv-gather geometry (X = conv input side, dY = conv output side); accumulates into dw[(tap*Cin + a)*Cout + b]
int tc_wgrad(const LaunchCtx& lc, const Geom& g, View x, View dy, float* dw);
int tc_gather_gemm(const LaunchCtx& lc, const Geom& g, View in, const void* w_packed, View out, double* stats);
size_t tc_packed_bytes(const Geom& g);
int tc_pack_weights(const LaunchCtx& lc, const Geom& g, const float* w, void* w_packed, int tile_width = 128);
// w2 != nullptr: the layer's weights are TWO parameter tensors side by side along the deconv's output-channel dimension (the
// output (C channels) and gate (1 channel) deconvs of a chain step run as one contraction): channels [0, n1) come from w, the
// rest from w2; that dimension is `co` for the forward form (w_out_major 1) and `ci` for the input-gradient form (w_out_major 0).
struct TcPackEntry { Geom g; const float* w; void* out; int KC, Cin_p, N_p, TW; long long total; const float* w2; int n1; };
TcPackEntry tc_pack_entry(const Geom& g, const float* w, void* w_packed, int tile_width = 128, const float* w2 = nullptr, int n1 = 0);
int tc_pack_batched(const LaunchCtx& lc, const void* dev_entries, int n, double total_elems);
// ---- tcgen05 fully-connected kernel (kernels_fc.cu): fp32 operands read directly, bf16 in shared memory -----------
bool tc_fc_supported(int K, int N);
// mode 0: out[M,N] = x[M,K] . w[K,N] ; 1: out[M,K] = x[M,N] . w[K,N]^T ; 2: out[K,N] = x[M,K]^T . w[M,N]
int tc_fc(const LaunchCtx& lc, int mode, const float* x, int ldx, const float* w, int ldw, float* out, int ldo, int M, int K,
          int N, int accumulate);
