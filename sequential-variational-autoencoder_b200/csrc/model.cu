// model.cu - static plan + executor of the Sequential-VAE chain and the C ABI (include/svae.h).
//
// The reference builds a TF graph (construct_network, sequential_vae.py:877-984) and runs it with Session.run; here
// the graph is a static plan built once from svae_config: per chain step a recognition net (inference_ladder,
// :1537-1630), a chain encoder (compute_encodings, :1745-1777) and a ladder decoder (generator_ladder, :1636-1739),
// each a list of "contraction + batch-norm + activation" blocks over a bump-allocated activation arena.  Channel
// concats are channel windows of shared buffers (no copies), the reshape between fc and conv layers is free (NHWC,
// H,W,C flatten order).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <memory>
#include <vector>

#include "../../include/svae.h"
#include "common.cuh"
#include "nccl_dl.h"

// Every non-kernel operation enqueued by this file ends the "previous node is a kernel" state that programmatic
// dependent launch relies on (common.cuh, launch_k): the next kernel then takes an ordinary full dependency.
#define cudaMemsetAsync(...) (h->pdl_prev = 0, cudaMemsetAsync(__VA_ARGS__))
#define cudaMemcpyAsync(...) (h->pdl_prev = 0, cudaMemcpyAsync(__VA_ARGS__))
#define cudaStreamWaitEvent(s_, e_, f_) ((void)(((s_) == h->stream) ? (h->pdl_prev = 0) : 0), cudaStreamWaitEvent(s_, e_, f_))

// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
std::string& svae_global_error() { return g_err; }
void svae_set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_err = buf;
}

namespace {

struct Param {
  std::string name;
  int ndim;
  int shape[4];
  int64_t numel, offset;
  int step, flags;
};

static_assert(sizeof(TieRun::off) / sizeof(int64_t) == SVAE_MAX_STEPS, "TieRun holds one offset per chain step");

struct PubParam {   // a variable as callers see it: one TF name, >= 1 per-step slices (see svae_handle::pub)
  std::string name;
  std::vector<int> members;   // indices into svae_handle::params, ascending chain step
};

// contraction + batch-norm + activation (conv2d_bn_lrelu / conv2d_t_bn[_relu] / fc_bn_lrelu, abstract_network.py:17-71)
struct Block {
  Geom g;          // forward geometry, B patched per call
  int w = -1, beta = -1;
  int act = ACT_NONE;
  int rpi = 1;     // BN rows per image (H*W for 4-D BN, 1 for 2-D BN)
  int feats = 0;   // BN features
  float* y = nullptr;        // pre-BN output [rows, feats]
  double* stats = nullptr;   // [2*feats] sum, sumsq (forward-zeroed)
  double* S = nullptr;       // [2*feats] backward sums (backward-zeroed)
  unsigned* fwd_bar = nullptr;   // grid-barrier counter of the fused conv + batch-norm forward kernel (forward-zeroed)
  FeatView out{};            // activated output location
  FeatView res{};            // residual added before the activation (ladder shortcut)
  void* w_packed = nullptr;      // tcgen05 packed weights (forward form)
  void* w_packed_d = nullptr;    // tcgen05 packed weights (dgrad form)
  bool tc_fwd = false, tc_dgrad = false, tc_wgrad = false;
  int dy_slot = -1;              // index of this block's d(pre-BN output) buffer inside a GradSet
  // TMA-fed kernels (tc2): bf16 planar copy of the block's input (written by whoever produces that tensor), and where
  // this block's own activated output is additionally written for its consumer
  BfAct in_bf{}; bool tc2_fwd = false, tc2_dgrad = false, tc2_wgrad = false;
  BfDst out_bf{};
  int tw_f = 128, tw_d = 128;    // tile width of the packed forward / input-gradient weights (tc2_pick_ntw at plan time)
  // the fp32 copy of the activated output is not written when every reader takes the bf16 planar copy (TMA-fed consumer,
  // and in training its TMA-fed weight gradient): 4 of the 10 bytes per element this block's batch-norm kernel moves
  bool skip_f32 = false;
  // set by the input-gradient kernel that produced this block's dL/d(activated output) when it already turned it into
  // g = da * act' and reduced the two backward sums (BnBwdFuse): the block's backward then skips pass 1
  bool g_fused = false;
  // what the last forward / backward of this block read (svae_debug_block_tensor: local-replay parity tests)
  View dbg_in{}; FeatView dbg_da{}; int dbg_gs = -1;
  bool y_zeroed = false;       // plan: y lives in the forward-zeroed arena (split-K fc output: no memset node on the chain)
  bool in_f32_valid = true;    // plan: the fp32 tensor this block reads is materialised (its producer does not skip_f32)
  bool dbg_dy_f32 = false;     // last backward: the fp32 dL/dy was written (some consumer is not TMA-fed)
};

// Gradient scratch of ONE chain step's backward.  Two sets (step parity) let the side streams (weight gradients,
// recognition net) of step t still read their buffers while the main stream already runs step t-1.
struct GradSet {
  std::vector<float*> d_c, d_dcat, d_e, d_inf, dy;
  std::vector<BfAct> dy_bf;      // bf16 planar copies of dy for the TMA-fed input-gradient kernels (p == nullptr: none)
  BfAct du_out_bf{}, du_gate_bf{};
  float *d_fca = nullptr, *d_cc = nullptr, *d_z = nullptr, *d_mu_pre = nullptr, *d_sd_pre = nullptr, *d_u = nullptr;
};

struct GraphEntry {   // a captured train step, valid for exactly these arguments
  const float *x, *tgt, *eps;
  int B;
  cudaGraphExec_t exec;
  int64_t launches;   // kernels inside the graph (svae_launch_count keeps counting them)
  bool bucket = false;   // captured with the per-chain-step Adam + repack (no repack at the start of the graph)
};

struct Head {  // layers.fully_connected head (sequential_vae.py:1592,1594,1607,1609)
  int w, b, n, col, is_sd;
};

struct Step {
  int t;
  // recognition net
  std::vector<Block> inf;                 // 2(L-1) conv blocks
  std::vector<std::vector<Head>> heads;   // per level 0..L-2
  float *mu_pre, *sd_pre, *mu, *sd, *z, *eps;
  // chain encoder (t>=1)
  std::vector<Block> enc;                 // 2(L-1)+1 conv blocks
  Block encfc;
  // decoder
  std::vector<Block> lat;                 // L latent projections
  Block decfc;
  std::vector<Block> ta, tb;              // per level: stride-2 deconv + BN (+shortcut, relu) ; stride-1 deconv + BN + relu
  std::vector<float*> dcat;               // concat buffers [B,S,S,2F]
  float* cc;                              // concat(e[L], P_{L-1}) [B, ncc]
  int ncc;
  int w_out, b_out, w_gate, b_gate;       // output deconvs (live bias)
  Geom g_out, g_gate;
  Block outb, gateb;                      // the same two deconvs as contraction blocks (no batch-norm) for TC routing
  // t >= 1, TMA-fed kernels: the output (C columns) and gate (1 column) deconvs read the same input and are each padded to 16
  // columns - ONE contraction with the two weight tensors packed side by side writes all C + 1 columns of u, one input
  // gradient reads the C + 1 channel copy of d_u, one weight gradient scatters its rows to the two parameter tensors: two
  // dependent kernels less per chain step on the chain, one less beside it (sequential_vae.py:1720,1727 share cur_sample).
  Block ogb; Geom g_og; bool merge_og = false;
  float *u, *xt;
  float* xs = nullptr;                    // add_noise_to_chain: the sample x_t + noise that the next step reads (own buffer per step)
  double* lat_mom = nullptr;              // [L][64] scratch: second moments of z_t per latent group (lat_fwd_fused)
  BfAct xt_bf{};                          // bf16 copy of x_t for the next step's chain encoder
  int64_t p_begin, p_end;                 // parameter range (all-reduce bucket)
  int64_t p_rec_end;                      // [p_begin, p_rec_end): recognition net, [p_rec_end, p_end): chain encoder + decoder
};

struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* get(size_t n) {
    size_t o = (off + 255) & ~(size_t)255;
    off = o + n * sizeof(T);
    return base ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};

Geom conv_geom(int H, int W, int Ci, int Co, int stride) {
  Geom g{};
  g.B = 0; g.Hin = H; g.Win = W; g.Cin = Ci; g.Hout = H / stride; g.Wout = W / stride; g.Cout = Co;
  g.KH = g.KW = 4; g.stride = stride; g.pad = 1; g.mode = 0; g.w_out_major = 0; g.accumulate = 0;
  return g;
}
Geom deconv_geom(int H, int W, int Ci, int Co, int stride) {
  Geom g{};
  g.B = 0; g.Hin = H; g.Win = W; g.Cin = Ci; g.Hout = H * stride; g.Wout = W * stride; g.Cout = Co;
  g.KH = g.KW = 4; g.stride = stride; g.pad = 1; g.mode = 1; g.w_out_major = 1; g.accumulate = 0;
  return g;
}
Geom fc_geom(int K, int N) {
  Geom g{};
  g.B = 0; g.Hin = g.Win = g.Hout = g.Wout = 1; g.Cin = K; g.Cout = N;
  g.KH = g.KW = 1; g.stride = 1; g.pad = 0; g.mode = 0; g.w_out_major = 0; g.accumulate = 0;
  return g;
}
// geometry that maps d(out) -> d(in) for a forward geometry
Geom dgrad_geom(const Geom& f) {
  Geom g = f;
  g.Hin = f.Hout; g.Win = f.Wout; g.Cin = f.Cout;
  g.Hout = f.Hin; g.Wout = f.Win; g.Cout = f.Cin;
  g.mode = 1 - f.mode;
  g.w_out_major = 1 - f.w_out_major;
  g.accumulate = 0;
  return g;
}

}  // namespace

struct svae_handle {
  svae_config cfg;
  int device = 0;
  int sm_count = 148;
  std::string err;
  cudaStream_t own_stream = nullptr, stream = nullptr, comm_stream = nullptr;
  int64_t launches = 0;
  // parameters
  std::vector<Param> params;
  int64_t arena_numel = 0;
  // Homogeneous chains (share_theta_weights / share_phi_weights, sequential_vae.py:213-214): the engine keeps one parameter
  // slice per chain step (`params`, above) and TIES the slices of a shared variable - they start equal (svae_param_set /
  // svae_adam_set write every member), receive the same summed gradient (tie_reduce after the backward) and therefore the
  // same deterministic element-wise Adam update, so they stay bit-identical.  `pub` is the variable table callers see:
  // one entry per TF variable ("phi/inference_network/...", "theta/generative_network/..."), members = its step slices.
  std::vector<PubParam> pub;
  std::vector<TieRun> tie_runs;      // contiguous runs of tied slices (what tie_reduce sums)
  bool tied = false;
  float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr;
  int64_t adam_t = 0;
  bool weights_dirty = true;
  // plan
  int L = 0, T = 0, Z = 0, D = 0, C = 0;
  std::vector<int> zoff;
  std::vector<Step> steps;      // T entries (parameters differ per step)
  int act_sets = 0;             // T when train_capacity else 1
  char* act_base = nullptr; size_t act_bytes = 0;
  char* zf_base = nullptr; size_t zf_bytes = 0;   // forward-zeroed (stats, loss sums)
  char* zb_base = nullptr; size_t zb_bytes = 0;   // backward-zeroed (S sums)
  double* loss_sums = nullptr;  // [2T] recon_sum, kl_sum
  double* loss_host = nullptr;  // pinned
  // gradient scratch (one step's worth)
  // Gradient scratch sets, used round-robin by the chain steps of a backward pass.  With NS sets the chain only has to wait
  // for the side streams (weight gradients, recognition net) of step t + NS before it reuses their buffers; with 2 sets a
  // timeline of the step showed the chain stalled ~0.4 ms per chain step on exactly that wait (the side streams lag behind
  // under contention).  NS = min(T, SVAE_GRAD_SETS, default 8): no wait at all for the reference's T = 8.
  static constexpr int MAX_GS = 8;
  GradSet gs[MAX_GS];
  int n_gs = 2;
  bool bwd_prezero = false;   // d_z / d_cc / d_e[last] of every scratch set are zeroed with the backward arena (n_gs >= T)
  int n_dy_slots = 0;
  std::vector<size_t> dy_slot_elems;
  float* gx[2] = {nullptr, nullptr};
  // channel stride of the dL/dx_t buffers: 4 for 3-channel images on the TMA-fed kernels, so that the first chain-encoder conv's
  // input gradient accumulates with 16-byte accesses (as a 4-column output whose fourth column has zero weights) instead of
  // scalar read-modify-writes at a 12-byte stride
  int gxld = 0;
  char* grad_base = nullptr; size_t grad_bytes = 0;
  char* bf_base = nullptr; size_t bf_bytes = 0;   // bf16 planar activation copies (zero-initialised once: the padding stays zero)
  BfAct x_bf{};                                   // copy of the input batch for the recognition nets' first conv
  bool use_tc2 = true;
  // SVAE_TIMELINE=1 (eager mode only): timing events at phase boundaries of every stream, printed by svae_sync
  bool timeline = false;
  struct Mark { cudaEvent_t e; const char* what; int t; };
  std::vector<Mark> marks;
  int bf_last_B = -1;                             // batch the copies were last written with (stale rows are zeroed on change)
  // execution: main stream (`stream`) + side streams; `cur` is where the next launch goes
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};   // 0: chain weight gradients, 1: recognition / latent branch, 2: its weight gradients, 3: latent branch when the recognition backward is batched
  cudaStream_t cur = nullptr;
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  bool use_streams = true, use_graph = true, capturing = false;
  int fork_mask = 15;   // debugging / ablation: 1 chain wgrads, 2 recognition branch, 4 its wgrads, 8 forward recognition
  // Bucketed update (svae_train_step only): the clipped-Adam update and the bf16 operand repack of chain step t's parameters
  // run on `upd_stream` as soon as that step's gradients are final (after its all-reduce when data parallel), overlapped with
  // the backward of the earlier steps - every step owns its parameters (inhomogeneous chain), so nothing still reads them.
  cudaStream_t upd_stream = nullptr;
  bool bucket_update = false;        // set by the train-step entry points around backward_impl
  bool use_bucket_update = true;     // SVAE_BUCKET_UPDATE=0: one Adam launch + one repack launch after the backward
  std::vector<int> pack_step_begin;  // pack-table entry range of every chain step (T + 1 offsets)
  std::vector<double> pack_step_elems;
  std::vector<int> pack_rec_end;     // end of the recognition net's entries inside every step's range
  std::vector<double> pack_entry_elems;
  int pdl_prev = 0;     // 1: the last node enqueued on the chain stream is a kernel (see launch_k)
  bool use_pdl = true;  // SVAE_PDL=0 disables programmatic dependent launch
  // SVAE_COOP_BN=1: batch-norm backward as ONE cooperative kernel (reduce, grid barrier, apply) instead of two.  Correct and
  // 1.2 ms/step less kernel time in isolation, but measured SLOWER inside the step (14.5 vs 12.6 ms): a cooperative grid has to
  // become resident all at once, which drains the SMs the side streams were filling.  Off by default.
  bool use_coop_bn = false;
  // SVAE_FUSE=1: batch-norm backward pass 1 inside the epilogue of the producing input-gradient kernel.  Correct (the GPU
  // suite passes with it) but measured SLOWER on B200 (13.5 vs 12.7 ms/step): the epilogue has 4 warps per SM for work a
  // standalone kernel spreads over 64, and it sits on the chain's critical path.  Off by default.
  bool use_fuse = false;
  bool no_bnf = false;   // set around the recognition nets: they run beside the chain (side streams) and keep the two-kernel blocks
  int ablate = 0;       // SVAE_ABLATE, TIMING EXPERIMENTS ONLY (results are wrong): 1 skip weight gradients, 2 skip the recognition / latent backward
  int eager_steps = 0;
  std::vector<GraphEntry> graphs;
  // per-iteration scalars (device copy + pinned ring)
  SvaeDyn dyn_host{}; SvaeDyn* dyn_dev = nullptr; SvaeDyn* dyn_ring = nullptr; int dyn_slot = 0;
  std::vector<cudaEvent_t> dyn_ev;
  // io staging
  float *in_x = nullptr, *in_tgt = nullptr, *in_eps = nullptr, *io_steps = nullptr, *gen_prev = nullptr;
  float *pin_x = nullptr, *pin_tgt = nullptr, *pin_eps = nullptr;
  size_t io_steps_elems = 0;
  // tcgen05 packed weights
  char* pack_base = nullptr; size_t pack_bytes = 0;
  void* pack_table = nullptr; int pack_entries = 0;
  int tc_layers = 0;
  // last forward
  int last_B = 0; float last_reg = 1.f; bool have_fwd = false;
  const float *last_x = nullptr, *last_tgt = nullptr;
  uint64_t iteration = 0;
  // nccl
  NcclApi* nccl = nullptr; void* comm = nullptr; int rank = 0, nranks = 1;
  std::vector<cudaEvent_t> bucket_ev; cudaEvent_t comm_done = nullptr;

  // Batched launches of the T recognition nets (MultiRec, common.cuh): q(z_t | x) depends on x only (sequential_vae.py:1022),
  // so the T nets + latent projections are recorded net by net and issued as ONE launch per layer kernel over a group of
  // chain steps (blockIdx.z = step), forward before / beside the chain and backward in groups behind it.
  bool use_multi = true;             // SVAE_MULTI=0: per-step launches on the side streams (round-1 behaviour)
  MultiRec* multi = nullptr;         // recording in progress (see lc())
  // SMs every recorded item sizes its grid for: a batched launch carries `items` grids side by side, so each is sized for
  // ~1/items of the machine (fewer, longer-lived CTAs per item: resident weights are fetched once per CTA, the weight
  // gradient flushes fewer partial sums) instead of `items` full-machine grids queueing behind each other
  int multi_sm = 148;
  float multi_sm_factor = 0.5f;      // SVAE_MULTI_SM_FACTOR (measured: 0.5 -> 10.70, 1 -> 10.83, 2 -> 11.03, full grids -> 11.30 ms/step)
  int wgrad_sm = 0;                  // SVAE_WGRAD_SM: SM budget of the chain's weight-gradient launches (side stream), 0 = all
  // Deferred chain weight gradients: while set, the TMA-fed weight gradients of the chain's blocks are RECORDED (one item per
  // chain step) instead of being launched on the weight-gradient stream; backward_impl issues them per group of chain steps as
  // one launch per layer (blockIdx.z = chain step), every item cut for wrec_sm SMs (fewer pixel splits = fewer partial-sum flushes).
  MultiRec* wrec = nullptr;
  int wrec_sm = 148;
  std::vector<int> wgrad_groups;     // chain steps per deferred weight-gradient group, backward order (SVAE_WGRAD_GROUPS; 0 = off)
  int upd_sm = 37;                   // SVAE_UPD_SM: SM budget of the per-bucket Adam + repack launches (update stream), 0 = all
  // exposed: nothing runs beside this group (first forward group: the chain waits for it; last backward group: the chain is
  // done) - its items share the whole machine
  void begin_multi(MultiRec* m, int items, bool exposed) {
    multi = m;
    int v = (int)((exposed ? 1.f : multi_sm_factor) * (float)sm_count / (float)(items > 0 ? items : 1) + 0.999f);
    multi_sm = v < 8 ? 8 : (v > sm_count ? sm_count : v);
  }
  std::vector<int> rec_fwd_groups;   // chain steps per forward group (SVAE_REC_FWD_GROUPS, default 2,6,6,...)
  std::vector<int> rec_bwd_groups;   // chain steps per backward group, in backward order (SVAE_REC_BWD_GROUPS, default 5,2,1)

  // chain noise (cfg.add_noise_to_chain): injected draws [T, chain_noise_B, H, W, C] or Philox
  float* chain_noise_dev = nullptr; int chain_noise_B = 0;
  bool noisy() const { return cfg.add_noise_to_chain != 0; }
  const float* chain_in(int t) const { return noisy() ? steps[t].xs : steps[t].xt; }   // what chain step t + 1 reads

  Profiler prof;
  LaunchCtx lc() {
    LaunchCtx c{cur ? cur : stream, &launches, sm_count, &prof};
    if (multi != nullptr) { c.multi = multi; c.sm_count = multi_sm; return c; }
    if (use_pdl && !prof.enabled && (cur == nullptr || cur == stream)) c.pdl_state = &pdl_prev;
    return c;
  }
  float* pw(int idx) { return P + params[idx].offset; }
  float* pg(int idx) { return G + params[idx].offset; }
};

namespace {

int fail(svae_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_err = msg;
  return code;
}
#define H_TRY(expr)                                   \
  do {                                                \
    int _r = (expr);                                  \
    if (_r != 0) {                                    \
      if (h->err.empty() || _r == -3) h->err = g_err; \
      return _r;                                      \
    }                                                 \
  } while (0)
#define H_CUDA(expr)                                                \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) {                                        \
      svae_set_cuda_error(_e, #expr, __FILE__, __LINE__);           \
      h->err = g_err;                                               \
      return SVAE_ECUDA;                                            \
    }                                                               \
  } while (0)

// ---- parameter table in TF creation order -------------------------------------------------------
struct Namer {
  std::string prefix;
  int conv = 0, convt = 0, bn = 0, fc = 0;
  static std::string nm(const std::string& base, int n) { return n == 0 ? base : base + "_" + std::to_string(n); }
  std::string next_conv() { return prefix + "/" + nm("Conv", conv++); }
  std::string next_convt() { return prefix + "/" + nm("Conv2d_transpose", convt++); }
  std::string next_bn() { return prefix + "/" + nm("BatchNorm", bn++); }
  std::string next_fc() { return prefix + "/" + nm("fully_connected", fc++); }
};

int add_param(svae_handle* h, const std::string& name, std::vector<int> shape, int step, int flags) {
  Param p;
  p.name = name;
  p.ndim = (int)shape.size();
  p.numel = 1;
  for (int i = 0; i < 4; ++i) p.shape[i] = i < p.ndim ? shape[i] : 1;
  for (int s : shape) p.numel *= s;
  p.offset = h->arena_numel;
  p.step = step;
  p.flags = flags;
  h->arena_numel += (p.numel + 3) / 4 * 4;  // keep every tensor 16-byte aligned
  h->params.push_back(p);
  return (int)h->params.size() - 1;
}

// conv2d_bn_lrelu-style block: weights, inert biases, BN beta
void add_block_params(svae_handle* h, Namer& nm, Block& b, int kind /*0 conv,1 convT,2 fc*/, std::vector<int> wshape,
                      int nout, int step, int flags) {
  std::string l = kind == 0 ? nm.next_conv() : kind == 1 ? nm.next_convt() : nm.next_fc();
  std::string bn = nm.next_bn();
  b.w = add_param(h, l + "/weights", wshape, step, flags);
  add_param(h, l + "/biases", {nout}, step, flags | SVAE_PF_INERT);
  b.beta = add_param(h, bn + "/beta", {nout}, step, flags);
}

void build_params(svae_handle* h) {
  const svae_config& c = h->cfg;
  const int L = c.levels, T = c.mc_steps, C = c.channels;
  const int* F = c.filter_sizes;
  std::vector<int> S(L + 1);
  for (int i = 0; i <= L; ++i) S[i] = c.height >> i;
  h->steps.resize(T);
  for (int t = 0; t < T; ++t) {
    Step& s = h->steps[t];
    s.t = t;
    s.p_begin = h->arena_numel;
    // --- recognition net: phi/inference_step_t (sequential_vae.py:1573-1630)
    {
      Namer nm{"phi/inference_step_" + std::to_string(t)};
      s.inf.resize(2 * (L - 1));
      s.heads.resize(L - 1);
      int col = 0;
      for (int l = 0; l < L - 1; ++l) {
        Block& a = s.inf[2 * l];
        a.g = conv_geom(S[l], S[l], F[l], F[l + 1], 2);
        add_block_params(h, nm, a, 0, {4, 4, F[l], F[l + 1]}, F[l + 1], t, 0);
        Block& b = s.inf[2 * l + 1];
        b.g = conv_geom(S[l + 1], S[l + 1], F[l + 1], F[l + 1], 1);
        add_block_params(h, nm, b, 0, {4, 4, F[l + 1], F[l + 1]}, F[l + 1], t, 0);
        int nflat = S[l + 1] * S[l + 1] * F[l + 1];
        for (int k = 0; k < 2; ++k) {
          std::string fc = nm.next_fc();
          Head hd;
          hd.n = c.latent_dims[l]; hd.col = col; hd.is_sd = k;
          hd.w = add_param(h, fc + "/weights", {nflat, hd.n}, t, SVAE_PF_XAVIER);
          hd.b = add_param(h, fc + "/biases", {hd.n}, t, 0);
          s.heads[l].push_back(hd);
        }
        col += c.latent_dims[l];
      }
      // dead branch (:1602-1605): variables exist, never read (Q3)
      {
        std::string cv = nm.next_conv(), bn = nm.next_bn();
        add_param(h, cv + "/weights", {4, 4, F[L - 1], F[L - 1]}, t, SVAE_PF_DEAD);
        add_param(h, cv + "/biases", {F[L - 1]}, t, SVAE_PF_DEAD | SVAE_PF_INERT);
        add_param(h, bn + "/beta", {F[L - 1]}, t, SVAE_PF_DEAD);
        std::string fc = nm.next_fc(), bn2 = nm.next_bn();
        add_param(h, fc + "/weights", {S[L] * S[L] * F[L - 1], F[L]}, t, SVAE_PF_DEAD);
        add_param(h, fc + "/biases", {F[L]}, t, SVAE_PF_DEAD | SVAE_PF_INERT);
        add_param(h, bn2 + "/beta", {F[L]}, t, SVAE_PF_DEAD);
      }
      // last heads read the stale level L-2 features (:1607,1609)
      int nflat = S[L - 1] * S[L - 1] * F[L - 1];
      for (int k = 0; k < 2; ++k) {
        std::string fc = nm.next_fc();
        Head hd;
        hd.n = c.latent_dims[L - 1]; hd.col = col; hd.is_sd = k;
        hd.w = add_param(h, fc + "/weights", {nflat, hd.n}, t, SVAE_PF_XAVIER);
        hd.b = add_param(h, fc + "/biases", {hd.n}, t, 0);
        s.heads[L - 2].push_back(hd);
      }
    }
    s.p_rec_end = h->arena_numel;
    // --- chain encoder: theta/generative_encoder_step_t (:1757-1777), t >= 1
    if (t > 0) {
      Namer nm{"theta/generative_encoder_step_" + std::to_string(t)};
      s.enc.resize(2 * (L - 1) + 1);
      for (int l = 0; l < L - 1; ++l) {
        Block& a = s.enc[2 * l];
        a.g = conv_geom(S[l], S[l], F[l], F[l + 1], 2);
        add_block_params(h, nm, a, 0, {4, 4, F[l], F[l + 1]}, F[l + 1], t, SVAE_PF_THETA);
        Block& b = s.enc[2 * l + 1];
        b.g = conv_geom(S[l + 1], S[l + 1], F[l + 1], F[l + 1], 1);
        add_block_params(h, nm, b, 0, {4, 4, F[l + 1], F[l + 1]}, F[l + 1], t, SVAE_PF_THETA);
      }
      Block& cl = s.enc[2 * (L - 1)];
      cl.g = conv_geom(S[L - 1], S[L - 1], F[L - 1], F[L - 1], 2);
      add_block_params(h, nm, cl, 0, {4, 4, F[L - 1], F[L - 1]}, F[L - 1], t, SVAE_PF_THETA);
      s.encfc.g = fc_geom(S[L] * S[L] * F[L - 1], F[L]);
      add_block_params(h, nm, s.encfc, 2, {S[L] * S[L] * F[L - 1], F[L]}, F[L], t, SVAE_PF_THETA);
    }
    // --- decoder: theta/generative_step_t (:1683-1729)
    {
      Namer nm{"theta/generative_step_" + std::to_string(t)};
      s.lat.resize(L);
      for (int i = 0; i < L - 1; ++i) {
        int n = S[i + 1] * S[i + 1] * F[i + 1];
        s.lat[i].g = fc_geom(c.latent_dims[i], n);
        add_block_params(h, nm, s.lat[i], 2, {c.latent_dims[i], n}, n, t, SVAE_PF_THETA);
      }
      s.lat[L - 1].g = fc_geom(c.latent_dims[L - 1], F[L + 1]);
      add_block_params(h, nm, s.lat[L - 1], 2, {c.latent_dims[L - 1], F[L + 1]}, F[L + 1], t, SVAE_PF_THETA);
      s.ncc = t > 0 ? F[L] + F[L + 1] : F[L + 1];
      int nfc = S[L] * S[L] * F[L];
      s.decfc.g = fc_geom(s.ncc, nfc);
      add_block_params(h, nm, s.decfc, 2, {s.ncc, nfc}, nfc, t, SVAE_PF_THETA);
      s.ta.resize(L - 1);
      s.tb.resize(L - 1);
      int cin = F[L];
      for (int l = L - 2; l >= 0; --l) {
        s.ta[l].g = deconv_geom(S[l + 2], S[l + 2], cin, F[l + 1], 2);
        add_block_params(h, nm, s.ta[l], 1, {4, 4, F[l + 1], cin}, F[l + 1], t, SVAE_PF_THETA);
        s.tb[l].g = deconv_geom(S[l + 1], S[l + 1], 2 * F[l + 1], F[l + 1], 1);
        add_block_params(h, nm, s.tb[l], 1, {4, 4, F[l + 1], 2 * F[l + 1]}, F[l + 1], t, SVAE_PF_THETA);
        cin = F[l + 1];
      }
      std::string o = nm.next_convt();
      s.w_out = add_param(h, o + "/weights", {4, 4, C, F[1]}, t, SVAE_PF_THETA | SVAE_PF_XAVIER);
      s.b_out = add_param(h, o + "/biases", {C}, t, SVAE_PF_THETA);
      s.g_out = deconv_geom(S[1], S[1], F[1], C, 2);
      s.outb.g = s.g_out; s.outb.w = s.w_out;
      s.w_gate = s.b_gate = -1;
      if (t > 0) {
        std::string r = nm.next_convt();
        s.w_gate = add_param(h, r + "/weights", {4, 4, 1, F[1]}, t, SVAE_PF_THETA | SVAE_PF_XAVIER);
        s.b_gate = add_param(h, r + "/biases", {1}, t, SVAE_PF_THETA);
        s.g_gate = deconv_geom(S[1], S[1], F[1], 1, 2);
        s.gateb.g = s.g_gate; s.gateb.w = s.w_gate;
        s.g_og = deconv_geom(S[1], S[1], F[1], C + 1, 2);
        s.ogb.g = s.g_og; s.ogb.w = s.w_out;
      }
    }
    s.p_end = h->arena_numel;
  }
}

// The caller-visible variable table (see svae_handle::pub).  Shared scopes follow the reference's scope names:
// "phi/inference_network" (sequential_vae.py:1573-1577), "theta/generative_encoder_network" (:1757-1761) and
// "theta/generative_network" for steps t >= 1 (:1683-1687; step 0 has no chain input and keeps "theta/generative_step_0").
// Entries appear in order of first use, which is the order TF creates the variables in (construct_network, :934-975).
void build_pub(svae_handle* h) {
  const bool st = h->cfg.share_theta_weights != 0, sp = h->cfg.share_phi_weights != 0;
  h->pub.clear();
  h->tie_runs.clear();
  std::map<std::string, int> index;
  for (int i = 0; i < (int)h->params.size(); ++i) {
    const Param& p = h->params[i];
    std::string name = p.name;
    auto rescoped = [&](const char* per_step, const char* shared) {
      const std::string pre = std::string(per_step) + std::to_string(p.step) + "/";
      if (name.compare(0, pre.size(), pre) == 0) { name = std::string(shared) + "/" + name.substr(pre.size()); return true; }
      return false;
    };
    if (sp) rescoped("phi/inference_step_", "phi/inference_network");
    if (st && !rescoped("theta/generative_encoder_step_", "theta/generative_encoder_network") && p.step >= 1)
      rescoped("theta/generative_step_", "theta/generative_network");
    auto it = index.find(name);
    if (it == index.end()) {
      index[name] = (int)h->pub.size();
      h->pub.push_back(PubParam{name, {i}});
    } else {
      h->pub[it->second].members.push_back(i);
    }
  }
  // contiguous runs of tied slices: consecutive variables whose members are laid out back to back in every step
  for (const PubParam& q : h->pub) {
    if (q.members.size() < 2) continue;
    const Param& p0 = h->params[q.members[0]];
    const int64_t padded = (p0.numel + 3) / 4 * 4;
    bool extend = !h->tie_runs.empty() && h->tie_runs.back().members == (int)q.members.size();
    if (extend) {
      const TieRun& r = h->tie_runs.back();
      for (size_t m = 0; m < q.members.size(); ++m)
        if (r.off[m] + r.n != h->params[q.members[m]].offset) extend = false;
    }
    if (extend) {
      h->tie_runs.back().n += padded;
    } else {
      TieRun r{};
      r.n = padded;
      r.members = (int)q.members.size();
      for (size_t m = 0; m < q.members.size(); ++m) r.off[m] = h->params[q.members[m]].offset;
      h->tie_runs.push_back(r);
    }
  }
  h->tied = !h->tie_runs.empty();
}

// ---- activation arena ----------------------------------------------------------------------------------------------
void place_block(Block& b, Arena& act, Arena& zf, Arena& zb, int64_t maxB, bool plane2d, size_t& max_y, bool y_in_zf = false) {
  if (plane2d) { b.rpi = 1; b.feats = b.g.Cout; } else { b.rpi = b.g.Hout * b.g.Wout; b.feats = b.g.Cout; }
  size_t n = (size_t)maxB * b.rpi * b.feats;
  // y_in_zf: the split-K fully-connected kernels accumulate their partial sums with atomics; an output that lives in the
  // forward-zeroed arena (one memset at the start of the step) needs no memset node in front of every launch on the chain
  b.y = y_in_zf ? zf.get<float>(n) : act.get<float>(n);
  b.y_zeroed = y_in_zf;
  b.stats = zf.get<double>(2 * (size_t)b.feats);
  b.fwd_bar = zf.get<unsigned>(4);
  b.S = zb.get<double>(2 * (size_t)b.feats + 1);   // + the grid-barrier counter of the fused backward kernel
  if (n > max_y) max_y = n;
}
FeatView fv4(float* p, int ld, int coff, int inner) { return FeatView{p, ld, coff, inner, 1}; }

void build_buffers(svae_handle* h, Arena& act, Arena& zf, Arena& zb, Arena& gr, Arena& bfa, size_t& max_y) {
  const svae_config& c = h->cfg;
  const int L = h->L, T = h->T, C = h->C, Z = h->Z;
  const int* F = c.filter_sizes;
  const int64_t B = c.max_batch;
  std::vector<int> S(L + 1);
  for (int i = 0; i <= L; ++i) S[i] = c.height >> i;
  for (int t = 0; t < T; ++t) {
    Step& s = h->steps[t];
    const bool share = (t >= h->act_sets) && t >= 2;  // steps >= 2 alias step 1's activations when not training
    if (share) {
      const Step& r = h->steps[1];
      auto alias = [](Block& b, const Block& q) {
        b.rpi = q.rpi; b.feats = q.feats; b.act = q.act; b.y = q.y; b.stats = q.stats; b.S = q.S; b.out = q.out; b.res = q.res; b.fwd_bar = q.fwd_bar;
        b.in_bf = q.in_bf; b.tc2_fwd = q.tc2_fwd; b.out_bf = q.out_bf;
      };
      for (size_t i = 0; i < s.inf.size(); ++i) alias(s.inf[i], r.inf[i]);
      for (size_t i = 0; i < s.enc.size(); ++i) alias(s.enc[i], r.enc[i]);
      alias(s.encfc, r.encfc);
      for (size_t i = 0; i < s.lat.size(); ++i) alias(s.lat[i], r.lat[i]);
      alias(s.decfc, r.decfc);
      for (size_t i = 0; i < s.ta.size(); ++i) alias(s.ta[i], r.ta[i]);
      for (size_t i = 0; i < s.tb.size(); ++i) alias(s.tb[i], r.tb[i]);
      s.dcat = r.dcat; s.cc = r.cc; s.u = r.u; s.xt = r.xt; s.lat_mom = r.lat_mom;
      s.mu_pre = r.mu_pre; s.sd_pre = r.sd_pre; s.mu = r.mu; s.sd = r.sd; s.z = r.z; s.eps = r.eps;
      continue;
    }
    // recognition
    for (int k = 0; k < 2 * (L - 1); ++k) {
      Block& b = s.inf[k];
      place_block(b, act, zf, zb, B, false, max_y);
      b.act = ACT_LRELU;
      float* a = act.get<float>((size_t)B * b.rpi * b.feats);
      b.out = fv4(a, b.feats, 0, b.feats);
    }
    s.mu_pre = act.get<float>(2 * (size_t)B * Z); s.sd_pre = s.mu_pre + (size_t)B * Z;   // one block: zeroed together
    s.mu = act.get<float>((size_t)B * Z); s.sd = act.get<float>((size_t)B * Z);
    s.z = act.get<float>((size_t)B * Z); s.eps = act.get<float>((size_t)B * Z);
    // decoder concat buffers first (encoder / projections write into them)
    s.dcat.resize(L - 1);
    for (int l = 0; l < L - 1; ++l) s.dcat[l] = act.get<float>((size_t)B * S[l + 1] * S[l + 1] * 2 * F[l + 1]);
    s.cc = act.get<float>((size_t)B * s.ncc);
    // encoder
    if (t > 0) {
      for (int k = 0; k <= 2 * (L - 1); ++k) {
        Block& b = s.enc[k];
        place_block(b, act, zf, zb, B, false, max_y);
        b.act = ACT_LRELU;
        float* a = act.get<float>((size_t)B * b.rpi * b.feats);
        b.out = fv4(a, b.feats, 0, b.feats);
      }
      place_block(s.encfc, act, zf, zb, B, true, max_y, c.train_capacity != 0);
      s.encfc.act = ACT_LRELU;
      s.encfc.out = fv4(s.cc, s.ncc, 0, F[L]);           // e[L] is the first window of the concat (:1696-1697,1834)
    }
    // latent projections (split_latent, :1796-1806)
    for (int i = 0; i < L; ++i) {
      Block& b = s.lat[i];
      place_block(b, act, zf, zb, B, true, max_y);
      b.act = ACT_LRELU;
      if (i < L - 1) b.out = FeatView{s.dcat[i], 2 * F[i + 1], F[i + 1], F[i + 1], S[i + 1] * S[i + 1]};  // second window of the level concat (:1716)
      else b.out = fv4(s.cc, s.ncc, t > 0 ? F[L] : 0, F[L + 1]);
    }
    // dec.fc (:1704-1705)
    place_block(s.decfc, act, zf, zb, B, true, max_y, c.train_capacity != 0);
    s.decfc.act = ACT_LRELU;
    {
      float* a = act.get<float>((size_t)B * s.decfc.feats);
      s.decfc.out = fv4(a, s.decfc.feats, 0, s.decfc.feats);
    }
    for (int l = L - 2; l >= 0; --l) {
      Block& a = s.ta[l];
      place_block(a, act, zf, zb, B, false, max_y);
      a.act = ACT_RELU;                                   // relu after the shortcut add (:1713-1714)
      a.out = fv4(s.dcat[l], 2 * F[l + 1], 0, F[l + 1]);
      if (t > 0) a.res = s.enc[2 * l + 1].out;            // e[l+1]
      Block& b = s.tb[l];
      place_block(b, act, zf, zb, B, false, max_y);
      b.act = ACT_RELU;
      float* cbuf = act.get<float>((size_t)B * b.rpi * b.feats);
      b.out = fv4(cbuf, b.feats, 0, b.feats);
    }
    s.u = act.get<float>((size_t)B * h->D * h->D * (C + 1));
    s.xt = act.get<float>((size_t)B * h->D * h->D * C);
    s.lat_mom = act.get<double>((size_t)L * 64);
  }
  // ---- bf16 planar copies for the TMA-fed kernels: every tensor a tcgen05 conv reads gets one, in its consumer's layout
  const bool tc2 = h->use_tc2 && c.operand_dtype == SVAE_OPERAND_BF16;
  auto want = [&](Geom g) { g.B = (int)B; return tc2 && g.KH == 4 && tc_supported(g) && tc2_supported(g); };
  auto mk = [&](const Geom& consumer, int H, int W, int Cc) {
    BfAct a = bf_act_describe(tc2_input_kind(consumer), (int)B, H, W, Cc);
    a.p = bfa.get<__nv_bfloat16>(bf_act_bytes(a) / 2);
    return a;
  };
  // consumer `cb` reads the tensor produced by block `pb` (or by nobody: pb == nullptr) through a fresh copy
  auto feed = [&](Block& cb, Block* pb) {
    if (!want(cb.g)) return;
    cb.in_bf = mk(cb.g, cb.g.Hin, cb.g.Win, cb.g.Cin);
    cb.tc2_fwd = true;
    if (pb) pb->out_bf = BfDst{cb.in_bf, 0, 0, 0};
  };
  if (tc2) {
    Block& first = h->steps[0].inf[0];
    if (want(first.g)) h->x_bf = mk(first.g, first.g.Hin, first.g.Win, first.g.Cin);
    for (int t = 0; t < T; ++t) {
      Step& s = h->steps[t];
      if ((t >= h->act_sets) && t >= 2) continue;   // aliased onto step 1 above
      for (int k = 0; k < (int)s.inf.size(); ++k) {
        if (k == 0) { if (want(s.inf[0].g)) { s.inf[0].in_bf = h->x_bf; s.inf[0].tc2_fwd = true; } }
        else feed(s.inf[k], &s.inf[k - 1]);
      }
      for (int k = 1; k < (int)s.enc.size(); ++k) feed(s.enc[k], &s.enc[k - 1]);
      // x_t -> chain encoder of step t+1 (same geometry in every step >= 1)
      if (T > 1 && want(h->steps[1].enc[0].g)) {
        const Geom& eg = h->steps[1].enc[0].g;
        s.xt_bf = mk(eg, eg.Hin, eg.Win, eg.Cin);
      }
      feed(s.ta[L - 2], &s.decfc);
      if (s.ta[L - 2].tc2_fwd) { s.decfc.out_bf.inner = F[L]; s.decfc.out_bf.ppr = S[L] * S[L]; }   // [B, S_L*S_L*F_L] -> [B,S_L,S_L,F_L]
      for (int l = L - 2; l >= 0; --l) {
        feed(s.tb[l], &s.ta[l]);                                                  // dcat_l: [relu(bn(ta)+e), P_l]
        if (s.tb[l].tc2_fwd) s.lat[l].out_bf = BfDst{s.tb[l].in_bf, F[l + 1], 0, 0};
        if (l > 0) feed(s.ta[l - 1], &s.tb[l]);
      }
      // both output deconvs read c_0 = tb[0].out through one copy
      if (want(s.outb.g) && (t == 0 || want(s.gateb.g))) {
        feed(s.outb, &s.tb[0]);
        if (t > 0) { s.gateb.in_bf = s.outb.in_bf; s.gateb.tc2_fwd = true; }
        static const bool merge = !(getenv("SVAE_MERGE_OG") && getenv("SVAE_MERGE_OG")[0] == '0');
        if (t > 0 && merge && want(s.ogb.g)) { s.ogb.in_bf = s.outb.in_bf; s.ogb.tc2_fwd = true; s.merge_og = !c.train_capacity; }
      }
    }
    for (int t = 1; t < T; ++t) {   // enc[0] of step t reads the copy of x_{t-1}
      Step& s = h->steps[t];
      Step& pr = h->steps[t - 1];
      if ((t >= h->act_sets) && t >= 2) {   // forward-only handles: steps >= 2 share step 1's buffers, copies included
        const Step& r = h->steps[1];
        auto alias_bf = [](Block& b, const Block& q) { b.in_bf = q.in_bf; b.tc2_fwd = q.tc2_fwd; b.out_bf = q.out_bf; };
        for (size_t i = 0; i < s.inf.size(); ++i) alias_bf(s.inf[i], r.inf[i]);
        for (size_t i = 0; i < s.enc.size(); ++i) alias_bf(s.enc[i], r.enc[i]);
        alias_bf(s.encfc, r.encfc); alias_bf(s.decfc, r.decfc);
        for (size_t i = 0; i < s.lat.size(); ++i) alias_bf(s.lat[i], r.lat[i]);
        for (size_t i = 0; i < s.ta.size(); ++i) { alias_bf(s.ta[i], r.ta[i]); alias_bf(s.tb[i], r.tb[i]); }
        alias_bf(s.outb, r.outb); alias_bf(s.gateb, r.gateb); alias_bf(s.ogb, r.ogb); s.merge_og = r.merge_og;
        s.xt_bf = r.xt_bf;
      }
      if (want(s.enc[0].g)) { s.enc[0].in_bf = pr.xt_bf; s.enc[0].tc2_fwd = true; }
    }
  }
  // gradient scratch: two sets (step parity) of one step's worth; every block has its own d(pre-BN output) buffer so
  // that its weight gradient can run on a side stream while the chain moves on
  if (c.train_capacity) {
    const int NI = 2 * (L - 1), NE = 2 * (L - 1) + 1;
    h->n_dy_slots = NI + NE + 2 + L + 2 * (L - 1);
    h->dy_slot_elems.assign(h->n_dy_slots, 0);
    auto slot = [&](Block& b, int id) {
      b.dy_slot = id;
      const size_t n = (size_t)B * b.rpi * b.feats;
      if (n > h->dy_slot_elems[id]) h->dy_slot_elems[id] = n;
    };
    for (int t = 0; t < T; ++t) {
      Step& s = h->steps[t];
      for (int k = 0; k < (int)s.inf.size(); ++k) slot(s.inf[k], k);
      for (int k = 0; k < (int)s.enc.size(); ++k) slot(s.enc[k], NI + k);
      if (t > 0) slot(s.encfc, NI + NE - 1 + 1);
      for (int i2 = 0; i2 < L; ++i2) slot(s.lat[i2], NI + NE + 1 + i2);
      slot(s.decfc, NI + NE + 1 + L);
      for (int l = 0; l < L - 1; ++l) { slot(s.ta[l], NI + NE + 2 + L + l); slot(s.tb[l], NI + NE + 2 + L + (L - 1) + l); }
    }
    const Step& s1 = h->steps[T > 1 ? 1 : 0];
    {
      int want = svae_handle::MAX_GS;
      const char* e = getenv("SVAE_GRAD_SETS");
      if (e) want = atoi(e);
      if (want < 2) want = 2;
      if (want > svae_handle::MAX_GS) want = svae_handle::MAX_GS;
      h->n_gs = T < want ? (T < 2 ? 2 : T) : want;
    }
    // Every chain step has its own scratch set: the buffers that are accumulated into with atomics (d_z by the latent
    // projections, d_cc / d_e[last] by the split-K fully-connected input gradients) live in the backward-zeroed arena - one
    // memset at the start of the backward instead of five memset nodes per chain step in front of kernels of the chain.
    h->bwd_prezero = h->n_gs >= T;
    Arena& za = h->bwd_prezero ? zb : gr;
    for (int p = 0; p < h->n_gs; ++p) {
      GradSet& g = h->gs[p];
      g.d_c.resize(L - 1); g.d_dcat.resize(L - 1);
      for (int l = 0; l < L - 1; ++l) {
        g.d_c[l] = gr.get<float>((size_t)B * S[l + 1] * S[l + 1] * F[l + 1]);
        g.d_dcat[l] = gr.get<float>((size_t)B * S[l + 1] * S[l + 1] * 2 * F[l + 1]);
      }
      g.d_fca = gr.get<float>((size_t)B * S[L] * S[L] * F[L]);
      g.d_cc = za.get<float>((size_t)B * (F[L] + F[L + 1]));
      g.d_z = za.get<float>((size_t)B * Z);
      g.d_mu_pre = gr.get<float>((size_t)B * Z);
      g.d_sd_pre = gr.get<float>((size_t)B * Z);
      g.d_u = gr.get<float>((size_t)B * h->D * h->D * (C + 1));
      g.d_e.assign(2 * (L - 1) + 1, nullptr);
      g.d_inf.assign(2 * (L - 1), nullptr);
      for (int k = 0; k < 2 * (L - 1); ++k) {
        const Block& b = h->steps[0].inf[k];
        g.d_inf[k] = gr.get<float>((size_t)B * b.rpi * b.feats);
      }
      if (T > 1)
        for (int k = 0; k <= 2 * (L - 1); ++k) {
          const Block& b = s1.enc[k];
          g.d_e[k] = (k == 2 * (L - 1) ? za : gr).get<float>((size_t)B * b.rpi * b.feats);
        }
      g.dy.assign(h->n_dy_slots, nullptr);
      for (int k = 0; k < h->n_dy_slots; ++k)
        if (h->dy_slot_elems[k] > 0) g.dy[k] = gr.get<float>(h->dy_slot_elems[k]);
      // bf16 copies of dy for the TMA-fed input-gradient kernels (one per slot: the geometry is the same in every step)
      g.dy_bf.assign(h->n_dy_slots, BfAct{});
      auto wgeom = [&](const Block& b) {   // conv-gather geometry of the weight gradient (tc_wgrad convention)
        Geom wg = b.g;
        if (b.g.mode == 1) { wg = dgrad_geom(b.g); wg.mode = 0; }
        wg.B = (int)B;
        return wg;
      };
      auto dyfeed = [&](Block& b, bool needs_din) {
        if (b.dy_slot < 0) return;
        Geom dg = dgrad_geom(b.g);
        const bool for_dgrad = needs_din && want(dg);
        // the weight gradient reads dy in the same layout as the input gradient does
        const bool for_wgrad = tc2 && b.tc2_fwd && b.g.KH == 4 && tc_wgrad_supported(b.g) && tc2_wgrad_supported(wgeom(b));
        if (!for_dgrad && !for_wgrad) return;
        if (g.dy_bf[b.dy_slot].Cpad == 0) { Geom dk = dg; g.dy_bf[b.dy_slot] = mk(dk, dg.Hin, dg.Win, dg.Cin); }
        if (for_dgrad) b.tc2_dgrad = true;
        if (for_wgrad) b.tc2_wgrad = true;
      };
      for (int t = 0; t < T; ++t) {
        Step& s = h->steps[t];
        for (int k = 0; k < (int)s.inf.size(); ++k) dyfeed(s.inf[k], k > 0);
        for (int k = 0; k < (int)s.enc.size(); ++k) dyfeed(s.enc[k], true);
        for (int l = 0; l < L - 1; ++l) { dyfeed(s.ta[l], true); dyfeed(s.tb[l], true); }
      }
      {
        Step& s1b = h->steps[T > 1 ? 1 : 0];
        Geom dgo = dgrad_geom(s1b.g_out);
        if (want(dgo)) g.du_out_bf = mk(dgo, dgo.Hin, dgo.Win, dgo.Cin);
        if (T > 1) { Geom dgg = dgrad_geom(s1b.g_gate); if (want(dgg) && g.du_out_bf.Cpad) g.du_gate_bf = mk(dgg, dgg.Hin, dgg.Win, dgg.Cin); }
        for (int t = 0; t < T; ++t) {   // output deconvs: X = d_u copy, dY = copy of c_0
          Step& s = h->steps[t];
          Geom wo = dgrad_geom(s.g_out); wo.mode = 0; wo.B = (int)B;
          s.outb.tc2_wgrad = g.du_out_bf.Cpad && s.outb.tc2_fwd && tc_wgrad_supported(s.outb.g) && tc2_wgrad_supported(wo);
          if (t > 0) {
            Geom wgt = dgrad_geom(s.g_gate); wgt.mode = 0; wgt.B = (int)B;
            s.gateb.tc2_wgrad = g.du_gate_bf.Cpad && s.gateb.tc2_fwd && tc_wgrad_supported(s.gateb.g) && tc2_wgrad_supported(wgt);
            // merged form: the gate's gradient is channel C of the d_u copy (Cpad = 16 holds it)
            Geom d4 = dgrad_geom(s.g_og), w4 = dgrad_geom(s.g_og); w4.mode = 0; w4.B = (int)B;
            s.ogb.tc2_dgrad = s.ogb.tc2_fwd && g.du_out_bf.Cpad >= 16 && want(d4);
            s.ogb.tc2_wgrad = s.ogb.tc2_dgrad && s.outb.tc2_wgrad && s.gateb.tc2_wgrad && tc_wgrad_supported(s.ogb.g) && tc2_wgrad_supported(w4);
            s.merge_og = s.ogb.tc2_fwd && s.ogb.tc2_dgrad && s.ogb.tc2_wgrad;
          }
        }
      }
    }
    {
      static const bool wide = !(getenv("SVAE_GX4") && getenv("SVAE_GX4")[0] == '0');
      h->gxld = (wide && tc2 && C == 3 && T > 1 && h->steps[1].enc[0].tc2_dgrad) ? 4 : C;
    }
    h->gx[0] = gr.get<float>((size_t)B * h->D * h->D * h->gxld);
    h->gx[1] = gr.get<float>((size_t)B * h->D * h->D * h->gxld);
  }
  // ---- which blocks may skip the fp32 copy of their activated output (see Block::skip_f32) -----------------------------
  {
    const bool train = c.train_capacity != 0;
    static const bool enabled = !(getenv("SVAE_SKIP_F32") && getenv("SVAE_SKIP_F32")[0] == '0');
    auto bf_only = [&](const Block& consumer) {   // this consumer (and its weight gradient) never touches the fp32 tensor
      return enabled && tc2 && consumer.tc2_fwd && consumer.in_bf.p != nullptr && (!train || consumer.tc2_wgrad);
    };
    for (int t = 0; t < T; ++t) {
      Step& s = h->steps[t];
      if ((t >= h->act_sets) && t >= 2) {   // aliased onto step 1: same decisions
        const Step& r = h->steps[1];
        for (size_t i = 0; i < s.inf.size(); ++i) s.inf[i].skip_f32 = r.inf[i].skip_f32;
        for (size_t i = 0; i < s.enc.size(); ++i) s.enc[i].skip_f32 = r.enc[i].skip_f32;
        for (size_t i = 0; i < s.lat.size(); ++i) s.lat[i].skip_f32 = r.lat[i].skip_f32;
        s.decfc.skip_f32 = r.decfc.skip_f32;
        for (size_t i = 0; i < s.ta.size(); ++i) { s.ta[i].skip_f32 = r.ta[i].skip_f32; s.tb[i].skip_f32 = r.tb[i].skip_f32; }
        continue;
      }
      // recognition net: even blocks feed the next conv only; odd blocks also feed the heads (fp32)
      for (int k = 0; k + 1 < (int)s.inf.size(); k += 2) s.inf[k].skip_f32 = bf_only(s.inf[k + 1]);
      // chain encoder: even blocks feed the next conv only; odd blocks are the decoder's shortcuts, the last one feeds enc.fc
      for (int k = 0; k + 1 < (int)s.enc.size(); k += 2) s.enc[k].skip_f32 = bf_only(s.enc[k + 1]);
      // decoder: dec.fc -> ta[L-2]; ta[l], P_l -> tb[l]; tb[l] -> ta[l-1]; tb[0] -> the two output deconvs
      s.decfc.skip_f32 = bf_only(s.ta[L - 2]);
      for (int l = L - 2; l >= 0; --l) {
        s.ta[l].skip_f32 = bf_only(s.tb[l]);
        s.lat[l].skip_f32 = bf_only(s.tb[l]) && s.lat[l].out_bf.a.p != nullptr;
        if (l > 0) s.tb[l].skip_f32 = bf_only(s.ta[l - 1]);
      }
      s.tb[0].skip_f32 = bf_only(s.outb) && (t == 0 || bf_only(s.gateb));
      for (Block* b : {&s.decfc}) if (b->out_bf.a.p == nullptr) b->skip_f32 = false;
      for (Block& b : s.inf) if (b.out_bf.a.p == nullptr) b.skip_f32 = false;
      for (Block& b : s.enc) if (b.out_bf.a.p == nullptr) b.skip_f32 = false;
      for (Block& b : s.ta) if (b.out_bf.a.p == nullptr) b.skip_f32 = false;
      for (Block& b : s.tb) if (b.out_bf.a.p == nullptr) b.skip_f32 = false;
    }
    // consumers of a skipped fp32 tensor (debug probe: which representation of its input a block can be shown)
    for (int t = 0; t < T; ++t) {
      Step& s = h->steps[t];
      for (int k = 0; k + 1 < (int)s.inf.size(); ++k) s.inf[k + 1].in_f32_valid = !s.inf[k].skip_f32;
      for (int k = 0; k + 1 < (int)s.enc.size(); ++k) s.enc[k + 1].in_f32_valid = !s.enc[k].skip_f32;
      s.ta[L - 2].in_f32_valid = !s.decfc.skip_f32;
      for (int l = L - 2; l >= 0; --l) {
        s.tb[l].in_f32_valid = !s.ta[l].skip_f32 && !s.lat[l].skip_f32;
        if (l > 0) s.ta[l - 1].in_f32_valid = !s.tb[l].skip_f32;
      }
      s.outb.in_f32_valid = s.gateb.in_f32_valid = !s.tb[0].skip_f32;
    }
  }
  // io staging (host-buffer entry points)
  h->in_x = act.get<float>((size_t)B * h->D * h->D * C);
  h->in_tgt = act.get<float>((size_t)B * h->D * h->D * C);
  h->in_eps = act.get<float>((size_t)T * B * Z);
  h->gen_prev = act.get<float>((size_t)B * h->D * h->D * C);
  if (c.add_noise_to_chain)   // one sample buffer per chain step, also on forward-only handles (svae_read_chain_samples_host)
    for (int t = 0; t < T; ++t) h->steps[t].xs = act.get<float>((size_t)B * h->D * h->D * C);
}

// ---- contraction dispatch ------------------------------------------------------------------------------------------
int contract(svae_handle* h, Geom g, int B, View in, const float* w, const void* w_packed, bool use_tc, View out,
             double* stats) {
  g.B = B;
  LaunchCtx lc = h->lc();
  if (use_tc && g.KH == 1) {
    // fully connected: forward (w_out_major 0: w = [Cin, Cout]) or input gradient (w_out_major 1: w = [Cout, Cin])
    int r = g.w_out_major == 0
                ? tc_fc(lc, 0, in.p + in.coff, in.ld, w, g.Cout, out.p + out.coff, out.ld, B, g.Cin, g.Cout, g.accumulate)
                : tc_fc(lc, 1, in.p + in.coff, in.ld, w, g.Cin, out.p + out.coff, out.ld, B, g.Cout, g.Cin, g.accumulate);
    if (r == 0 && stats != nullptr) {
      if (out.ld != g.Cout || out.coff != 0) { svae_global_error() = "fc statistics need a dense output"; return -1; }
      r = col_stats(lc, out.p, B, g.Cout, stats);
    }
    return r;
  }
  if (use_tc) return tc_gather_gemm(lc, g, in, w_packed, out, stats);
  return simt_gather_gemm(lc, g, in, w, out, stats);
}

// ---- streams: the main stream carries the chain; side streams carry work that is off the chain's critical path ------
struct OnStream {   // RAII: route launches to `s` for the lifetime of the scope
  svae_handle* h; cudaStream_t prev;
  OnStream(svae_handle* h_, cudaStream_t s) : h(h_), prev(h_->cur) { h->cur = s; }
  ~OnStream() { h->cur = prev; }
};
cudaStream_t cur_stream(svae_handle* h) { return h->cur ? h->cur : h->stream; }
cudaEvent_t next_event(svae_handle* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}
void tl_mark(svae_handle* h, cudaStream_t s, const char* what, int t) {
  if (!h->timeline || h->capturing) return;
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  h->marks.push_back(svae_handle::Mark{e, what, t});
}
// work enqueued on `to` after this call waits for everything enqueued on `from` so far
int link(svae_handle* h, cudaStream_t from, cudaStream_t to) {
  if (from == to) return 0;
  cudaEvent_t e = next_event(h);
  H_CUDA(cudaEventRecord(e, from));
  H_CUDA(cudaStreamWaitEvent(to, e, 0));
  return 0;
}
// forked execution is used for training-capacity handles outside profiling (the profiler wants serialised kernels)
bool forked(const svae_handle* h) { return h->use_streams && h->cfg.train_capacity && !h->prof.enabled && h->side[0] != nullptr; }

// contraction through the TMA-fed kernel when the input has a bf16 planar copy, else the SIMT-staged / fp32 kernels
int contract_bf(svae_handle* h, Geom g, int B, bool tc2, const BfAct& in_bf, View in, const float* w, const void* w_packed,
                bool use_tc, View out, double* stats, const BnBwdFuse* fuse = nullptr, int w_tile_width = 128) {
  if (tc2) {
    g.B = B;
    LaunchCtx lc = h->lc();
    return tc2_gather_gemm(lc, g, in_bf, 0, w_packed, out, stats, fuse, w_tile_width);
  }
  if (fuse != nullptr) { svae_global_error() = "fused batch-norm backward requested on a non-tc2 contraction"; return -1; }
  if (w_tile_width != 128 && use_tc) { svae_global_error() = "weights packed for the TMA-fed kernel reached the SIMT-staged one"; return -1; }
  return contract(h, g, B, in, w, w_packed, use_tc, out, stats);
}

// Describe pass 1 of block `up`'s batch-norm backward for fusion into the input-gradient kernel (geometry dg, TMA-fed path)
// that writes `din`, whose first up.feats channels are dL/d(up's activated output).  false: not fusable, run it separately.
bool make_fuse(svae_handle* h, Block* up, int B, float* dres, int dres_acc, Geom dg, View din, BnBwdFuse& fz) {
  if (up == nullptr || !h->use_fuse || up->g.KH != 4 || up->rpi <= 1) return false;   // 4-D batch norm only
  if (up->res.p != nullptr && !(up->res.ppr == 1 && up->res.inner == up->feats && up->res.ld % 4 == 0 && up->res.coff % 4 == 0 &&
                                ((uintptr_t)up->res.p & 15) == 0))
    return false;
  dg.B = B;
  if (up->res.p != nullptr && dg.accumulate) return false;   // the fused epilogue carries one auxiliary operand per element
  if ((int64_t)dg.Hout * dg.Wout != up->rpi || !tc2_fuse_supported(dg, din, up->feats)) return false;
  if (((uintptr_t)up->y & 15) != 0 || (dres != nullptr && ((uintptr_t)dres & 15) != 0)) return false;
  fz = BnBwdFuse{up->y, up->stats, h->pw(up->beta), up->res.p, up->res.ld, up->res.coff, dres, dres_acc, up->S, up->feats, up->act,
                 (long long)B * up->rpi};
  return true;
}

// fully-connected block with a 2-D batch norm (enc.fc, dec.fc): statistics + normalisation in one kernel (bn2d_fwd / bn2d_bwd)
bool is_fc2d(const Block& b, int B) { return b.g.KH == 1 && b.rpi == 1 && b.res.p == nullptr && B <= 2048; }

// conv / deconv + batch norm (+ shortcut) + activation as ONE kernel (BnFwdFuse): on the chain stream only - the kernel holds
// its CTAs at a grid barrier, and two such kernels from different streams could each keep the other's remaining CTAs out.
bool bnf_fusable(svae_handle* h, const Block& b, int B) {
  if (!b.tc2_fwd || b.rpi <= 1 || b.g.KH != 4 || b.fwd_bar == nullptr || h->no_bnf) return false;
  if (cur_stream(h) != h->stream) return false;
  auto aligned4 = [](const float* p, int ld, int coff) { return ld % 4 == 0 && coff % 4 == 0 && (((uintptr_t)p) & 15) == 0; };
  if (b.res.p != nullptr && !(b.res.ppr == 1 && b.res.inner == b.feats && aligned4(b.res.p, b.res.ld, b.res.coff))) return false;
  const bool want_f32 = b.out.p != nullptr && !b.skip_f32;
  if (want_f32 && !(b.out.ppr == 1 && b.out.inner == b.feats && aligned4(b.out.p, b.out.ld, b.out.coff))) return false;
  if (b.out_bf.a.p != nullptr && (b.out_bf.inner != 0 || (b.out_bf.coff & 7))) return false;
  if (!want_f32 && b.out_bf.a.p == nullptr) return false;
  Geom g = b.g; g.B = B;
  return tc2_bnf_supported(g, b.tw_f, h->sm_count, mkview(b.y, b.feats, 0), b.feats);
}

int block_fwd(svae_handle* h, Block& b, int B, View in) {
  const bool fc2d = is_fc2d(b, B);
  b.dbg_in = in;
  if (bnf_fusable(h, b, B)) {
    const bool want_f32 = b.out.p != nullptr && !b.skip_f32;
    BnFwdFuse ff{h->pw(b.beta), b.res.p, b.res.ld, b.res.coff, want_f32 ? b.out.p : nullptr, b.out.ld, b.out.coff, b.out_bf,
                 b.fwd_bar, b.act, (long long)B * b.rpi};
    Geom g = b.g; g.B = B;
    LaunchCtx lc = h->lc();
    return tc2_gather_gemm(lc, g, b.in_bf, 0, b.w_packed, mkview(b.y, b.feats, 0), b.stats, nullptr, b.tw_f, &ff);
  }
  Geom gf = b.g;
  if (b.y_zeroed && b.g.KH == 1) gf.accumulate = 1;   // y was zeroed with the forward arena: += is =, without a memset node
  H_TRY(contract_bf(h, gf, B, b.tc2_fwd, b.in_bf, in, h->pw(b.w), b.w_packed, b.tc_fwd, mkview(b.y, b.feats, 0),
                    fc2d ? nullptr : b.stats, nullptr, b.tw_f));
  LaunchCtx lc = h->lc();
  FeatView out = b.out;
  if (b.skip_f32) out.p = nullptr;   // only the bf16 planar copy is read downstream
  if (fc2d) H_TRY(bn2d_fwd(lc, b.y, h->pw(b.beta), B, b.feats, b.act, b.stats, out, b.out_bf));
  else H_TRY(bn_act_fwd(lc, b.y, b.stats, h->pw(b.beta), (int64_t)B * b.rpi, b.feats, b.act, b.res, out, b.out_bf));
  return 0;
}

// backward of a block on the current stream: da -> dy, beta gradient, optional input gradient; the weight gradient goes
// to `wst` (a side stream, or the current one)
int block_bwd(svae_handle* h, GradSet& gs, Block& b, int B, FeatView da, View in, float* dres, int dres_acc, const View* din,
              int din_acc, cudaStream_t wst, Block* up = nullptr, float* up_dres = nullptr, int up_dres_acc = 0) {
  LaunchCtx lc = h->lc();
  const int64_t rows = (int64_t)B * b.rpi;
  b.dbg_da = da; b.dbg_gs = (int)(&gs - h->gs);
  float* dy = gs.dy[b.dy_slot];
  const bool have_bf = gs.dy_bf[b.dy_slot].p != nullptr;
  const bool tc2d = b.tc2_dgrad && din != nullptr && have_bf;
  const bool tc2w = b.tc2_wgrad && have_bf && b.in_bf.p != nullptr;
  const BfDst dy_bf = (tc2d || tc2w) ? BfDst{gs.dy_bf[b.dy_slot], 0, 0, 0} : BfDst{};
  b.dbg_dy_f32 = !tc2w || (din != nullptr && !tc2d) || dy_bf.a.p == nullptr;
  if (b.g_fused) {
    // pass 1 ran inside the kernel that produced da (it holds g now); the fp32 dy is only written for consumers that are
    // not TMA-fed
    b.g_fused = false;
    const bool need_f32 = !tc2w || (din != nullptr && !tc2d);
    H_TRY(bn_bwd_apply_from(lc, mkview(da.p, da.ld, da.coff), need_f32 ? dy : nullptr, b.y, b.stats, b.S, rows, b.feats,
                            h->pg(b.beta), dy_bf));
  } else if (is_fc2d(b, B) && dres == nullptr && dy_bf.a.p == nullptr) {
    H_TRY(bn2d_bwd(lc, da, b.y, b.stats, h->pw(b.beta), B, b.feats, b.act, dy, h->pg(b.beta)));
  } else {
    int fr = 1;   // 1: not fused
    if (h->use_coop_bn && h->multi == nullptr && cur_stream(h) == h->stream) {   // chain stream only: one grid barrier at a time
      const bool need_f32 = !tc2w || (din != nullptr && !tc2d);
      fr = bn_bwd_fused(lc, da, b.y, b.stats, h->pw(b.beta), rows, b.feats, b.act, b.res, need_f32 ? dy : nullptr, b.S, dres,
                        dres_acc, h->pg(b.beta), dy_bf);
      if (fr < 0) { if (h->err.empty()) h->err = g_err; return fr; }
    }
    if (fr == 1) {
      H_TRY(bn_bwd_reduce(lc, da, b.y, b.stats, h->pw(b.beta), rows, b.feats, b.act, b.res, dy, b.S, dres, dres_acc));
      // the fp32 dy is only written for consumers that are not TMA-fed (they read the bf16 planar copy)
      const bool need_f32 = !tc2w || (din != nullptr && !tc2d);
      if (!need_f32 && dy_bf.a.p != nullptr && b.feats % 8 == 0 && b.feats <= 2048 && b.rpi > 1)
        H_TRY(bn_bwd_apply_from(lc, mkview(dy, b.feats, 0), nullptr, b.y, b.stats, b.S, rows, b.feats, h->pg(b.beta), dy_bf));
      else
        H_TRY(bn_bwd_apply(lc, dy, b.y, b.stats, b.S, rows, b.feats, h->pg(b.beta), dy_bf));
    }
  }
  View dyv = mkview(dy, b.feats, 0);
  const bool wdefer = h->wrec != nullptr && h->multi == nullptr && tc2w;
  // dy is final here: the weight gradient (side stream) depends on this point only, not on the input gradient below
  if (!wdefer) H_TRY(link(h, cur_stream(h), wst));
  if (din != nullptr) {
    Geom g = dgrad_geom(b.g);
    g.accumulate = din_acc;
    // 3-channel image gradient in a 4-channel buffer (svae_handle::gxld): written as 4 columns - the packed weights of the
    // fourth are zero - so that the epilogue takes its 16-byte path
    if (tc2d && g.Cout == 3 && din->ld == 4 && din->coff == 0) g.Cout = 4;
    BnBwdFuse fz;
    const bool fuse = tc2d && h->multi == nullptr && make_fuse(h, up, B, up_dres, up_dres_acc, g, *din, fz);
    H_TRY(contract_bf(h, g, B, tc2d, gs.dy_bf[b.dy_slot], dyv, h->pw(b.w), b.w_packed_d, b.tc_dgrad, *din, nullptr,
                      fuse ? &fz : nullptr, b.tw_d));
    if (fuse) up->g_fused = true;
  }
  if (wdefer) {
    LaunchCtx lw = h->lc();
    lw.multi = h->wrec; lw.sm_count = h->wrec_sm; lw.pdl_state = nullptr;
    if (b.g.mode == 0) { Geom g = b.g; g.B = B; H_TRY(tc2_wgrad(lw, g, b.in_bf, gs.dy_bf[b.dy_slot], h->pg(b.w))); }
    else { Geom g = dgrad_geom(b.g); g.B = B; g.mode = 0; H_TRY(tc2_wgrad(lw, g, gs.dy_bf[b.dy_slot], b.in_bf, h->pg(b.w))); }
  } else if (!(h->ablate & 1)) {
    OnStream os(h, wst);
    struct Lane { MultiRec* m; Lane(MultiRec* m_) : m(m_) { if (m) m->lane = 1; } ~Lane() { if (m) m->lane = 0; } } lane(h->multi);
    LaunchCtx lw = h->lc();
    if (h->wgrad_sm > 0 && h->multi == nullptr && wst != h->stream) lw.sm_count = std::min(lw.sm_count, h->wgrad_sm);
    if (b.g.mode == 0) {
      Geom g = b.g; g.B = B;
      if (tc2w) H_TRY(tc2_wgrad(lw, g, b.in_bf, gs.dy_bf[b.dy_slot], h->pg(b.w)));
      else if (b.tc_wgrad) H_TRY(tc_wgrad(lw, g, in, dyv, h->pg(b.w))); else H_TRY(simt_wgrad(lw, g, in, dyv, h->pg(b.w)));
    } else {
      Geom g = dgrad_geom(b.g); g.B = B; g.mode = 0;  // conv geometry from the deconv's output grid to its input grid
      if (tc2w) H_TRY(tc2_wgrad(lw, g, gs.dy_bf[b.dy_slot], b.in_bf, h->pg(b.w)));
      else if (b.tc_wgrad) H_TRY(tc_wgrad(lw, g, dyv, in, h->pg(b.w))); else H_TRY(simt_wgrad(lw, g, dyv, in, h->pg(b.w)));
    }
  }
  return 0;
}

int skinny_block_bwd(svae_handle* h, GradSet& gs, Block& b, int B, FeatView da, View zin, int K, View dz_out) {
  // latent projection (K <= 32 inputs): dW via thread-per-feature, dz via row-wise dot with the [K,N] weights
  if (h->ablate & 2) return 0;
  LaunchCtx lc = h->lc();
  b.dbg_da = da; b.dbg_gs = (int)(&gs - h->gs);
  if (lat_fused_supported(B, K))
    return lat_bwd_fused(lc, da, b.y, b.stats, h->pw(b.beta), zin, h->pw(b.w), B, K, b.feats, b.act, h->pg(b.w), h->pg(b.beta),
                         dz_out);
  float* dy = gs.dy[b.dy_slot];
  H_TRY(bn_bwd_reduce(lc, da, b.y, b.stats, h->pw(b.beta), B, b.feats, b.act, FeatView{}, dy, b.S, nullptr, 0));
  H_TRY(bn_bwd_apply(lc, dy, b.y, b.stats, b.S, B, b.feats, h->pg(b.beta)));
  H_TRY(lat_wgrad(lc, zin, dy, B, K, b.feats, h->pg(b.w)));
  H_TRY(lat_dz(lc, dy, h->pw(b.w), B, K, b.feats, dz_out));   // d_z is zeroed at the start of the step's backward
  return 0;
}

int zero_region(svae_handle* h, void* p, size_t bytes) {
  if (bytes == 0) return 0;
  H_CUDA(cudaMemsetAsync(p, 0, bytes, cur_stream(h)));
  return 0;
}

int repack_if_dirty(svae_handle* h);

// The bf16 copies are indexed by image: when the batch size changes, rows of images beyond the new batch would keep stale
// values that the reductions over the padded pixel space (weight gradients) must not see.  Rare, so: wipe everything.
int bf_guard(svae_handle* h, int B) {
  if (h->bf_base != nullptr && h->bf_last_B != B) {
    H_CUDA(cudaMemsetAsync(h->bf_base, 0, h->bf_bytes, h->stream));
    h->bf_last_B = B;
  }
  return 0;
}

// per-iteration scalars -> device (pinned ring slot -> dyn_dev on the main stream).  Never captured into a graph: the
// graph is replayed with whatever values the copy enqueued just before it delivered.
int dyn_push(svae_handle* h) {
  if (h->capturing) return 0;
  const int slot = h->dyn_slot;
  h->dyn_slot = (h->dyn_slot + 1) % (int)h->dyn_ev.size();
  if (h->dyn_ev[slot] != nullptr) H_CUDA(cudaEventSynchronize(h->dyn_ev[slot]));   // the copy that last used this slot is done
  else H_CUDA(cudaEventCreateWithFlags(&h->dyn_ev[slot], cudaEventDisableTiming));
  h->dyn_ring[slot] = h->dyn_host;
  H_CUDA(cudaMemcpyAsync(h->dyn_dev, &h->dyn_ring[slot], sizeof(SvaeDyn), cudaMemcpyHostToDevice, h->stream));
  H_CUDA(cudaEventRecord(h->dyn_ev[slot], h->stream));
  return 0;
}

// ---- batched launches (MultiRec, common.cuh) --------------------------------------------------------------------------
// Issue a recorded sequence on the current stream: one launch per slot with grid.z = items (the argument blocks travel in the
// kernel parameter space, see multi_launch_chunks).
int multi_flush(svae_handle* h, MultiRec& m, cudaStream_t wst = nullptr) {
  if (m.pending && m.err.empty()) m.err = "batched launch: a kernel wrapper without batched-launch support ran inside a recorded sequence";
  if (!m.err.empty()) return fail(h, SVAE_ESTATE, m.err);
  if (m.items == 0 || m.slots.empty()) return 0;
  for (const MultiSlot& sl : m.slots)
    if (sl.count != m.items) return fail(h, SVAE_ESTATE, "batched launch: an item recorded fewer kernels than the first one");
  cudaStream_t st = cur_stream(h);
  // Lane-1 slots (weight gradients) only consume what earlier lane-0 slots produced and nothing in the sequence consumes
  // them: they go to `wst`, ordered behind the lane-0 slots recorded before them.
  bool need_link = true;
  for (size_t i = 0; i < m.slots.size(); ++i) {
    MultiSlot& sl = m.slots[i];
    cudaStream_t ls = st;
    if (sl.lane == 1 && wst != nullptr && wst != st) {
      if (need_link) { H_TRY(link(h, st, wst)); need_link = false; }
      ls = wst;
    } else {
      need_link = true;
    }
    OnStream os(h, ls);
    LaunchCtx lc = h->lc();
    lc.pdl_state = nullptr;
    ProfScope ps(lc, sl.kc, sl.flops, sl.bytes);
    h->launches += (m.items - 1) / MULTI_MAX;   // more than MULTI_MAX items: one launch per chunk
    sl.launch(sl.host.data(), m.items, sl.grid, sl.block, sl.smem, ls);
    H_CUDA(cudaGetLastError());
  }
  if (st == h->stream) h->pdl_prev = 0;
  return 0;
}

// ---- forward of one chain step --------------------------------------------------------------------------------------
int encoder_fwd(svae_handle* h, Step& s, int B, const float* xprev) {
  const int L = h->L;
  View cur = mkview(const_cast<float*>(xprev), h->C, 0);
  for (int k = 0; k <= 2 * (L - 1); ++k) {
    H_TRY(block_fwd(h, s.enc[k], B, cur));
    cur = mkview(s.enc[k].out.p, s.enc[k].feats, 0);
  }
  View flat = mkview(cur.p, s.encfc.g.Cin, 0);
  H_TRY(block_fwd(h, s.encfc, B, flat));
  return 0;
}

// latent projections P_i = fc_bn_lrelu(z_i) (split_latent, sequential_vae.py:1796-1806): depend on z_t only
int latent_fwd(svae_handle* h, Step& s, int B, const float* z) {
  for (int i = 0; i < h->L; ++i) {
    Block& b = s.lat[i];
    View zv = mkview(const_cast<float*>(z), h->Z, h->zoff[i]);
    if (b.g.Cin <= 8 || lat_fused_supported(B, b.g.Cin)) {
      LaunchCtx lc = h->lc();
      FeatView out = b.out;
      if (b.skip_f32) out.p = nullptr;
      // forward-only handles never run the backward: the pre-BN tensor and the statistics are not stored either
      const bool keep = h->cfg.train_capacity != 0;
      H_TRY(lat_fwd_fused(lc, zv, h->pw(b.w), h->pw(b.beta), B, b.g.Cin, b.feats, b.act, keep ? b.y : nullptr, keep ? b.stats : nullptr,
                          out, b.out_bf, s.lat_mom ? s.lat_mom + (size_t)i * 64 : nullptr));
    } else {
      H_TRY(block_fwd(h, b, B, zv));
    }
  }
  return 0;
}

int decoder_fwd(svae_handle* h, Step& s, int B, const float* xprev, const float* tgt, float* xt_out, double* recon_sum) {
  const int L = h->L, C = h->C;
  LaunchCtx lc = h->lc();
  H_TRY(block_fwd(h, s.decfc, B, mkview(s.cc, s.ncc, 0)));
  View cur = mkview(s.decfc.out.p, s.ta[L - 2].g.Cin, 0);
  for (int l = L - 2; l >= 0; --l) {
    H_TRY(block_fwd(h, s.ta[l], B, cur));
    H_TRY(block_fwd(h, s.tb[l], B, mkview(s.dcat[l], 2 * s.tb[l].feats, 0)));
    cur = mkview(s.tb[l].out.p, s.tb[l].feats, 0);
  }
  const int has_gate = s.t > 0 ? 1 : 0;
  const int ldu = C + has_gate;
  if (has_gate && s.merge_og) {
    H_TRY(contract_bf(h, s.g_og, B, true, s.ogb.in_bf, cur, h->pw(s.w_out), s.ogb.w_packed, true, mkview(s.u, ldu, 0), nullptr, nullptr,
                      s.ogb.tw_f));
  } else {
  H_TRY(contract_bf(h, s.g_out, B, s.outb.tc2_fwd, s.outb.in_bf, cur, h->pw(s.w_out), s.outb.w_packed, s.outb.tc_fwd,
                    mkview(s.u, ldu, 0), nullptr));
  if (has_gate)
    H_TRY(contract_bf(h, s.g_gate, B, s.gateb.tc2_fwd, s.gateb.in_bf, cur, h->pw(s.w_gate), s.gateb.w_packed, s.gateb.tc_fwd,
                      mkview(s.u, ldu, C), nullptr));
  }
  OutMixParams p{(int64_t)B * h->D * h->D, C, has_gate, h->cfg.range_lo, h->cfg.range_hi, h->cfg.min_highway,
                 h->cfg.max_highway};
  H_TRY(out_mix_fwd(lc, p, s.u, h->pw(s.b_out), has_gate ? h->pw(s.b_gate) : nullptr, xprev, tgt, xt_out, recon_sum,
                    (s.xt_bf.p && !h->noisy()) ? BfDst{s.xt_bf, 0, 0, 0} : BfDst{}));
  return 0;
}

// add_noise_to_chain (sequential_vae.py:1088-1091): s.xs = xt + reg * noise_stddevs[t] * N(0, I), and the bf16 copy the next
// step's chain encoder reads.  dyn != nullptr: reg and the Philox key / counter base come from the per-iteration scalars (training
// forward); else reg_fixed (generation: the placeholder's default 1.0) and `seed`.
int chain_noise_step(svae_handle* h, Step& s, int B, const float* xt, const SvaeDyn* dyn, float reg_fixed, uint64_t seed) {
  const int64_t pixels = (int64_t)B * h->D * h->D;
  const uint64_t img = (uint64_t)h->cfg.max_batch * h->D * h->D * h->C;
  const float* inj = nullptr;
  if (h->chain_noise_dev != nullptr) {
    if (h->chain_noise_B != B) return fail(h, SVAE_EINVAL, "injected chain noise was set for a different batch size");
    inj = h->chain_noise_dev + (size_t)s.t * B * h->D * h->D * h->C;
  }
  LaunchCtx lc = h->lc();
  return chain_noise(lc, xt, inj, dyn, reg_fixed, h->cfg.noise_stddevs[s.t], 0xC4A17015E5EEDull ^ seed, (uint64_t)s.t * img, (uint64_t)h->T * img,
                     s.xs, pixels, h->C, s.xt_bf.p ? BfDst{s.xt_bf, 0, 0, 0} : BfDst{});
}

HeadSet make_headset(svae_handle* h, Step& s, int l) {
  HeadSet hs{};
  hs.nheads = (int)s.heads[l].size();
  for (int i = 0; i < hs.nheads; ++i) {
    const Head& hd = s.heads[l][i];
    hs.w[i] = h->pw(hd.w); hs.b[i] = h->pw(hd.b);
    hs.gw[i] = h->G ? h->pg(hd.w) : nullptr; hs.gb[i] = h->G ? h->pg(hd.b) : nullptr;
    hs.n[i] = hd.n; hs.col[i] = hd.col; hs.is_sd[i] = hd.is_sd;
  }
  return hs;
}

int recognition_fwd_impl(svae_handle* h, Step& s, int B, const float* x, const float* eps, double* kl_sum);
int recognition_fwd(svae_handle* h, Step& s, int B, const float* x, const float* eps, double* kl_sum) {
  h->no_bnf = true;
  const int r = recognition_fwd_impl(h, s, B, x, eps, kl_sum);
  h->no_bnf = false;
  return r;
}
int recognition_fwd_impl(svae_handle* h, Step& s, int B, const float* x, const float* eps, double* kl_sum) {
  const int L = h->L;
  LaunchCtx lc = h->lc();
  // the heads accumulate into mu_pre / sd_pre (adjacent in the arena) with atomics
  H_CUDA(cudaMemsetAsync(s.mu_pre, 0, sizeof(float) * 2 * (size_t)h->cfg.max_batch * h->Z, cur_stream(h)));
  View cur = mkview(const_cast<float*>(x), h->C, 0);
  for (int k = 0; k < 2 * (L - 1); ++k) {
    H_TRY(block_fwd(h, s.inf[k], B, cur));
    cur = mkview(s.inf[k].out.p, s.inf[k].feats, 0);
    if (k & 1) {
      const int l = k / 2;
      const int K = s.inf[k].rpi * s.inf[k].feats;
      HeadSet hs = make_headset(h, s, l);
      H_TRY(heads_fwd(lc, cur.p, B, K, hs, s.mu_pre, s.sd_pre, h->Z));
    }
  }
  ReparamParams rp{B, h->Z, h->cfg.latent_mean_clip, h->cfg.prior_stddev};
  H_TRY(reparam_fwd(lc, rp, s.mu_pre, s.sd_pre, eps, h->dyn_dev, h->T, s.t, (uint64_t)(h->cfg.max_batch * h->Z), s.eps, s.mu,
                    s.sd, s.z, kl_sum));
  return 0;
}

// Philox key of the in-kernel eps.  Data-parallel replicas must draw INDEPENDENT noise for their shards (the effective batch
// of noise samples is N*B, not B): the rank is folded into the key (splitmix64 of the rank, so that neighbouring ranks get
// unrelated keys); rank 0 / single GPU keeps the caller's seed unchanged.
uint64_t rank_seed(const svae_handle* h, uint64_t seed) {
  if (h->rank == 0) return seed;
  uint64_t z = (uint64_t)h->rank * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return seed ^ (z ^ (z >> 31));
}

float step_coef(const svae_handle* h, int t) { return t == 0 ? h->cfg.first_step_loss_coeff : 1.f; }
bool step_has_recon(const svae_handle* h, int t) { return h->cfg.intermediate_reconstruction || t == h->T - 1; }
bool step_has_kl(const svae_handle* h, int t) { return (h->cfg.regularized_mask >> t) & 1ull; }

// The recognition nets + latent projections of all chain steps can run as batched launches when every kernel on that path
// records (TMA-fed convs / input gradients / weight gradients, fused latent kernels, separate batch-norm passes) and every
// chain step owns its buffers and gradient scratch set.
bool rec_multi_ok(const svae_handle* h, int B) {
  if (!h->use_multi || !h->cfg.train_capacity || h->act_sets != h->T || h->T < 2) return false;
  if ((h->ablate & 3) || h->timeline) return false;
  for (const Step& s : h->steps) {
    for (size_t k = 0; k < s.inf.size(); ++k) {
      const Block& b = s.inf[k];
      if (!b.tc2_fwd || b.in_bf.p == nullptr || !b.tc2_wgrad || (k > 0 && !b.tc2_dgrad)) return false;
    }
    for (const Block& b : s.lat)
      if (!lat_fused_supported(B, b.g.Cin) || (size_t)B * 32 * 8 > 40 * 1024) return false;
  }
  return true;
}

int forward_impl(svae_handle* h, const float* x, const float* tgt, int B, const float* eps, uint64_t seed, float reg,
                 float* mu_out, float* sd_out, float* xs_out) {
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  h->cur = h->stream;
  h->ev_used = 0;
  tl_mark(h, h->stream, "main: step start", -1);
  h->dyn_host.reg = reg; h->dyn_host.seed = rank_seed(h, seed); h->dyn_host.iteration = h->iteration;
  H_TRY(dyn_push(h));
  H_TRY(repack_if_dirty(h));
  H_TRY(zero_region(h, h->zf_base, h->zf_bytes));
  H_TRY(zero_region(h, h->loss_sums, sizeof(double) * 2 * h->T));
  const size_t img = (size_t)B * h->D * h->D * h->C;
  const size_t bz = (size_t)B * h->Z;
  if (h->x_bf.p != nullptr) {   // bf16 planar copy of the batch for the recognition nets' first conv (all T steps read it)
    LaunchCtx lc = h->lc();
    BfAct xb = h->x_bf;
    xb.B = B;                     // only the first B images exist in the caller's buffer; the rest of the copy is zeroed
    H_TRY(bf_act_fill(lc, xb, mkview(const_cast<float*>(x), h->C, 0), h->C));
  }
  const bool fork = forked(h) && h->act_sets == h->T && (h->fork_mask & 8);
  std::vector<cudaEvent_t> rec_ev(h->T, nullptr);
  const bool multi = rec_multi_ok(h, B);
  if (multi) {
    // Batched: the nets of a GROUP of chain steps are recorded one after the other and issued as one launch per layer kernel
    // (blockIdx.z = step).  The first group is small so that the chain can start early; the groups run side by side on the
    // side streams (on the main stream, before the chain, when the step is not forked).
    int t0 = 0;
    for (int gi = 0; t0 < h->T; ++gi) {
      const int n = h->rec_fwd_groups[std::min((size_t)gi, h->rec_fwd_groups.size() - 1)];
      const int t1 = std::min(h->T, t0 + n);
      cudaStream_t sd = fork ? h->side[gi % 3] : h->stream;
      if (fork && gi < 3) H_TRY(link(h, h->stream, sd));
      OnStream os(h, sd);
      MultiRec rec;
      h->begin_multi(&rec, t1 - t0, t0 == 0);
      int r = 0;
      for (int t = t0; t < t1 && r == 0; ++t) {
        Step& s = h->steps[t];
        rec.begin_item();
        r = recognition_fwd(h, s, B, x, eps ? eps + t * bz : nullptr, h->loss_sums + h->T + t);
        if (r == 0) r = latent_fwd(h, s, B, s.z);
      }
      h->multi = nullptr;
      H_TRY(r);
      H_TRY(multi_flush(h, rec));
      if (fork) {
        cudaEvent_t e = next_event(h);
        H_CUDA(cudaEventRecord(e, sd));
        for (int t = t0; t < t1; ++t) rec_ev[t] = e;
        tl_mark(h, sd, "side: recognition fwd group done", t0);
      }
      t0 = t1;
    }
  } else if (fork) {
    // The recognition nets read only x (sequential_vae.py:1011-1025) and the latent projections only z_t: all T of them
    // run on the side streams, off the chain's critical path.
    for (int t = 0; t < h->T; ++t) {
      Step& s = h->steps[t];
      cudaStream_t sd = h->side[t % 3];
      if (t < 3) H_TRY(link(h, h->stream, sd));
      OnStream os(h, sd);
      H_TRY(recognition_fwd(h, s, B, x, eps ? eps + t * bz : nullptr, h->loss_sums + h->T + t));
      H_TRY(latent_fwd(h, s, B, s.z));
      rec_ev[t] = next_event(h);
      H_CUDA(cudaEventRecord(rec_ev[t], sd));
      tl_mark(h, sd, "side: recognition fwd done", t);
    }
  }
  const float* prev = nullptr;
  for (int t = 0; t < h->T; ++t) {
    Step& s = h->steps[t];
    // forward-only handles alias steps >= 2 onto step 1's buffers: restart their batch-norm statistics
    if (t >= 2 && h->act_sets < h->T) H_TRY(zero_region(h, h->zf_base, h->zf_bytes));
    if (!fork && !multi) H_TRY(recognition_fwd(h, s, B, x, eps ? eps + t * bz : nullptr, h->loss_sums + h->T + t));
    if (t > 0) H_TRY(encoder_fwd(h, s, B, prev));
    if (fork) { if (t == 0 || rec_ev[t] != rec_ev[t - 1]) H_CUDA(cudaStreamWaitEvent(h->stream, rec_ev[t], 0)); }
    else if (!multi) H_TRY(latent_fwd(h, s, B, s.z));
    H_TRY(decoder_fwd(h, s, B, prev, tgt, s.xt, h->loss_sums + t));
    tl_mark(h, h->stream, "main: fwd chain step done", t);
    if (xs_out) H_CUDA(cudaMemcpyAsync(xs_out + t * img, s.xt, img * 4, cudaMemcpyDeviceToDevice, h->stream));
    if (mu_out) H_CUDA(cudaMemcpyAsync(mu_out + t * bz, s.mu, bz * 4, cudaMemcpyDeviceToDevice, h->stream));
    if (sd_out) H_CUDA(cudaMemcpyAsync(sd_out + t * bz, s.sd, bz * 4, cudaMemcpyDeviceToDevice, h->stream));
    if (h->noisy()) H_TRY(chain_noise_step(h, s, B, s.xt, h->dyn_dev, 0.f, 0));
    prev = h->chain_in(t);
    if (!h->noisy() && h->act_sets < h->T && t >= 1 && t + 1 < h->T) {
      // the next step overwrites s.xt (aliased): keep x_t in the staging target buffer (samples have one buffer per step)
      H_CUDA(cudaMemcpyAsync(h->gen_prev, s.xt, img * 4, cudaMemcpyDeviceToDevice, h->stream));
      prev = h->gen_prev;
    }
  }
  tl_mark(h, h->stream, "main: forward done", -1);
  h->last_B = B; h->last_reg = reg; h->have_fwd = true; h->last_x = x; h->last_tgt = tgt;
  return 0;
}

// ---- backward ------------------------------------------------------------------------------------------------------
// Streams of one step's backward: the chain (decoder, chain encoder: bn backward + input gradients) stays on the main
// stream; weight gradients of the chain go to side[0]; the latent projections' backward and the whole recognition net go
// to side[1] with its weight gradients on side[2].
struct BwdStreams { cudaStream_t chain, w, rec, recw, lat; };   // lat: the latent projections' backward (the recognition stream unless its work is batched)

int decoder_bwd(svae_handle* h, GradSet& gs, const BwdStreams& st, Step& s, int B, const float* gx_in, float* gx_prev,
                const float* xprev) {
  const int L = h->L, C = h->C;
  const int* F = h->cfg.filter_sizes;
  LaunchCtx lc = h->lc();
  const int has_gate = s.t > 0 ? 1 : 0;
  const int ldu = C + has_gate;
  s.outb.dbg_gs = s.gateb.dbg_gs = (int)(&gs - h->gs);
  const float coef = step_has_recon(h, s.t)
                         ? 16.f * step_coef(h, s.t) * 2.f / ((float)B * h->D * h->D * C) : 0.f;   // :1146,1163,1168
  OutMixParams p{(int64_t)B * h->D * h->D, C, has_gate, h->cfg.range_lo, h->cfg.range_hi, h->cfg.min_highway,
                 h->cfg.max_highway, h->gxld};
  if (!h->bwd_prezero) H_CUDA(cudaMemsetAsync(gs.d_z, 0, sizeof(float) * (size_t)B * h->Z, st.chain));   // lat_dz accumulates with atomics
  H_TRY(out_mix_bwd(lc, p, s.u, h->pw(s.b_out), has_gate ? h->pw(s.b_gate) : nullptr, xprev, h->last_tgt, s.xt, gx_in,
                    coef, gs.d_u, gx_prev, h->pg(s.b_out), has_gate ? h->pg(s.b_gate) : nullptr,
                    gs.du_out_bf.p ? BfDst{gs.du_out_bf, 0, 0, 0} : BfDst{},
                    (has_gate && gs.du_gate_bf.p && !s.merge_og) ? BfDst{gs.du_gate_bf, 0, 0, 0} : BfDst{}, (has_gate && s.merge_og) ? 1 : 0));
  // output deconvs: dgrad into d_c[0], wgrads
  View c0 = mkview(s.tb[0].out.p, F[1], 0);
  View dc0 = mkview(gs.d_c[0], F[1], 0);
  if (has_gate && s.merge_og) {
    // merged output + gate deconv: one input gradient from the C + 1 channel copy of d_u, one weight gradient whose rows go to
    // the two parameter tensors
    Geom g4 = dgrad_geom(s.g_og);
    H_TRY(contract_bf(h, g4, B, true, gs.du_out_bf, mkview(gs.d_u, ldu, 0), h->pw(s.w_out), s.ogb.w_packed_d, true, dc0, nullptr, nullptr,
                      s.ogb.tw_d));
    Geom gw4 = dgrad_geom(s.g_og); gw4.B = B; gw4.mode = 0;
    if (h->wrec != nullptr) {
      LaunchCtx lr = h->lc();
      lr.multi = h->wrec; lr.sm_count = h->wrec_sm; lr.pdl_state = nullptr;
      H_TRY(tc2_wgrad(lr, gw4, gs.du_out_bf, s.ogb.in_bf, h->pg(s.w_out), h->pg(s.w_gate), C));
    } else if (!(h->ablate & 1)) {
      H_TRY(link(h, st.chain, st.w));
      OnStream os(h, st.w);
      LaunchCtx lw = h->lc();
      if (h->wgrad_sm > 0 && st.w != h->stream) lw.sm_count = std::min(lw.sm_count, h->wgrad_sm);
      H_TRY(tc2_wgrad(lw, gw4, gs.du_out_bf, s.ogb.in_bf, h->pg(s.w_out), h->pg(s.w_gate), C));
    }
  } else
  {
    // d_c[0] is complete after the LAST of the two input gradients: that one carries pass 1 of tb[0]'s batch-norm backward
    Geom g = dgrad_geom(s.g_out);
    const bool tc2o = gs.du_out_bf.p != nullptr && s.outb.tc_dgrad;
    const bool tc2g = has_gate && gs.du_gate_bf.p != nullptr && s.gateb.tc_dgrad;
    BnBwdFuse fz;
    const bool fuse_o = !has_gate && tc2o && make_fuse(h, &s.tb[0], B, nullptr, 0, g, dc0, fz);
    H_TRY(contract_bf(h, g, B, tc2o, gs.du_out_bf, mkview(gs.d_u, ldu, 0), h->pw(s.w_out), s.outb.w_packed_d, s.outb.tc_dgrad, dc0,
                      nullptr, fuse_o ? &fz : nullptr));
    if (fuse_o) s.tb[0].g_fused = true;
    if (has_gate) {
      Geom g2 = dgrad_geom(s.g_gate); g2.accumulate = 1;
      const bool fuse_g = tc2g && make_fuse(h, &s.tb[0], B, nullptr, 0, g2, dc0, fz);
      H_TRY(contract_bf(h, g2, B, gs.du_gate_bf.p != nullptr && s.gateb.tc_dgrad, gs.du_gate_bf, mkview(gs.d_u, ldu, C),
                        h->pw(s.w_gate), s.gateb.w_packed_d, s.gateb.tc_dgrad, dc0, nullptr, fuse_g ? &fz : nullptr));
      if (fuse_g) s.tb[0].g_fused = true;
    }
    const bool wdef_o = h->wrec != nullptr && s.outb.tc2_wgrad && gs.du_out_bf.p;
    const bool wdef_g = h->wrec != nullptr && has_gate && s.gateb.tc2_wgrad && gs.du_gate_bf.p;
    LaunchCtx lr = h->lc();
    lr.multi = h->wrec; lr.sm_count = h->wrec_sm; lr.pdl_state = nullptr;
    if (!(wdef_o && (wdef_g || !has_gate))) H_TRY(link(h, st.chain, st.w));
    OnStream os(h, st.w);
    LaunchCtx lw = h->lc();
    if (h->wgrad_sm > 0 && st.w != h->stream) lw.sm_count = std::min(lw.sm_count, h->wgrad_sm);
    Geom gw = dgrad_geom(s.g_out); gw.B = B; gw.mode = 0;
    if (wdef_o) H_TRY(tc2_wgrad(lr, gw, gs.du_out_bf, s.outb.in_bf, h->pg(s.w_out)));
    else if (s.outb.tc2_wgrad && gs.du_out_bf.p) H_TRY(tc2_wgrad(lw, gw, gs.du_out_bf, s.outb.in_bf, h->pg(s.w_out)));
    else if (s.outb.tc_wgrad) H_TRY(tc_wgrad(lw, gw, mkview(gs.d_u, ldu, 0), c0, h->pg(s.w_out)));
    else H_TRY(simt_wgrad(lw, gw, mkview(gs.d_u, ldu, 0), c0, h->pg(s.w_out)));
    if (has_gate) {
      Geom gw2 = dgrad_geom(s.g_gate); gw2.B = B; gw2.mode = 0;
      if (wdef_g) H_TRY(tc2_wgrad(lr, gw2, gs.du_gate_bf, s.gateb.in_bf, h->pg(s.w_gate)));
      else if (s.gateb.tc2_wgrad && gs.du_gate_bf.p) H_TRY(tc2_wgrad(lw, gw2, gs.du_gate_bf, s.gateb.in_bf, h->pg(s.w_gate)));
      else if (s.gateb.tc_wgrad) H_TRY(tc_wgrad(lw, gw2, mkview(gs.d_u, ldu, C), c0, h->pg(s.w_gate)));
      else H_TRY(simt_wgrad(lw, gw2, mkview(gs.d_u, ldu, C), c0, h->pg(s.w_gate)));
    }
  }
  for (int l = 0; l <= L - 2; ++l) {
    const int Fl = F[l + 1];
    // c_l = relu(bn(deconv_s1(dcat_l)))
    View dcat_in = mkview(s.dcat[l], 2 * Fl, 0);
    View d_dcat = mkview(gs.d_dcat[l], 2 * Fl, 0);
    H_TRY(block_bwd(h, gs, s.tb[l], B, fv4(gs.d_c[l], Fl, 0, Fl), dcat_in, nullptr, 0, &d_dcat, 0, st.w, &s.ta[l],
                    has_gate ? gs.d_e[2 * l + 1] : nullptr, 0));
    // P_l = lrelu(bn(fc(z_l))): second channel window of d_dcat, on the recognition stream
    {
      H_TRY(link(h, st.chain, st.lat));
      OnStream os(h, st.lat);
      const Block& lb = s.lat[l];
      FeatView da{gs.d_dcat[l], 2 * Fl, Fl, Fl, lb.out.ppr};
      H_TRY(skinny_block_bwd(h, gs, s.lat[l], B, da, mkview(s.z, h->Z, h->zoff[l]), h->cfg.latent_dims[l],
                             mkview(gs.d_z, h->Z, h->zoff[l])));
    }
    // d = relu(bn(deconv_s2(c_{l+1})) + e[l+1])
    View ta_in = l < L - 2 ? mkview(s.tb[l + 1].out.p, F[l + 2], 0) : mkview(s.decfc.out.p, s.ta[l].g.Cin, 0);
    View d_next = l < L - 2 ? mkview(gs.d_c[l + 1], F[l + 2], 0) : mkview(gs.d_fca, s.ta[l].g.Cin, 0);
    H_TRY(block_bwd(h, gs, s.ta[l], B, fv4(gs.d_dcat[l], 2 * Fl, 0, Fl), ta_in, has_gate ? gs.d_e[2 * l + 1] : nullptr, 0,
                    &d_next, 0, st.w, l < L - 2 ? &s.tb[l + 1] : nullptr));
  }
  // dec.fc
  {
    View d_cc = mkview(gs.d_cc, s.ncc, 0);
    H_TRY(block_bwd(h, gs, s.decfc, B, fv4(gs.d_fca, s.decfc.feats, 0, s.decfc.feats), mkview(s.cc, s.ncc, 0), nullptr, 0,
                    &d_cc, h->bwd_prezero ? 1 : 0, st.w));   // pre-zeroed: += is = (no memset node for the split-K kernel)
  }
  // P_{L-1}
  {
    H_TRY(link(h, st.chain, st.lat));
    OnStream os(h, st.lat);
    const int coffP = s.t > 0 ? F[L] : 0;
    FeatView da{gs.d_cc, s.ncc, coffP, F[L + 1], 1};
    H_TRY(skinny_block_bwd(h, gs, s.lat[L - 1], B, da, mkview(s.z, h->Z, h->zoff[L - 1]), h->cfg.latent_dims[L - 1],
                           mkview(gs.d_z, h->Z, h->zoff[L - 1])));
  }
  return 0;
}

int encoder_bwd(svae_handle* h, GradSet& gs, const BwdStreams& st, Step& s, int B, float* gx_prev, const float* xprev) {
  const int L = h->L;
  const int* F = h->cfg.filter_sizes;
  const int last = 2 * (L - 1);
  // e[L] = lrelu(bn(fc(flat)))
  {
    View flat = mkview(s.enc[last].out.p, s.encfc.g.Cin, 0);
    View d_flat = mkview(gs.d_e[last], s.encfc.g.Cin, 0);
    H_TRY(block_bwd(h, gs, s.encfc, B, fv4(gs.d_cc, s.ncc, 0, F[L]), flat, nullptr, 0, &d_flat, h->bwd_prezero ? 1 : 0, st.w));
  }
  for (int k = last; k >= 0; --k) {
    Block& b = s.enc[k];
    View in = k > 0 ? mkview(s.enc[k - 1].out.p, s.enc[k - 1].feats, 0) : mkview(const_cast<float*>(xprev), h->C, 0);
    View din = k > 0 ? mkview(gs.d_e[k - 1], s.enc[k - 1].feats, 0) : mkview(gx_prev, h->gxld, 0);
    // d_e[k-1] already holds the decoder shortcut gradient when k-1 is odd (e[l+1] = enc[2l+1]); gx_prev always holds
    // the highway gradient
    const int acc = k > 0 ? ((k - 1) & 1) : 1;
    H_TRY(block_bwd(h, gs, b, B, fv4(gs.d_e[k], b.feats, 0, b.feats), in, nullptr, 0, &din, acc, st.w, k > 0 ? &s.enc[k - 1] : nullptr));
  }
  return 0;
}

// runs on the current stream (the recognition stream when forked); weight gradients on `wst`
int recognition_bwd(svae_handle* h, GradSet& gs, Step& s, int B, cudaStream_t wst) {
  const int L = h->L;
  LaunchCtx lc = h->lc();
  ReparamParams rp{B, h->Z, h->cfg.latent_mean_clip, h->cfg.prior_stddev};
  const float kl_scale = step_has_kl(h, s.t) ? step_coef(h, s.t) / ((float)B * h->Z) : 0.f;   // :1156-1164,1172
  H_TRY(reparam_bwd(lc, rp, gs.d_z, s.mu_pre, s.mu, s.sd, s.eps, h->dyn_dev, kl_scale, gs.d_mu_pre, gs.d_sd_pre));
  for (int l = 0; l < L - 1; ++l) {
    const Block& fb = s.inf[2 * l + 1];
    const int K = fb.rpi * fb.feats;
    HeadSet hs = make_headset(h, s, l);
    H_TRY(heads_dgrad(lc, hs, gs.d_mu_pre, gs.d_sd_pre, B, h->Z, K, gs.d_inf[2 * l + 1]));
    if (h->multi) h->multi->lane = 1;
    const int rw = heads_wgrad(lc, fb.out.p, hs, gs.d_mu_pre, gs.d_sd_pre, B, h->Z, K);
    if (h->multi) h->multi->lane = 0;
    H_TRY(rw);
  }
  for (int k = 2 * (L - 1) - 1; k >= 0; --k) {
    Block& b = s.inf[k];
    View in = k > 0 ? mkview(s.inf[k - 1].out.p, s.inf[k - 1].feats, 0) : mkview(const_cast<float*>(h->last_x), h->C, 0);
    if (k > 0) {
      View din = mkview(gs.d_inf[k - 1], s.inf[k - 1].feats, 0);
      H_TRY(block_bwd(h, gs, b, B, fv4(gs.d_inf[k], b.feats, 0, b.feats), in, nullptr, 0, &din, (k - 1) & 1, wst, &s.inf[k - 1]));
    } else {
      H_TRY(block_bwd(h, gs, b, B, fv4(gs.d_inf[k], b.feats, 0, b.feats), in, nullptr, 0, nullptr, 0, wst));
    }
  }
  return 0;
}

// part: 0 = the whole slice of chain step t, 1 = its chain encoder + decoder, 2 = its recognition net (the two halves are
// used when the recognition backward is deferred to a batched group: the chain half is final long before the other)
void bucket_range(const svae_handle* h, int t, int part, int64_t& b, int64_t& e) {
  const Step& s = h->steps[t];
  b = part == 1 ? s.p_rec_end : s.p_begin;
  e = part == 2 ? s.p_rec_end : s.p_end;
}

int allreduce_bucket(svae_handle* h, int t, const BwdStreams& st, int part = 0) {
  if (h->comm == nullptr) return 0;
  int64_t pb, pe;
  bucket_range(h, t, part, pb, pe);
  // the bucket is complete when every stream that produced gradients of step t has finished its part
  H_TRY(link(h, st.chain, h->comm_stream));
  H_TRY(link(h, st.w, h->comm_stream));
  H_TRY(link(h, st.rec, h->comm_stream));
  H_TRY(link(h, st.recw, h->comm_stream));
  h->pdl_prev = 0;
  int r = h->nccl->AllReduce(h->G + pb, h->G + pb, (size_t)(pe - pb), /*ncclFloat*/ 7, /*ncclSum*/ 0, h->comm, h->comm_stream);
  if (r != 0) return fail(h, SVAE_ENCCL, std::string("ncclAllReduce failed: ") + h->nccl->GetErrorString(r));
  return 0;
}

int ensure_pack_table(svae_handle* h);

// clipped Adam + operand repack of chain step t's parameter slice (see svae_handle::upd_stream)
int update_bucket(svae_handle* h, int t, const BwdStreams& st, int part = 0) {
  if (!h->bucket_update) return 0;
  int64_t pb, pe;
  bucket_range(h, t, part, pb, pe);
  cudaStream_t us = h->comm != nullptr ? h->comm_stream : h->upd_stream;
  if (h->comm == nullptr) {   // with a communicator the all-reduce of this bucket was just enqueued on the same stream
    H_TRY(link(h, st.chain, us));
    H_TRY(link(h, st.w, us));
    H_TRY(link(h, st.rec, us));
    H_TRY(link(h, st.recw, us));
  }
  OnStream os(h, us);
  LaunchCtx lc = h->lc();
  // The update runs beside the chain's backward of the earlier steps.  Sized for the whole machine, the Adam kernel (1 184
  // persistent 256-thread blocks = every thread slot of every SM for ~60 us) and the repack kept the chain's next kernels waiting
  // for a free slot: 70-80 us per chain step in the kernel timeline (profiles/r2_step_analysis.md).  Sized for upd_sm SMs'
  // worth of blocks they stream at a fraction of the HBM rate - there is ~1 ms until the next bucket - and leave the slots free.
  if (h->upd_sm > 0 && us != h->stream) lc.sm_count = std::min(lc.sm_count, h->upd_sm);
  const int64_t n = pe - pb;
  H_TRY(adam_update(lc, h->P + pb, h->G + pb, h->M + pb, h->V + pb, n, h->dyn_dev, h->dyn_host.lr_t,
                    h->cfg.adam_beta1, h->cfg.adam_beta2, h->cfg.adam_eps, h->cfg.clip_value, 1.f / (float)h->nranks));
  if (h->cfg.operand_dtype == SVAE_OPERAND_BF16 && h->pack_table != nullptr) {
    const int b0 = part == 1 ? h->pack_rec_end[t] : h->pack_step_begin[t];
    const int b1 = part == 2 ? h->pack_rec_end[t] : h->pack_step_begin[t + 1];
    double elems = 0;
    for (int i = b0; i < b1; ++i) elems += h->pack_entry_elems[i];
    if (b1 > b0) H_TRY(tc_pack_batched(lc, reinterpret_cast<const TcPackEntry*>(h->pack_table) + b0, b1 - b0, elems));
  }
  return 0;
}

int backward_impl(svae_handle* h) {
  if (!h->cfg.train_capacity) return fail(h, SVAE_ESTATE, "handle created without train_capacity");
  if (!h->have_fwd) return fail(h, SVAE_ESTATE, "svae_backward requires a preceding svae_forward");
  const int B = h->last_B, T = h->T;
  h->cur = h->stream;
  H_TRY(zero_region(h, h->zb_base, h->zb_bytes));
  const bool fork = forked(h);
  BwdStreams st{h->stream, h->stream, h->stream, h->stream, h->stream};
  if (fork) {
    if (h->fork_mask & 1) st.w = h->side[0];
    if (h->fork_mask & 2) st.rec = h->side[1];
    st.recw = (h->fork_mask & 4) ? h->side[2] : st.rec;
  }
  st.lat = st.rec;
  // Gradient arena: 4 B x all parameters (335 MB for CelebA T = 8: ~55 us of memset).  Only the slice of the LAST chain step is
  // needed at once; the slices of the earlier steps are zeroed on the weight-gradient stream beside the first chain step.
  cudaEvent_t g_zeroed = nullptr;
  if (fork && T > 1 && st.w != st.chain) {
    const int64_t last = h->steps[T - 1].p_begin;
    H_TRY(zero_region(h, h->G + last, (size_t)(h->arena_numel - last) * 4));
    H_TRY(link(h, st.chain, st.w));
    H_CUDA(cudaMemsetAsync(h->G, 0, (size_t)last * 4, st.w));
    g_zeroed = next_event(h);
    H_CUDA(cudaEventRecord(g_zeroed, st.w));
  } else {
    H_TRY(zero_region(h, h->G, (size_t)h->arena_numel * 4));
  }
  // side_done[t][i]: side stream i has finished step t's work (its scratch set may be reused two steps later)
  std::vector<cudaEvent_t> side_done((size_t)T * 3, nullptr);
  int cur = 0;
  const float* gx_in = nullptr;  // dL/dx_t from later steps
  // Deferred recognition backward (rec_multi_ok): the latent projections' and the recognition nets' backward of a GROUP of chain
  // steps is recorded step by step once the chain has passed the group's last decoder and issued as one launch per layer
  // kernel (blockIdx.z = step) on the recognition stream, beside the chain's backward of the earlier steps.  Every step of
  // a group keeps its own gradient scratch set (group size <= n_gs).  The step's parameter slice is then all-reduced /
  // updated in two halves: chain encoder + decoder right after the chain step, recognition net after its group.
  const bool multi = rec_multi_ok(h, B);
  // batched: side[1] / side[2] carry the groups (input-gradient sequence / weight gradients), side[3] the per-step latent
  // projections (their parameters belong to the decoder's half of the slice)
  if (multi && fork && h->side[3] != nullptr && (h->fork_mask & 2)) st.lat = h->side[3];
  size_t gi = 0;
  auto group_size = [&]() { return std::max(1, std::min(h->rec_bwd_groups[std::min(gi, h->rec_bwd_groups.size() - 1)], h->n_gs)); };
  int group = group_size();
  std::vector<int> pending;
  struct Unmulti { svae_handle* h; ~Unmulti() { h->multi = nullptr; h->wrec = nullptr; } } unmulti{h};
  // Deferred chain weight gradients (svae_handle::wrec): recorded per chain step, issued per group on the weight-gradient
  // stream; step 0 (no chain encoder, no gate: a different kernel sequence) is always a group of its own.
  // (single GPU only: with a communicator the chain halves of a group would reach their all-reduce late - measured at N = 2:
  //  11.40 vs 11.01 ms/step)
  const bool wdefer = multi && fork && st.w != st.chain && !h->wgrad_groups.empty() && h->wgrad_groups[0] > 0 && !(h->ablate & 1) &&
                      h->comm == nullptr;
  size_t wgi = 0;
  int wgroup = 0;
  std::vector<int> wpending;
  std::unique_ptr<MultiRec> wrec;
  for (int t = T - 1; t >= 0; --t) {
    Step& s = h->steps[t];
    const int NS = h->n_gs;
    GradSet& gs = h->gs[t % NS];
    if ((fork || multi) && t + NS < T)
      for (int i = 0; i < 3; ++i)
        if (side_done[(size_t)(t + NS) * 3 + i] != nullptr) H_CUDA(cudaStreamWaitEvent(h->stream, side_done[(size_t)(t + NS) * 3 + i], 0));
    if (t == T - 2 && g_zeroed != nullptr) H_CUDA(cudaStreamWaitEvent(h->stream, g_zeroed, 0));
    if (wdefer) {
      if (wpending.empty()) {
        const int want = h->wgrad_groups[std::min(wgi, h->wgrad_groups.size() - 1)];
        wgroup = t == 0 ? 1 : std::max(1, std::min(std::min(want, t), h->n_gs));   // steps t .. t - wgroup + 1, never step 0 with others
        wrec.reset(new MultiRec);
        h->wrec = wrec.get();
        const int v = (int)((t == 0 ? 1.f : h->multi_sm_factor) * (float)h->sm_count / (float)wgroup + 0.999f);
        h->wrec_sm = v < 8 ? 8 : (v > h->sm_count ? h->sm_count : v);
        ++wgi;
      }
      wrec->begin_item();
      wpending.push_back(t);
    }
    const float* xprev = t > 0 ? h->chain_in(t - 1) : nullptr;   // d sample / d mle = 1: the chain gradient is unchanged
    float* gx_prev = h->gx[cur ^ 1];
    H_TRY(decoder_bwd(h, gs, st, s, B, gx_in, gx_prev, xprev));
    if (!multi) {
      // recognition net of this step: needs d_z (complete once the last latent projection's backward has run on st.rec)
      OnStream os(h, st.rec);
      if (!(h->ablate & 2)) H_TRY(recognition_bwd(h, gs, s, B, st.recw));
    } else {
      pending.push_back(t);
      if ((int)pending.size() == group || t == 0) {
        H_TRY(link(h, st.lat, st.rec));   // d_z of every step of the group is final
        OnStream os(h, st.rec);
        MultiRec rec;
        h->begin_multi(&rec, (int)pending.size(), t == 0);
        int r = 0;
        for (size_t i = 0; i < pending.size() && r == 0; ++i) {
          Step& ps = h->steps[pending[i]];
          GradSet& pgs = h->gs[pending[i] % NS];
          rec.begin_item();
          r = recognition_bwd(h, pgs, ps, B, st.rec);
        }
        h->multi = nullptr;
        H_TRY(r);
        H_TRY(multi_flush(h, rec, st.recw));
        tl_mark(h, st.rec, "R: recognition bwd group done", t);
      }
    }
    if (t > 0) H_TRY(encoder_bwd(h, gs, st, s, B, gx_prev, xprev));
    std::vector<int> chain_ready;          // chain steps whose encoder + decoder gradients are final as of this iteration
    if (!wdefer) {
      chain_ready.push_back(t);
    } else if ((int)wpending.size() == wgroup) {
      h->wrec = nullptr;
      H_TRY(link(h, st.chain, st.w));       // every dy / activation copy of the group is final
      {
        OnStream os(h, st.w);
        H_TRY(multi_flush(h, *wrec));
      }
      cudaEvent_t e = next_event(h);
      H_CUDA(cudaEventRecord(e, st.w));
      for (int pt : wpending) side_done[(size_t)pt * 3] = e;
      chain_ready.swap(wpending);
      wpending.clear();
      wrec.reset();
    }
    tl_mark(h, st.chain, "main: bwd chain step done", t);
    if (fork) { tl_mark(h, st.w, "W: chain wgrads done", t); tl_mark(h, st.rec, "R: recognition bwd done", t); tl_mark(h, st.recw, "W2: recognition wgrads done", t); }
    if (!multi) {
      if (fork) {
        cudaStream_t sd[3] = {st.w, st.rec, st.recw};
        for (int i = 0; i < 3; ++i) {
          side_done[(size_t)t * 3 + i] = next_event(h);
          H_CUDA(cudaEventRecord(side_done[(size_t)t * 3 + i], sd[i]));
        }
      }
      H_TRY(allreduce_bucket(h, t, st));
      H_TRY(update_bucket(h, t, st));
    } else if (!wdefer && pending.size() == 1 && pending[0] == t && (group == 1 || t == 0)) {
      // a group of one, issued in this iteration: the whole slice is final together, as without batching
      if (st.lat != st.rec) H_TRY(link(h, st.lat, st.rec));
      cudaStream_t sd[3] = {st.w, st.rec, st.recw};
      for (int i = 0; i < 3; ++i)
        if (sd[i] != st.chain) {
          side_done[(size_t)t * 3 + i] = next_event(h);
          H_CUDA(cudaEventRecord(side_done[(size_t)t * 3 + i], sd[i]));
        }
      H_TRY(allreduce_bucket(h, t, st));
      H_TRY(update_bucket(h, t, st));
      pending.clear();
      ++gi;
      group = group_size();
    } else {
      // chain half of the slice: final once the chain, its weight-gradient stream and the latent projections are done with step t
      BwdStreams ch{st.chain, st.w, st.lat, st.w, st.lat};
      if (!wdefer && st.w != st.chain) {
        side_done[(size_t)t * 3] = next_event(h);
        H_CUDA(cudaEventRecord(side_done[(size_t)t * 3], st.w));
      }
      if (st.lat != st.chain) {
        side_done[(size_t)t * 3 + 2] = next_event(h);
        H_CUDA(cudaEventRecord(side_done[(size_t)t * 3 + 2], st.lat));
      }
      for (int pt : chain_ready) {
        H_TRY(allreduce_bucket(h, pt, ch, 1));
        H_TRY(update_bucket(h, pt, ch, 1));
      }
      if ((int)pending.size() == group || t == 0) {
        // recognition halves of the group that was just issued on st.rec
        BwdStreams rs{st.rec, st.recw, st.rec, st.recw, st.rec};
        if (st.recw != st.rec) H_TRY(link(h, st.recw, st.rec));   // the scratch sets of the group are free when both lanes are done
        cudaEvent_t e = nullptr;
        if (st.rec != st.chain) { e = next_event(h); H_CUDA(cudaEventRecord(e, st.rec)); }
        for (int pt : pending) {
          side_done[(size_t)pt * 3 + 1] = e;
          H_TRY(allreduce_bucket(h, pt, rs, 2));
          H_TRY(update_bucket(h, pt, rs, 2));
        }
        pending.clear();
        ++gi;
        group = group_size();
      }
    }
    gx_in = gx_prev;
    cur ^= 1;
  }
  if (fork)   // join: everything that follows on the main stream (Adam) sees all gradients
  {
    H_TRY(link(h, st.w, h->stream));
    H_TRY(link(h, st.rec, h->stream));
    H_TRY(link(h, st.recw, h->stream));
    H_TRY(link(h, st.lat, h->stream));
  }
  if (h->comm != nullptr) {
    H_CUDA(cudaEventRecord(h->comm_done, h->comm_stream));
    H_CUDA(cudaStreamWaitEvent(h->stream, h->comm_done, 0));
  } else if (h->bucket_update) {
    H_TRY(link(h, h->upd_stream, h->stream));   // the next step reads the updated weights and operand copies
  }
  if (h->bucket_update) h->weights_dirty = false;
  // homogeneous chain: every slice of a shared variable receives the gradient summed over the chain steps (and, when data
  // parallel, over the ranks: the per-step buckets above were all-reduced first; both sums commute)
  if (h->tied) {
    h->cur = h->stream;
    LaunchCtx lc = h->lc();
    for (const TieRun& r : h->tie_runs) H_TRY(tie_reduce(lc, h->G, r));
  }
  h->have_fwd = false;
  return 0;
}

int adam_impl(svae_handle* h, float lr) {
  h->adam_t += 1;
  const double b1 = h->cfg.adam_beta1, b2 = h->cfg.adam_beta2;
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow(b2, (double)h->adam_t)) / (1.0 - pow(b1, (double)h->adam_t)));
  h->cur = h->stream;
  h->dyn_host.lr_t = lr_t;
  H_TRY(dyn_push(h));
  LaunchCtx lc = h->lc();
  H_TRY(adam_update(lc, h->P, h->G, h->M, h->V, h->arena_numel, h->dyn_dev, lr_t, h->cfg.adam_beta1, h->cfg.adam_beta2,
                    h->cfg.adam_eps, h->cfg.clip_value, 1.f / (float)h->nranks));
  h->weights_dirty = true;
  tl_mark(h, h->stream, "main: adam done", -1);
  return 0;
}

int read_losses_impl(svae_handle* h, svae_losses* out) {
  const int T = h->T;
  H_CUDA(cudaMemcpyAsync(h->loss_host, h->loss_sums, sizeof(double) * 2 * T, cudaMemcpyDeviceToHost, h->stream));
  H_CUDA(cudaStreamSynchronize(h->stream));
  memset(out, 0, sizeof *out);
  const int B = h->last_B;
  double total = 0;
  for (int t = 0; t < T; ++t) {
    double recon = h->loss_host[t] / ((double)B * h->D * h->D * h->C);
    double kl = h->loss_host[T + t] / ((double)B * h->Z);
    out->recon[t] = (float)recon;
    out->kl[t] = (float)kl;
    if (step_has_recon(h, t)) total += 16.0 * recon;
    if (step_has_kl(h, t)) total += (double)h->last_reg * kl;
    if (t == 0) total *= h->cfg.first_step_loss_coeff;
  }
  out->total = (float)total;
  out->final_recon = out->recon[T - 1];
  return 0;
}

// ---- tcgen05 weight packing ------------------------------------------------------------------------------------------
void for_each_block(svae_handle* h, void (*fn)(svae_handle*, Block&, void*), void* ctx) {
  for (Step& s : h->steps) {
    for (Block& b : s.inf) fn(h, b, ctx);
    for (Block& b : s.enc) fn(h, b, ctx);
    if (s.t > 0) fn(h, s.encfc, ctx);
    for (Block& b : s.lat) fn(h, b, ctx);
    fn(h, s.decfc, ctx);
    for (Block& b : s.ta) fn(h, b, ctx);
    for (Block& b : s.tb) fn(h, b, ctx);
    fn(h, s.outb, ctx);
    if (s.t > 0) fn(h, s.gateb, ctx);
    if (s.t > 0 && s.merge_og) fn(h, s.ogb, ctx);
  }
}

void plan_pack(svae_handle* h, Block& b, void* ctx) {
  Arena* a = (Arena*)ctx;
  if (h->cfg.operand_dtype != SVAE_OPERAND_BF16) return;
  Geom f = b.g; f.B = h->cfg.max_batch;
  const bool is_fc = f.KH == 1;   // fc layers read the fp32 master weights directly (kernels_fc.cu): nothing to pack
  if (tc_supported(f)) {
    b.tc_fwd = true;
    if (!is_fc) b.w_packed = a->get<char>(tc_packed_bytes(f));
    if (!is_fc && b.tc2_fwd) b.tw_f = tc2_pick_ntw(f, h->sm_count);
    if (a->base) h->tc_layers++;
  }
  if (tc_wgrad_supported(f)) {
    b.tc_wgrad = true;
    if (a->base) h->tc_layers++;
  }
  Geom d = dgrad_geom(b.g); d.B = h->cfg.max_batch;
  if (tc_supported(d)) {
    b.tc_dgrad = true;
    if (!is_fc) b.w_packed_d = a->get<char>(tc_packed_bytes(d));
    if (!is_fc && b.tc2_dgrad) b.tw_d = tc2_pick_ntw(d, h->sm_count);
    if (a->base) h->tc_layers++;
  }
}

void collect_pack(svae_handle* h, Block& b, void* ctx) {
  std::vector<TcPackEntry>* v = (std::vector<TcPackEntry>*)ctx;
  if (b.tc_fwd && b.w_packed) { Geom f = b.g; f.B = 1; v->push_back(tc_pack_entry(f, h->pw(b.w), b.w_packed, b.tw_f)); }
  if (b.tc_dgrad && b.w_packed_d) { Geom d = dgrad_geom(b.g); d.B = 1; v->push_back(tc_pack_entry(d, h->pw(b.w), b.w_packed_d, b.tw_d)); }
}

int ensure_pack_table(svae_handle* h) {
  if (h->pack_table != nullptr || h->pack_bytes == 0) return 0;
  std::vector<TcPackEntry> v;
  h->pack_step_begin.assign(h->T + 1, 0);
  h->pack_step_elems.assign(h->T, 0.0);
  h->pack_rec_end.assign(h->T, 0);
  for (int t = 0; t < h->T; ++t) {     // for_each_block order, one step at a time: every step's entries are contiguous
    Step& s = h->steps[t];
    h->pack_step_begin[t] = (int)v.size();
    for (Block& b : s.inf) collect_pack(h, b, &v);
    h->pack_rec_end[t] = (int)v.size();
    for (Block& b : s.enc) collect_pack(h, b, &v);
    if (s.t > 0) collect_pack(h, s.encfc, &v);
    for (Block& b : s.lat) collect_pack(h, b, &v);
    collect_pack(h, s.decfc, &v);
    for (Block& b : s.ta) collect_pack(h, b, &v);
    for (Block& b : s.tb) collect_pack(h, b, &v);
    if (s.t > 0 && s.merge_og) {   // one packed operand from the two parameter tensors (the separate forms are not used)
      Block& b = s.ogb;
      if (b.tc_fwd && b.w_packed) { Geom f = b.g; f.B = 1; v.push_back(tc_pack_entry(f, h->pw(s.w_out), b.w_packed, b.tw_f, h->pw(s.w_gate), h->C)); }
      if (b.tc_dgrad && b.w_packed_d) { Geom d = dgrad_geom(b.g); d.B = 1; v.push_back(tc_pack_entry(d, h->pw(s.w_out), b.w_packed_d, b.tw_d, h->pw(s.w_gate), h->C)); }
    } else {
      collect_pack(h, s.outb, &v);
      if (s.t > 0) collect_pack(h, s.gateb, &v);
    }
    for (int i = h->pack_step_begin[t]; i < (int)v.size(); ++i) h->pack_step_elems[t] += (double)v[i].total;
  }
  h->pack_step_begin[h->T] = (int)v.size();
  h->pack_entries = (int)v.size();
  h->pack_entry_elems.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) h->pack_entry_elems[i] = (double)v[i].total;
  if (v.empty()) return 0;
  H_CUDA(cudaMalloc(&h->pack_table, sizeof(TcPackEntry) * v.size()));
  H_CUDA(cudaMemcpy(h->pack_table, v.data(), sizeof(TcPackEntry) * v.size(), cudaMemcpyHostToDevice));
  return 0;
}

// fp32 master weights -> packed bf16 operand copies of every tensor-core layer, ONE launch
int repack_if_dirty(svae_handle* h) {
  if (!h->weights_dirty) return 0;
  if (h->cfg.operand_dtype == SVAE_OPERAND_BF16 && h->pack_bytes > 0) {
    H_TRY(ensure_pack_table(h));
    LaunchCtx lc = h->lc();
    H_TRY(tc_pack_batched(lc, h->pack_table, h->pack_entries, (double)h->pack_bytes / 2));
  }
  h->weights_dirty = false;
  return 0;
}

void drop_graph(svae_handle* h, GraphEntry& e) { (void)h; cudaGraphExecDestroy(e.exec); }
void drop_graphs(svae_handle* h) {
  for (GraphEntry& e : h->graphs) drop_graph(h, e);
  h->graphs.clear();
}

void destroy_impl(svae_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int i = 0; i < 4; ++i) if (h->side[i]) cudaStreamSynchronize(h->side[i]);
  if (h->upd_stream) cudaStreamSynchronize(h->upd_stream);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  // captured graphs hold the communicator's collectives: they must go before the communicator does
  drop_graphs(h);
  if (h->comm && h->nccl) h->nccl->CommDestroy(h->comm);
  if (h->chain_noise_dev) cudaFree(h->chain_noise_dev);
  for (cudaEvent_t e : h->bucket_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->dyn_ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < 4; ++i) if (h->side[i]) cudaStreamDestroy(h->side[i]);
  if (h->upd_stream) cudaStreamDestroy(h->upd_stream);
  cudaFree(h->dyn_dev);
  if (h->dyn_ring) cudaFreeHost(h->dyn_ring);
  if (h->comm_done) cudaEventDestroy(h->comm_done);
  cudaFree(h->P); cudaFree(h->G); cudaFree(h->M); cudaFree(h->V);
  cudaFree(h->bf_base); cudaFree(h->act_base); cudaFree(h->zf_base); cudaFree(h->zb_base); cudaFree(h->grad_base); cudaFree(h->pack_base); cudaFree(h->pack_table);
  cudaFree(h->io_steps); cudaFree(h->loss_sums);
  if (h->loss_host) cudaFreeHost(h->loss_host);
  if (h->pin_x) cudaFreeHost(h->pin_x);
  if (h->pin_tgt) cudaFreeHost(h->pin_tgt);
  if (h->pin_eps) cudaFreeHost(h->pin_eps);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  delete h;
}

int ensure_io_steps(svae_handle* h, size_t elems) {
  if (h->io_steps_elems >= elems) return 0;
  if (h->io_steps) cudaFree(h->io_steps);
  h->io_steps = nullptr; h->io_steps_elems = 0;
  H_CUDA(cudaMalloc(&h->io_steps, elems * 4));
  h->io_steps_elems = elems;
  return 0;
}

int stage_in(svae_handle* h, float* dev, float** pin, const float* host, size_t elems, size_t cap_elems) {
  // Caller's buffer already page-locked (cudaHostAlloc / cudaHostRegister, e.g. a torch pinned tensor viewed as numpy): the
  // DMA engine reads it directly.  Pageable memory goes through the handle's pinned staging buffer, which keeps the copy
  // asynchronous but costs one host memcpy (~1 ms for the 9.8 MB of a CelebA B=100 input + target).
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
    H_CUDA(cudaMemcpyAsync(dev, host, elems * 4, cudaMemcpyHostToDevice, h->stream));
    return 0;
  }
  cudaGetLastError();
  if (*pin == nullptr) H_CUDA(cudaMallocHost((void**)pin, cap_elems * 4));
  memcpy(*pin, host, elems * 4);
  H_CUDA(cudaMemcpyAsync(dev, *pin, elems * 4, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

}  // namespace

// ======================================================= C ABI ========================================================
extern "C" {

const char* svae_version(void) { return "svae-b200 0.1 (sm_100a)"; }

const char* svae_last_error(const svae_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

static int validate_cfg(const svae_config* cfg) {
  const int L = cfg->levels;
  if (L < 2 || L > SVAE_MAX_LEVELS || cfg->mc_steps < 1 || cfg->mc_steps > SVAE_MAX_STEPS || cfg->max_batch < 1 ||
      cfg->height != cfg->width || cfg->height % (1 << L) != 0 || cfg->channels < 1 || cfg->channels > 4)
    return fail(nullptr, SVAE_EINVAL, "unsupported configuration (levels/mc_steps/batch/image size/channels)");
  for (int i = 0; i < L; ++i)
    if (cfg->latent_dims[i] < 1 || cfg->latent_dims[i] > 32)
      return fail(nullptr, SVAE_EINVAL, "latent_dims entries must be in [1,32]");
  for (int i = 1; i < L + 2; ++i)
    if (cfg->filter_sizes[i] < 1) return fail(nullptr, SVAE_EINVAL, "filter_sizes entries must be positive");
  if (cfg->filter_sizes[0] != cfg->channels) return fail(nullptr, SVAE_EINVAL, "filter_sizes[0] must equal channels");
  if (cfg->add_noise_to_chain)
    for (int t = 0; t < cfg->mc_steps; ++t)
      if (!(cfg->noise_stddevs[t] >= 0.f)) return fail(nullptr, SVAE_EINVAL, "noise_stddevs entries must be >= 0");
  return 0;
}

static void fill_info(const svae_handle* h, int i, svae_param_info* o) {
  const Param& p = h->params[h->pub[i].members[0]];
  memset(o, 0, sizeof *o);
  snprintf(o->name, SVAE_NAME_LEN, "%s", h->pub[i].name.c_str());
  o->ndim = p.ndim;
  for (int k = 0; k < 4; ++k) o->shape[k] = p.shape[k];
  o->numel = p.numel; o->offset = p.offset; o->step = p.step; o->flags = p.flags;
}

int svae_param_table(const svae_config* cfg, svae_param_info* out, int capacity) {
  if (!cfg) return fail(nullptr, SVAE_EINVAL, "null argument");
  int vr = validate_cfg(cfg);
  if (vr != 0) return vr;
  svae_handle* h = new svae_handle();
  h->cfg = *cfg;
  h->L = cfg->levels; h->T = cfg->mc_steps; h->D = cfg->height; h->C = cfg->channels;
  build_params(h);
  build_pub(h);
  const int n = (int)h->pub.size();
  if (out)
    for (int i = 0; i < n && i < capacity; ++i) fill_info(h, i, &out[i]);
  delete h;
  return n;
}

int svae_create(const svae_config* cfg, int device, svae_handle** out) {
  if (!cfg || !out) return fail(nullptr, SVAE_EINVAL, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(nullptr, SVAE_ENODEVICE, "no usable CUDA device (libsvae has no CPU fallback)");
  }
  { int vr = validate_cfg(cfg); if (vr != 0) return vr; }
  const int L = cfg->levels;
  svae_handle* h = new svae_handle();
  h->cfg = *cfg;
  h->device = device;
  h->L = L; h->T = cfg->mc_steps; h->D = cfg->height; h->C = cfg->channels;
  h->Z = 0;
  for (int i = 0; i < L; ++i) { h->zoff.push_back(h->Z); h->Z += cfg->latent_dims[i]; }
  h->act_sets = cfg->train_capacity ? h->T : (h->T > 1 ? 2 : 1);
#define C_CUDA(expr)                                                  \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) {                                          \
      svae_set_cuda_error(_e, #expr, __FILE__, __LINE__);             \
      int code = _e == cudaErrorMemoryAllocation ? SVAE_ENOMEM : SVAE_ECUDA; \
      destroy_impl(h);                                                \
      return code;                                                    \
    }                                                                 \
  } while (0)
  C_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  C_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  if (cfg->operand_dtype == SVAE_OPERAND_BF16 && prop.major != 10) {
    destroy_impl(h);
    return fail(nullptr, SVAE_ENODEVICE, "SVAE_OPERAND_BF16 needs an sm_100a (B200) device");
  }
  // the chain (main stream) is the critical path: it gets the highest priority, the side streams the lowest, so that
  // their kernels fill the SMs the chain leaves idle instead of competing with it
  int prio_lo = 0, prio_hi = 0;
  C_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  { const char* e = getenv("SVAE_PRIO"); if (e && e[0] == '0') prio_lo = prio_hi = 0; }
  C_CUDA(cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, prio_hi));
  h->stream = h->own_stream;
  {
    const char* e1 = getenv("SVAE_STREAMS"); const char* e2 = getenv("SVAE_GRAPH");
    h->use_streams = !(e1 && e1[0] == '0');
    h->use_graph = !(e2 && e2[0] == '0');
    const char* e4 = getenv("SVAE_TIMELINE");
    h->timeline = e4 && e4[0] == '1';
    if (h->timeline) h->use_graph = false;
    const char* e3 = getenv("SVAE_FORK_MASK");
    if (e3) h->fork_mask = atoi(e3);
    const char* e7 = getenv("SVAE_FUSE");
    h->use_fuse = e7 && e7[0] == '1';
    const char* e9 = getenv("SVAE_COOP_BN");
    h->use_coop_bn = e9 && e9[0] == '1';
    const char* e6 = getenv("SVAE_PDL");
    h->use_pdl = !(e6 && e6[0] == '0');
    const char* e5 = getenv("SVAE_ABLATE");
    if (e5) h->ablate = atoi(e5);
  }
  if (cfg->train_capacity) {
    for (int i = 0; i < 4; ++i) C_CUDA(cudaStreamCreateWithPriority(&h->side[i], cudaStreamNonBlocking, prio_lo));
    C_CUDA(cudaStreamCreateWithPriority(&h->upd_stream, cudaStreamNonBlocking, prio_lo));
    const char* e8 = getenv("SVAE_BUCKET_UPDATE");
    h->use_bucket_update = !(e8 && e8[0] == '0');
    // chain steps per group, the last entry repeats.  Forward: the chain needs net t at step t, so the first groups are
    // small; backward (steps T-1 .. 0): whatever is issued behind the chain's last step is exposed, so the LAST groups are small.
    auto parse = [](const char* e, const char* dflt, std::vector<int>& out) {
      for (const char* q = e ? e : dflt; *q;) {
        const int v = atoi(q);
        if (v > 0) out.push_back(v);
        while (*q && *q != ',') ++q;
        if (*q == ',') ++q;
      }
      if (out.empty()) out.push_back(1);
    };
    parse(getenv("SVAE_REC_FWD_GROUPS"), "1,1,6", h->rec_fwd_groups);
    parse(getenv("SVAE_REC_BWD_GROUPS"), "5,2,1", h->rec_bwd_groups);
    { const char* eg = getenv("SVAE_WGRAD_GROUPS");   // deferred, batched chain weight gradients (default 4,3: steps 7..4, 3..1, 0); 0: per-step launches
      if (eg && atoi(eg) <= 0) h->wgrad_groups.assign(1, 0); else parse(eg, "4,3", h->wgrad_groups); }
    const char* e15 = getenv("SVAE_UPD_SM");
    if (e15) h->upd_sm = atoi(e15);
    const char* e14 = getenv("SVAE_WGRAD_SM");
    if (e14) h->wgrad_sm = atoi(e14);
    const char* e13 = getenv("SVAE_MULTI_SM_FACTOR");
    if (e13 && atof(e13) > 0) h->multi_sm_factor = (float)atof(e13);
    const char* e10 = getenv("SVAE_MULTI");
    h->use_multi = !(e10 && e10[0] == '0') && cfg->operand_dtype == SVAE_OPERAND_BF16;
  }
  C_CUDA(cudaMalloc((void**)&h->dyn_dev, sizeof(SvaeDyn)));
  C_CUDA(cudaMemset(h->dyn_dev, 0, sizeof(SvaeDyn)));
  C_CUDA(cudaMallocHost((void**)&h->dyn_ring, sizeof(SvaeDyn) * 256));
  h->dyn_ev.assign(256, nullptr);
  build_params(h);
  build_pub(h);
  // tied slices only hold their final gradient after the whole backward: no per-chain-step update
  if (h->tied) h->use_bucket_update = false;

  const size_t pbytes = (size_t)h->arena_numel * 4;
  C_CUDA(cudaMalloc(&h->P, pbytes));
  C_CUDA(cudaMemset(h->P, 0, pbytes));
  if (cfg->train_capacity) {
    C_CUDA(cudaMalloc(&h->G, pbytes)); C_CUDA(cudaMemset(h->G, 0, pbytes));
    C_CUDA(cudaMalloc(&h->M, pbytes)); C_CUDA(cudaMemset(h->M, 0, pbytes));
    C_CUDA(cudaMalloc(&h->V, pbytes)); C_CUDA(cudaMemset(h->V, 0, pbytes));
  }
  // two-pass bump allocation
  {
    Arena act, zf, zb, gr, bfa; size_t max_y = 0;
    { const char* e = getenv("SVAE_TC2"); h->use_tc2 = !(e && e[0] == '0'); }
    build_buffers(h, act, zf, zb, gr, bfa, max_y);
    h->act_bytes = act.off + 256; h->zf_bytes = zf.off + 256; h->zb_bytes = zb.off + 256; h->grad_bytes = gr.off + 256;
    h->bf_bytes = bfa.off + 256;
    C_CUDA(cudaMalloc(&h->bf_base, h->bf_bytes));
    C_CUDA(cudaMemset(h->bf_base, 0, h->bf_bytes));
    C_CUDA(cudaMalloc(&h->act_base, h->act_bytes));
    C_CUDA(cudaMalloc(&h->zf_base, h->zf_bytes));
    C_CUDA(cudaMalloc(&h->zb_base, h->zb_bytes));
    C_CUDA(cudaMalloc(&h->grad_base, h->grad_bytes));
    C_CUDA(cudaMemset(h->act_base, 0, h->act_bytes));
    Arena act2, zf2, zb2, gr2, bfa2; size_t my2 = 0;
    act2.base = h->act_base; zf2.base = h->zf_base; zb2.base = h->zb_base; gr2.base = h->grad_base; bfa2.base = h->bf_base;
    build_buffers(h, act2, zf2, zb2, gr2, bfa2, my2);
  }
  {
    Arena pk;
    for_each_block(h, plan_pack, &pk);
    h->pack_bytes = pk.off;
    if (h->pack_bytes > 0) {
      C_CUDA(cudaMalloc(&h->pack_base, h->pack_bytes + 256));
      Arena pk2; pk2.base = h->pack_base;
      for_each_block(h, plan_pack, &pk2);
    }
  }
  C_CUDA(cudaMallocHost((void**)&h->loss_host, sizeof(double) * 2 * SVAE_MAX_STEPS));
  C_CUDA(cudaMalloc(&h->loss_sums, sizeof(double) * 2 * SVAE_MAX_STEPS));
  C_CUDA(cudaMemset(h->loss_sums, 0, sizeof(double) * 2 * SVAE_MAX_STEPS));
  *out = h;
  return SVAE_OK;
#undef C_CUDA
}

int svae_destroy(svae_handle* h) { destroy_impl(h); return SVAE_OK; }

int svae_set_stream(svae_handle* h, void* s) {
  if (!h) return SVAE_EINVAL;
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  h->cur = nullptr;
  drop_graphs(h);   // captured on the previous stream's dependencies
  return SVAE_OK;
}
int svae_sync(svae_handle* h) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 4; ++i) if (h->side[i]) H_CUDA(cudaStreamSynchronize(h->side[i]));
  if (h->comm_stream) H_CUDA(cudaStreamSynchronize(h->comm_stream));
  if (h->timeline && !h->marks.empty()) {
    // print the most recent step only (marks since the last "step start")
    size_t first = 0;
    for (size_t i = 0; i < h->marks.size(); ++i) if (h->marks[i].t == -1 && h->marks[i].what[6] == 's' && h->marks[i].what[7] == 't') first = i;
    for (size_t i = first; i < h->marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->marks[first].e, h->marks[i].e);
      fprintf(stderr, "TIMELINE %8.3f ms  %s %d\n", ms, h->marks[i].what, h->marks[i].t);
    }
    for (auto& m : h->marks) cudaEventDestroy(m.e);
    h->marks.clear();
  }
  return SVAE_OK;
}

int svae_param_count(const svae_handle* h) { return h ? (int)h->pub.size() : SVAE_EINVAL; }
int svae_param_info_get(const svae_handle* h, int i, svae_param_info* o) {
  if (!h || !o || i < 0 || i >= (int)h->pub.size()) return SVAE_EINVAL;
  fill_info(h, i, o);
  return SVAE_OK;
}
int svae_param_slices(const svae_handle* h, int i, int64_t* offsets_out, int capacity) {
  if (!h || i < 0 || i >= (int)h->pub.size()) return SVAE_EINVAL;
  const PubParam& q = h->pub[i];
  for (int m = 0; m < (int)q.members.size() && m < capacity && offsets_out; ++m) offsets_out[m] = h->params[q.members[m]].offset;
  return (int)q.members.size();
}
// `i` indexes the caller-visible table (svae_handle::pub).  Uploads go to every tied slice; reads take the first one (the
// slices of a shared variable are bit-identical, and after svae_backward each holds the gradient summed over the chain).
static int param_copy(svae_handle* h, float* arena, int i, float* host, bool to_dev) {
  if (!h || !host || i < 0 || i >= (int)h->pub.size() || !arena) return fail(h, SVAE_EINVAL, "bad parameter index or arena");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaStreamSynchronize(h->stream));
  if (h->upd_stream) H_CUDA(cudaStreamSynchronize(h->upd_stream));
  const PubParam& q = h->pub[i];
  if (to_dev) {
    for (int m : q.members) {
      const Param& p = h->params[m];
      H_CUDA(cudaMemcpy(arena + p.offset, host, p.numel * 4, cudaMemcpyHostToDevice));
    }
  } else {
    const Param& p = h->params[q.members[0]];
    H_CUDA(cudaMemcpy(host, arena + p.offset, p.numel * 4, cudaMemcpyDeviceToHost));
  }
  return SVAE_OK;
}
int svae_param_set(svae_handle* h, int i, const float* src) {
  int r = param_copy(h, h ? h->P : nullptr, i, const_cast<float*>(src), true);
  if (r == 0) h->weights_dirty = true;
  return r;
}
int svae_param_get(svae_handle* h, int i, float* dst) { return param_copy(h, h ? h->P : nullptr, i, dst, false); }
int svae_grad_get(svae_handle* h, int i, float* dst) { return param_copy(h, h ? h->G : nullptr, i, dst, false); }
int svae_adam_get(svae_handle* h, int i, float* m, float* v) {
  int r = param_copy(h, h ? h->M : nullptr, i, m, false);
  return r ? r : param_copy(h, h->V, i, v, false);
}
int svae_adam_set(svae_handle* h, int i, const float* m, const float* v) {
  int r = param_copy(h, h ? h->M : nullptr, i, const_cast<float*>(m), true);
  return r ? r : param_copy(h, h->V, i, const_cast<float*>(v), true);
}
int64_t svae_adam_step_count(const svae_handle* h) { return h ? h->adam_t : 0; }
int svae_adam_set_step_count(svae_handle* h, int64_t t) { if (!h) return SVAE_EINVAL; h->adam_t = t; return SVAE_OK; }
void* svae_param_arena(svae_handle* h) { return h ? h->P : nullptr; }
void* svae_grad_arena(svae_handle* h) { return h ? h->G : nullptr; }
int64_t svae_arena_numel(const svae_handle* h) { return h ? h->arena_numel : 0; }
int svae_arena_read(svae_handle* h, int which, int64_t offset, int64_t n, float* host_dst) {
  if (!h || !host_dst || which < 0 || which > 3 || offset < 0 || n < 0 || offset + n > h->arena_numel)
    return fail(h, SVAE_EINVAL, "svae_arena_read: bad arena or range");
  float* arenas[4] = {h->P, h->G, h->M, h->V};
  if (!arenas[which]) return fail(h, SVAE_ESTATE, "svae_arena_read: arena not allocated (handle created without train_capacity)");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaStreamSynchronize(h->stream));
  H_CUDA(cudaMemcpy(host_dst, arenas[which] + offset, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return SVAE_OK;
}

int svae_forward(svae_handle* h, const float* x, const float* tgt, int B, const float* eps, uint64_t seed, float reg,
                 float* mu_out, float* sd_out, float* xs_out) {
  if (!h || !x || !tgt) return fail(h, SVAE_EINVAL, "null argument");
  H_CUDA(cudaSetDevice(h->device));
  H_TRY(bf_guard(h, B));
  return forward_impl(h, x, tgt, B, eps, seed, reg, mu_out, sd_out, xs_out);
}
int svae_backward(svae_handle* h) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  return backward_impl(h);
}
int svae_adam_step(svae_handle* h, float lr) {
  if (!h) return SVAE_EINVAL;
  if (!h->cfg.train_capacity) return fail(h, SVAE_ESTATE, "handle created without train_capacity");
  H_CUDA(cudaSetDevice(h->device));
  return adam_impl(h, lr);
}
// One captured CUDA graph per (x, target, eps, batch) argument set: ~1300 kernel launches on four streams become one
// graph launch.  Everything that changes between iterations (step size, KL coefficient, Philox key/counter) is read by
// the kernels from h->dyn_dev, which a small copy enqueued just before the graph refreshes.
static int train_step_graph(svae_handle* h, const float* x, const float* tgt, int B, const float* eps, uint64_t seed, float lr,
                            float reg) {
  GraphEntry* ge = nullptr;
  for (GraphEntry& e : h->graphs)
    if (e.x == x && e.tgt == tgt && e.eps == eps && e.B == B) ge = &e;
  // host-side state the eager path would have updated
  h->iteration += 1;
  h->adam_t += 1;
  const double b1 = h->cfg.adam_beta1, b2 = h->cfg.adam_beta2;
  h->dyn_host.lr_t = (float)((double)lr * sqrt(1.0 - pow(b2, (double)h->adam_t)) / (1.0 - pow(b1, (double)h->adam_t)));
  h->dyn_host.reg = reg; h->dyn_host.seed = rank_seed(h, seed); h->dyn_host.iteration = h->iteration;
  H_TRY(dyn_push(h));
  if (ge == nullptr) {
    if (h->graphs.size() >= 8) {   // callers that rotate many buffers: drop the oldest
      drop_graph(h, h->graphs.front());
      h->graphs.erase(h->graphs.begin());
    }
    const int64_t adam_t0 = h->adam_t, launches0 = h->launches;
    const bool bucket = h->use_bucket_update && h->upd_stream != nullptr;
    if (!bucket) h->adam_t -= 1;    // adam_impl increments it again below
    // one Adam + one repack at the end: the captured step starts by refreshing the packed operand copies.  Bucketed update:
    // the copies are refreshed per chain step inside the backward, and are fresh when the graph starts (see below).
    H_TRY(repack_if_dirty(h));
    h->weights_dirty = !bucket;
    std::vector<cudaEvent_t> eager_pool;
    eager_pool.swap(h->ev_pool);    // events recorded during capture are kept apart from the eager ones
    h->capturing = true;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    int r = ce == cudaSuccess ? 0 : SVAE_ECUDA;
    if (r == 0) r = forward_impl(h, x, tgt, B, eps, seed, reg, nullptr, nullptr, nullptr);
    h->bucket_update = bucket;
    if (r == 0) r = backward_impl(h);
    h->bucket_update = false;
    if (r == 0 && !bucket) r = adam_impl(h, lr);
    cudaError_t ee = cudaStreamEndCapture(h->stream, &graph);
    h->capturing = false;
    h->ev_pool.swap(eager_pool);
    for (cudaEvent_t e : eager_pool) cudaEventDestroy(e);
    h->adam_t = adam_t0;
    if (r != 0 || ee != cudaSuccess || graph == nullptr) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (r == 0) { svae_set_cuda_error(ee != cudaSuccess ? ee : ce, "stream capture of the train step", __FILE__, __LINE__); h->err = g_err; r = SVAE_ECUDA; }
      return r;
    }
    GraphEntry e{x, tgt, eps, B, nullptr};
    e.bucket = bucket;
    e.launches = h->launches - launches0;
    h->launches = launches0;
    cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { svae_set_cuda_error(ie, "cudaGraphInstantiate", __FILE__, __LINE__); h->err = g_err; return SVAE_ECUDA; }
    h->graphs.push_back(e);
    ge = &h->graphs.back();
  }
  if (ge->bucket) H_TRY(repack_if_dirty(h));   // e.g. after svae_set_params / svae_adam_step: this graph does not repack at its start
  H_CUDA(cudaGraphLaunch(ge->exec, h->stream));
  h->launches += ge->launches;
  h->last_B = B; h->last_reg = reg; h->last_x = x; h->last_tgt = tgt;
  h->have_fwd = false;
  h->weights_dirty = !ge->bucket;
  return SVAE_OK;
}

int svae_train_step(svae_handle* h, const float* x, const float* tgt, int B, const float* eps, uint64_t seed, float lr,
                    float reg) {
  if (!h || !x || !tgt) return fail(h, SVAE_EINVAL, "null argument");
  if (!h->cfg.train_capacity) return fail(h, SVAE_ESTATE, "handle created without train_capacity");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  H_TRY(bf_guard(h, B));
  // the first steps run eagerly (one-off allocations, function attributes); profiling needs per-kernel events
  if (h->use_graph && !h->prof.enabled && h->eager_steps >= 1) return train_step_graph(h, x, tgt, B, eps, seed, lr, reg);
  h->eager_steps += 1;
  h->iteration += 1;
  const bool bucket = h->use_bucket_update && h->upd_stream != nullptr && !h->prof.enabled;
  if (bucket) {   // the step size must be on the device before the first bucket's update: pushed with the forward's scalars
    h->adam_t += 1;
    const double b1 = h->cfg.adam_beta1, b2 = h->cfg.adam_beta2;
    h->dyn_host.lr_t = (float)((double)lr * sqrt(1.0 - pow(b2, (double)h->adam_t)) / (1.0 - pow(b1, (double)h->adam_t)));
    H_TRY(ensure_pack_table(h));
  }
  H_TRY(forward_impl(h, x, tgt, B, eps, seed, reg, nullptr, nullptr, nullptr));
  h->bucket_update = bucket;
  int rb = backward_impl(h);
  h->bucket_update = false;
  H_TRY(rb);
  if (!bucket) H_TRY(adam_impl(h, lr));
  return SVAE_OK;
}
int svae_train_step_host(svae_handle* h, const float* x, const float* tgt, int B, const float* eps, uint64_t seed,
                         float lr, float reg, svae_losses* losses) {
  if (!h || !x || !tgt) return fail(h, SVAE_EINVAL, "null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  const size_t img = (size_t)B * h->D * h->D * h->C, cap = (size_t)h->cfg.max_batch * h->D * h->D * h->C;
  H_TRY(stage_in(h, h->in_x, &h->pin_x, x, img, cap));
  const float* dtgt = h->in_x;
  if (tgt != x) { H_TRY(stage_in(h, h->in_tgt, &h->pin_tgt, tgt, img, cap)); dtgt = h->in_tgt; }
  const float* deps = nullptr;
  if (eps) {
    H_TRY(stage_in(h, h->in_eps, &h->pin_eps, eps, (size_t)h->T * B * h->Z, (size_t)h->T * h->cfg.max_batch * h->Z));
    deps = h->in_eps;
  }
  H_TRY(svae_train_step(h, h->in_x, dtgt, B, deps, seed, lr, reg));
  if (losses) H_TRY(read_losses_impl(h, losses)); else H_CUDA(cudaStreamSynchronize(h->stream));
  return SVAE_OK;
}
int svae_apply_noise(svae_handle* h, const float* x, float* out, int64_t n, float pepper, float salt, float scale, float lo,
                     float hi, uint64_t seed, float* draws) {
  if (!h || !x || !out || n < 0) return fail(h, SVAE_EINVAL, "null argument");
  if (!(pepper >= 0.f && pepper <= 1.f && salt >= 0.f && salt <= 1.f && scale >= 0.f && lo <= hi))
    return fail(h, SVAE_EINVAL, "svae_apply_noise: probabilities must be in [0,1], scale >= 0, clip_lo <= clip_hi");
  if (n == 0) return SVAE_OK;
  H_CUDA(cudaSetDevice(h->device));
  h->cur = h->stream;
  h->pdl_prev = 0;
  LaunchCtx lc = h->lc();
  lc.pdl_state = nullptr;
  return apply_noise(lc, x, out, n, pepper, salt, scale, lo, hi, seed, draws);
}
int svae_apply_noise_host(svae_handle* h, const float* x, float* out, int64_t n, float pepper, float salt, float scale,
                          float lo, float hi, uint64_t seed, float* draws) {
  if (!h || !x || !out || n < 0) return fail(h, SVAE_EINVAL, "null argument");
  if (n == 0) return SVAE_OK;
  H_CUDA(cudaSetDevice(h->device));
  float* buf = nullptr;
  H_CUDA(cudaMallocAsync((void**)&buf, (size_t)n * 4 * (draws ? 4 : 1), h->stream));
  int r = SVAE_OK;
  cudaError_t e = cudaMemcpyAsync(buf, x, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    r = svae_apply_noise(h, buf, buf, n, pepper, salt, scale, lo, hi, seed, draws ? buf + n : nullptr);
    if (r == 0) e = cudaMemcpyAsync(out, buf, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (r == 0 && e == cudaSuccess && draws) e = cudaMemcpyAsync(draws, buf + n, (size_t)n * 12, cudaMemcpyDeviceToHost, h->stream);
  }
  cudaFreeAsync(buf, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (r != 0) return r;
  H_CUDA(e);
  return SVAE_OK;
}
int svae_train_step_host_denoise(svae_handle* h, const float* x, int B, const float* eps, uint64_t seed, float lr, float reg,
                                 float pepper, float salt, float scale, uint64_t noise_seed, float* noisy_out,
                                 svae_losses* losses) {
  if (!h || !x) return fail(h, SVAE_EINVAL, "null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  const size_t img = (size_t)B * h->D * h->D * h->C, cap = (size_t)h->cfg.max_batch * h->D * h->D * h->C;
  H_TRY(stage_in(h, h->in_tgt, &h->pin_tgt, x, img, cap));          // the clean batch: target of every step's ELBO
  H_TRY(svae_apply_noise(h, h->in_tgt, h->in_x, (int64_t)img, pepper, salt, scale, h->cfg.range_lo, h->cfg.range_hi,
                         noise_seed, nullptr));                      // its corruption: the network input
  const float* deps = nullptr;
  if (eps) {
    H_TRY(stage_in(h, h->in_eps, &h->pin_eps, eps, (size_t)h->T * B * h->Z, (size_t)h->T * h->cfg.max_batch * h->Z));
    deps = h->in_eps;
  }
  if (noisy_out) H_CUDA(cudaMemcpyAsync(noisy_out, h->in_x, img * 4, cudaMemcpyDeviceToHost, h->stream));
  H_TRY(svae_train_step(h, h->in_x, h->in_tgt, B, deps, seed, lr, reg));
  if (losses) H_TRY(read_losses_impl(h, losses)); else H_CUDA(cudaStreamSynchronize(h->stream));
  return SVAE_OK;
}
int svae_forward_host(svae_handle* h, const float* x, const float* tgt, int B, const float* eps, uint64_t seed,
                      float reg, float* mu_out, float* sd_out, float* xs_out, float* last_out, svae_losses* losses) {
  if (!h || !x || !tgt) return fail(h, SVAE_EINVAL, "null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  const size_t img = (size_t)B * h->D * h->D * h->C, cap = (size_t)h->cfg.max_batch * h->D * h->D * h->C;
  const size_t bz = (size_t)B * h->Z;
  H_TRY(stage_in(h, h->in_x, &h->pin_x, x, img, cap));
  const float* dtgt = h->in_x;
  if (tgt != x) { H_TRY(stage_in(h, h->in_tgt, &h->pin_tgt, tgt, img, cap)); dtgt = h->in_tgt; }
  const float* deps = nullptr;
  if (eps) {
    H_TRY(stage_in(h, h->in_eps, &h->pin_eps, eps, (size_t)h->T * bz, (size_t)h->T * h->cfg.max_batch * h->Z));
    deps = h->in_eps;
  }
  float *dmu = nullptr, *dsd = nullptr, *dxs = nullptr;
  if (mu_out || sd_out || xs_out) {
    H_TRY(ensure_io_steps(h, (size_t)h->T * (img + 2 * bz)));
    dxs = h->io_steps; dmu = h->io_steps + h->T * img; dsd = dmu + h->T * bz;
  }
  H_TRY(bf_guard(h, B));
  H_TRY(forward_impl(h, h->in_x, dtgt, B, deps, seed, reg, mu_out ? dmu : nullptr, sd_out ? dsd : nullptr,
                     xs_out ? dxs : nullptr));
  if (xs_out) H_CUDA(cudaMemcpyAsync(xs_out, dxs, h->T * img * 4, cudaMemcpyDeviceToHost, h->stream));
  if (mu_out) H_CUDA(cudaMemcpyAsync(mu_out, dmu, h->T * bz * 4, cudaMemcpyDeviceToHost, h->stream));
  if (sd_out) H_CUDA(cudaMemcpyAsync(sd_out, dsd, h->T * bz * 4, cudaMemcpyDeviceToHost, h->stream));
  if (last_out) H_CUDA(cudaMemcpyAsync(last_out, h->steps[h->T - 1].xt, img * 4, cudaMemcpyDeviceToHost, h->stream));
  if (losses) H_TRY(read_losses_impl(h, losses)); else H_CUDA(cudaStreamSynchronize(h->stream));
  return SVAE_OK;
}
int svae_read_losses(svae_handle* h, svae_losses* out) {
  if (!h || !out) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  return read_losses_impl(h, out);
}

int svae_generate(svae_handle* h, int B, const float* z, uint64_t seed, float* out) {
  if (!h || !out) return fail(h, SVAE_EINVAL, "null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  h->cur = h->stream;
  H_TRY(bf_guard(h, B));
  H_TRY(repack_if_dirty(h));
  H_TRY(zero_region(h, h->zf_base, h->zf_bytes));
  LaunchCtx lc = h->lc();
  const size_t img = (size_t)B * h->D * h->D * h->C;
  const size_t bz = (size_t)B * h->Z;
  if (z == nullptr) {  // np.random.normal(size=(batch, latent_dim)) per step (sequential_vae.py:1417-1418), drawn on device
    H_TRY(fill_normal(lc, h->in_eps, (int64_t)h->T * bz, seed, 0));
    z = h->in_eps;
  }
  const float* prev = nullptr;
  for (int t = 0; t < h->T; ++t) {
    Step& s = h->steps[t];
    if (t >= 2 && h->act_sets < h->T) {
      // steps >= 2 reuse step 1's buffers: their batch-norm statistics must restart from zero
      H_TRY(zero_region(h, h->zf_base, h->zf_bytes));
    }
    if (t > 0) H_TRY(encoder_fwd(h, s, B, prev));
    H_TRY(latent_fwd(h, s, B, z + t * bz));
    H_TRY(decoder_fwd(h, s, B, prev, nullptr, out + t * img, nullptr));
    prev = out + t * img;
    if (h->noisy()) {   // generative_sample = generative_mle + reg_coeff (placeholder default 1.0) * stddev * N(0, I), :1090
      H_TRY(chain_noise_step(h, s, B, out + t * img, nullptr, 1.f, seed));
      prev = s.xs;
    }
  }
  h->have_fwd = false;
  h->last_B = B;
  return SVAE_OK;
}
int svae_set_chain_noise_host(svae_handle* h, const float* noise, int B) {
  if (!h) return SVAE_EINVAL;
  if (!h->noisy()) return fail(h, SVAE_ESTATE, "handle created without add_noise_to_chain");
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaStreamSynchronize(h->stream));
  if (h->chain_noise_dev) { cudaFree(h->chain_noise_dev); h->chain_noise_dev = nullptr; h->chain_noise_B = 0; }
  drop_graphs(h);   // captured steps hold the old pointer (or the Philox branch)
  if (noise == nullptr) return SVAE_OK;
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  const size_t n = (size_t)h->T * B * h->D * h->D * h->C;
  H_CUDA(cudaMalloc((void**)&h->chain_noise_dev, n * sizeof(float)));
  H_CUDA(cudaMemcpy(h->chain_noise_dev, noise, n * sizeof(float), cudaMemcpyHostToDevice));
  h->chain_noise_B = B;
  return SVAE_OK;
}
int svae_read_chain_samples_host(svae_handle* h, float* out, int B) {
  if (!h || !out) return fail(h, SVAE_EINVAL, "null argument");
  if (!h->noisy()) return fail(h, SVAE_ESTATE, "handle created without add_noise_to_chain");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  const size_t img = (size_t)B * h->D * h->D * h->C;
  for (int t = 0; t < h->T; ++t) H_CUDA(cudaMemcpyAsync(out + t * img, h->steps[t].xs, img * 4, cudaMemcpyDeviceToHost, h->stream));
  H_CUDA(cudaStreamSynchronize(h->stream));
  return SVAE_OK;
}
int svae_generate_host(svae_handle* h, int B, const float* z, uint64_t seed, float* out) {
  if (!h || !out) return fail(h, SVAE_EINVAL, "null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, SVAE_EINVAL, "batch exceeds max_batch");
  H_CUDA(cudaSetDevice(h->device));
  const size_t img = (size_t)B * h->D * h->D * h->C;
  const float* dz = nullptr;
  if (z) {
    H_TRY(stage_in(h, h->in_eps, &h->pin_eps, z, (size_t)h->T * B * h->Z, (size_t)h->T * h->cfg.max_batch * h->Z));
    dz = h->in_eps;
  }
  H_TRY(ensure_io_steps(h, (size_t)h->T * img));
  H_TRY(svae_generate(h, B, dz, seed, h->io_steps));
  H_CUDA(cudaMemcpyAsync(out, h->io_steps, h->T * img * 4, cudaMemcpyDeviceToHost, h->stream));
  H_CUDA(cudaStreamSynchronize(h->stream));
  return SVAE_OK;
}

int svae_nccl_unique_id(char id_out[128], const char* path) {
  NcclApi* api = nccl_load(path);
  if (!api) return fail(nullptr, SVAE_ENCCL, "libnccl not found: " + nccl_load_error());
  int r = api->GetUniqueId(id_out);
  if (r != 0) return fail(nullptr, SVAE_ENCCL, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
  return SVAE_OK;
}
int svae_comm_init(svae_handle* h, int rank, int nranks, const char id[128], const char* path) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, SVAE_EINVAL, "bad rank/nranks");
  if (!h->cfg.train_capacity) return fail(h, SVAE_ESTATE, "communicator needs a train_capacity handle");
  H_CUDA(cudaSetDevice(h->device));
  h->nccl = nccl_load(path);
  if (!h->nccl) return fail(h, SVAE_ENCCL, "libnccl not found: " + nccl_load_error());
  drop_graphs(h);   // captured without the all-reduces
  int r = h->nccl->CommInitRank(&h->comm, nranks, id, rank);
  if (r != 0) { h->comm = nullptr; return fail(h, SVAE_ENCCL, std::string("ncclCommInitRank: ") + h->nccl->GetErrorString(r)); }
  h->rank = rank; h->nranks = nranks;
  H_CUDA(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  h->bucket_ev.resize(h->T);
  for (int t = 0; t < h->T; ++t) H_CUDA(cudaEventCreateWithFlags(&h->bucket_ev[t], cudaEventDisableTiming));
  H_CUDA(cudaEventCreateWithFlags(&h->comm_done, cudaEventDisableTiming));
  return SVAE_OK;
}
int svae_comm_destroy(svae_handle* h) {
  if (!h) return SVAE_EINVAL;
  if (h->comm && h->nccl) {
    svae_sync(h);
    drop_graphs(h);   // they contain this communicator's collectives
    h->nccl->CommDestroy(h->comm);
  }
  h->comm = nullptr; h->nranks = 1; h->rank = 0;
  return SVAE_OK;
}

int64_t svae_launch_count(const svae_handle* h) { return h ? h->launches : 0; }

int svae_profile_enable(svae_handle* h, int on) {
  if (!h) return SVAE_EINVAL;
  h->prof.enabled = on != 0;
  return SVAE_OK;
}
int svae_profile_read(svae_handle* h, svae_kernel_stats* out, int capacity) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  H_CUDA(cudaStreamSynchronize(h->stream));
  Profiler& p = h->prof;
  const char* trace = getenv("SVAE_TRACE");
  for (const Profiler::Rec& r : p.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) p.ms[r.kc] += ms;
    if (trace && trace[0] == '1' && r.geo[0] > 0)
      fprintf(stderr, "TRACE %s B=%d Hin=%d Cin=%d Hout=%d Cout=%d k=%d s=%d mode=%d ms=%.4f\n", kKClassNames[r.kc],
              r.geo[0], r.geo[1], r.geo[2], r.geo[3], r.geo[4], r.geo[5], r.geo[6], r.geo[7], ms);
  }
  cudaGetLastError();
  p.recs.clear();
  p.pool_used = 0;
  int n = 0;
  for (int k = 0; k < KC_COUNT; ++k) {
    if (out && n < capacity) {
      memset(&out[n], 0, sizeof out[n]);
      snprintf(out[n].name, sizeof out[n].name, "%s", kKClassNames[k]);
      out[n].launches = p.launches[k]; out[n].total_ms = p.ms[k]; out[n].flops = p.flops[k]; out[n].bytes = p.bytes[k];
    }
    ++n;
    p.launches[k] = 0; p.ms[k] = 0; p.flops[k] = 0; p.bytes[k] = 0;
  }
  return n;
}
int64_t svae_activation_bytes(const svae_handle* h) { return h ? (int64_t)(h->act_bytes + h->grad_bytes) : 0; }
int svae_tc_layers(const svae_handle* h) { return h ? h->tc_layers : 0; }

// ---- layer-level entry points --------------------------------------------------------------------------------------
static int op_contract(svae_handle* h, Geom g, int B, const float* x, int ldx, const float* w, float* y, int ldy,
                       double* stats, int operand) {
  g.B = B;
  LaunchCtx lc = h->lc();
  if (stats) H_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * g.Cout, h->stream));
  if (operand == SVAE_OPERAND_BF16) {
    if (!tc_supported(g)) return fail(h, SVAE_EINVAL, "shape not supported by the tcgen05 kernels");
    void* packed = nullptr;
    H_CUDA(cudaMalloc(&packed, tc_packed_bytes(g)));
    const char* e2 = getenv("SVAE_TC2");
    const bool use2 = !(e2 && e2[0] == '0') && tc2_supported(g);
    const int tw = use2 ? tc2_pick_ntw(g, h->sm_count) : 128;
    int r = tc_pack_weights(lc, g, w, packed, tw);
    if (r == 0 && use2) {
      // TMA-fed kernel: stage the fp32 input as a padded bf16 copy first (inside the chain the producing kernel writes it)
      BfAct a = bf_act_describe(tc2_input_kind(g), B, g.Hin, g.Win, g.Cin);
      void* abuf = nullptr;
      H_CUDA(cudaMalloc(&abuf, bf_act_bytes(a)));
      a.p = reinterpret_cast<__nv_bfloat16*>(abuf);
      r = bf_act_fill(lc, a, mkview(const_cast<float*>(x), ldx, 0), g.Cin);
      if (r == 0) r = tc2_gather_gemm(lc, g, a, 0, packed, mkview(y, ldy, 0), stats, nullptr, tw);
      cudaStreamSynchronize(h->stream);
      cudaFree(abuf);
    } else if (r == 0) {
      r = tc_gather_gemm(lc, g, mkview(const_cast<float*>(x), ldx, 0), packed, mkview(y, ldy, 0), stats);
    }
    cudaStreamSynchronize(h->stream);
    cudaFree(packed);
    if (r != 0) { h->err = g_err; return r; }
    return 0;
  }
  H_TRY(simt_gather_gemm(lc, g, mkview(const_cast<float*>(x), ldx, 0), w, mkview(y, ldy, 0), stats));
  return 0;
}


// Weight gradient of the layer-level entry points: the SVAE_OPERAND_BF16 family takes the PRODUCTION kernel (tc2_wgrad, both
// operands staged as bf16 planar copies exactly as the chain's producers write them) whenever the plan would; the
// SIMT-staged tc_wgrad otherwise.  g: conv-gather geometry (X = conv input side, dY = conv output side), B set.
static int op_wgrad(svae_handle* h, const Geom& g, const float* xrole, const float* yrole, float* dw, int operand, const Geom& fwd) {
  LaunchCtx lc = h->lc();
  H_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 16 * g.Cin * g.Cout, h->stream));
  View xv = mkview(const_cast<float*>(xrole), g.Cin, 0), yv = mkview(const_cast<float*>(yrole), g.Cout, 0);
  if (operand != SVAE_OPERAND_BF16 || !tc_wgrad_supported(fwd)) { H_TRY(simt_wgrad(lc, g, xv, yv, dw)); return 0; }
  const char* e2 = getenv("SVAE_TC2");
  if (!(e2 && e2[0] == '0') && tc2_wgrad_supported(g)) {
    BfAct ax = bf_act_describe(tc2_input_kind(g), g.B, g.Hin, g.Win, g.Cin);
    BfAct ay = bf_act_describe(g.stride == 1 ? 0 : 1, g.B, g.Hout, g.Wout, g.Cout);
    void *bx = nullptr, *by = nullptr;
    H_CUDA(cudaMalloc(&bx, bf_act_bytes(ax)));
    if (cudaMalloc(&by, bf_act_bytes(ay)) != cudaSuccess) { cudaFree(bx); return fail(h, SVAE_ENOMEM, "op_wgrad: staging buffer"); }
    ax.p = reinterpret_cast<__nv_bfloat16*>(bx); ay.p = reinterpret_cast<__nv_bfloat16*>(by);
    int r = bf_act_fill(lc, ax, xv, g.Cin);
    if (r == 0) r = bf_act_fill(lc, ay, yv, g.Cout);
    if (r == 0) r = tc2_wgrad(lc, g, ax, ay, dw);
    cudaStreamSynchronize(h->stream);
    cudaFree(bx); cudaFree(by);
    if (r != 0) { h->err = g_err; return r; }
    return 0;
  }
  H_TRY(tc_wgrad(lc, g, xv, yv, dw));
  return 0;
}

int svae_op_conv2d(svae_handle* h, const float* x, const float* w, float* y, double* stats, int B, int H, int W, int Ci,
                   int Co, int stride, int operand) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  Geom g = conv_geom(H, W, Ci, Co, stride);
  return op_contract(h, g, B, x, Ci, w, y, Co, stats, operand);
}
int svae_op_conv2d_transpose(svae_handle* h, const float* x, const float* w, float* y, double* stats, int B, int H,
                             int W, int Ci, int Co, int stride, int operand) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  Geom g = deconv_geom(H, W, Ci, Co, stride);
  return op_contract(h, g, B, x, Ci, w, y, Co, stats, operand);
}
int svae_op_conv2d_backward(svae_handle* h, const float* x, const float* w, const float* dy, float* dx, float* dw, int B,
                            int H, int W, int Ci, int Co, int stride, int operand) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  Geom f = conv_geom(H, W, Ci, Co, stride);
  if (dx) { int r = op_contract(h, dgrad_geom(f), B, dy, Co, w, dx, Ci, nullptr, operand); if (r) return r; }
  if (dw) {
    Geom g = f; g.B = B;
    H_TRY(op_wgrad(h, g, x, dy, dw, operand, f));
  }
  return 0;
}
int svae_op_conv2d_transpose_backward(svae_handle* h, const float* x, const float* w, const float* dy, float* dx,
                                      float* dw, int B, int H, int W, int Ci, int Co, int stride, int operand) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  Geom f = deconv_geom(H, W, Ci, Co, stride);
  if (dx) { int r = op_contract(h, dgrad_geom(f), B, dy, Co, w, dx, Ci, nullptr, operand); if (r) return r; }
  if (dw) {
    Geom g = dgrad_geom(f); g.B = B; g.mode = 0;   // conv geometry from the deconv's output grid (X role: dy) to its input grid
    H_TRY(op_wgrad(h, g, dy, x, dw, operand, f));
  }
  return 0;
}
/* fully_connected data path (abstract_network.py:65; sequential_vae.py:1704,1775): y[B,N] = x[B,K] . w[K,N] */
int svae_op_fc(svae_handle* h, const float* x, const float* w, float* y, int B, int K, int N, int operand) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  Geom g = fc_geom(K, N);
  const bool tc = operand == SVAE_OPERAND_BF16;
  if (tc && !tc_supported(g)) return fail(h, SVAE_EINVAL, "shape not supported by the tcgen05 kernels");
  H_TRY(contract(h, g, B, mkview(const_cast<float*>(x), K, 0), w, nullptr, tc, mkview(y, N, 0), nullptr));
  return 0;
}
int svae_op_fc_backward(svae_handle* h, const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int K,
                        int N, int operand) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  Geom f = fc_geom(K, N);
  const bool tc = operand == SVAE_OPERAND_BF16;
  if (tc && !tc_supported(f)) return fail(h, SVAE_EINVAL, "shape not supported by the tcgen05 kernels");
  if (dx) H_TRY(contract(h, dgrad_geom(f), B, mkview(const_cast<float*>(dy), N, 0), w, nullptr, tc, mkview(dx, K, 0), nullptr));
  if (dw) {
    LaunchCtx lc = h->lc();
    H_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * N, h->stream));
    Geom g = f; g.B = B;
    if (tc) H_TRY(tc_wgrad(lc, g, mkview(const_cast<float*>(x), K, 0), mkview(const_cast<float*>(dy), N, 0), dw));
    else H_TRY(simt_wgrad(lc, g, mkview(const_cast<float*>(x), K, 0), mkview(const_cast<float*>(dy), N, 0), dw));
  }
  return 0;
}

/* Debug probe (parity tests): one tensor of one block of the LAST forward / backward as a dense fp32 host array, exactly
 * as the kernels consumed or produced it (a tensor that only exists as a bf16 planar copy is returned with its bf16 values). */
int svae_debug_block_tensor(svae_handle* h, int t, int net, int index, int which, float* host_dst, int64_t capacity,
                            int32_t dims_out[4]) {
  if (!h || t < 0 || t >= h->T || !dims_out) return fail(h, SVAE_EINVAL, "svae_debug_block_tensor: bad argument");
  H_CUDA(cudaSetDevice(h->device));
  Step& s = h->steps[t];
  const int B = h->last_B, L = h->L, C = h->C;
  if (B <= 0) return fail(h, SVAE_ESTATE, "svae_debug_block_tensor: no forward has run");
  Block* b = nullptr;
  switch (net) {
    case 0: if (index >= 0 && index < (int)s.inf.size()) b = &s.inf[index]; break;
    case 1: if (index >= 0 && index < (int)s.enc.size()) b = &s.enc[index]; break;
    case 2: if (t > 0) b = &s.encfc; break;
    case 3: if (index >= 0 && index < (int)s.lat.size()) b = &s.lat[index]; break;
    case 4: b = &s.decfc; break;
    case 5: if (index >= 0 && index < (int)s.ta.size()) b = &s.ta[index]; break;
    case 6: if (index >= 0 && index < (int)s.tb.size()) b = &s.tb[index]; break;
    case 7: b = &s.outb; break;
    case 8: if (t > 0) b = &s.gateb; break;
    default: break;
  }
  if (b == nullptr) return fail(h, SVAE_EINVAL, "svae_debug_block_tensor: no such block");
  const bool head = net == 7 || net == 8;
  const Geom& g = b->g;
  const int64_t rows = head ? (int64_t)B * g.Hout * g.Wout : (int64_t)B * b->rpi;
  const int feats = head ? g.Cout : b->feats;
  h->cur = h->stream;
  LaunchCtx lc = h->lc();
  lc.pdl_state = nullptr;
  h->pdl_prev = 0;
  // every stream that may still be writing the tensors
  H_TRY(svae_sync(h));
  if (h->upd_stream) H_CUDA(cudaStreamSynchronize(h->upd_stream));
  int64_t n = 0;
  int d[4] = {0, 0, 0, 0};
  enum { SRC_NONE, SRC_BF, SRC_FV } kind = SRC_NONE;
  BfAct bf{}; int bf_coff = 0, bf_C = 0; FeatView fv{}; int64_t fv_rows = 0; int fv_feats = 0;
  const bool bwd = which == 3 || which == 4;
  if (bwd && (!h->cfg.train_capacity || b->dbg_gs < 0)) return fail(h, SVAE_ESTATE, "svae_debug_block_tensor: no backward has run for this block");
  GradSet* gs = bwd ? &h->gs[b->dbg_gs] : nullptr;
  const int ldu = C + (t > 0 ? 1 : 0);
  switch (which) {
    case 0:   // IN: the contraction's input
      if (b->tc2_fwd && b->in_bf.p != nullptr && !b->in_f32_valid) { kind = SRC_BF; bf = b->in_bf; bf_coff = 0; bf_C = g.Cin; d[0] = B; d[1] = g.Hin; d[2] = g.Win; d[3] = g.Cin; }
      else if (head) { kind = SRC_FV; fv = FeatView{s.tb[0].out.p, s.tb[0].feats, 0, s.tb[0].feats, 1}; fv_rows = (int64_t)B * g.Hin * g.Win; fv_feats = g.Cin; d[0] = B; d[1] = g.Hin; d[2] = g.Win; d[3] = g.Cin; }
      else if (b->dbg_in.p != nullptr) { kind = SRC_FV; fv = FeatView{b->dbg_in.p, b->dbg_in.ld, b->dbg_in.coff, g.Cin, 1}; fv_rows = (int64_t)B * g.Hin * g.Win; fv_feats = g.Cin; d[0] = B; d[1] = g.Hin; d[2] = g.Win; d[3] = g.Cin; }
      break;
    case 1:   // Y: pre-BN contraction output (heads: pre-sigmoid, without the bias)
      if (head) { kind = SRC_FV; fv = FeatView{s.u, ldu, net == 8 ? C : 0, feats, 1}; }
      else if (b->y != nullptr) { kind = SRC_FV; fv = FeatView{b->y, feats, 0, feats, 1}; }
      fv_rows = rows; fv_feats = feats; d[0] = B; d[1] = g.Hout; d[2] = g.Wout; d[3] = g.Cout;
      break;
    case 2:   // OUT: activated output
      if (head) break;
      if (b->out.p != nullptr && !b->skip_f32) { kind = SRC_FV; fv = b->out; fv_rows = rows; fv_feats = feats; }
      else if (b->out_bf.a.p != nullptr) {
        kind = SRC_BF; bf = b->out_bf.a; bf_coff = b->out_bf.coff;
        bf_C = b->out_bf.inner ? b->out_bf.inner : b->out.inner;
      }
      d[0] = B; d[1] = g.Hout; d[2] = g.Wout; d[3] = g.Cout;
      if (kind == SRC_BF) { d[1] = bf.H; d[2] = bf.W; d[3] = bf_C; }
      break;
    case 3:   // DA: dL/d(activated output) as the block's backward read it
      if (head || b->dbg_da.p == nullptr) break;
      kind = SRC_FV; fv = b->dbg_da; fv_rows = rows; fv_feats = feats; d[0] = B; d[1] = g.Hout; d[2] = g.Wout; d[3] = g.Cout;
      break;
    case 4: { // DY: dL/d(pre-BN output) as the input / weight gradient kernels read it
      d[0] = B; d[1] = g.Hout; d[2] = g.Wout; d[3] = g.Cout;
      if (head) {
        kind = SRC_FV; fv = FeatView{gs->d_u, ldu, net == 8 ? C : 0, feats, 1}; fv_rows = rows; fv_feats = feats;   // always written in fp32
        break;
      }
      if (b->dy_slot < 0) break;
      const bool have_bf = gs->dy_bf[b->dy_slot].p != nullptr && (b->tc2_dgrad || b->tc2_wgrad) && !b->dbg_dy_f32;
      if (have_bf) { kind = SRC_BF; bf = gs->dy_bf[b->dy_slot]; bf_coff = 0; bf_C = g.Cout; }
      else if (gs->dy[b->dy_slot] != nullptr && !(net == 3 && lat_fused_supported(B, g.Cin))) {
        kind = SRC_FV; fv = FeatView{gs->dy[b->dy_slot], feats, 0, feats, 1}; fv_rows = rows; fv_feats = feats;
      }
      break;
    }
    case 5:   // RES: tensor added before the activation
      if (!head && b->res.p != nullptr) { kind = SRC_FV; fv = b->res; fv_rows = rows; fv_feats = feats; d[0] = B; d[1] = g.Hout; d[2] = g.Wout; d[3] = g.Cout; }
      break;
    default: break;
  }
  if (kind == SRC_NONE) return fail(h, SVAE_EINVAL, "svae_debug_block_tensor: this tensor is not materialised for this block");
  n = kind == SRC_BF ? (int64_t)B * bf.H * bf.W * bf_C : fv_rows * fv_feats;
  for (int i = 0; i < 4; ++i) dims_out[i] = d[i];
  if (host_dst == nullptr) return SVAE_OK;                       // size query
  if (capacity < n) return fail(h, SVAE_EINVAL, "svae_debug_block_tensor: destination too small");
  float* tmp = nullptr;
  H_CUDA(cudaMalloc(&tmp, (size_t)n * 4));
  int r = kind == SRC_BF ? probe_bf_unpack(lc, bf, bf_coff, bf_C, B, tmp) : probe_fv_gather(lc, fv, fv_rows, fv_feats, tmp);
  cudaError_t e = cudaSuccess;
  if (r == 0) e = cudaMemcpyAsync(host_dst, tmp, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  if (r != 0) { h->err = g_err; return r; }
  H_CUDA(e);
  return SVAE_OK;
}
extern void* g_tc_debug_buffer;
int svae_debug_set_buffer(void* dev_buffer) { g_tc_debug_buffer = dev_buffer; return 0; }

int svae_op_tc_supported(int transposed, int H, int W, int Ci, int Co, int stride, int direction) {
  Geom f = transposed == 2 ? fc_geom(Ci, Co) : transposed ? deconv_geom(H, W, Ci, Co, stride) : conv_geom(H, W, Ci, Co, stride);
  f.B = 1;
  if (direction == 0) return tc_supported(f) ? 1 : 0;
  if (direction == 1) { Geom d = dgrad_geom(f); d.B = 1; return tc_supported(d) ? 1 : 0; }
  return tc_wgrad_supported(f) ? 1 : 0;
}
/* 1 when the SVAE_OPERAND_BF16 family runs this contraction on the TMA-fed production kernels (tc2_conv_kernel for direction
 * 0 / 1, tc2_wgrad_kernel for direction 2) at batch B, i.e. when the layer-level entry points exercise exactly the kernels of
 * the chain; 0 when it takes the SIMT-staged tcgen05 variant or the fp32 kernels. */
int svae_op_tc2_supported(int transposed, int B, int H, int W, int Ci, int Co, int stride, int direction) {
  if (transposed == 2) return 0;
  { const char* e2 = getenv("SVAE_TC2"); if (e2 && e2[0] == '0') return 0; }
  Geom f = transposed ? deconv_geom(H, W, Ci, Co, stride) : conv_geom(H, W, Ci, Co, stride);
  f.B = B;
  if (direction == 0) return tc_supported(f) && tc2_supported(f) ? 1 : 0;
  if (direction == 1) { Geom d = dgrad_geom(f); d.B = B; return tc_supported(d) && tc2_supported(d) ? 1 : 0; }
  Geom g = f;
  if (transposed) { g = dgrad_geom(f); g.mode = 0; g.B = B; }
  return tc_wgrad_supported(f) && tc2_wgrad_supported(g) ? 1 : 0;
}
int svae_op_bn_act(svae_handle* h, const float* y, const float* beta, float* out, int64_t rows, int C, int act) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  double* stats = nullptr;
  H_CUDA(cudaMalloc(&stats, sizeof(double) * 2 * C));
  H_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, h->stream));
  LaunchCtx lc = h->lc();
  int r = col_stats(lc, y, rows, C, stats);
  if (r == 0) r = bn_act_fwd(lc, y, stats, beta, rows, C, act, FeatView{}, FeatView{out, C, 0, C, 1});
  cudaStreamSynchronize(h->stream);
  cudaFree(stats);
  if (r != 0) { h->err = g_err; return r; }
  return 0;
}
int svae_op_bn_act_backward(svae_handle* h, const float* da, const float* y, const float* beta, const float* residual,
                            float* dy, float* dbeta, float* dres, int dres_accumulate, int64_t rows, int C, int act) {
  if (!h || !da || !y || !beta || !dy) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  double* stats = nullptr;   // [2C] forward statistics, [2C] backward sums
  H_CUDA(cudaMalloc(&stats, sizeof(double) * 4 * C));
  H_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 4 * C, h->stream));
  LaunchCtx lc = h->lc();
  int r = col_stats(lc, y, rows, C, stats);
  const FeatView resv = residual ? FeatView{const_cast<float*>(residual), C, 0, C, 1} : FeatView{};
  if (r == 0)
    r = bn_bwd_reduce(lc, FeatView{const_cast<float*>(da), C, 0, C, 1}, y, stats, beta, rows, C, act, resv, dy, stats + 2 * C,
                      dres, dres_accumulate);
  if (r == 0) r = bn_bwd_apply(lc, dy, y, stats, stats + 2 * C, rows, C, dbeta);
  cudaStreamSynchronize(h->stream);
  cudaFree(stats);
  if (r != 0) { h->err = g_err; return r; }
  return 0;
}
int svae_op_adam(svae_handle* h, float* p, const float* g, float* m, float* v, int64_t n, float lr, int64_t t, float b1,
                 float b2, float eps, float clip, float gscale) {
  if (!h) return SVAE_EINVAL;
  H_CUDA(cudaSetDevice(h->device));
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  LaunchCtx lc = h->lc();
  H_TRY(adam_update(lc, p, g, m, v, n, nullptr, lr_t, b1, b2, eps, clip, gscale));
  return 0;
}

}  // extern "C"
