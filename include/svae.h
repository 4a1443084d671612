/* svae.h - C ABI of libsvae.so: the B200-native Sequential-VAE hot path (train step, forward probes, generation).
 *
 * The reference (MWPainter/Sequential-Variational-Autoencoder) is pure Python on TensorFlow 1.x and has no FFI of
 * its own; the boundary this library sits behind is the Python class surface that main.py / trainer.py call, whose
 * device work is a single `Session.run`.  Each entry point below names the reference interface it replaces
 * (file:line in /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types; every function returns 0 on success or a negative
 *     SVAE_E* code; svae_last_error() returns a human-readable message for the last failure on that handle.
 *   - tensors are fp32, NHWC, dense.  "dev" pointers are CUDA device pointers owned by the caller; "host"
 *     pointers are ordinary host memory (pinned or pageable).
 *   - one handle per device per process; a handle is not thread-safe.  All device work is ordered on the handle's
 *     stream (svae_set_stream); only the *_host entry points, svae_sync and svae_read_losses synchronise.
 *   - there is NO CPU fallback: without a CUDA device svae_create fails with SVAE_ENODEVICE.
 */
#ifndef SVAE_H_
#define SVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVAE_MAX_LEVELS 8
#define SVAE_MAX_STEPS 64
#define SVAE_NAME_LEN 128

enum {
  SVAE_OK = 0,
  SVAE_EINVAL = -1,     /* bad argument / unsupported configuration */
  SVAE_ENODEVICE = -2,  /* no usable CUDA device */
  SVAE_ECUDA = -3,      /* CUDA runtime error (message in svae_last_error) */
  SVAE_ENOMEM = -4,     /* device allocation failed */
  SVAE_ENCCL = -5,      /* NCCL missing or failed */
  SVAE_ESTATE = -6      /* call sequence error (e.g. backward without forward) */
};

/* operand precision of the conv/deconv contractions (accumulation is always fp32) */
enum {
  SVAE_OPERAND_FP32 = 0, /* fp32 SIMT kernels: strict-parity mode */
  SVAE_OPERAND_BF16 = 1  /* bf16 operands on tcgen05 tensor cores, fp32 TMEM accumulators */
};

/* Hyper-parameters of one SequentialVAE instance.
 * Replaces: the attribute block of SequentialVAE.__init__, sequential_vae.py:197-258 (+ netname rows :281-862). */
typedef struct svae_config {
  int32_t height, width, channels;          /* dataset.data_dims                         sequential_vae.py:197   */
  int32_t levels;                           /* vlae_levels                               :202                    */
  int32_t latent_dims[SVAE_MAX_LEVELS];     /* vlae_latent_dims                          :203                    */
  int32_t filter_sizes[SVAE_MAX_LEVELS + 2];/* filter_sizes (levels+2 entries)           :207                    */
  int32_t mc_steps;                         /* mc_steps                                  :219                    */
  int32_t intermediate_reconstruction;      /* :221                                                              */
  uint64_t regularized_mask;                /* bit t set <=> t in regularized_steps      :224                    */
  float first_step_loss_coeff;              /* :226                                                              */
  float latent_mean_clip;                   /* :229 (INFINITY = no clip)                                         */
  float prior_stddev;                       /* latent_prior_stddev :231                                          */
  float min_highway, max_highway;           /* :240-241                                                          */
  float range_lo, range_hi;                 /* dataset.range                             :1721                   */
  float clip_value;                         /* clip_grad_value (<=0: clip_grads False)   :257-258                */
  float adam_beta1, adam_beta2, adam_eps;   /* tf.train.AdamOptimizer defaults           :1267                   */
  int32_t max_batch;                        /* largest batch any call will use                                   */
  int32_t train_capacity;                   /* 1: keep per-step activations for backward; 0: forward/generate only */
  int32_t operand_dtype;                    /* SVAE_OPERAND_*                                                    */
  int32_t share_theta_weights;              /* :213  homogeneous chain: one chain encoder and one decoder for all
                                               steps t >= 1 (step 0 keeps its own decoder, :1683-1687,1757-1761)  */
  int32_t share_phi_weights;                /* :214  one recognition net for all steps (:1573-1577)              */
  int32_t add_noise_to_chain;               /* :233  x_t fed to step t+1 = mle_t + reg_coeff * noise_stddevs[t] * N(0,I)
                                               (:1088-1091; netnames c_sample_images, c_homog_sample_images :761-767)  */
  int32_t reserved[5];
  float noise_stddevs[SVAE_MAX_STEPS];      /* :239  fixed per-step stddev of the chain noise (mc_steps entries used)   */
} svae_config;

/* Per-step ELBO terms of the last forward.  Replaces the scalars TF lets callers fetch: self.loss, self.final_loss
 * (sequential_vae.py:1168-1176,1204) and the per-step summaries reconstruction_loss_step_t / regularization_loss_step_t
 * (:1207-1208). */
typedef struct svae_losses {
  float total;                  /* self.loss                                   */
  float final_recon;            /* self.final_loss = recon[T-1]                */
  float recon[SVAE_MAX_STEPS];  /* mean_b mean_hwc (x_t - target)^2            */
  float kl[SVAE_MAX_STEPS];     /* mean_b mean_j KL(N(mu,sigma) || N(0,prior)) */
} svae_losses;

typedef struct svae_param_info {
  char name[SVAE_NAME_LEN];     /* TF variable name, e.g. "phi/inference_step_0/Conv/weights"  */
  int32_t ndim;
  int32_t shape[4];
  int64_t numel;
  int64_t offset;               /* element offset inside the flat parameter / gradient / Adam arenas */
  int32_t step;                 /* chain step that owns it (gradient all-reduce bucket) */
  int32_t flags;                /* SVAE_PF_* */
} svae_param_info;

enum {
  SVAE_PF_THETA = 1,  /* generative (theta/...) variable, else recognition (phi/...)                               */
  SVAE_PF_INERT = 2,  /* BN-shadowed bias: kept for checkpoint compatibility, never read, gradient exactly 0 (Q2)   */
  SVAE_PF_DEAD = 4,   /* belongs to the reference's dead recognition branch (Q3): never read, no gradient          */
  SVAE_PF_XAVIER = 8  /* reference initialiser is xavier-uniform (heads, output deconvs) instead of N(0,0.02)        */
};

typedef struct svae_handle svae_handle;

/* ---- lifetime -------------------------------------------------------------------------------------------------
 * Replaces SequentialVAE.__init__ -> construct_network/init_network (sequential_vae.py:81,864-870;
 * abstract_network.py:85-107,139-152): builds the static plan, allocates weights (zero-filled; the host uploads
 * initial values with svae_param_set), Adam slots and the activation arena. */
int svae_create(const svae_config* cfg, int device, svae_handle** out);
int svae_destroy(svae_handle* h);
const char* svae_last_error(const svae_handle* h); /* h may be NULL: last creation error */
const char* svae_version(void);

/* Use `cuda_stream` (a cudaStream_t cast to void*) for all subsequent work; NULL restores the handle's own stream. */
int svae_set_stream(svae_handle* h, void* cuda_stream);
int svae_sync(svae_handle* h);

/* ---- parameters (tf.trainable_variables(), sequential_vae.py:1225; Saver save/restore abstract_network.py:124-152) */
int svae_param_count(const svae_handle* h);
/* Device-free: the parameter table a handle built from `cfg` would have.  Returns the number of parameters (or a
 * negative SVAE_E* code) and fills up to `capacity` entries of `out` (may be NULL). */
int svae_param_table(const svae_config* cfg, svae_param_info* out, int capacity);
int svae_param_info_get(const svae_handle* h, int index, svae_param_info* out);
/* Homogeneous chains (share_theta_weights / share_phi_weights): a shared variable appears ONCE in the table, under the
 * reference's shared scope name ("phi/inference_network/...", "theta/generative_encoder_network/...",
 * "theta/generative_network/..."; sequential_vae.py:1573-1577,1683-1687,1757-1761).  Inside the arenas it is kept as one
 * slice per chain step that uses it; the slices are tied (equal values, equal Adam slots, each receives the gradient
 * summed over the chain).  svae_param_slices returns the number of slices of variable `index` and writes up to `capacity`
 * of their arena element offsets (the first equals svae_param_info.offset); 1 for a variable that is not shared. */
int svae_param_slices(const svae_handle* h, int index, int64_t* offsets_out, int capacity);
int svae_param_set(svae_handle* h, int index, const float* host_src);   /* reference layout (HWIO / [kh,kw,out,in] / [in,out]) */
int svae_param_get(svae_handle* h, int index, float* host_dst);
int svae_grad_get(svae_handle* h, int index, float* host_dst);          /* gradient of the last svae_backward */
int svae_adam_get(svae_handle* h, int index, float* host_m, float* host_v);
int svae_adam_set(svae_handle* h, int index, const float* host_m, const float* host_v);
int64_t svae_adam_step_count(const svae_handle* h);
int svae_adam_set_step_count(svae_handle* h, int64_t t);
void* svae_param_arena(svae_handle* h); /* device base pointers of the flat fp32 arenas (offsets from param_info) */
void* svae_grad_arena(svae_handle* h);
int64_t svae_arena_numel(const svae_handle* h);
/* Copy `n` floats starting at element `offset` of arena `which` (0 parameters, 1 gradients, 2 Adam m, 3 Adam v) to host
 * memory; synchronises.  With svae_param_slices this reads an individual slice of a shared variable. */
int svae_arena_read(svae_handle* h, int which, int64_t offset, int64_t n, float* host_dst);

/* ---- training-mode chain ----------------------------------------------------------------------------------------
 * svae_forward replaces sess.run(self.training_mles / training_samples / loss) (sequential_vae.py:1381-1391,
 * 1434-1455): runs all mc_steps steps - recognition net, z = mu + sigma*eps, chain encoder, decoder, per-step ELBO.
 *   x_in_dev, x_tgt_dev : [B,H,W,C]
 *   eps_dev             : [T,B,Z] injected noise (sequential_vae.py:1023); NULL => counter-based Philox N(0,1)
 *                         keyed by (seed, iteration, t, b, j)
 *   mu_out_dev, sigma_out_dev : optional [T,B,Z]; x_steps_out_dev : optional [T,B,H,W,C] (training_mles)
 */
int svae_forward(svae_handle* h, const float* x_in_dev, const float* x_tgt_dev, int batch, const float* eps_dev,
                 uint64_t seed, float reg_coeff, float* mu_out_dev, float* sigma_out_dev, float* x_steps_out_dev);
/* Reverse-mode through the whole chain (optimizer.compute_gradients, sequential_vae.py:1273).  Gradients are left in
 * the gradient arena (svae_grad_get).  Requires a preceding svae_forward on a train_capacity handle. */
int svae_backward(svae_handle* h);
/* clip_by_value(+-clip_value) + TensorFlow-formulation Adam (sequential_vae.py:18-25,1275-1276); all-reduces the
 * gradients first when a communicator is attached. */
int svae_adam_step(svae_handle* h, float learning_rate);
/* One full training iteration = the train_op of sess.run at sequential_vae.py:1365 (forward + backward + [all-reduce]
 * + clipped Adam), device-resident inputs, asynchronous. */
int svae_train_step(svae_handle* h, const float* x_in_dev, const float* x_tgt_dev, int batch, const float* eps_dev,
                    uint64_t seed, float learning_rate, float reg_coeff);
/* Same, through HOST buffers: the feed_dict path of SequentialVAE.train (sequential_vae.py:1355-1365): copies
 * input/target (and eps when given) host->device, runs the step, copies the losses back and synchronises. */
int svae_train_step_host(svae_handle* h, const float* x_in_host, const float* x_tgt_host, int batch,
                         const float* eps_host, uint64_t seed, float learning_rate, float reg_coeff,
                         svae_losses* losses_out);
/* ---- denoising corruption (the host step in front of train/test) -------------------------------------------------------
 * Replaces NoisyTrainer.apply_noise (trainer.py:56-78; constants pepper_prob = salt_prob = gaussian_noise_scale = 0.1,
 * trainer.py:16-18):  out = clip(x * Bernoulli(1 - pepper_prob) + Bernoulli(salt_prob) + N(0, gaussian_scale), lo, hi)
 * element-wise over `n` floats, drawn on the device with counter-based Philox4x32-10 keyed by `seed` (the reference uses
 * the unseeded numpy global RNG on the host).  In place (out_dev == x_dev) is allowed.  draws_out_dev: optional [3,n]
 * buffer receiving the three random fields (keep mask, salt mask, Gaussian term) so that a checker can replay
 * trainer.py:69-78 on the same draws. */
int svae_apply_noise(svae_handle* h, const float* x_dev, float* out_dev, int64_t n, float pepper_prob, float salt_prob,
                     float gaussian_scale, float clip_lo, float clip_hi, uint64_t seed, float* draws_out_dev);
/* Same through host buffers (upload, corrupt, download; synchronises). draws_out_host: optional [3,n]. */
int svae_apply_noise_host(svae_handle* h, const float* x_host, float* out_host, int64_t n, float pepper_prob,
                          float salt_prob, float gaussian_scale, float clip_lo, float clip_hi, uint64_t seed,
                          float* draws_out_host);
/* One denoising training iteration, trainer.py:100-104 with --denoise_train: the clean batch is uploaded ONCE and is the
 * target; the network input is its corruption, produced on the device (half the host->device bytes of the reference's
 * feed, no host RNG).  Clips to the configured range_lo/range_hi (dataset.range, trainer.py:78).  x_noisy_out_host: optional
 * [B,H,W,C] copy of the corrupted input (what plot_reconstruction shows). */
int svae_train_step_host_denoise(svae_handle* h, const float* x_clean_host, int batch, const float* eps_host,
                                 uint64_t seed, float learning_rate, float reg_coeff, float pepper_prob, float salt_prob,
                                 float gaussian_scale, uint64_t noise_seed, float* x_noisy_out_host,
                                 svae_losses* losses_out);
/* SequentialVAE.test / training_mc_samples through host buffers (sequential_vae.py:1381-1391,1434-1455):
 * x_steps_out_host [T,B,H,W,C] (may be NULL), last_out_host [B,H,W,C] (may be NULL). */
int svae_forward_host(svae_handle* h, const float* x_in_host, const float* x_tgt_host, int batch,
                      const float* eps_host, uint64_t seed, float reg_coeff, float* mu_out_host,
                      float* sigma_out_host, float* x_steps_out_host, float* last_out_host, svae_losses* losses_out);
/* Synchronise and fetch the ELBO terms of the last forward. */
int svae_read_losses(svae_handle* h, svae_losses* out);

/* ---- generation-mode chain --------------------------------------------------------------------------------------
 * Replaces sess.run(self.generative_samples) (generate_mc_samples, sequential_vae.py:1397-1428): decoder chain only,
 * z fed per step; BN uses the statistics of the generated batch (Q1).
 *   z_dev   : [T,B,Z] or NULL => Philox N(0,1) from `seed`
 *   out_dev : [T,B,H,W,C]  (x_1..x_T; the reference's leading uniform-noise x_0 is produced by the Python wrapper) */
int svae_generate(svae_handle* h, int batch, const float* z_dev, uint64_t seed, float* out_dev);
int svae_generate_host(svae_handle* h, int batch, const float* z_host, uint64_t seed, float* out_host);

/* ---- chain noise (svae_config.add_noise_to_chain) ------------------------------------------------------------------
 * Replaces the tf.random_normal(image_batch_shape) of create_generator_network (sequential_vae.py:1088-1091).  By default the
 * draws are counter-based Philox inside the kernel (keyed by the step's seed; a different stream than the latent eps).
 * svae_set_chain_noise_host injects them instead - [T,B,H,W,C] standard-normal values used by every following forward /
 * train step / generation with that batch size - so that a parity test can feed the oracle the same draws; NULL restores
 * Philox.  svae_read_chain_samples_host returns the samples x_t + noise of the last forward or generation, [T,B,H,W,C]
 * (training_samples / generative_samples; the mles are what svae_forward / svae_generate return). */
int svae_set_chain_noise_host(svae_handle* h, const float* noise_host_or_null, int batch);
int svae_read_chain_samples_host(svae_handle* h, float* out_host, int batch);

/* ---- data parallel ----------------------------------------------------------------------------------------------
 * The reference has no multi-device path for this model.  Batch-sharded DP: per-replica BN, gradients
 * summed over ranks with NCCL (one bucket per chain step, issued as soon as that step's backward ends) and scaled by
 * 1/nranks inside the Adam kernel. */
int svae_nccl_unique_id(char id_out[128], const char* libnccl_path_or_null);
int svae_comm_init(svae_handle* h, int rank, int nranks, const char id[128], const char* libnccl_path_or_null);
int svae_comm_destroy(svae_handle* h);

/* ---- introspection ---------------------------------------------------------------------------------------------- */
int64_t svae_launch_count(const svae_handle* h);   /* kernels launched by this handle so far */
int64_t svae_activation_bytes(const svae_handle* h);
int svae_tc_layers(const svae_handle* h);          /* number of contractions per train step routed to tcgen05 kernels */

/* Per-kernel-class CUDA-event profile (what bench.py's roofline numbers are computed from): while enabled, every
 * kernel launch is bracketed by events on the handle's stream and its algorithmic flops / bytes are accumulated.
 * svae_profile_read synchronises, fills up to `capacity` classes, resets the counters and returns the class count. */
typedef struct svae_kernel_stats {
  char name[32];
  int64_t launches;
  double total_ms;   /* sum of per-launch event durations */
  double flops;      /* algorithmic flops of those launches */
  double bytes;      /* algorithmic HBM bytes of those launches */
} svae_kernel_stats;
int svae_profile_enable(svae_handle* h, int on);
int svae_profile_read(svae_handle* h, svae_kernel_stats* out, int capacity);

/* ---- layer-level entry points (unit parity against abstract_network.py:12-71; tests and micro-benchmarks only) ----
 * All pointers are device pointers; weights in the reference layout.  `operand_dtype` selects the kernel family. */
/* convolution2d data path (abstract_network.py:18): x [B,H,W,Ci], w [4,4,Ci,Co] -> y [B,H/s,W/s,Co], SAME padding.
 * stats_out (may be NULL): [2*Co] doubles = per-channel sum and sum of squares of y. */
int svae_op_conv2d(svae_handle* h, const float* x, const float* w, float* y, double* stats_out, int B, int H, int W,
                   int Ci, int Co, int stride, int operand_dtype);
/* convolution2d_transpose data path (abstract_network.py:37,56): x [B,H,W,Ci], w [4,4,Co,Ci] -> y [B,H*s,W*s,Co] */
int svae_op_conv2d_transpose(svae_handle* h, const float* x, const float* w, float* y, double* stats_out, int B, int H,
                             int W, int Ci, int Co, int stride, int operand_dtype);
/* gradients of the conv: dy [B,H/s,W/s,Co] -> dx [B,H,W,Ci] (may be NULL), dw [4,4,Ci,Co] (may be NULL) */
int svae_op_conv2d_backward(svae_handle* h, const float* x, const float* w, const float* dy, float* dx, float* dw,
                            int B, int H, int W, int Ci, int Co, int stride, int operand_dtype);
/* gradients of the transposed conv: dy [B,H*s,W*s,Co] -> dx [B,H,W,Ci], dw [4,4,Co,Ci] */
int svae_op_conv2d_transpose_backward(svae_handle* h, const float* x, const float* w, const float* dy, float* dx,
                                      float* dw, int B, int H, int W, int Ci, int Co, int stride, int operand_dtype);
/* fully_connected data path of fc_bn_lrelu (abstract_network.py:65): y[B,N] = x[B,K] . w[K,N]; gradients dx[B,K]
 * (may be NULL), dw[K,N] (may be NULL).  SVAE_OPERAND_BF16 needs K, N >= 64 (svae_op_tc_supported(2, ...)). */
int svae_op_fc(svae_handle* h, const float* x, const float* w, float* y, int B, int K, int N, int operand_dtype);
int svae_op_fc_backward(svae_handle* h, const float* x, const float* w, const float* dy, float* dx, float* dw, int B,
                        int K, int N, int operand_dtype);
/* 1 when the SVAE_OPERAND_BF16 kernel family runs this contraction on tensor cores (operands rounded to bf16), else 0
 * (it then runs on the fp32 SIMT kernels).  transposed: 0 conv, 1 transposed conv, 2 fully connected (Ci -> Co, H = W = 1);
 * direction: 0 forward, 1 input gradient, 2 weight gradient. */
int svae_op_tc_supported(int transposed, int H, int W, int Ci, int Co, int stride, int direction);
/* 1 when the SVAE_OPERAND_BF16 family runs this contraction at batch B on the TMA-fed PRODUCTION kernels of the chain
 * (tc2_conv_kernel for direction 0 / 1, tc2_wgrad_kernel for direction 2): the layer-level entry points above then exercise
 * exactly the kernels a train step launches.  0: SIMT-staged tcgen05 variant or fp32 kernels. */
int svae_op_tc2_supported(int transposed, int B, int H, int W, int Ci, int Co, int stride, int direction);
/* Parity probe (tests only; TF lets callers fetch any tensor of the graph, sequential_vae.py:919-923): one tensor of one
 * block of the LAST svae_forward / svae_backward as a dense fp32 host array, exactly as the kernels consumed or produced it
 * (a tensor kept only as a bf16 copy is returned with its bf16 values).
 *   net:   0 recognition conv block (index 0..2(L-1)-1), 1 chain-encoder conv block (0..2(L-1)), 2 chain-encoder fc block,
 *          3 latent projection (0..L-1), 4 decoder fc block, 5 stride-2 deconv of level `index`, 6 stride-1 deconv of level
 *          `index`, 7 output deconv, 8 highway-gate deconv
 *   which: 0 contraction input, 1 pre-batch-norm contraction output, 2 activated output, 3 dL/d(activated output),
 *          4 dL/d(pre-batch-norm output), 5 tensor added before the activation (ladder shortcut)
 * dims_out receives [B,H,W,C] of the tensor; host_dst == NULL only queries the dims.  Synchronises. */
int svae_debug_block_tensor(svae_handle* h, int step, int net, int index, int which, float* host_dst, int64_t capacity,
                            int32_t dims_out[4]);
/* Development aid: when non-NULL, the tcgen05 conv kernel writes per-CTA phase timestamps ([cta][8] uint64 ns) here. */
int svae_debug_set_buffer(void* dev_buffer);
/* batch_norm (training mode, no gamma, eps 1e-3) + activation (0 none, 1 lrelu(0.1), 2 relu), rows x channels */
int svae_op_bn_act(svae_handle* h, const float* y, const float* beta, float* out, int64_t rows, int C, int act);
/* backward of the same block (autodiff of abstract_network.py:22-23 / :41-42 / :69-70): da = dL/d(out) [rows,C], y = the
 * pre-BN contraction output, residual = tensor added before the activation (ladder shortcut, sequential_vae.py:1713; may be
 * NULL).  Writes dy = dL/dy [rows,C], dbeta[C] (may be NULL) and, when dres != NULL, the shortcut gradient
 * (dres = or += da * act'). */
int svae_op_bn_act_backward(svae_handle* h, const float* da, const float* y, const float* beta, const float* residual,
                            float* dy, float* dbeta, float* dres, int dres_accumulate, int64_t rows, int C, int act);
/* fused clip + TF-Adam on n elements (sequential_vae.py:1275-1276) */
int svae_op_adam(svae_handle* h, float* p, const float* g, float* m, float* v, int64_t n, float lr, int64_t t,
                 float beta1, float beta2, float eps, float clip, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* SVAE_H_ */
