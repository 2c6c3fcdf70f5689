/*
 * racb200 -- C ABI of the B200-native planning hot path of penn-pal-lab/robot_aware_control.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below names the reference
 * Python interface it replaces (file:line relative to the reference checkout). The Python package
 * `robot_aware_control_b200` binds these with ctypes and mirrors the reference classes on top.
 *
 * Conventions: every function returns 0 on success or a negative rac_status; nothing throws across the boundary;
 * rac_last_error() gives the message of the last failure on a handle. All tensor pointers are DEVICE pointers owned
 * by the caller unless a parameter is documented as host memory. All work is enqueued on the caller's stream
 * (cudaStream_t passed as void*); no call synchronises the device. One handle per (device, host thread); the
 * recurrent state lives in the handle exactly as it lives on the reference module (lstm.py:216,255).
 */
#ifndef RACB200_H_
#define RACB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAC_ABI_VERSION 8

typedef enum {
  RAC_OK = 0,
  RAC_ERR_INVALID = -1,     /* bad argument / unsupported configuration (reference: ValueError, dynamics.py:470-473) */
  RAC_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
  RAC_ERR_STATE = -3,       /* call order violated (weights not loaded, batch not prepared, ...) */
  RAC_ERR_UNSUPPORTED = -4  /* reference: NotImplementedError */
} rac_status;

typedef struct rac_handle rac_handle;

/* Model configuration: the subset of the reference argparse Namespace the SVG path reads
 * (src/config/__init__.py:165-249, dynamics.py:467-516). */
typedef struct {
  int image_height;            /* 48 */
  int image_width;             /* 64 */
  int g_dim;                   /* multiple of 64 */
  int z_dim;                   /* <= 64 */
  int action_dim;              /* model action channels (5 in BASELINE) */
  int robot_dim;               /* 5 */
  int use_mask;                /* cfg.model_use_mask */
  int use_future_mask;         /* cfg.model_use_future_mask */
  int use_robot_state;         /* cfg.model_use_robot_state */
  int use_future_robot_state;  /* cfg.model_use_future_robot_state */
  int conv_impl;               /* 0 = tcgen05/TMA product path; 1 = SIMT cross-check kernel (tests only) */
  int lstm_group_norm;         /* cfg.lstm_group_norm: NormConvLSTMCell instead of ConvLSTMCell (lstm.py:151-198,206) */
} rac_config;

/* Packed-layer ids: one per convolution of SVGConvModel (state_dict prefixes in SURVEY.md 8(a)). */
enum {
  RAC_L_ENC_C1_0 = 0, RAC_L_ENC_C1_1, RAC_L_ENC_C2_0, RAC_L_ENC_C2_1, RAC_L_ENC_C3_0, RAC_L_ENC_C3_1, RAC_L_ENC_C3_2,
  RAC_L_ENC_C4_0, RAC_L_ENC_C4_1, RAC_L_ENC_C4_2,
  RAC_L_PRIOR_IN, RAC_L_PRIOR_LSTM0, RAC_L_PRIOR_LSTM1, RAC_L_PRIOR_GAUSS,
  RAC_L_FP_IN, RAC_L_FP_LSTM0, RAC_L_FP_LSTM1,
  RAC_L_DEC_UPC2_0, RAC_L_DEC_UPC2_1, RAC_L_DEC_UPC2_2, RAC_L_DEC_UPC3_0, RAC_L_DEC_UPC3_1, RAC_L_DEC_UPC3_2,
  RAC_L_DEC_UPC4_0, RAC_L_DEC_UPC4_1, RAC_L_DEC_UPC5_0, RAC_L_DEC_UPC5_1,
  RAC_L_POST_IN, RAC_L_POST_LSTM0, RAC_L_POST_LSTM1, RAC_L_POST_GAUSS,
  RAC_L_COUNT
};
/* cfg.lstm_group_norm only: the RAC_L_*_LSTMk ids above are then the `ih_gates.0` convolutions (g -> 4g) and the ids
 * below the `hh_gates.0` convolutions of the same cells (lstm.py:163-171). */
enum {
  RAC_L_PRIOR_LSTM0_HH = RAC_L_COUNT, RAC_L_PRIOR_LSTM1_HH, RAC_L_FP_LSTM0_HH, RAC_L_FP_LSTM1_HH,
  RAC_L_POST_LSTM0_HH, RAC_L_POST_LSTM1_HH,
  RAC_L_COUNT_GN
};

int rac_abi_version(void);

/* SVGConvModel.__init__ (dynamics.py:460-534). */
int rac_create(const rac_config* cfg, rac_handle** out);
int rac_destroy(rac_handle* h);
const char* rac_last_error(const rac_handle* h);

/* Expected packed sizes of a layer for this configuration: weight elements (bf16; fp32 for RAC_L_ENC_C1_0),
 * bias elements (fp32), packed K (input channels incl. padding x taps) and packed N. */
int rac_layer_shape(const rac_handle* h, int layer, int64_t* w_elems, int64_t* bias_elems, int* k_packed,
                    int* n_packed);

/* nn.Module.load_state_dict (widowx_VMPC_controller.py:98-101): one packed layer; the library copies into its own
 * device memory. `w` / `bias` may be host or device pointers (cudaMemcpyDefault). */
int rac_load_layer(rac_handle* h, int layer, const void* w, int64_t w_elems, const float* bias, int64_t bias_elems);

/* cfg.lstm_group_norm: GroupNorm affine parameters of one NormConvLSTMCell (lstm.py:163-172), `layer` = the cell's
 * RAC_L_*_LSTMk id. `packed` (host or device, 18*g floats): [ih gamma | ih beta | hh gamma | hh beta], 4g each, in
 * packed gate-column order (channel * 4 + gate), then [c_norm gamma | c_norm beta], g each. */
int rac_load_lstm_norm(rac_handle* h, int layer, const float* packed, int64_t elems);

/* Allocate activation workspace, recurrent state and TMA descriptors for `batch` candidates (idempotent). */
int rac_prepare(rac_handle* h, int batch);

/* SVGConvModel.init_hidden (dynamics.py:536-542): zero (h, c) of prior / posterior / frame predictor. */
int rac_init_hidden(rac_handle* h, int batch, void* stream);

/* SVGConvModel.forward (dynamics.py:544-644), eval-mode BatchNorm. */
typedef struct {
  int n;                    /* batch */
  const float* image;       /* (n,3,H,W) fp32 NCHW in [0,1] */
  const float* mask;        /* (n, 1 or 2, H, W) fp32 or NULL: channel 0 = mask_t, channel 1 = mask_{t+1} */
  const float* robot;       /* (n, robot_dim) or NULL */
  const float* robot_next;  /* (n, robot_dim) or NULL (model_use_future_robot_state) */
  const float* action;      /* (n, action_dim) */
  const float* eps;         /* (n, z_dim, H/8, W/8) prior noise or NULL -> Philox(seed, noise_ctr) */
  unsigned long long seed;
  unsigned int noise_ctr;
  int sample_mean;          /* forward(..., sample_mean=) */
  int use_posterior;        /* next_image given: run the posterior branch (dynamics.py:613-629) */
  const float* next_robot;  /* (n, robot_dim) r_target or NULL */
  const float* eps_post;    /* posterior noise or NULL */
  int force_use_prior;
  int keep_skip;            /* decoder uses the skip tensors already held by the handle (last_frame_skip False) */
  float* x_pred;            /* out (n,4,H,W) fp32 NCHW */
  float* mu_p;              /* out (n,z,H/8,W/8) or NULL */
  float* logvar_p;
  float* mu;                /* posterior outputs or NULL */
  float* logvar;
} rac_step;
int rac_forward(rac_handle* h, const rac_step* s, void* stream);

/* TrajectorySampler.generate_model_rollouts (trajectory_sampler.py:35-199) for one shard of candidates:
 * autoregressive rollout + per-step RobotWorldCost accumulation, no host round trip. */
typedef struct {
  int n;                       /* candidates in this call */
  int steps;                   /* rollout length L (= horizon - 1 in the reference) */
  int cand_offset;             /* global id of candidate 0 (Philox noise is keyed on global ids) */
  const float* actions;        /* (n, steps, action_dim) fp32 */
  const uint8_t* start_img;    /* (H,W,3) uint8 */
  const uint8_t* goal_imgs;    /* (G,H,W,3) uint8 */
  int num_goals;
  const float* goal_masks;     /* (G,H,W) fp32 {0,1} or NULL */
  const float* states;         /* (steps+1, *, robot_dim) fp32 or NULL; pointer at this shard's first candidate */
  int64_t state_t_stride;      /* elements between consecutive time steps of `states` */
  const float* masks;          /* (steps+1, *, H, W) fp32 {0,1} or NULL; pointer at this shard's first candidate */
  int64_t mask_t_stride;       /* elements between consecutive time steps of `masks` */
  const float* eps;            /* (steps, n, z_dim, H/8, W/8) supplied prior noise or NULL -> Philox */
  unsigned long long seed;
  unsigned int noise_ctr_base; /* step t uses noise_ctr_base + t */
  int sample_mean;             /* cfg.sample_mean */
  int zero_robot;              /* "dontcare" in cfg.reconstruction_loss or cfg.black_robot_input */
  int dontcare_cost;           /* cfg.reward_type == "dontcare" */
  int sparse_cost;             /* cfg.sparse_cost */
  float world_cost_weight;     /* cfg.world_cost_weight */
  float* obs_out;              /* (steps, n, H, W, 4) fp32 (rgb + pad) or NULL */
  float* step_cost_out;        /* (steps, n) fp32 or NULL */
  double* sum_cost;            /* (n) fp64 out */
  /* Multi-GPU (candidates sharded over the GPUs of one NVLink / NVSwitch node): when peer_world > 0 the kernel that
   * finishes the per-candidate cost of the LAST step also stores it straight into the gathered cost vector of every
   * rank through peer memory -- peer_cost_bufs = DEVICE array of peer_world pointers to each rank's (N) fp64 buffer
   * mapped into this process (e.g. torch symmetric memory), element peer_offset + i for local candidate i. This is
   * the all-gather of cem.py:96 fused into the cost kernel; rac_peer_barrier() orders it against the consumers. */
  double* const* peer_cost_bufs;
  int peer_world;
  int64_t peer_offset;
} rac_rollout;
int rac_rollout_cost(rac_handle* h, const rac_rollout* r, void* stream);

/* Cross-GPU barrier after the fused cost exchange: every rank stores `seq` into slot [rank] of every peer's signal pad
 * (system-scope release), then waits until all peer_world slots of its own pad hold `seq` (acquire). signal_pads =
 * DEVICE array of peer_world pointers to the ranks' uint32 pads; words [slot_base, slot_base + peer_world) of each pad
 * are used (zero-initialised; seq must increase by one per call). A rank that never arrives makes the kernel trap after ~60 s instead of hanging for ever. */
int rac_peer_barrier(uint32_t* const* signal_pads, int slot_base, int rank, int peer_world, uint32_t seq, void* stream);

/* CEMPolicy.get_action pieces (cem.py:76-104). */
int rac_cem_sample(const float* mean, const float* stdv, const float* noise, unsigned long long seed, int iter,
                   int n_total, int steps, int action_dim_model, int cand_offset, int n_local, float clamp,
                   float* act2_out /* (n_total, steps, 2) */, float* act_model_out /* (n_local, steps, action_dim) */,
                   void* stream);
/* costs.topk(K) (cem.py:97): K largest, ties -> lowest index, output sorted (value desc, index asc). */
int rac_topk(const double* costs, int n, int k, int64_t* idx_out, double* val_out_or_null, void* stream);
/* torch.std_mean(top_act_seq, dim=0) + std floor (cem.py:101-104). */
int rac_cem_refit(const float* act2, int steps, const int64_t* elite_idx, int k, float std_floor, float* mean_out,
                  float* std_out, void* stream);

/* ---- Robot state / mask producer: the device-side stand-in for
 * `robot_model.predict_batch(data, thick=True)` of TrajectorySampler.generate_model_rollouts
 * (trajectory_sampler.py:86-109), so that a robot-aware plan has no host round trip per CEM iteration.
 *   states: planar end-effector integration of WX250sAnalyticalModel.predict_batch / predict_trajectory
 *           (src/dataset/wx250s/wx250s_model.py:57-66,98-117,121-182) or FrankaAnalyticalModel.predict_batch
 *           (src/dataset/franka/franka_model.py:30-80) with the reference's float32 / float64 mix: bit-equal.
 *   masks:  the reference renders MuJoCo segmentation images of the arm meshes after an IK call (neither available
 *           outside the authors' lab setup); here a capsule model of the arm is posed by closed-form planar IK and
 *           rasterised through a pinhole camera. NOT reference-pinned (oracle: oracle/robot_oracle.py). Callers with
 *           their own mask source keep passing `masks` to rac_rollout and use rac_predict_states alone. */
enum { RAC_ROBOT_WX250S = 0, RAC_ROBOT_FRANKA = 1 };
typedef struct {
  int kind;               /* RAC_ROBOT_WX250S / RAC_ROBOT_FRANKA: which reference integration rule */
  float low[5], high[5];  /* state normalisation bounds (trajectory_sampler.py:22-23) */
  double frame_diff[2];   /* LOCO_WX250S_DIFF / LOCO_FRANKA_DIFF (src/utils/camera_calibration.py:176-177) */
  double push_height;     /* WX250s: z of every predicted state (wx250s_model.py:65) */
  /* capsule rasteriser (robot base frame, metres) */
  float cam_center[3];    /* camera centre */
  float cam_minv[9];      /* row-major 3x3: pixel (u, v, 1) of the H x W image -> ray direction, = (K R)^-1 */
  float shoulder_z;       /* height of the shoulder joint above the base */
  float l_upper, l_fore, l_wrist; /* link lengths: shoulder-elbow, elbow-wrist, wrist-finger tip */
  float pitch;            /* gripper pitch (default_pitch of the controller), > 0 = pointing down */
  float radius[4];        /* capsule radii: column, upper arm, forearm, wrist + gripper */
} rac_robot_model;
/* start_state: DEVICE (5) normalised start state (states[0, i] of the reference's `data`); actions (n, steps,
 * action_dim); states_out (steps+1, *, 5) with `state_t_stride` elements between time steps. */
int rac_predict_states(const rac_robot_model* m, const float* start_state, const float* actions, int n, int steps,
                       int action_dim, float* states_out, int64_t state_t_stride, void* stream);
/* states (steps+1, *, 5) normalised -> masks_out (steps+1, *, H, W) float {0,1}; extra_radius >= 0 thickens every
 * capsule (the reference's thick=True masks). */
int rac_render_masks(const rac_robot_model* m, const float* states, int64_t state_t_stride, int n, int steps, int H,
                     int W, float extra_radius, float* masks_out, int64_t mask_t_stride, void* stream);

/* Whole single-GPU plan: `iters` x (sample -> [robot states + masks] -> rollout+cost -> top-k -> refit), device
 * resident. */
typedef struct {
  int n;               /* action_candidates */
  int steps;           /* horizon - 1 */
  int iters;           /* opt_iter */
  int topk;            /* K */
  float init_std;
  float clamp;         /* 0.05 */
  float std_floor;     /* 0.001 */
  const float* noise;  /* (iters, n, steps, 2) standard normals or NULL -> Philox(seed) */
  rac_rollout rollout; /* template: actions / sum_cost / n / steps are filled by the library */
  /* optional device-side robot model: when set, every iteration predicts the candidates' robot states (and, with
   * robot_render_masks, their masks) from the sampled actions on the device and feeds them to the rollout;
   * rollout.states / rollout.masks are then ignored (masks only when rendered here) */
  const rac_robot_model* robot; /* HOST pointer or NULL */
  const float* robot_start_state; /* DEVICE (5) normalised start state */
  int robot_render_masks;
  float robot_extra_radius;
} rac_cem;
int rac_cem_plan(rac_handle* h, const rac_cem* c, float* mean_out /* (steps,2) */, float* std_out /* (steps,2) */,
                 int64_t* elite_idx_out /* (topk) or NULL */, double* last_costs_out /* (n) or NULL */, void* stream);

/* ImgL2Cost._call_tensor / ImgDontcareCost._call_tensor (losses.py:224-235,244-263) in the reference tensor layout:
 * curr (n,3,H,W), goal (3,H,W), curr_mask (n,1,H,W) or NULL, goal_mask (1,H,W) or NULL -> out (n) = -dist. */
int rac_masked_cost(const float* curr, const float* goal, const float* curr_mask, const float* goal_mask,
                    int dontcare, float* out, int n, int hw, void* stream);

/* Training criteria, forward value (losses.py:13-19,35-50,97-106); out = 1 fp32 scalar. */
int rac_l1_loss(const float* pred, const float* target, float* out, int64_t numel, void* stream);
int rac_dontcare_l1_loss(const float* pred, const float* target, const float* mask, float robot_weight, float* out,
                         int n, int hw, void* stream);
/* All four reconstruction criteria of PredictionTrainer._recon_loss (trainer.py:149-161; losses.py:11-50) in one entry:
 * kind 0 = l1_criterion, 1 = dontcare_l1_criterion, 2 = mse_criterion (nn.MSELoss), 3 = dontcare_mse_criterion.
 * mask (n,1,H,W) for the dontcare kinds; batch_weight (n) or NULL (the l1 kinds' movement weighting);
 * per_sample: caller scratch of n floats (the per-sample terms, summed in index order); out = 1 fp32 scalar. */
int rac_recon_loss(const float* pred, const float* target, const float* mask, const float* batch_weight, int kind,
                   float robot_weight, float* per_sample, float* out, int n, int hw, void* stream);
/* robot_mse_criterion / world_mse_criterion (losses.py:52-78): out2[0] += robot, out2[1] += world (accumulating). */
int rac_robot_world_mse(const float* pred, const float* target, const float* mask, float* out2, int n, int hw,
                        void* stream);
int rac_kl_loss(const float* mu1, const float* logvar1, const float* mu2, const float* logvar2, float* out,
                int64_t numel, int batch, void* stream);

/* Evaluation metrics of PredictionTrainer._eval_step (trainer.py:685-700). `mask` (n,1,H,W) or NULL: robot pixels
 * of BOTH images are zeroed first (zero_robot_region with the true mask, image.py:5-20).
 * rac_psnr: psnr(estimates, targets) of src/utils/metrics.py:57-78 (data_dims 3; inputs pass through (x+1)/2),
 *           clamp01 != 0 applies clamp(0, 1) after the masking (trainer.py:689) -> out (n).
 * rac_world_psnr: world_psnr_criterion (losses.py:80-94) -> out (n).
 * rac_ssim: ssim(img1, img2, window_size=11) of src/utils/metrics.py:22-54: map_out (n,c,H,W) and / or
 *           plane_mean_out (n*c) = mean of the map over (H,W); either may be NULL. */
int rac_psnr(const float* est, const float* target, const float* mask, int clamp01, float* out, int n, int c, int hw,
             void* stream);
int rac_world_psnr(const float* pred, const float* target, const float* mask, float* out, int n, int hw, void* stream);
int rac_ssim(const float* img1, const float* img2, const float* mask, float* map_out, float* plane_mean_out, int n,
             int c, int h, int w, void* stream);

/* Compositing of the decoder output (trainer.py:653-654, trajectory_sampler.py:149-150): x_pred4 (n,4,H,W) = rgb +
 * blend mask, x_j (n,3,H,W) -> out (n,3,H,W) = (1 - m) * x_j + m * rgb. */
int rac_composite(const float* x_pred4, const float* x_j, float* out, int n, int hw, void* stream);

/* ---- training-data path --------------------------------------------------------------------
 * Replaces the per-clip CPU work of RoboNetDataset._preprocess_images_masks (src/dataset/robonet/robonet_dataset.py:
 * 257-300: ToTensor, optional random crop + bilinear resize back to H x W, shuffled colour jitter :546-573, masks cast
 * back to {0, 1}) and the batch-first -> time-first transposition of process_batch (:434-451) with one launch over the
 * raw uint8 frames. The random draws stay on the host (data.py::sample_augment consumes the generators exactly as the
 * reference does); the library gets their values. */
typedef struct {
  int crop_i, crop_j, crop_h, crop_w; /* F.crop window (top, left, height, width); (0, 0, H, W) = no crop / resize */
  double factor[4];                   /* brightness, contrast, saturation, hue factors */
  int order[4];                       /* colour transforms in application order: 0 brightness, 1 contrast,
                                         2 saturation, 3 hue; -1 = none */
} rac_augment;
/* frames: device uint8 (B, T, H, W, 3) as stored / collated; masks: device (B, T, H, W) float32 (mask_is_u8 = 0) or
 * uint8 (1), or NULL; aug: DEVICE rac_augment[B] (one per clip) or NULL (no augmentation);
 * images_out (T, B, 3, 48, 64), masks_out (T, B, 1, 48, 64) float32, time-first. H x W = the STORED frame size:
 * 48 x 64 (ToTensor only, bit-exact value / 255; `frames` 16-byte aligned) or anything else (e.g. RoboNet's 240 x 320),
 * in which case the dataset's `tf.Resize((48, 64))` runs first (robonet_dataset.py:58: bilinear, align_corners False,
 * no antialiasing -- the behaviour of the torchvision 0.8 / 0.9 the reference pins). */
int rac_process_batch(const uint8_t* frames, const void* masks, int mask_is_u8, int B, int T, int H, int W,
                      const rac_augment* aug, float* images_out, float* masks_out, void* stream);

/* RoboNetDataset._preprocess_states (robonet_dataset.py:302-334), the autograsp column of _load_actions (:173-194) and
 * process_batch's time-first layout (:434-451) for a collated batch. One record per clip (bounds after
 * _preprocess_bounds, :222-255 -- computed by the host glue, data.py): */
typedef struct {
  int kind;             /* 0 = RoboNet file (states normalised in the bounds: denormalised first), 1 = locobot (raw
                           metres), 2 = franka (raw metres, mapped onto the locobot frame: xy += frame_diff, z = 0.14) */
  int camera;           /* 1 = "camera" in cfg.preprocess_action: end-effector position -> camera frame */
  int grip_col;         /* column of the stored states that holds the gripper reading (the last stored column) */
  int pad_;
  double low[5], high[5];    /* normalisation bounds (camera-space box when camera) */
  double world2cam[16];      /* row-major 4 x 4 (src/utils/camera_calibration.py world_to_camera_dict) */
  double frame_diff[2];      /* LOCO_FRANKA_DIFF (robonet_dataset.py:22) */
  double grip_low, grip_high;/* raw_low[4], raw_high[4] of _load_bounds: the two autograsp action values */
} rac_clip_calib;
/* states (B, T, R) float32 as loaded (padded to R = cfg.robot_dim), actions (B, T-1, A_in) or NULL; calib: DEVICE
 * rac_clip_calib[B]; states_out (T, B, R), actions_out (T-1, B, A_out) time-first. A_out == A_in + 1 imputes the
 * autograsp action column (cfg.impute_autograsp_action). */
int rac_preprocess_states(const float* states, const float* actions, const rac_clip_calib* calib, int B, int T, int R,
                          int A_in, int A_out, float* states_out, float* actions_out, void* stream);

/* Test / inspection hook: device pointer and element count of a named internal buffer of the prepared workspace
 * ("h1".."h4", "prior_in", "z", "h_pred", "img", ...). */
int rac_debug_buffer(rac_handle* h, const char* name, void** ptr, int64_t* elems, int* elem_bytes);
/* Number of kernels the library has launched on this handle since creation (bench.py gpu_launches). */
int64_t rac_launch_count(const rac_handle* h);

/* ---- Training step: PredictionTrainer._train_step (trainer.py:326-465), Adam from _init_models (:109-122).
 * Parameters / BatchNorm running statistics / gradients / Adam moments are flat fp32 device buffers owned by the
 * caller; the per-layer tables say where each convolution's tensors live in them. */
typedef struct {
  const long long* row_off;  /* device int64[n_packed]: params offset of packed output column n (weight row), -1 = zero */
  const int* col_off;        /* device int32[ctot]: offset of packed input channel c inside a weight row, -1 = zero */
  const long long* bias_off; /* device int64[n_packed]: params offset of the bias of column n, or NULL (no bias) */
  long long gamma_off, beta_off;   /* BatchNorm affine in params, -1 if the layer has no BatchNorm */
  long long rmean_off, rvar_off;   /* BatchNorm running statistics in buffers */
  long long w_off;                 /* RAC_L_ENC_C1_0 only: offset of its (64, cin, 3, 3) weight */
  int flip;                        /* 1: ConvTranspose2d weight (taps stored flipped) */
  /* cfg.lstm_group_norm (NormConvLSTMCell, lstm.py:151-198), RAC_L_*_LSTMk = ih_gates.0 and RAC_L_*_LSTMk_HH =
   * hh_gates.0 entries only: gamma_off / beta_off above are then the GroupNorm(16, 4g) affine of ih_gates.1 /
   * hh_gates.1, and on the ih entry these two are the cell's c_norm weight / bias */
  long long cnorm_gamma_off, cnorm_beta_off;
  /* the contiguous range of the flat gradient buffer that holds this convolution's weight (+ bias) gradient and
   * nothing else: final as soon as the layer's weight gradient has been unpacked (see rac_train_batch.grads_ready) */
  long long grad_off, grad_count;
  /* > 0: the layer's weight is ONE tensor of w_count elements at w_off and every element of it appears in the packing
   * tables -- such a layer can take the fused optimizer step (rac_train_batch.defer_unpack); 0 otherwise */
  long long w_count;
} rac_train_layer;

typedef struct {
  int batch;                 /* B (multiple of 4) */
  int steps;                 /* n_past + n_future - 1 predicted frames */
  float lr, beta1, beta2, adam_eps;
  float kl_beta;             /* cfg.beta */
  float robot_pixel_weight;  /* cfg.robot_pixel_weight */
  int recon_kind;            /* cfg.reconstruction_loss (trainer.py:149-161): 0 = l1, 1 = dontcare_l1, 2 = mse (the
                                argparse default), 3 = dontcare_mse */
  int zero_robot;            /* "dontcare" in reconstruction_loss or black_robot_input */
  long long n_params, n_buffers;
  int fixed_skip;            /* 1 = cfg.last_frame_skip False (the config default, src/config/__init__.py:217-222): every
                                step decodes with the skips of the clip's FIRST frame (trainer.py:370-371,409-411;
                                dynamics.py:586-588,644); 0 = skips of the step's own input frame */
} rac_train_config;

typedef struct {
  const float* images;     /* (steps+1, B, 3, H, W) time-first, as the reference loader (robonet_dataset.py:434-451) */
  const float* masks;      /* (steps+1, B, 1, H, W) or NULL */
  const float* states;     /* (steps+1, B, robot_dim) or NULL */
  const float* actions;    /* (steps, B, action_dim) */
  const float* eps_prior;  /* (steps, B, z_dim, H/8, W/8) or NULL -> Philox(seed) */
  const float* eps_post;
  unsigned long long seed;
  float* losses;           /* out, device float[4]: sums over steps of reconstruction loss, KL term, and (when masks
                              are given) the logged robot_mse / world_mse metrics (trainer.py:436-439) */
  const int* true_token;   /* HOST int[steps] or NULL: 0 at step t >= 1 = feed the model's own previous prediction
                              (scheduled sampling, trainer.py:132-147,353-356); NULL = always the ground-truth frame */
  unsigned long long noise_step; /* Philox counter base of the reparameterisation noise when eps_* are NULL: the
                              caller's global training step (the "step" of the reference checkpoint, trainer.py:829-837),
                              so that a resumed run or a re-created state never replays earlier noise */
  const float* batch_weight; /* (B) or NULL: cfg.load_movement_info weighting of the l1 / dontcare_l1 loss
                              (trainer.py:426-429; losses.py:13-19,35-50) */
  /* Data-parallel overlap (NULL = off): called on the host, while rac_train_forward_backward is still enqueueing,
   * right after the kernels that FINISH grads[off, off + count) have been put on the stream -- the weight gradient
   * of a large layer (>= grads_ready_min elements; the ConvLSTM gate convolutions: 89 % of all parameters), which
   * BPTT completes long before the encoder's. The caller starts its all-reduce of `count` floats at `ptr` there (stream-ordered
   * after everything enqueued so far) and it runs underneath the rest of the backward pass. */
  void (*grads_ready)(void* user, float* ptr, long long count, long long flat_off, long long flat_count);
  void* grads_ready_user;
  long long grads_ready_min;
  /* 1: the caller will run rac_train_adam_step right after this call and does not need the flat gradient of the large
   * single-tensor convolutions (w_count >= defer_min): their packed weight gradient stays where the wgrad GEMM left
   * it and rac_train_adam_step updates such a layer in ONE pass (packed gradient -> Adam on its parameter / moment
   * slices -> bf16 operand of the next step) instead of unpack + flat Adam + re-pack: 30 instead of 46 bytes per
   * weight. grads_ready then hands over ptr = the packed gradient buffer (count elements, to be all-reduced in place)
   * and flat_off / flat_count = the flat range that is NOT going to be written (exclude it from the remainder);
   * without defer_unpack ptr = grads + flat_off, count = flat_count. rac_train_unpack_deferred writes the flat
   * gradient of the deferred layers after all (inspection, tests). */
  int defer_unpack;
  long long defer_min;
} rac_train_batch;

int rac_train_create(rac_handle* h, const rac_train_config* cfg,
                     const rac_train_layer* layers /* [RAC_L_COUNT], [RAC_L_COUNT_GN] with lstm_group_norm */,
                     float* params, float* buffers, float* grads, float* adam_m, float* adam_v);
int rac_train_destroy(rac_handle* h);
/* forward (train-mode BatchNorm, posterior) + BPTT backward: fills `grads` (caller may all-reduce it) */
int rac_train_forward_backward(rac_handle* h, const rac_train_batch* batch, void* stream);
/* test hook: device pointer of a gradient accumulator ("G_d5", "G_cat5", ...) or of a saved forward tensor ("cat5",
 * "d5", "raw18", ...) of time step `step` */
int rac_train_debug_buffer(rac_handle* h, const char* name, int step, void** ptr);
/* params <- Adam(params, grads) */
int rac_train_adam_step(rac_handle* h, void* stream);
/* Factor applied to every gradient inside the Adam kernel (1 / world size after a SUM all-reduce: saves a pass over
 * the gradient buffer). Stays in force until changed; rac_train_create resets it to 1. */
int rac_train_set_grad_scale(rac_handle* h, float scale);
/* flat gradient of the layers a defer_unpack step left packed (before rac_train_adam_step) */
int rac_train_unpack_deferred(rac_handle* h, void* stream);
/* the caller changed parameters behind the library's back (load_state_dict, an external optimizer): the bf16 operands
 * kept from the last fused optimizer step are stale and are re-packed at the next step */
int rac_train_invalidate_packed(rac_handle* h);
/* Number of Adam steps already taken (the t of the bias corrections). rac_train_create starts at 0: a caller that
 * re-creates the training state (another batch shape) or resumes from a checkpoint (torch.optim.Adam state "step",
 * trainer.py:829-896) restores it here; the moments themselves live in the caller's adam_m / adam_v. */
int rac_train_set_adam_step(rac_handle* h, int steps_taken);

/* ---- Step API: the same training tape driven ONE time step per call, so that a train-mode SVGConvModel.forward
 * (dynamics.py:544-644) can sit inside the caller's torch autograd graph and the reference's own _train_step body
 * (trainer.py:326-465: compositing, _recon_loss, kl_criterion, loss.backward(), torch.optim.Adam) runs unchanged.
 * rac_train_step_begin = model.init_hidden() in train mode (zero recurrent state, re-pack the bf16 operands from the
 * current parameters, zero `grads`); _forward appends one step to the tape (posterior z, batch-statistics BatchNorm,
 * running statistics updated as by two encoder passes); _backward must be called for the steps in reverse order
 * (BPTT), with the gradients w.r.t. that step's outputs; after step 0 `grads` holds dL/d(parameters). */
typedef struct {
  const float* image;       /* (B, 3, H, W): the frame the model sees (robot pixels already zeroed by the caller) */
  const float* mask;        /* (B, 1, H, W), or (B, 2, H, W) = cat([m_j, m_i]) with model_use_future_mask; NULL if unused */
  const float* robot;       /* (B, robot_dim) r_j or NULL */
  const float* next_robot;  /* (B, robot_dim) r_i: the future-state channels and the posterior's input, or NULL */
  const float* action;      /* (B, action_dim) */
  const float* eps_prior;   /* (B, z_dim, H/8, W/8) or NULL -> Philox(seed, noise_step) */
  const float* eps_post;
  unsigned long long seed, noise_step;
  int keep_skip;            /* cfg.fixed_skip: 0 at the first step, 1 afterwards (decode with the first step's skips) */
  float *x_pred;            /* out (B, 4, H, W): sigmoid(decoder logits), channel 3 = compositing mask */
  float *mu, *logvar, *mu_p, *logvar_p; /* out (B, z_dim, H/8, W/8) */
} rac_train_step;
int rac_train_step_begin(rac_handle* h, void* stream);
int rac_train_step_forward(rac_handle* h, const rac_train_step* step, void* stream);
/* step: the same input pointers as the forward call of time step t (outputs ignored). d_*: gradients w.r.t. the
 * outputs (NULL = zero). d_image: out (B, 3, H, W) gradient w.r.t. `image`, or NULL when not needed. */
int rac_train_step_backward(rac_handle* h, int t, const rac_train_step* step, const float* d_x_pred, const float* d_mu,
                            const float* d_logvar, const float* d_mu_p, const float* d_logvar_p, float* d_image,
                            void* stream);

/* Live timing of one kernel family for the roofline line of bench.py: CUDA event pairs are recorded on the launch
 * stream around every convolution launch whose layer name contains `name_substr` (e.g. "lstm.0"), up to
 * `max_launches`. rac_profile_end waits for the recorded events and returns their count and summed duration. */
int rac_profile_begin(rac_handle* h, const char* name_substr, int max_launches);
int rac_profile_end(rac_handle* h, int64_t* launches, double* total_ms);

#ifdef __cplusplus
}
#endif
#endif /* RACB200_H_ */
