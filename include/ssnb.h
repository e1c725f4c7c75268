/* ssnb.h — C ABI of libssn_b200.so: the B200-native SSN forward/backward hot path.
 *
 * The reference (yjxiong/action-detection) has no FFI: its hot path is Python calling stock
 * PyTorch ops.  This ABI is what a binding for that path binds instead; each entry names the
 * reference interface it replaces (file:line relative to the reference root).  Plain pointers
 * and sizes only — no torch types.  All device pointers are CUDA device memory owned by the
 * caller; `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success and a
 * non-zero code otherwise (never throws, never aborts); ssnb_last_error() returns the message.
 * All work is enqueued on `stream`; no call synchronises the device.
 */
#ifndef SSNB_H
#define SSNB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssnb_engine* ssnb_handle;

enum { SSNB_OK = 0, SSNB_EINVAL = 1, SSNB_ECUDA = 2, SSNB_ESTATE = 3, SSNB_ENOSUPPORT = 4 };

/* precision modes of the backbone */
enum {
  SSNB_EXACT_FP32 = 0, /* fp32 storage, fp32 SIMT FMA: end-to-end parity mode */
  SSNB_FAST_FP16 = 1,  /* fp16 storage, tcgen05 kind::f16 MMA with fp32 TMEM accumulators */
  SSNB_EXACT_TC = 2    /* fp32 storage; every convolution product as error-compensated split fp16 operands
                          (x = hi + lo, three tcgen05 MMAs a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate):
                          fp32-grade results (layer_factory.py:25-39 computes in fp32) on the tensor cores */
};

typedef struct {
  int32_t in_channels; /* 3 (RGB) or 10 (Flow 2x5), ssn_models.py:260 sample_len */
  int32_t frames;      /* F = proposals * segments processed per call */
  int32_t precision;   /* SSNB_EXACT_FP32 | SSNB_FAST_FP16 | SSNB_EXACT_TC */
  int32_t training;    /* 1: keep activations + allocate gradient buffers */
  float grad_scale;    /* power-of-two loss scale: FAST scales dfeat (fp16 gradient storage); EXACT_TC scales the
                          fp16 operand planes of the output gradients (fp32 gradients themselves are unscaled) */
  int32_t bn1_train;   /* 1: the FIRST BatchNorm2d (conv1's) runs in training mode -- batch statistics, running-stat update, gradients for
                          its weight / bias: bn_mode='partial' (ssn_models.py:95-105,156-174).  EXACT_FP32 / EXACT_TC only. */
  int32_t reserved[2];
} ssnb_config;

/* ---- engine lifetime ---------------------------------------------------------------------- */
int ssnb_create(const ssnb_config* cfg, ssnb_handle* out);
int ssnb_destroy(ssnb_handle h);
/* h may be NULL: returns the calling thread's last error from a handle-less function. */
const char* ssnb_last_error(ssnb_handle h);
const char* ssnb_version(void);

/* ---- BNInception backbone: replaces model_zoo.BNInception (pytorch_load.py:8-61) as used by
 *      SSN.train_forward / test_forward (ssn_models.py:266, :298) ------------------------------ */
/* the 69 convolutions in graph order (bn_inception.yaml); lets the host check its own table */
int ssnb_num_convs(void);
int ssnb_conv_info(int idx, int in_channels, char* name, int name_cap, int* cin, int* cout, int* k,
                   int* stride, int* pad);
/* caller-owned scratch (activations, gradients, packed weights, split-K partials) */
size_t ssnb_workspace_bytes(ssnb_handle h);
int ssnb_set_workspace(ssnb_handle h, void* dev_ptr, size_t bytes);
/* Fold frozen BatchNorm2d (ssn_models.py:156-174) into each conv and re-layout for the kernels.
 * Arrays of ssnb_num_convs() device pointers in graph order, reference shapes: w [cout,cin,k,k],
 * b/gamma/beta/mean/var [cout].  Call after every optimizer step. */
int ssnb_pack_weights(ssnb_handle h, const float* const* w, const float* const* b, const float* const* gamma,
                      const float* const* beta, const float* const* mean, const float* const* var, void* stream);
/* bn1_train engines: the first BatchNorm2d's tensors (device pointers, 64 floats each).  gamma / beta / running stats are read
 * (running stats also updated) by ssnb_backbone_fwd; dgamma / dbeta (may be NULL) are written by ssnb_backbone_bwd
 * (added to when grad accumulation is on).  conv1's weights are then packed WITHOUT a BatchNorm fold. */
int ssnb_set_bn1(ssnb_handle h, const float* gamma, const float* beta, float* running_mean, float* running_var, float* dgamma,
                 float* dbeta, float momentum, float eps);
/* input [F, C, 224, 224] fp32 NCHW (the reference's frame tensor after input.view(-1, C, H, W),
 * ssn_models.py:266) -> feat [F, 1024] fp32 (global_pool output, fc replaced by Identity/Dropout) */
int ssnb_backbone_fwd(ssnb_handle h, const float* input_nchw, float* feat, void* stream);
/* dfeat [F,1024] fp32 -> dw[i] [cout,cin,k,k], db[i] [cout] fp32 in reference layout (what autograd
 * leaves in Conv2d.weight.grad / .bias.grad; BN params are frozen and get none).  Overwrites. */
int ssnb_backbone_bwd(ssnb_handle h, const float* dfeat, float* const* dw, float* const* db, void* stream);
/* The same backward in pieces, for overlapping the gradient exchange with it: runs the reverse schedule for ops
 * op_hi .. op_lo (indices of ssnb_op_info; op_hi < 0 = the last op, the global pool, which reads dfeat) and finalises the
 * weight / bias gradients of exactly those ops, so after the call dw[i] / db[i] of their convolutions are complete and a
 * bucket all-reduce can be issued on another stream while the next range runs.  Ranges must be issued from the top down. */
int ssnb_backbone_bwd_range(ssnb_handle h, const float* dfeat, float* const* dw, float* const* db, int op_hi, int op_lo,
                            void* stream);
/* 0 (default): ssnb_backbone_bwd overwrites dw/db; 1: it adds to them (what autograd's AccumulateGrad does
 * with Conv2d.weight.grad), so the caller can hand in the live .grad tensors */
int ssnb_set_grad_accumulate(ssnb_handle h, int accumulate);
/* bind the gradient outputs used by ssnb_run_op(backward=1) without running the whole backward */
int ssnb_bind_grads(ssnb_handle h, float* const* dw, float* const* db);

/* introspection for per-layer parity tests: named values are the reference's blob names
 * ("data", "conv1_7x7_s2_bn", "pool1_3x3_s2", "inception_3a_output", ...). */
int ssnb_num_ops(ssnb_handle h);
int ssnb_op_info(ssnb_handle h, int op, char* kind, int kind_cap, char* in_name, int in_cap, char* out_name, int out_cap);
int ssnb_value_shape(ssnb_handle h, const char* name, int* c, int* hh, int* ww);
int ssnb_value_write(ssnb_handle h, const char* name, int grad, const float* src_nchw, void* stream);
int ssnb_value_read(ssnb_handle h, const char* name, int grad, float* dst_nchw, void* stream);
int ssnb_run_op(ssnb_handle h, int op, int backward, void* stream);
/* Loss-scale guard: 1 when a gradient left the fp16 range under grad_scale since the last clear (EXACT_TC: an operand plane
 * saw |dz * grad_scale| > 65504 or NaN; FAST: a weight-gradient sum came out inf / NaN), 0 otherwise, -1 on error.
 * Synchronises the device: poll it every N steps and skip / rescale like a dynamic loss scaler would. */
int ssnb_grad_overflow(ssnb_handle h, int clear);

/* Per-launch device timing for the roofline figures (bench.py): ssnb_timing_begin opens a session on the calling thread
 * (every library launch then records a CUDA event on its stream); ssnb_timing_report closes it, waits for the last
 * launch and returns lines "kernel\tphase\tlaunches\tms\talgorithmic_flop\n" aggregated by kernel and pass
 * (phase 0 forward, 1 data gradient, 2 weight gradient, 3 other).  The string is owned by the library (thread-local). */
int ssnb_timing_begin(void* stream);
const char* ssnb_timing_report(void);
/* per-kernel-family launch counters since creation (bench.py's gpu_launches claim) */
int64_t ssnb_launch_count(ssnb_handle h);
int64_t ssnb_global_launch_count(void);

/* ---- STPP: replaces StructuredTemporalPyramidPooling.forward (ops/ssn_ops.py:39-70) ----------
 * ft [n*n_seg, D]; scaling [n,2]; parts: n_parts entries (lo, hi, norm, scale_col) over segment
 * indices (host mirrors the tick arithmetic of :49-53); course = mean over [course_lo, course_hi).
 * out: course_ft [n,D], stpp_ft [n, n_parts*D]. */
int ssnb_stpp_fwd(const float* ft, const float* scaling, int n, int n_seg, int D, int n_parts, const int* part_lo,
                  const int* part_hi, const int* part_norm, const int* part_scale_col, int course_lo, int course_hi,
                  float* course_ft, float* stpp_ft, void* stream);
int ssnb_stpp_bwd(const float* d_course, const float* d_stpp, const float* scaling, int n, int n_seg, int D,
                  int n_parts, const int* part_lo, const int* part_hi, const int* part_norm,
                  const int* part_scale_col, int course_lo, int course_hi, float* d_ft, void* stream);
/* fused tail of the backbone: 7x7 global average pool (bn_inception.yaml:552) + optional dropout
 * mask + STPP, reading the engine's 5b output directly.  feat [F,1024] is also written. */
int ssnb_gpool_stpp_fwd(ssnb_handle h, const float* drop_mask, const float* scaling, int n_seg, int n_parts,
                        const int* part_lo, const int* part_hi, const int* part_norm, const int* part_scale_col,
                        int course_lo, int course_hi, float* feat, float* course_ft, float* stpp_ft, void* stream);

/* ---- STPPReorgainzed.forward (ops/ssn_ops.py:109-170), standalong_classifier + regression ----
 * scores [T, D] with D = act_len + M*comp_len + M*reg_len; ticks [N,4] int32; scaling [N,2];
 * stage_parts: 3 stages, parts-per-level lists flattened: level_counts[3] + levels[]. */
int ssnb_stpp_reorg(const float* scores, int T, int D, const int32_t* ticks, const float* scaling, int N,
                    int act_len, int comp_len, int reg_len, const int* level_counts, const int* levels,
                    float* out_act, float* out_comp, float* out_reg, void* stream);

/* Same result through one fp64 exclusive column scan of scores (workspace: (T+1)*D doubles, ssnb_stpp_reorg_workspace_bytes)
 * + one gather per proposal: every part costs two loads instead of its row count (1000 heavily overlapping proposals/video). */
size_t ssnb_stpp_reorg_workspace_bytes(int T, int D);
int ssnb_stpp_reorg_prefix(const float* scores, int T, int D, const int32_t* ticks, const float* scaling, int N, int act_len,
                           int comp_len, int reg_len, const int* level_counts, const int* levels, float* out_act,
                           float* out_comp, float* out_reg, void* workspace, void* stream);

/* ---- heads: activity_fc / completeness_fc / regressor_fc (ssn_models.py:272-283) ------------- */
int ssnb_linear_fwd(const float* x, const float* w, const float* b, int n, int in_dim, int out_dim, float* y,
                    void* stream);
/* Test-time scores of one chunk of a video with the 10-crop mean folded into the folded FC (replaces
 * `rst, _ = net(input); sc = rst.view(num_crop, -1, D).mean(0)`, ssn_test.py:83-84, with test_fc from
 * SSN.prepare_test_fc, ssn_models.py:176-201): feat [crops*nt, in_dim] crop-major -> y [nt, out_dim]. */
int ssnb_test_fc_cropmean(const float* feat, const float* w, const float* b, int crops, int nt, int in_dim, int out_dim,
                          float* y, void* stream);
/* dy [n,out] -> dx [n,in] (may be NULL), dw [out,in], db [out]; overwrite */
int ssnb_linear_bwd(const float* x, const float* w, const float* dy, int n, int in_dim, int out_dim, float* dx,
                    float* dw, float* db, void* stream);

/* ---- losses ------------------------------------------------------------------------------------
 * OHEMHingeLoss.forward/backward (ops/ssn_ops.py:180-213): pred [m,K], labels [m] int64 1-based
 * (0 wraps to class K-1); groups of group_size rows; keep_num = int(group_size*ratio) computed by
 * the caller.  loss[1]; kept [m] uint8 marks the rows whose gradient is written; slopes has room for
 * 2*m floats (slopes [m] followed by the per-row hinge losses [m]). */
int ssnb_ohem_hinge_fwd(const float* pred, const int64_t* labels, int m, int K, int is_positive, int group_size,
                        int keep_num, float* loss, uint8_t* kept, float* slopes, void* stream);
int ssnb_ohem_hinge_bwd(const int64_t* labels, const uint8_t* kept, const float* slopes, const float* grad_out,
                        int m, int K, float* grad_pred, void* stream);
/* ClassWiseRegressionLoss.forward (ops/ssn_ops.py:251-258) and its gradient */
int ssnb_classwise_reg_fwd(const float* pred, const int64_t* labels, const float* targets, int n, int K, float* loss,
                           void* stream);
int ssnb_classwise_reg_bwd(const float* pred, const int64_t* labels, const float* targets, const float* grad_out,
                           int n, int K, float* grad_pred, void* stream);

/* Fused classifier + multi-task loss forward AND backward in one kernel (ssn_models.py:272-289 +
 * ssn_train.py:210-214): three heads, row selection by prop_type, CE + w_comp*completeness(OHEM) +
 * w_reg*class-wise smooth-L1, and all gradients (d course_ft, d stpp_ft, dW, db of the heads). */
typedef struct {
  int32_t n;             /* proposals (videos * props_per_video) */
  int32_t props_per_video;
  int32_t num_class;     /* K */
  int32_t feat_dim;      /* 1024 */
  int32_t feat_mult;     /* M */
  int32_t fg_per_video;  /* sample_split (ssn_train.py:189) */
  int32_t comp_group;    /* fg + incomplete per video (ssn_train.py:190) */
  int32_t global_videos; /* videos in the GLOBAL batch: completeness denominator */
  int32_t keep_neg;      /* int((comp_group - fg_per_video) * ohem_ratio), host-computed (ops/ssn_ops.py:191) */
  float comp_denom;      /* (pos_cnt + int(neg_cnt * ohem_ratio)) of the GLOBAL batch (ops/ssn_ops.py:236-239)
                            divided by the number of data-parallel ranks, so that rank-averaged losses and
                            gradients equal the global-batch ones */
  float comp_w, reg_w;   /* 0.1, 0.1 (ssn_opts.py:35-37) */
  float loss_scale;      /* multiplies every gradient (1/world_size for data parallel) */
} ssnb_heads_cfg;
size_t ssnb_heads_loss_workspace_bytes(const ssnb_heads_cfg* cfg);
/* prop_type/target int64 [n]; reg_target [n,2].  outputs: raw_act [n,K+1], raw_comp [n,K],
 * raw_reg [n,2K] (all rows, unselected), losses[4] = {act, comp, reg, total}. */
int ssnb_heads_loss_fwd_bwd(const ssnb_heads_cfg* cfg, const float* course_ft, const float* stpp_ft,
                            const float* act_w, const float* act_b, const float* comp_w, const float* comp_b,
                            const float* reg_w, const float* reg_b, const int64_t* prop_type, const int64_t* target,
                            const float* reg_target, float* raw_act, float* raw_comp, float* raw_reg, float* losses,
                            float* d_course_ft, float* d_stpp_ft, float* d_act_w, float* d_act_b, float* d_comp_w,
                            float* d_comp_b, float* d_reg_w, float* d_reg_b, void* workspace, void* stream);

/* ---- detection post-processing of one video (eval_detection_results.py:91-183, ops/utils.py:38-40,56-82) --------------
 * rel_props [N,2] (start, end in [0,1]), act_scores [N,K+1], comp_scores [N,K], reg_scores [N,K,2] (already de-normalised,
 * ssn_test.py:89-92) ->  per class c: combined score softmax(act)[:,c+1] * exp(comp[:,c]), greedy temporal NMS at
 * nms_thresh in descending score order, then (regress != 0) the location regression of the survivors.
 * detections [K, N, 5] rows (t0, t1, score, loc, dur) in kept order, counts [K] int32; combined_ws: N*K floats of scratch
 * (ssnb_detect_workspace_bytes).  N <= 8192.  act_scores == NULL: combined_ws already holds the [N,K] scores to rank by
 * (plain class-wise temporal_nms, ops/utils.py:56-82). */
size_t ssnb_detect_workspace_bytes(int n_props, int num_class);
int ssnb_detect_postprocess(const float* rel_props, const float* act_scores, const float* comp_scores, const float* reg_scores,
                            int n_props, int num_class, double nms_thresh, int regress, float* detections, int* counts,
                            float* combined_ws, void* stream);

/* fused SGD-momentum step over flat fp32 buffers (ssn_train.py:141-144 torch.optim.SGD semantics):
 * g = grad*grad_mult + wd*p; buf = mom*buf + g; p -= lr*buf */
int ssnb_sgd_step(float* param, const float* grad, float* momentum_buf, size_t n, float lr, float momentum,
                  float weight_decay, float grad_mult, void* stream);

/* The whole model in ONE launch: the flat buffers are cut into n_seg <= 512 segments (one per parameter tensor; seg_end =
 * cumulative element ends, device int64) with their parameter group's learning rate and weight decay (device fp32 arrays):
 * the per-group lr_mult / decay_mult of SSN.get_optim_policies (ssn_models.py:203-251) as applied by
 * adjust_learning_rate (ssn_train.py:391-398).  Same update rule as ssnb_sgd_step. */
int ssnb_sgd_step_groups(float* param, const float* grad, float* momentum_buf, size_t n, const int64_t* seg_end,
                         const float* seg_lr, const float* seg_wd, int n_seg, float momentum, float grad_mult, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSNB_H */
