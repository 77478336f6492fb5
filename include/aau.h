/* libaau -- C ABI of the B200-native AttentionASPPUNet inference engine.
 *
 * The reference (vivi-git188/ATT-ASPP-UNET) is pure Python and has no FFI of its own: its boundary for this hot
 * path is the duck-typed torch.nn.Module `AttentionASPPUNet` plus two numpy helpers.  Every entry point below
 * replaces one of those reference interfaces (file:line relative to /root/reference) and is what the Python
 * host layer in att-aspp-unet_b200/ binds with ctypes.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions: plain pointers and sizes only (no torch types); every function returns 0 on success and a
 * negative aau_status on failure, never throws or aborts; `aau_last_error` gives the message.  All device work
 * is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL = legacy default stream) with no
 * hidden synchronisation unless stated.  The caller owns input / output / workspace buffers (device memory);
 * the library owns its packed weights.  One handle per (device, host thread): there is no global mutable state.
 */
#ifndef AAU_H_
#define AAU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aau_handle aau_handle;

typedef enum {
    AAU_OK = 0,
    AAU_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    AAU_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
    AAU_ERR_STATE = -3,       /* call order (e.g. forward before weights were committed) */
    AAU_ERR_WEIGHTS = -4,     /* missing / mis-shaped state_dict entry */
    AAU_ERR_WORKSPACE = -5,   /* workspace too small or misaligned */
    AAU_ERR_DEVICE = -6       /* a kernel reported a pipeline fault (see aau_device_fault) */
} aau_status;

enum { AAU_VARIANT_PIPELINE = 0, AAU_VARIANT_ABLATION = 1 };
enum { AAU_ACT_BF16 = 0, AAU_ACT_FP16 = 1 };
enum { AAU_X_F32 = 0, AAU_X_U8 = 1 };
enum { AAU_IN_LOGITS = 0, AAU_IN_PROB = 1, AAU_IN_U8 = 2, AAU_IN_LOGIT_CUT = 3 };

/* Constructor arguments.
 * Replaces AttentionASPPUNet.__init__(in_channels=1, num_classes=1, base_c=32)
 *   attention_aspp_unet_pipeline_stage.py:111-122                       (variant = AAU_VARIANT_PIPELINE)
 * and AttentionASPPUNet.__init__(..., use_att, use_aspp, att_depth)
 *   test_ablation.py:168-203                                            (variant = AAU_VARIANT_ABLATION). */
typedef struct {
    int32_t in_channels;   /* must be 1 */
    int32_t num_classes;   /* must be 1 */
    int32_t base_c;        /* 16, 32, 48 or 64 (reference uses 16, 32, 48) */
    int32_t variant;       /* AAU_VARIANT_* */
    int32_t use_att;       /* ablation only */
    int32_t use_aspp;      /* ablation only */
    int32_t att_depth;     /* ablation only */
    int32_t act_dtype;     /* AAU_ACT_*: storage type of activations / packed weights (accumulation is fp32) */
} aau_config;

/* Replaces `AttentionASPPUNet(...).to(device)` (model_attention_aspp.py:36). */
int aau_create(const aau_config* cfg, int device, aau_handle** out);
int aau_destroy(aau_handle* h);
/* Message of the last failure on this handle (or on creation when h == NULL).  Never NULL. */
const char* aau_last_error(const aau_handle* h);

/* Replaces `load_state_dict(sd, strict=False)` (model_attention_aspp.py:37,
 * attention_aspp_unet_pipeline_stage.py:134-141): feed every floating-point state_dict entry by its reference
 * key (fp32, contiguous, host memory), then commit.  Unknown keys are ignored and reported through
 * aau_unexpected_count (strict=False semantics; the legacy spellings `.W_g.` / `.W_x.` are renamed as the
 * reference does).  Keys never loaded are counted by aau_missing_count: BatchNorm entries then take the module
 * defaults (gamma 1, beta 0, mean 0, var 1), while a missing convolution weight / bias makes the commit fail
 * with AAU_ERR_WEIGHTS (the host layer always feeds its own initialised parameters).  aau_commit_weights folds BatchNorm (eps 1e-5) in fp32, rounds ONCE to the
 * activation dtype, re-lays weights K-major for the tensor cores and uploads them (synchronous). */
int aau_load_tensor(aau_handle* h, const char* key, const float* data, int64_t numel);
int aau_commit_weights(aau_handle* h);
int aau_missing_count(const aau_handle* h);
int aau_unexpected_count(const aau_handle* h);
/* Number of state_dict keys this configuration owns and the i-th key / its element count (layout contract,
 * SURVEY.md section 8 a9). */
int aau_num_keys(const aau_handle* h);
const char* aau_key_name(const aau_handle* h, int i);
int64_t aau_key_numel(const aau_handle* h, int i);

/* Bytes of device scratch `aau_forward` needs for a batch of B frames of H x W (0 on invalid arguments). */
size_t aau_workspace_bytes(const aau_handle* h, int B, int H, int W);

/* Replaces `AttentionASPPUNet.forward(x)` (attention_aspp_unet_pipeline_stage.py:123-127; ablation twin
 * test_ablation.py:205-218).
 *   x       : device, [B,1,H,W] float32 in [0,1] (AAU_X_F32) or [B,H,W] uint8 (AAU_X_U8, normalised as
 *             float(u8)/255.0f like model_attention_aspp.py:17)
 *   logits  : device, float32 [B,1,H,W]
 *   psi3/psi2 : ablation variant only, device float32 [B,1,H/8,W/8] and [B,1,H/4,W/4]; may be NULL
 *   workspace : device, 256-byte aligned, at least aau_workspace_bytes(h,B,H,W)
 * H, W >= 16.  Asynchronous on `stream`. */
int aau_forward(aau_handle* h, const void* x, int x_dtype, int B, int H, int W, float* logits, float* psi3,
                float* psi2, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces the selection head: `torch.sigmoid(...)` (model_attention_aspp.py:54), `(prob > thr)` (:71),
 * `bin_.sum((1,2)).argmax()` (:74) and `select_fetal_abdomen_mask_and_frame` (:91-97).
 *   values : device float32 [N,H,W]: logits (AAU_IN_LOGITS: `sigmoid(x) > prob_thr` is decided as `x > cut`, cut = the largest
 *            fp32 whose correctly rounded fp32 sigmoid is <= prob_thr, found by bisection on the host -- no transcendental
 *            per pixel, bit exact; AAU_IN_LOGIT_CUT: `prob_thr` already IS that cutoff in logit space, e.g. the one of the
 *            caller's own sigmoid implementation), probabilities (AAU_IN_PROB, compared as they are), or a uint8 mask volume
 *            (AAU_IN_U8: `areas` receives the per-frame sum of byte values as `mask_3d.sum((1,2))` does, and
 *            `mask` receives `v > 0`; prob_thr is ignored);
 *   prob_thr: threshold on the probability (0.05 in the reference)
 *   areas  : device int32 [N] (overwritten);  best : device int32 [2] = {first arg-max index, its area}
 *   mask   : optional device uint8 [N,H,W] receiving the {0,1} volume, or NULL
 * Asynchronous on `stream`. */
int aau_frame_scores(aau_handle* h, const void* values, int input_kind, int N, int H, int W, float prob_thr,
                     int32_t* areas, int32_t* best, uint8_t* mask, void* stream);

/* First index of the maximum of `areas` (device int32 [N]) -> best[0], its value -> best[1]: the
 * `areas.argmax()` of model_attention_aspp.py:74,94 (numpy tie-break: lowest index).  Used after the per-batch
 * aau_frame_scores calls of a sweep, and on the host-gathered scores of a multi-GPU run.  Asynchronous. */
int aau_best_frame(aau_handle* h, const int32_t* areas, int N, int32_t* best, void* stream);

/* {0,1} mask (device uint8 [H,W]) of the frame an earlier aau_best_frame / aau_frame_scores call selected: `best` is that call's
 * device int32 [2] result, read ON THE DEVICE, so the index never has to visit the host between the two calls; an area of 0
 * gives the all-zero mask (model_attention_aspp.py:75-76).  `values` / `input_kind` / `prob_thr` as in aau_frame_scores
 * (float kinds only).  With it a sweep's whole device part is enqueued without a host synchronisation.  Asynchronous. */
int aau_best_frame_mask(aau_handle* h, const float* values, int input_kind, int N, int H, int W, float prob_thr,
                        const int32_t* best, uint8_t* mask, void* stream);

/* prob[i] = 1 / (1 + exp(-logits[i])), device fp32 -> device fp32, n elements: the `torch.sigmoid(self.net(x))` of
 * model_attention_aspp.py:54 for callers that want the probability volume itself (`predict`).  Asynchronous. */
int aau_sigmoid(aau_handle* h, const float* logits, int64_t n, float* prob, void* stream);

/* Flip test-time augmentation of the pipeline CLI (attention_aspp_unet_pipeline_stage.py:336-338, test_ablation.py:365-371):
 * `sigmoid((net(x) + flip(net(flip(x, [-1])), [-1])) / 2)`.  aau_flip_w mirrors `rows` rows of W elements (uint8 or fp32,
 * x_dtype as in aau_forward; not in place); aau_tta_prob combines the logits of the plain pass with those of the mirrored
 * pass (read mirrored back) into probabilities.  Device pointers, asynchronous on `stream`. */
int aau_flip_w(aau_handle* h, const void* x, int x_dtype, int64_t rows, int W, void* y, void* stream);
int aau_tta_prob(aau_handle* h, const float* logits, const float* logits_of_flipped, int64_t rows, int W, float* prob, void* stream);

/* Per-slice head and tail of the pipeline CLI's slice loop around the network (attention_aspp_unet_pipeline_stage.py:492-498,
 * test_ablation.py:826-834).
 * aau_resize_u8 replaces `Resize(IMG_SIZE, IMG_SIZE)` = cv2.resize(uint8, INTER_LINEAR) of the conditioned frame, bit exact
 *   with OpenCV's fixed-point scheme:  src device uint8 [N,SH,SW] -> dst device uint8 [N,DH,DW] (not in place).
 * aau_tail_masks replaces `cv2.resize(prob, (W, H))`, `cv2.GaussianBlur(prob, (5, 5), 0)` and `(prob > THR)`:
 *   prob device float32 [N,PH,PW] (the TTA probabilities) -> mask device uint8 [N,H,W] in {0,1}, areas device int32 [N]
 *   (overwritten) = pixels set per frame.  fp32 arithmetic; equal to OpenCV's up to its last-bit summation order.
 * Asynchronous on `stream`. */
int aau_resize_u8(aau_handle* h, const uint8_t* src, int N, int SH, int SW, uint8_t* dst, int DH, int DW, void* stream);
int aau_tail_masks(aau_handle* h, const float* prob, int N, int PH, int PW, int H, int W, float thr, uint8_t* mask,
                   int32_t* areas, void* stream);

/* Frame conditioning of the reference wrapper on the device, bit exact with the OpenCV calls it makes
 * (model_attention_aspp.py:11-17, inference.py:147-190): per frame `cv2.normalize(NORM_MINMAX, 0, 255)` -> uint8,
 * `cv2.createCLAHE(clipLimit=1.0, tileGridSize=(8,8)).apply`, `cv2.medianBlur(3)`.
 *   frames : device uint8 [N,H,W];  out : device uint8 [N,H,W] (not in place), to be fed to aau_forward as AAU_X_U8
 *            (which applies the reference's `/ 255.0`)
 *   workspace : device scratch of aau_condition_workspace_bytes(h, N) bytes.  Asynchronous on `stream`. */
size_t aau_condition_workspace_bytes(const aau_handle* h, int N);
int aau_condition_frames(aau_handle* h, const uint8_t* frames, int N, int H, int W, uint8_t* out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Host-only helpers (no device needed), exported for tests and for callers that prepare thresholds / weights themselves:
 *   aau_logit_cutoff: the largest fp32 x with sigmoid(x) <= prob_thr for a correctly rounded fp32 sigmoid -- what
 *     aau_frame_scores(AAU_IN_LOGITS) compares logits against (+inf for prob_thr >= 1, -inf for prob_thr < 0);
 *   aau_round_window_keep_sum: the rounding aau_commit_weights applies to the nine BN-folded taps of one (out, in) channel pair
 *     of a 3x3 convolution: nearest 16-bit values (fp16 != 0: IEEE half, else bfloat16 bit patterns), then single-ulp moves
 *     until the rounded taps sum to within half an ulp of the true window sum (every tap stays within one ulp of its value). */
float aau_logit_cutoff(float prob_thr);
int aau_round_window_keep_sum(const double* window9, int fp16, uint16_t* out9);

/* Device-side fault flag raised by a kernel whose internal pipeline wait timed out (0 = none).  Every mbarrier wait in the
 * kernels is bounded; on a timeout the kernel records a code (which role was waiting for what) and traps instead of hanging
 * the GPU.  The code is also written to mapped pinned host memory, so it is reported even though the trapped context can no
 * longer be read.  Synchronises the device. */
int aau_device_fault(aau_handle* h);

/* Debug / measurement aids. */
int aau_num_launches(const aau_handle* h);          /* kernels launched by the last aau_forward */
int aau_last_forward_was_graph(const aau_handle* h);   /* 1 when the last aau_forward replayed its kernels as ONE CUDA-graph launch */
int aau_num_ops(const aau_handle* h);               /* stream operations of the last aau_forward: its kernels plus stream-ordering
                                                        steps (the side-stream join of the ASPP image-pooling branch) */
/* Per-operation record of the LAST aau_forward: i in [0, aau_num_ops).  `layer` = the reference layer the
 * launch implements, `kernel` = kernel symbol, `flops` / `bytes` = its algorithmic work (2*MAC with dense tap
 * count; activations in + out + weights once).  `ms` is the CUDA-event time between this launch and the next on
 * the caller's stream when aau_set_option("profile", 1) was on during that forward, else -1 (synchronises). */
int aau_op_profile(aau_handle* h, int i, const char** layer, const char** kernel, float* ms, double* flops, double* bytes);
/* Copy one named intermediate of the LAST forward (NHWC, activation dtype) to `dst` (device, element count
 * returned through numel/C); names: x1 x2 x3 x4 p4 bridge d4 d3 d2 (used by layer-by-layer parity tests). */
int aau_debug_tensor(aau_handle* h, const char* name, void** ptr, int* B, int* H, int* W, int* C, int* ld, int* choff);
/* Planner options (test / measurement aids; every combination computes the same function, parity tests force them all):
 *   "amode"    A-operand staging of the 3x3 convolutions: -1 auto, 0 per-tap boxes, 1 halo slabs, 2 dx-stacked, 3 row-shifted
 *   "rs" "rs_mt" "mt" "mt_shape" "slab_max_bn" "dxn_full"   individual staging choices (row-shifted taps, stacked M-blocks,
 *              tile-shape search, un-split dx-stacked weights)
 *   "resident" 1/0 allow / forbid keeping a layer's whole weight matrix in shared memory
 *   "ctas" (0 auto, 1..2 CTAs per SM), "ng" (0 auto, 2 / 4 epilogue groups = TMEM stages), "cslots", "convt_batch", "lean"
 *   "pair"     CTA pairs (cta_group::2): bit 0 slab-staged layers, bit 1 per-tap staged layers, bit 2 resident-weight
 *              small-N layers, bit 3 resident-weight transposed convolutions (default 7)
 *   "tb"       1/0 per-tap staged tiles of small images may span two frames
 *   "spec"     1/0 use the kernel instantiations specialised per (staging mode, epilogue, K step, fused pool)
 *   "stem_tc"  1/0 uint8 frames run d1.0 on the tensor cores (0: packed-fp32 stem, as float frames always do)
 *   "stem_lo"  1/0 the tensor-core stem carries its weights as two 16-bit terms (hi + lo, two MMAs): ~22-bit first-layer weights
 *   "keep_sum" 1/0 3x3 weights are rounded to 16 bits with the rule that preserves every (out, in) window's tap sum
 *              (both re-run the weight preparation: they synchronise the device)
 *   "fusepool" "fusefix" "fixcc"   MaxPool2d / bilinear fix-up fused into the producing GEMM's epilogue
 *   "side" "pdl" "titer"   side stream for the ASPP pooling branch, programmatic dependent launch, incremental tile walk
 *   "fault_inject" 1: TEST HOOK -- the first tensor-core launch of the next plan gets a silent TMA producer in CTA 0, so that its
 *              MMA issuer runs into the bounded mbarrier wait: the kernel traps and aau_device_fault reports the code (the CUDA
 *              context is lost afterwards: use a scratch process)
 *   "fixcompact" 1/0 the fused transposed-conv + fix-up GEMM keeps and multiplies only the non-zero blocks of its weight tile
 *   "tapskip"  1/0 per-tap staged 3x3 layers (the dilated ASPP branches) skip taps whose whole box lies outside the image
 *   "aspp_merge" 1/0 ASPP blocks.0 (1x1) rides in the dilated branches' launch as the centre tap of a 3x3 (needs "tapskip")
 *   "graph"    CUDA-graph replay of the forward's launch sequence: -1 auto (batches of at most "graph_max_px" = B*H*W pixels,
 *              default 4 * 562 * 744: the launch-bound small-batch regime), 0 never, 1 always.  The sequence is captured once
 *              per (shape, workspace, input type) against library-owned input / output buffers in the workspace; a replay is
 *              copy-in, ONE graph launch, copy-out on a library-owned stream joined to the caller's stream by events.
 *   "profile"  0/1 record CUDA events around every launch of the following forwards (aau_op_profile). */
int aau_set_option(aau_handle* h, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* AAU_H_ */
