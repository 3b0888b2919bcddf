/* yolox_b200 — C ABI of the B200-native YOLOX inference hot path.
 *
 * The reference (aiha-lab/COCO-dataset-based-light-weight-fast-object-detection-model) has no
 * FFI/plugin interface: its hot-path boundary is the Python call surface of L2.
 * Each entry point below is what a reference-side binding for that surface would call; the
 * `replaces:` line cites the reference code (relative to the reference root) whose device work it
 * performs.  Python shims in the package mirror the reference signatures on top of these
 * (see INTEGRATION.md for the ctypes stubs).
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host sync inside
 *    (except yx_engine_profile / yx_engine_create, which are setup/diagnostic calls).
 *  - return 0 on success, a negative yx_status otherwise; yx_last_error() gives the message.
 *  - caller owns all buffers; the library owns only the opaque engine object.
 *  - activations: NHWC fp16 inside the engine; public tensors keep the reference layouts.
 */
#ifndef YOLOX_B200_H_
#define YOLOX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YX_ABI_VERSION 4

typedef enum yx_status {
  YX_OK = 0,
  YX_ERR_INVALID = -1,   /* bad argument / unsupported shape            */
  YX_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed         */
  YX_ERR_UNSUPPORTED = -3
} yx_status;

typedef enum yx_act {  /* replaces: get_activation, yolox/models/network_blocks.py:12-24 */
  YX_ACT_NONE = 0,
  YX_ACT_SILU = 1,
  YX_ACT_HSWISH = 2,
  YX_ACT_RELU = 3,
  YX_ACT_LRELU = 4     /* LeakyReLU(0.1) */
} yx_act;

typedef enum yx_dtype {
  YX_F16 = 0,
  YX_F32 = 1,
  YX_U8 = 2   /* images only (engine input, yx_preprocess_batch output): pixel values 0..255, which is what the reference's
                 float batches hold (PIL uint8 pixels); the engine evaluates them exactly as the same values in fp16 */
} yx_dtype;

typedef enum yx_op_kind {
  YX_OP_CONV = 0,      /* BaseConv.fused_forward: conv(k in {1,3}, stride in {1,2}; k = 4 with stride 2; pad (k-1)/2)+bias+act(+residual)
                          replaces: network_blocks.py:73-84,199-205 ; yolox_infer/models/blocks.py:21-49 */
  YX_OP_S2D = 1,       /* Focus / FocusCustom space-to-depth of the NCHW input image into NHWC[.,.,.,16]
                          replaces: network_blocks.py:330-361 ; blocks.py:286-304 */
  YX_OP_SPP = 2,       /* SPPBottleneck max-pools 5/9/13 written into the concat buffer
                          replaces: network_blocks.py:239-246 */
  YX_OP_UPSAMPLE = 3,  /* nn.Upsample(2,"nearest") into a concat slice; replaces: yolo_pafpn_p6.py:153-164 */
  YX_OP_DWCONV = 4     /* depthwise kxk conv + bias + act; replaces: network_blocks.py:107-120 (dconv) */
} yx_op_kind;

/* A view of an NHWC fp16 tensor inside the activation arena.  Channel slices of a wider buffer
 * (concat elimination) are expressed with pitch > c. */
typedef struct yx_view {
  int64_t offset;   /* byte offset from the arena base to element (0,0,0,0) of the view */
  int64_t nstride;  /* elements between consecutive images (>= h*w*pitch; e.g. A*C for a pyramid
                       level of a [B,A,C] head output) */
  int32_t n, h, w, c;
  int32_t pitch;    /* elements between consecutive pixels of the parent buffer (>= c, multiple of 8) */
  int32_t _pad;
} yx_view;

typedef struct yx_op {
  int32_t kind;       /* yx_op_kind */
  int32_t ksize;      /* CONV: 1, 3, or 4 (stride 2 only); DWCONV: 3, 5 */
  int32_t stride;     /* CONV: 1 or 2 */
  int32_t act;        /* yx_act */
  yx_view src;        /* S2D: ignored (reads the external image) */
  yx_view dst;
  yx_view res;        /* residual added after the activation; res.c == 0 -> none */
  yx_view up;         /* CONV (1x1, stride 1): low-resolution tensor [n, h/2, w/2, c_up] whose nearest-neighbour x2
                         upsampling is CONCATENATED IN FRONT of src along channels (yolo_pafpn_p6.py:153-154): the
                         upsample and the torch.cat are folded into the conv's TMA loads (stride-0 tensor-map
                         dimensions repeat each pixel 2x2).  up.c == 0 -> none; up.c must be a multiple of 64.
                         Weights cover [up channels | src channels] in that order */
  int64_t w_offset;   /* byte offset into the weight blob: fp16 [cout_pad][k*k][cin_pad] (DWCONV: [k*k][c]) */
  int64_t b_offset;   /* byte offset into the bias blob: fp32 [cout_pad] */
  int32_t cin_pad;    /* multiple of 16 */
  int32_t cout_pad;   /* multiple of 16 */
  int32_t aux;        /* S2D: bit0 = channel order (0 Focus [TL,BL,TR,BR], 1 pixel_unshuffle), bit1 = padded rows
                         [0 | W/2 pixels | 0 0 0].  CONV: 1 = row-packed 3x3 over that padded 16-channel tensor
                         (weights [cout_pad][3 (dy)][48 = dx*16 + c]; cin_pad = 48) */
  int32_t _pad;
} yx_op;

typedef struct yx_engine yx_engine;

const char* yx_last_error(void);
int yx_abi_version(void);

/* Build an engine from a fully planned op list (the Python graph builder owns the topology).
 * arena / weights / biases must stay alive and at the same address for the engine's lifetime.
 * replaces: the nn.Module graph built by build_yolox, choijhanyangackr/main.py:31-59. */
int yx_engine_create(const yx_op* ops_host, int n_ops, void* arena, size_t arena_bytes,
                     const void* weights, size_t weights_bytes, const void* biases, size_t bias_bytes,
                     int in_h, int in_w, int batch, yx_engine** out);
void yx_engine_destroy(yx_engine* e);

/* Run every op on `stream`.  image: NCHW [batch,3,in_h,in_w], fp16 or fp32 (yx_dtype).
 * in_scale/in_shift: optional input affine applied while the image is read (the predict loop's
 * img.mul_(0.9).add_(11.4), main.py:164, evaluated in the image dtype); pass 1, 0 to disable.
 * use_graph != 0 replays a captured CUDA graph for everything after the first (image-reading) op.
 * replaces: YOLOXP6.forward / YOLOX.forward, yolox_infer/models/yolox_p6.py:31-34. */
int yx_engine_run(yx_engine* e, const void* image, int image_dtype, float in_scale, float in_shift, int use_graph,
                  void* stream);

/* Diagnostic: run ops [first, first+count) only (used by the per-op parity tests, which check every
 * op of a real network against a CPU evaluation of the SAME device inputs). */
int yx_engine_run_ops(yx_engine* e, const void* image, int image_dtype, float in_scale, float in_shift, int first,
                      int count, void* stream);

/* Diagnostic: per-op device time (ms, mean over iters, CUDA events on `stream`), host-synchronising.
 * Also reports algorithmic flops / bytes per op so callers can print roofline fractions. */
int yx_engine_profile(yx_engine* e, const void* image, int image_dtype, int iters, void* stream,
                      float* ms_host, double* flops_host, double* bytes_host, int n_ops);
int yx_engine_num_launches(const yx_engine* e); /* kernels launched by one yx_engine_run */

/* Launch-shape knobs of the conv kernel (see csrc/yx_conv.cu).  yx_engine_tune picks them per layer by timing;
 * yx_conv2d_ex lets the parity tests force every shape.  A shape that does not fit the layer is YX_ERR_INVALID. */
typedef struct yx_conv_tune {
  int32_t variant;             /* 1 = one TMA box per filter tap, 2 = 3x3/stride-1 halo tile + nine descriptors */
  int32_t n_tile;              /* output channels per tile: multiple of 64, or the whole padded Cout; <= 256 */
  int32_t ctas_per_sm;         /* 1 or 2 */
  int32_t halves;              /* variant 2: 128-pixel halves stacked per CTA (1 or 2) */
  int32_t epilogue_groups;     /* 1 or 2 epilogue warpgroups */
  int32_t staging_buffers;     /* 1 or 2 output staging buffers */
  int32_t second_producer;     /* 0 = one TMA producer warp per operand, 1 = add a second one */
  int32_t no_resident_weights; /* 1 = always stream the weights through the ring */
  int32_t cta_pair;            /* 1 = two-CTA clusters issuing cta_group::2 MMAs (M = 256), half of the weights per CTA */
  int32_t sparse;              /* 1 = the 2:4 sparse tensor-core variant (tcgen05.mma.sp, weights as the sparse A operand);
                                  only valid for a layer whose weights are 2:4-compliant along Cin */
  int32_t epilogue_alternate;  /* with two epilogue groups: 0 = both convert half of every tile's columns, 1 = the groups
                                  alternate tiles (each owns one accumulator and one staging buffer) */
  int32_t reserved;            /* 0 */
} yx_conv_tune;

/* Per-layer launch-shape selection by measurement (the counterpart of torch.backends.cudnn.benchmark = True, which
 * the reference sets in tools/eval.py:122).  Runs the network once on `image`, timing every candidate shape of every
 * conv on its real inputs (min of `iters` CUDA-event timings) and keeping the fastest; the arena holds a valid forward
 * result afterwards.  Host-synchronising setup call. */
int yx_engine_tune(yx_engine* e, const void* image, int image_dtype, float in_scale, float in_shift, int iters,
                   void* stream);
/* With YX_TUNE_CHECK=1 in the environment yx_engine_tune also verifies that every candidate shape reproduces the default
 * shape's output; returns the number of candidates that did not (and were rejected) and their descriptions. */
int yx_engine_tune_mismatches(const yx_engine* e, char* buf_host, int buf_len);
/* Launch shape currently selected for conv op i (YX_ERR_INVALID for a non-conv op) / force one.  Together they let the
 * caller PERSIST the tuner's choices (yolox_b200/plan.py keeps them in a cache file keyed by device, kernel revision and
 * layer geometry) so that launch shapes -- and with them the fp32 summation order, i.e. the low bits of the result -- are
 * reproducible from run to run.  yx_engine_set_tune re-plans the op; a shape that does not fit is YX_ERR_INVALID and
 * leaves the op unchanged.  yx_engine_mark_tuned tells the engine not to tune again. */
int yx_engine_get_tune(const yx_engine* e, int i, yx_conv_tune* out_host);
int yx_engine_set_tune(yx_engine* e, int i, const yx_conv_tune* tune_host);
/* Human-readable description of op i and of the launch shape chosen for it (diagnostics / profiles). */
int yx_engine_op_desc(const yx_engine* e, int i, char* buf_host, int buf_len);
/* 1 when op i is a conv whose weights are 2:4-compliant and were packed for the sparse tensor-core variant at creation (the
 * ingest contract of choijhanyangackr/main.py:52-55: masks arrive as zeros in the dense weights); the tuner then has more
 * candidates for this layer, so a persisted launch-shape choice must be keyed on it. */
int yx_engine_op_sparse_ok(const yx_engine* e, int i);

/* ---- stand-alone operators (used by tests and by the reference-style Python functions) ---------- */

/* One conv op outside an engine (same kernel the engine launches). Views are relative to `base`. */
int yx_conv2d(const yx_op* op_host, void* base, const void* weights, const void* biases, void* stream);

int yx_conv2d_ex(const yx_op* op_host, void* base, const void* weights, const void* biases, const yx_conv_tune* tune_host,
                 void* stream);

/* ---- head decode / candidate selection / NMS --------------------------------------------------- */

typedef struct yx_levels {
  int32_t n_levels;
  int32_t h[8], w[8], stride[8];
} yx_levels;

typedef enum yx_nms_mode {
  YX_NMS_TRICK = 0,     /* torchvision batched_nms coordinate trick: boxes + label*(max+1) */
  YX_NMS_VANILLA = 1,   /* per-class NMS on the raw boxes */
  YX_NMS_AGNOSTIC = 2,  /* class-agnostic torchvision.ops.nms */
  YX_NMS_AUTO = 3       /* per image, torchvision 0.26's CUDA dispatch: VANILLA when the candidate
                           boxes passed to batched_nms have numel() > 100000 (n > 25000), else TRICK */
} yx_nms_mode;

/* decode, infer flavour.  reg/obj/cls are raw logits with element strides (so permuted views and
 * the engine's packed [B,A,8] reg+obj buffer can be passed as they are); outputs fp32 contiguous
 * boxes[B,A,4] (xyxy), obj_conf[B,A], cls_conf[B,A,C] (= sigmoid(cls)*sigmoid(obj)).
 * replaces: yolox_postprocess_output_torch_batch, yolox_infer/postprocess_utils.py:27-52. */
int yx_decode_infer(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb, int64_t obj_sa,
                    const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B, int A, int C,
                    const yx_levels* lv_host, float* boxes, float* obj_conf, float* cls_conf, void* stream);

/* The same decode with the anchor geometry READ from the caller's device tensors grids (1,A,2) = (x, y) and
 * scales (1,A,1), of dtype geom_dtype (YX_F16 / YX_F32, promoted to fp32 as torch promotes them): the values the
 * caller built are the values used (a custom grid offset included), and no pyramid description has to be recovered.
 * replaces: yolox_postprocess_output_torch_batch(reg, obj, cls, grids, scales), postprocess_utils.py:27-52. */
int yx_decode_infer_grids(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                          int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B, int A,
                          int C, const void* grids, const void* scales, int geom_dtype, float* boxes, float* obj_conf,
                          float* cls_conf, void* stream);

/* Workspace (bytes) for the detection entry points below. */
size_t yx_detect_workspace_bytes(int B, int A);

/* Candidate selection + top-k + batched NMS + top-max_det from DECODED fp32 tensors.
 * det[B,max_det,7] = [x1,y1,x2,y2,obj,cls_conf,label], det_count[B]; det_anchor[B,max_det] (may be NULL).
 * max_nms <= 0: no candidate cap.  Rows beyond det_count are zero.
 * replaces: yolox_nms_torch_batch (default mode), yolox_infer/postprocess_utils.py:55-129 and the
 * torchvision.ops.batched_nms / nms it calls (yolox_infer/nms.py:19,40). */
int yx_nms_main(const float* boxes, const float* obj_conf, const float* cls_conf, int B, int A, int C,
                float conf_thr, float nms_thr, int max_nms, int max_det, int mode, void* workspace,
                size_t workspace_bytes, float* det, int32_t* det_count, int32_t* det_anchor, void* stream);

/* How yx_nms_main_ex forms candidates (yolox_infer/postprocess_utils.py:74-95). */
typedef enum yx_cand_mode {
  YX_CAND_MAX = 0,          /* :86-89 one per anchor: first-max class, kept when max >= conf_thr (every shipped config) */
  YX_CAND_MULTI_CLASS = 1,  /* :90-95 multi_class=True: one per (anchor, class) with cls_conf >= conf_thr */
  YX_CAND_RMMOP = 2         /* :74-84 rmmop=(r1, r2): top-1 class per anchor, kept when top1 >= top2*r1 and
                               obj^2 >= top1*r2; conf_thr is not applied; needs C >= 2 */
} yx_cand_mode;

/* Workspace (bytes) for yx_nms_main_ex: multi_class ranks up to A*C candidates per image. */
size_t yx_nms_workspace_bytes(int B, int A, int C, int cand_mode, int max_nms);

/* yx_nms_main with the candidate rule selectable.  In multi_class mode a detection's label is the candidate's class
 * and det rows default to A*C when max_det <= 0.
 * replaces: yolox_nms_torch_batch(..., multi_class=..., rmmop=...), yolox_infer/postprocess_utils.py:55-129. */
int yx_nms_main_ex(const float* boxes, const float* obj_conf, const float* cls_conf, int B, int A, int C,
                   float conf_thr, float nms_thr, int max_nms, int max_det, int mode, int cand_mode, float rmmop_r1,
                   float rmmop_r2, void* workspace, size_t workspace_bytes, float* det, int32_t* det_count,
                   int32_t* det_anchor, void* stream);

/* Fused decode + threshold + compaction + NMS straight from the raw head logits (no [B,A,C]
 * fp32 tensor is ever materialised).  Same results as yx_decode_infer followed by yx_nms_main.
 * replaces: main.py:180-188 (decode + NMS of the predict loop). */
int yx_detect_main(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb, int64_t obj_sa,
                   const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B, int A, int C,
                   const yx_levels* lv_host, float conf_thr, float nms_thr, int max_nms, int max_det, int mode,
                   void* workspace, size_t workspace_bytes, float* det, int32_t* det_count, int32_t* det_anchor,
                   void* stream);

/* ---- multi-GPU: detections gathered by the NMS kernel itself ------------------------------------ *
 * Images shard over ranks (one process per GPU); every rank needs all detections.  Instead of a separate
 * collective, the NMS kernel's tail stores each image's [max_det,7] rows and count straight into every rank's receive
 * window (peer memory mapped with CUDA IPC, NVLink / NVSwitch posted writes) and counts the image in that rank's arrival
 * counter; a one-warp kernel on the receiving stream then waits until all ranks' images of the step have arrived.
 * replaces: the pickled gloo gather of yolox/evaluators/coco_evaluator.py:127 + yolox/utils/dist.py:224-265.
 *
 * The window is ordinary device memory owned by the caller (a torch tensor); yx_ipc_export / yx_ipc_open turn it into
 * pointers valid in the other ranks' processes (the 64-byte handles travel through any host-side exchange). */
#define YX_MAX_PEERS 8
#define YX_IPC_HANDLE_BYTES 64

/* handle of the allocation that contains dev_ptr + dev_ptr's byte offset inside it */
int yx_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out);
/* maps a peer's allocation into this process (peer access enabled lazily); *base_out + offset = the peer's pointer */
int yx_ipc_open(const void* handle, void** base_out);
int yx_ipc_close(void* base);

enum { YX_PEER_WAIT_NONE = 0, YX_PEER_WAIT_AFTER = 1, YX_PEER_WAIT_BEFORE = 2 };

typedef struct yx_peer_out {
  int32_t world;                /* number of ranks (1..YX_MAX_PEERS), own rank included */
  int32_t wait_target;          /* value every local arrival counter reaches when the awaited step's images are all here:
                                   step_index * B as a wrapping 32-bit counter (counters are never reset) */
  int32_t timeout_ms;           /* bound of the device-side wait (<= 0: 10 s); on expiry *status = 1 + late rank */
  int32_t wait_mode;            /* YX_PEER_WAIT_AFTER: wait for THIS step's rows right after the NMS (lock-step);
                                   YX_PEER_WAIT_BEFORE: wait for an EARLIER step (wait_target) before this step's selection
                                   kernels: its rows arrived while this step's network ran, nobody waits for the slowest rank;
                                   YX_PEER_WAIT_NONE: no wait in this call (yx_peer_wait completes the step on demand) */
  void* det[YX_MAX_PEERS];      /* this rank's [B,max_det,7] fp32 block inside rank w's window */
  void* cnt[YX_MAX_PEERS];      /* this rank's [B] int32 block inside rank w's window */
  void* arrive[YX_MAX_PEERS];   /* rank w's int32 arrival counter for this rank */
  void* local_arrive;           /* this rank's own int32[world] counters */
  void* status;                 /* int32 in this rank's memory, 0 while healthy */
  void* wait_cnt;               /* [world,B] int32 counts of the AWAITED window in this rank's memory (a late rank's counts
                                   are zeroed on timeout), or NULL */
} yx_peer_out;

/* yx_detect_main + the gather described above.  Requires max_det > 0 and the same B on every rank.  The caller rotates
 * over THREE windows on consecutive steps: with YX_PEER_WAIT_BEFORE a rank may run one whole step ahead of a peer, and a
 * window is overwritten three steps later, after every reader of it has issued its next step. */
int yx_detect_main_gather(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb, int64_t obj_sa,
                          const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B, int A, int C,
                          const yx_levels* lv_host, float conf_thr, float nms_thr, int max_nms, int max_det, int mode,
                          void* workspace, size_t workspace_bytes, float* det, int32_t* det_count, int32_t* det_anchor,
                          const yx_peer_out* peer, void* stream);

/* Stream-ordered completion of one gathered step on the receiving side (the wait yx_detect_main_gather performs in its
 * AFTER / BEFORE modes, on demand): returns once every rank's B images counted by wait_target have arrived. */
int yx_peer_wait(const void* local_arrive, int world, int wait_target, void* status, int timeout_ms, void* wait_cnt, int B,
                 void* stream);

/* yolox-package head output: out[B,A,5+C] = [reg, sigmoid(obj), sigmoid(cls)] in `out_dtype`, then
 * (decode != 0) decoded in place like decode_outputs.
 * replaces: yolox/models/yolo_head.py:167-168,186-190,210-225. */
int yx_head_assemble(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb, int64_t obj_sa,
                     const void* cls, int64_t cls_sb, int64_t cls_sa, int B, int A, int C, const yx_levels* lv_host,
                     int decode, void* out, int out_dtype, void* stream);
/* In-place decode of outputs[..., :4]; replaces: YOLOXHead.decode_outputs, yolo_head.py:210-225. */
int yx_decode_outputs(void* outputs, int dtype, int B, int A, int C, const yx_levels* lv_host, void* stream);

/* yolox.utils.postprocess: converts prediction[:,:,:4] to xyxy IN PLACE, then threshold
 * (obj*class_conf >= conf_thr) and NMS without caps.  det[B,A,7] rows = [x1,y1,x2,y2,obj,class_conf,class_pred].
 * replaces: postprocess, yolox/utils/boxes.py:32-82. */
int yx_postprocess_yolox(void* prediction, int dtype, int B, int A, int C, float conf_thr, float nms_thr, int mode,
                         void* workspace, size_t workspace_bytes, float* det, int32_t* det_count,
                         int32_t* det_anchor, void* stream);

/* ---- the rows either side of the hot path ----------------------------------- */

/* Device-side pre-processing of decoded RGB uint8 images (packed HWC, image b at src + src_off[b]):
 * aspect-preserving Pillow-BILINEAR resize (bit-identical to PIL.Image.resize, two passes with uint8 rounding),
 * top-left paste into a 114-filled [B,3,Hp,Wp] batch, RGB -> BGR, NCHW, values 0..255 in fp16 / fp32.
 * geom[B][4] = (h, w, new_h, new_w); bounds_* / kk_* are Pillow's per-axis coefficient tables (first input index and
 * tap count per output index; 22-bit fixed-point weights, ks_* per output index), built by the Python shim exactly
 * like Resample.c's precompute_coeffs.  All pointers are device pointers.
 * replaces: yolox_load_one_image_pil + yolox_collate_batch, choijhanyangackr/yolox_infer/preprocess_utils.py:9-55. */
int yx_preprocess_batch(const void* src, const int64_t* src_off, const int32_t* geom, const int32_t* bounds_h,
                        const int32_t* kk_h, const int32_t* bounds_v, const int32_t* kk_v, int ks_h, int ks_v, int B,
                        int Hp, int Wp, void* out, int out_dtype, void* stream);

/* det[B,max_det,7] (+ det_count[B]) -> records[B,max_det,6] = [x, y, w, h, score, category_id]: corners divided by
 * scale[b] (fp32 division), xyxy -> xywh, score = det[4]*det[5], category_id = class_ids[(int)det[6]]; rows beyond
 * det_count are zero.  replaces: convert_to_coco_format, choijhanyangackr/common/utils.py:27-73. */
int yx_coco_records(const float* det, const int32_t* det_count, int B, int max_det, const float* scale,
                    const int32_t* class_ids, int n_classes, float* records, void* stream);

/* ---- COCO bounding-box evaluation (host code) ---------------------------------------- *
 * The computation behind COCOEvaluator.evaluate_prediction: pycocotools COCOeval(gt, dt, "bbox") evaluate() +
 * accumulate() + summarize() (ten IoU thresholds .50:.05:.95, 101 recall points, area ranges all/small/medium/large,
 * 1/10/100 detections per image).  Arrays are host memory; bboxes are [x, y, w, h] doubles; image_ids / category_ids
 * list what is evaluated (duplicates are ignored).  stats12 = the twelve numbers summarize() prints (AP, AP50, AP75,
 * APs, APm, APl, AR1, AR10, AR100, ARs, ARm, ARl; -1 where undefined).  precision_out [10][101][K][4][3] and
 * recall_out [10][K][4][3] are optional (NULL), K = number of distinct category ids, ascending.
 * replaces: yolox/evaluators/coco_evaluator.py:198-215 (pycocotools / yolox.layers.COCOeval_opt, csrc/cocoeval). */
int yx_cocoeval_bbox(const int64_t* gt_image, const int32_t* gt_category, const double* gt_bbox, const double* gt_area,
                     const int32_t* gt_iscrowd, int64_t n_gt, const int64_t* dt_image, const int32_t* dt_category,
                     const double* dt_bbox, const double* dt_score, int64_t n_dt, const int64_t* image_ids,
                     int64_t n_images, const int32_t* category_ids, int32_t n_categories, double* stats12,
                     double* precision_out, double* recall_out);

#ifdef __cplusplus
}
#endif
#endif /* YOLOX_B200_H_ */
