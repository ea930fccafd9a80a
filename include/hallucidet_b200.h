/* hallucidet_b200.h -- C ABI of libhallucidet_b200.so (sm_100a only).
 *
 * Drop-in boundary for the HalluciDet hot path: the reference is pure Python and
 * reaches its arithmetic through ATen/cuDNN operators; each entry point below replaces the operator the
 * reference's modules call at the cited site (paths relative to the reference repo; "TV:" = torchvision).
 * There is no CPU fallback: every compute entry point returns HD_ERR_CUDA when no sm_100 device is usable.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's allocator); the library never
 *     allocates, frees or retains them past the call; all work is enqueued on `stream`, no host sync;
 *   - activations: NHWC bf16, channel count a multiple of 16, base 16-byte aligned (hd_act);
 *   - module-edge tensors: NCHW fp32 (what the reference's modules exchange);
 *   - return value: 0 (HD_OK) or a negative hd_status; hd_last_error() gives a thread-local message.
 */
#ifndef HALLUCIDET_B200_H
#define HALLUCIDET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hd_stream; /* cudaStream_t */

typedef enum hd_status {
    HD_OK = 0,
    HD_ERR_BAD_ARG = -1,   /* shape / alignment / unsupported configuration */
    HD_ERR_CUDA = -2,      /* CUDA runtime or driver error (incl. no device, wrong arch) */
    HD_ERR_UNSUPPORTED = -3
} hd_status;

/* NHWC bf16 activation tensor [n][h][w][c], c contiguous. */
typedef struct hd_act {
    void* ptr;
    int32_t n, h, w, c;
} hd_act;

/* ---- library ------------------------------------------------------------------------------------ */
int hd_version(void);                 /* ABI version (1) */
const char* hd_last_error(void);      /* message of the last failing call on this thread */
int hd_device_ok(void);               /* HD_OK iff the current device is sm_100 */

/* ---- convolution as implicit GEMM on tcgen05 -------------------------------------------------------
 * Replaces nn.Conv2d forward / input-gradient / weight-gradient (cuDNN) at
 *   src/segmentation_models/base/modules.py:29-36 (Conv2dReLU conv), base/heads.py:24 (head conv),
 *   TV: models/resnet.py:59-105,108-163 (BasicBlock / Bottleneck convs), TV: ops/feature_pyramid_network.py:91-96.
 * Supported (kh,kw,stride): (1,1,1) (3,3,1) (1,1,2) (3,3,2); padding = k/2; dilation 1; groups 1.
 * The input may be the channel-concatenation of two tensors (x0 | x1) -- the U-Net skip concat
 * (decoders/unet/decoder.py:41) without materialising it; the dgrad output may be split the same way (y0 | y1).
 */
/* Train-mode BatchNorm finalize fused into the tail of the producing convolution (hd_conv_args.bn_fin): the LAST CTA of
 * hd_conv_fwd to finish sums the per-CTA statistics rows in a fixed order (fp64) and writes what hd_bn_finalize would:
 * mean, invstd, scale = gamma*invstd, shift = beta - mean*scale, and the running statistics (momentum, unbiased variance)
 * exactly as nn.BatchNorm2d (base/modules.py:42, TV: models/resnet.py:80-83).  Removes one latency-bound launch per layer. */
typedef struct hd_bn_fin {
    double count;             /* elements per channel (n*h*w of the conv output) */
    const float* gamma;
    const float* beta;
    float* running_mean;      /* may be NULL */
    float* running_var;       /* may be NULL */
    float* mean;              /* outputs, fp32 [channels] */
    float* invstd;
    float* scale;
    float* shift;
    uint32_t* counter;        /* one device word per layer, zero before the first launch; the kernel leaves it zero */
    float eps, momentum;
} hd_bn_fin;

typedef struct hd_conv_args {
    hd_act x0, x1;            /* fwd: input (x1.ptr NULL if unused).  dgrad: x0 = dY.  wgrad: input X */
    hd_act y0, y1;            /* fwd: output.  dgrad: dX (optionally split).  wgrad: y0 = dY (y1 unused) */
    const void* w;            /* fwd: bf16 [cout_pad][kh*kw*cin]; dgrad: bf16 [cin_pad][kh*kw*cout]; wgrad: unused */
    int32_t kh, kw, stride;
    /* epilogue (fwd / dgrad), applied in this order: +bias, +add, relu, mask, stats, stores */
    const float* bias;        /* [channels of y] or NULL */
    const void* add;          /* bf16 NHWC tensor shaped like y0 (single-output only), may alias y0.ptr (in-place) */
    const void* mask;         /* bf16 NHWC tensor shaped like y0: result zeroed where mask <= 0 (ReLU backward) */
    int32_t relu;
    int32_t sigmoid;          /* applied to the fp32 NCHW output only */
    float* stats;             /* fp32 [stats_replicas][2][channels]: partial sum / sum-of-squares of the (bf16-rounded) output, ONE ROW PER CTA of the persistent grid (each CTA adds up its own tiles in a fixed order), written (not accumulated), surplus rows zeroed: deterministic, no memset needed; NULL = off */
    int32_t stats_replicas;   /* rows available in `stats`; must be >= hd_conv_fwd_tiles(args) */
    float* out_f32_nchw;      /* optional fp32 NCHW copy of the output (first out_f32_channels channels) */
    int32_t out_f32_channels;
    int32_t store_bf16;       /* 1: write y0/y1 (bf16 NHWC); 0: only out_f32_nchw */
    int32_t phase_mask;       /* dgrad stride 2 only: bit (2*p+q) set = compute output phase (row parity p, col parity q); 0 = all */
    /* wgrad only */
    float* dw;                /* fp32 [cout][kh*kw][cin] accumulated with atomics (caller zeroes) */
    int32_t split_k;          /* 0 = auto */
    int32_t out_f32_nhwc;     /* fwd: 1 = the fp32 copy (out_f32_nchw) is channels-last [n][h][w][out_f32_channels] instead of NCHW
                                 (out_f32_channels % 16 == 0, no sigmoid): what cuDNN and the RoIAlign kernels read without a transpose */
    const hd_bn_fin* bn_fin;  /* fwd with stats: optional fused BatchNorm finalize (host pointer, copied at launch); NULL = off */
    void* workspace;          /* fwd / dgrad: optional device scratch of >= hd_conv_workspace_bytes() bytes, zero-initialised ONCE by the
                                 caller and then left to the library (flags are restored by the kernels); enables stream-K for
                                 problems whose tiles under-fill the SMs.  One workspace per stream of convolution launches. NULL = off */
    int64_t workspace_bytes;
} hd_conv_args;

int hd_conv_fwd(const hd_conv_args* a, hd_stream stream);
int hd_conv_fwd_tiles(const hd_conv_args* a);   /* statistics rows hd_conv_fwd needs for this problem (= CTAs of its grid; with y0.c unset: an upper bound); host-only, no launch */
int hd_conv_dgrad(const hd_conv_args* a, hd_stream stream);
int64_t hd_conv_workspace_bytes(void);           /* size of hd_conv_args.workspace */
int hd_conv_has_streamk(void);                   /* 1 if built with the (experimental, default-off) stream-K code paths */
/* Development aid (tools/conv_timeline.py): later hd_conv_fwd / hd_conv_dgrad launches write 8 %globaltimer stamps per CTA
 * into buf ([grid][8] int64, zeroed by the caller before each launch): kernel start, dependencies resolved, last TMA issued,
 * first operands landed, last MMA issued, last accumulator complete, epilogue done, CTA exit.  NULL = off (default). */
int hd_conv_debug_timestamps(void* buf);
int hd_conv_wgrad(const hd_conv_args* a, hd_stream stream);

/* Weight packing (once per optimizer step for the U-Net, once at load for the frozen detector).
 * w_oihw fp32 [cout][cin][kh][kw] (the reference's parameter layout); scale: optional per-cout multiplier
 * (frozen / eval BatchNorm fold, TV: ops/misc.py:54-63).  Any output pointer may be NULL.
 *   w_fwd   bf16 [cout_pad][k_pad]          k = (r*kw+s)*cin + ci          (zero padded)
 *   w_dgrad bf16 [cin_pad ][kh*kw*cout]     k = (r*kw+s)*cout + co
 *   w_t     bf16 [k_pad  ][cout_pad]        transpose of w_fwd (stem col2im GEMM)
 */
int hd_pack_conv_weight(const float* w_oihw, const float* scale, int cout, int cin, int kh, int kw,
                        void* w_fwd, int cout_pad, int k_pad, void* w_dgrad, int cin_pad, void* w_t, hd_stream stream);
/* dw_packed[co*row_stride + tap*tap_stride + ci] -> grad fp32 OIHW [cout][cin][kh][kw], multiplied by `scale`.
 * hd_conv_wgrad output: tap_stride = cin, row_stride = kh*kw*cin; stem GEMM output [cout][k_pad]: tap_stride = cin, row_stride = k_pad. */
int hd_unpack_wgrad(const float* dw_packed, float* grad_oihw, int cout, int cin, int kh, int kw, int tap_stride,
                    int row_stride, float scale, hd_stream stream);

/* Whole-network variants of the two calls above: one launch for every layer.  The descriptor tables live in DEVICE
 * memory; first_block is the running sum, over the preceding layers, of hd_pack_blocks(desc) for packing and of
 * hd_unpack_blocks(cout, cin, taps) for unpacking; total_blocks is the grand total. */
typedef struct hd_pack_desc {
    const float* w;       /* fp32 OIHW master weight */
    const float* scale;   /* optional per-cout scale (folded BN) */
    void* w_fwd;          /* [cout_pad][k_pad] bf16 */
    void* w_dgrad;        /* [cin_pad][kh*kw*cout] bf16 or NULL */
    void* w_t;            /* [k_pad][cout_pad] bf16 or NULL */
    int32_t cout, cin, kh, kw, cout_pad, k_pad, cin_pad, first_block;
    /* hd_adam_pack_conv_weights only: gradient and Adam moments of `w` (same OIHW layout), all NULL = pack only.  Allowed
     * only for layers of the tiled path (cout % 16 == 0, cin % 16 == 0, both bf16 layouts, no padding): there every master
     * weight is read exactly once per launch, so the update happens on that read. */
    const float* g;
    float* m;
    float* v;
} hd_pack_desc;
/* Adam hyper-parameters of one step (torch.optim.Adam, no weight decay / amsgrad; train_hallucidet.py:429-435,
 * src/config/config.py:215-219) with the gradient post-processing of the reference's loop fused in: g <- clamp(g * grad_scale,
 * -clip, clip) (grad_scale = 1/world after a summing all-reduce; clip_grad_value_(0.5), train_hallucidet.py:498-499; clip <= 0: off). */
typedef struct hd_adam_args {
    float lr, beta1, beta2, eps;
    float bias_correction1, bias_correction2;     /* 1 - beta^t of this step */
    float grad_scale, clip;
    float one_minus_beta1, one_minus_beta2;       /* computed in double by the caller (1 - 0.999f is 4.7e-5 off in fp32) */
} hd_adam_args;
typedef struct hd_adam_desc {   /* one small parameter tensor of hd_adam_multi */
    float* p;
    const float* g;
    float* m;
    float* v;
    int32_t n, first_block;
} hd_adam_desc;
typedef struct hd_unpack_desc {
    const float* dw;      /* packed fp32 weight gradient (see hd_conv_wgrad) */
    float* g;             /* fp32 OIHW gradient */
    int32_t cout, cin, taps, tap_stride, row_stride, first_block;
    float scale;
    int32_t pad_;
} hd_unpack_desc;
int hd_multi_blocks(int64_t elements);
int hd_pack_blocks(const hd_pack_desc* desc_host);   /* blocks one layer occupies in hd_pack_conv_weights (for first_block) */
int hd_pack_conv_weights(const hd_pack_desc* descs_dev, int n_layers, int total_blocks, hd_stream stream);
int hd_unpack_blocks(int cout, int cin, int taps);     /* blocks one layer occupies in hd_unpack_wgrads (for first_block) */
int hd_unpack_wgrads(const hd_unpack_desc* descs_dev, int n_layers, int total_blocks, hd_stream stream);
/* The optimizer tail in one pass over the parameters: hd_pack_conv_weights whose tiled layers ALSO apply
 * clip + Adam to the fp32 master weight as they read it (p, m, v written back, then both bf16 operand layouts from the new
 * value) -- instead of clamp, multi-tensor Adam and re-pack as three passes.  Layers with g == NULL are packed only. */
int hd_adam_pack_conv_weights(const hd_pack_desc* descs_dev, int n_layers, int total_blocks, const hd_adam_args* adam, hd_stream stream);
/* Element-wise Adam (same arithmetic) for the tensors that are not on the tiled path: BatchNorm weights / biases, the head, the stem.
 * first_block = running sum of hd_multi_blocks(n) over the preceding descriptors. */
int hd_adam_multi(const hd_adam_desc* descs_dev, int n_tensors, int total_blocks, const hd_adam_args* adam, hd_stream stream);

/* ---- stem 7x7 stride-2 pad-3 conv via explicit patches (cin = 3 is not TMA-addressable) --------------
 * Replaces encoder.conv1 (encoders/resnet.py:50) and body.conv1 (TV: models/resnet.py:197) im2col/col2im.
 * x fp32 NCHW [n][3][h][w] -> patches bf16 [n*ho*wo][k_pad], k = (r*7+s)*3 + c, ho = h/2, wo = w/2. */
int hd_stem_im2col(const float* x_nchw, void* patches, int n, int h, int w, int k_pad, hd_stream stream);
/* Single-channel stem input (the U-Net's IR plane, replicated x3 by src/utils/utils.py:52-53 -> conv with the channel-summed
 * filter): x [n][1][h][w] as fp32 (x_dtype 0) or uint8 (x_dtype 1), multiplied by `scale` (1/255 for the camera bytes,
 * src/dataloader/dataloader.py:13-73) -> patches bf16 [n*ho*wo][k_pad], k = r*7 + s (49 taps, zero padded to k_pad >= 56). */
int hd_stem_im2col_1ch(const void* x, int x_dtype, float scale, void* patches, int n, int h, int w, int k_pad, hd_stream stream);
/* Fused stem forward (csrc/stem_conv.cu): y = relu?(conv7x7/2(x * x_scale, w * w_scale[cout]) + bias) as bf16 NHWC [n][h/2][w/2][64],
 * straight from the fp32 OIHW master weight [64][3][7][7] -- no patch matrix in HBM.  cin = 3: x fp32 NCHW [n][3][h][w];
 * cin = 1: x is the single plane [n][1][h][w] (fp32, x_dtype 0, or uint8, x_dtype 1) of an input whose three channels are equal,
 * convolved with the channel-summed filter.  Optional BatchNorm statistics of the output (stats: fp32 [stats_rows][2][64], one
 * row per CTA, stats_rows >= hd_stem_fwd_rows) and fused finalize (bn_fin), as hd_conv_fwd. */
int hd_stem_fwd(const void* x, int x_dtype, float x_scale, int cin, const float* w_oihw, const float* w_scale, const float* bias,
                int relu, void* y, int n, int h, int w, float* stats, int stats_rows, const hd_bn_fin* bn_fin, hd_stream stream);
int hd_stem_fwd_rows(int n, int h, int w);
/* dpatches bf16 [n*ho*wo][k_pad] -> dx fp32 NCHW [n][3][h][w] (overwrites). */
int hd_stem_col2im(const void* dpatches, float* dx_nchw, int n, int h, int w, int k_pad, hd_stream stream);

/* ---- train-mode BatchNorm2d (U-Net; base/modules.py:42, TV: models/resnet.py:80-83) --------------------
 * stats rows (per-tile partials from hd_conv_fwd, summed in a fixed order) -> per-channel scale/shift
 * (+ running-stat update, momentum/unbiased var as nn.BatchNorm2d). */
int hd_bn_finalize(const float* stats, int stats_replicas, int channels, double count, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* mean_out, float* invstd_out, float* scale_out, float* shift_out, hd_stream stream);
/* y = relu?( z*scale+shift + (res ? (res_scale ? res*res_scale+res_shift : res) : 0) ), all bf16 NHWC, n_pix pixels. */
int hd_bn_apply(const void* z, const float* scale, const float* shift, const void* res, const float* res_scale,
                const float* res_shift, int relu, void* y, int64_t n_pix, int channels, hd_stream stream);
/* Backward, pass 1: g = dy masked by the ReLU that followed the BN: (y_relu > 0) when y_relu is given (residual blocks:
 * the block output), else (z*relu_scale+relu_shift > 0) recomputed from z when relu_scale is given, else no mask.
 * sums[0][c] = sum g, sums[1][c] = sum g*xhat (atomics; caller zeroes). */
int hd_bn_bwd_reduce(const void* dy, const void* y_relu, const float* relu_scale, const float* relu_shift, const void* z,
                     const float* mean, const float* invstd, float* sums, int64_t n_pix, int channels, hd_stream stream);
/* Backward, pass 2: dz = gamma*invstd*(g - sums0/count - xhat*sums1/count); optional g_out = g (masked dy);
 * dgamma = sums1, dbeta = sums0 are written (fp32, scaled by grad_scale) when non-NULL. */
int hd_bn_bwd_apply(const void* dy, const void* y_relu, const float* relu_scale, const float* relu_shift, const void* z,
                    const float* mean, const float* invstd, const float* gamma, const float* sums, double count, void* dz, void* g_out, float* dgamma,
                    float* dbeta, int64_t n_pix, int channels, hd_stream stream);

/* Both backward passes in ONE persistent kernel (one CTA per SM, grid barrier in the middle): same arithmetic as
 * hd_bn_bwd_reduce + hd_bn_bwd_apply; layers whose per-CTA slice of (g, z) fits in shared memory read them from HBM once.
 * sums: fp32 [2][channels], zeroed by the caller; barrier_words: two zero-initialised device words (count, generation), kept
 * by the caller across launches (the kernel restores count = 0 itself). */
int hd_bn_bwd_fused(const void* dy, const void* y_relu, const float* relu_scale, const float* relu_shift, const void* z,
                    const float* mean, const float* invstd, const float* gamma, float* sums, double count, void* dz, void* g_out,
                    float* dgamma, float* dbeta, int64_t n_pix, int channels, uint32_t* barrier_words, hd_stream stream);

/* ---- memory-bound glue ------------------------------------------------------------------------------ */
/* MaxPool2d(3,2,1) after the stem ReLU (encoders/resnet.py:51, TV: models/resnet.py:199). */
/* idx (optional, [n][h/2][w/2][c] bytes, 8-byte aligned): window position r*3+s of the first maximum of every output
 * element, 15 = "nobody" (mask_nonpositive: also when the maximum is <= 0, i.e. a following ReLU mask would kill it). */
int hd_maxpool_fwd(const hd_act* x, const hd_act* y, void* idx, int mask_nonpositive, hd_stream stream);
/* dx = (add ? add : 0) + scatter of dy to the arg-max taps (first max in window order, as ATen); optional ReLU mask
 * (x > 0).  With idx (from hd_maxpool_fwd; pass mask_nonpositive there instead of relu_mask here) x / y are not read. */
int hd_maxpool_bwd(const hd_act* x, const hd_act* y, const void* dy, const void* add, void* dx, int relu_mask,
                   const void* idx, hd_stream stream);
/* Nearest upsample x2 (decoders/unet/decoder.py:7-8): y[h][w] = x[h/2][w/2]; backward = 2x2 sum. */
int hd_upsample2x_fwd(const hd_act* x, const hd_act* y, hd_stream stream);
int hd_upsample2x_bwd(const hd_act* dy, const hd_act* dx, hd_stream stream);
/* FPN top-down (TV: ops/feature_pyramid_network.py:193-196): y += nearest_resize(x to y's size); backward dx = gather-sum(dy). */
int hd_add_nearest_fwd(const hd_act* x, const hd_act* y, hd_stream stream);
int hd_add_nearest_bwd(const hd_act* dy, const hd_act* dx, int accumulate, hd_stream stream);
/* Odd feature maps (detector size 300 -> 75 / 19 pixels): stride-2 convolutions run on an even zero-padded copy.
 * y = x zero-padded at the bottom / right to y's size;  dx = (crop(dxp) + add) masked by (mask > 0) (add / mask may be NULL). */
int hd_pad_hw(const hd_act* x, const hd_act* y, hd_stream stream);
int hd_crop_add_mask(const hd_act* dxp, const void* add, const void* mask, const hd_act* dx, hd_stream stream);
/* Layout / dtype converters at the module edges. */
int hd_nchw_f32_to_nhwc_bf16(const float* x, const hd_act* y, int channels, int accumulate, hd_stream stream);
int hd_nhwc_bf16_to_nchw_f32(const hd_act* x, float* y, int channels, hd_stream stream);
/* Segmentation head backward prologue: dlogits = dhal * hal * (1-hal) (fp32 NCHW [n][3][h][w]) -> bf16 NHWC, c = 16 (zero padded);
 * dbias[c] += sum (atomics, caller zeroes). */
int hd_sigmoid_bwd_pack(const float* dhal, const float* hal, const hd_act* dlogits, int channels, float* dbias, hd_stream stream);

/* ---- detector input transform (src/models/custom_generalized_transform.py:136-186, 52-100, 256-274) ------
 * y[b][c][i][j] = (x[b][c][src(i)][src(j)] - mean[c]) / std[c], src(d) = min(floor(d * fp32(in/out)), in-1). */
int hd_resize_nearest_fwd(const float* x, float* y, int n, int c, int h_in, int w_in, int h_out, int w_out,
                          const float* mean, const float* std, hd_stream stream);
/* dx = sum of dy over the output pixels that sampled each input pixel, divided by std[c]; overwrites dx (or accumulates). */
int hd_resize_nearest_bwd(const float* dy, float* dx, int n, int c, int h_in, int w_in, int h_out, int w_out,
                          const float* std, int accumulate, hd_stream stream);

/* ---- hallucination regulariser (src/losses/losses.py:28-48; train_hallucidet.py:173-176) -----------------
 * kind 0 = MSE, 1 = L1.  loss[0] += w_rgb*pixel(rgb,hal), loss[1] += w_ir*pixel(ir3,hal) (mean over n*3*h*w; caller zeroes),
 * dhal (fp32 NCHW, may be NULL) receives (accumulate ? += : =) the gradient of (loss[0]+loss[1]) * grad_scale.
 * ir is [n][1][h][w] (broadcast to 3 channels, src/utils/utils.py:52-53); rgb/hal are [n][3][h][w]. */
int hd_regulariser(int kind, const float* hal, const float* rgb, const float* ir, float w_rgb, float w_ir,
                   int n, int h, int w, float* loss, float* dhal, float grad_scale, int accumulate, hd_stream stream);

/* ---- greedy NMS of the detector's proposal filter / final detections -----------------------------------------
 * Replaces the kernels behind torchvision.ops.nms (torchvision csrc/ops/cuda/nms_kernel.cu: nms_kernel_impl +
 * gather_keep_from_mask), which the reference reaches through torchvision's RPN.filter_proposals and
 * RoIHeads.postprocess_detections (src/models/detector.py builds torchvision detectors).  Same IoU predicate (fp32, same
 * operation order), same greedy rule, so the keep set is identical.
 * boxes_sorted: [total][4] fp32 x1,y1,x2,y2 -- `problems` independent box lists back to back, each sorted by descending
 * score (the caller sorts: stable, descending, as torchvision does).  offsets: HOST array of problems+1 box offsets
 * (slots reserved per problem, n_p <= 8192).  counts_dev: optional DEVICE array with the number of boxes actually present
 * in each problem (<= its slots; the remaining slots get keep = 0), so a caller with data-dependent counts needs no sync.
 * mask_ws: device scratch of sum_p n_p*ceil(n_p/64) 64-bit words.  keep: [total] bytes, 1 = kept. */
int hd_nms(const float* boxes_sorted, const int* offsets, const int* counts_dev, int problems, float iou_threshold,
           void* mask_ws, unsigned char* keep, hd_stream stream);
/* hd_nms with valid_dev: optional DEVICE bytes, one per box slot; a box with 0 is never kept and suppresses nothing, wherever it
 * sits in its problem (the proposal filter's size / score tests of TV rpn.py:265-276 need no compaction before the NMS). */
int hd_nms_valid(const float* boxes_sorted, const int* offsets, const int* counts_dev, const unsigned char* valid_dev, int problems,
                 float iou_threshold, void* mask_ws, unsigned char* keep, hd_stream stream);

/* ---- Balanced positive / negative sampler (RPN loss, RoI heads) without a host round trip ----------------------
 * torchvision det_utils.BalancedPositiveNegativeSampler as the reference's detector uses it (TV models/detection/rpn.py
 * compute_loss, roi_heads.py subsample; reached from src/utils/eval_forward_fasterrcnn.py:62-99 and :112-128) for a batch
 * of label rows, drawn on the device from the Philox stream that torch.randperm would consume, so the selection is
 * bit-identical to torchvision's loop on the same generator state -- but the data-dependent randperm sizes never travel
 * to the host.  labels: [batch][n] float32 (labels_dtype 0) or int64 (1); >= 1 positive, 0 negative, anything else is
 * ignored.  seed / *offset_in: the CUDA generator's Philox seed and offset; *offset_out (a different device word) receives
 * the offset after the 2 * batch randperm calls.  sampled: [batch][n] bytes out, 1 = drawn positive, 2 = drawn negative,
 * 0 = not drawn.  counts: [batch][4] int32 out = (positives, negatives, drawn positives, drawn negatives).
 * batch <= 64, batch_size_per_image <= 1024.  workspace: hd_sample_balanced_workspace_bytes(batch) bytes, 16-byte aligned. */
int64_t hd_sample_balanced_workspace_bytes(int batch);
int hd_sample_balanced(const void* labels, int labels_dtype, int batch, int n, int batch_size_per_image, int num_pos_max,
                       uint64_t seed, const uint64_t* offset_in, uint64_t* offset_out, uint8_t* sampled, int32_t* counts,
                       void* workspace, int64_t workspace_bytes, hd_stream stream);

/* ---- RoI-head target assignment and sample gathering (two launches instead of ~80 element-wise ones) --------------
 * torchvision RoIHeads.select_training_samples (TV models/detection/roi_heads.py: add_gt_proposals, box_iou + Matcher,
 * labels, the gathers after the sampler, BoxCoder.encode) followed by MultiScaleRoIAlign's RoI format + LevelMapper
 * (TV ops/poolers.py:47-101), as the reference reaches them from src/utils/eval_forward_fasterrcnn.py:112-131 -- the same
 * fp32 operations in the same order, so the results are bit-identical to the operator chain.
 * hd_roi_match_labels: candidates of image b = props[b][0 .. slots) (live below n_props[b], a DEVICE count) followed by the
 * image's n_gt (<= 64, padded; gt_present marks real rows) ground-truth boxes; labels[b][c] = class of the matched box,
 * 0 = background (IoU below low_threshold), -1 = ignored (between the thresholds, or padding); matched[b][c] = matched
 * ground-truth row (clamped at 0).  Matcher without low-quality matches (RoI heads).
 * hd_roi_gather_samples: rows = batch * batch_size_per_image drawn candidates (flat = ascending positions b * (slots + n_gt)
 * + c, padded; counts = hd_sample_balanced's) -> their boxes, labels (-100 on padding rows), matched rows, image index,
 * regression targets (encode_boxes with `weights`), RoIs [rows][5] and pyramid level; out_n_drawn / out_per_image: the drawn
 * totals (device int64). */
typedef struct hd_roi_gather_args {
    const int64_t* flat;
    const int32_t* counts;
    const float* props;
    const float* gt;
    const int64_t* labels;
    const int64_t* matched;
    int batch, slots, n_gt, rows;
    float weights[4];
    float canonical_scale, canonical_level, eps, k_min, k_max;
    float* out_props;
    int64_t* out_labels;
    int64_t* out_matched;
    int64_t* out_image;
    float* out_targets;
    float* out_rois;
    int64_t* out_levels;
    int64_t* out_n_drawn;
    int64_t* out_per_image;
} hd_roi_gather_args;
int hd_roi_match_labels(const float* props, const int64_t* n_props, const float* gt, const uint8_t* gt_present,
                        const int64_t* gt_labels, int batch, int slots, int n_gt, float low_threshold, float high_threshold,
                        int64_t* labels, int64_t* matched, hd_stream stream);
int hd_roi_gather_samples(const hd_roi_gather_args* args, hd_stream stream);

/* ---- RPN predictor maps <-> flattened per-anchor tensors --------------------------------------------------------------
 * torchvision concat_box_prediction_layers (TV rpn.py:81-110, called from RegionProposalNetwork.forward; reference
 * src/utils/eval_forward_fasterrcnn.py:76-78) for channels-last predictor maps preds[l] = [batch][hw[l]][channel_pitch[l]]
 * fp32 whose channels are `anchors_per_pixel` objectness logits followed by 4 * anchors_per_pixel box deltas:
 * direction 0 writes objectness [batch][sum_l hw[l] * a] and deltas [batch][sum_l hw[l] * a][4] (level-major, then pixel,
 * then anchor -- torchvision's order); direction 1 is the adjoint (gradients back into the maps, padding channels = 0).
 * One launch for all levels instead of ~12 forward / ~30 backward (autograd's slice / cat backward). */
int hd_rpn_concat_preds(void* const* preds, const int* hw, const int* channel_pitch, int levels, int batch, int anchors_per_pixel,
                        float* objectness, float* deltas, int direction, hd_stream stream);

/* ---- RPN anchor targets ----------------------------------------------------------------------------------------------
 * RegionProposalNetwork.assign_targets_to_anchors + BoxCoder.encode (TV rpn.py:190-229, _utils.py:139-160; reference
 * src/utils/eval_forward_fasterrcnn.py:93-95) for a batch that shares one anchor set: box_iou(gt, anchors), Matcher with
 * (allow_low_quality != 0) or without low-quality matches, labels[b][a] = 1 foreground / 0 background / -1 ignored and the
 * regression target of every anchor against its matched box -- bit-identical to the operator chain.  gt [batch][n_gt][4] padded
 * (n_gt <= 64), gt_present marks real rows; coder_weights: HOST array of 4; highest_ws: batch * n_gt * 4 bytes of device scratch. */
int hd_rpn_assign_targets(const float* anchors, const float* gt, const uint8_t* gt_present, int batch, int n_anchors, int n_gt,
                          float low_threshold, float high_threshold, int allow_low_quality, const float* coder_weights,
                          void* highest_ws, float* labels, float* regression_targets, hd_stream stream);

/* ---- RPN proposal filter, element-wise front end for the selected candidates ------------------------------------------
 * BoxCoder.decode (TV models/detection/_utils.py:162-226) + sigmoid + clip_boxes_to_image + the size / score tests of
 * RegionProposalNetwork.filter_proposals (TV rpn.py:263-276; reference src/utils/eval_forward_fasterrcnn.py:80-84) for the
 * candidates idx[b][j] (anchor index of image b; the per-level top-k) only -- bit-identical to decoding every anchor first.
 * objectness [batch][anchors] logits, deltas [batch][anchors][4], anchors [anchors][4] (the same for every image),
 * coder_weights: HOST array of 4.  boxes [batch][selected][4], scores [batch][selected], valid [batch][selected] bytes. */
int hd_rpn_decode_selected(const float* objectness, const float* deltas, const float* anchors, const int64_t* idx, int batch,
                           int anchors_per_image, int selected, const float* coder_weights, float xform_clip, float img_w,
                           float img_h, float min_size, float score_thresh, float* boxes, float* scores, uint8_t* valid,
                           hd_stream stream);

/* ---- Detection losses, value and gradient in one launch -------------------------------------------------------------
 * hd_fastrcnn_loss: torchvision roi_heads.fastrcnn_loss (cross-entropy over the sampled proposals, mean; smooth-L1 with
 * `beta` of the matched class's box deltas over the foreground rows, summed, / number of sampled rows) as the reference's
 * roi_heads_eval calls it (src/utils/eval_forward_fasterrcnn.py:133-136).  labels [rows] int64: class, 0 = background,
 * -100 = padding row (ignored, as F.cross_entropy's ignore_index).  losses [2] = (classification, box regression);
 * grad_logits [rows][classes] / grad_box [rows][4*classes] = d(loss) / d(input), fully written.
 * hd_rpn_loss: RegionProposalNetwork.compute_loss (TV rpn.py; reference :96-99) on the sampled anchors given as a row list:
 * flat [rows] int64 = positions of the drawn anchors in the flattened [batch * anchors] arrays (rows past the drawn total
 * are padding), sampled = hd_sample_balanced's byte map (1 = positive), counts = its counts.  losses [2] = (objectness,
 * box regression); grad_objectness [batch * anchors] / grad_deltas [batch * anchors][4] must be zeroed by the caller. */
int hd_fastrcnn_loss(const float* class_logits, const float* box_regression, const int64_t* labels, const float* regression_targets,
                     int rows, int num_classes, float beta, float* losses, float* grad_logits, float* grad_box, hd_stream stream);
int hd_rpn_loss(const float* objectness, const float* pred_bbox_deltas, const float* labels, const float* regression_targets,
                const int64_t* flat, const uint8_t* sampled, const int32_t* counts, int batch, int rows, float beta, float* losses,
                float* grad_objectness, float* grad_deltas, hd_stream stream);

/* ---- RoIAlign backward (box head pooling over the FPN levels) ------------------------------------------------
 * grad_in_nhwc ([n][h][w][c] fp32 channels-last scratch, zeroed by the caller) += gradient of
 * torchvision.ops.roi_align(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, aligned = False) -- the op
 * torchvision's MultiScaleRoIAlign calls per FPN level inside the reference's roi_heads_eval
 * (src/utils/eval_forward_fasterrcnn.py:130) -- for grad_out [num_rois][c][pooled_h][pooled_w]; rois [num_rois][5] =
 * (batch index, x1, y1, x2, y2).  Same per-contribution arithmetic as torchvision's roi_align_backward_kernel_impl, issued
 * as coalesced 16-byte vector reductions over the channels.  c % 4 == 0, c <= 256, pooled_h*pooled_w <= 49, ratio 1..2.
 * hd_nhwc_to_nchw_f32 converts the scratch to the [n][c][h][w] layout autograd expects. */
int hd_roi_align_bwd_nhwc(const float* grad_out, const float* rois, float* grad_in_nhwc, int num_rois, int channels, int height,
                          int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, hd_stream stream);
int hd_nhwc_to_nchw_f32(const float* x_nhwc, float* y_nchw, int n, int channels, int height, int width, hd_stream stream);
/* Forward of the same op on a channels-last copy of the input (hd_nchw_to_nhwc_f32): out [num_rois][c][pooled_h][pooled_w].
 * Same expression per output element as torchvision's roi_align_forward_kernel_impl (samples in (iy, ix) order, then / count);
 * every sample read is a coalesced 16-byte-per-lane access over the channels instead of 16 scattered 4-byte gathers. */
int hd_roi_align_fwd_nhwc(const float* feat_nhwc, const float* rois, float* out, int num_rois, int channels, int height,
                          int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, hd_stream stream);
int hd_nchw_to_nhwc_f32(const float* x_nchw, float* y_nhwc, int n, int channels, int height, int width, hd_stream stream);
/* Whole-pyramid variants (torchvision.ops.MultiScaleRoIAlign, TV ops/poolers.py): RoI k is pooled from level
 * level_of_roi[k] (DEVICE int64 array, the output of torchvision's LevelMapper), so all levels are one launch and the
 * per-level index lists -- one host sync per level in torchvision -- disappear.  levels: HOST array of n_levels (<= 8)
 * entries; forward reads feat_nhwc, backward accumulates into grad_nhwc (zeroed by the caller). */
typedef struct hd_roi_level {
    const void* feat_nhwc;    /* [n][h][w][c] fp32 (hd_roi_align_ml_fwd) or bf16 (hd_roi_align_ml_fwd_bf16) */
    float* grad_nhwc;         /* [n][h][w][c] fp32 */
    int32_t h, w;
    float scale;              /* spatial_scale of the level */
    int32_t pad_;
} hd_roi_level;
int hd_roi_align_ml_fwd(const hd_roi_level* levels, int n_levels, const float* rois, const int64_t* level_of_roi, float* out,
                        int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio, hd_stream stream);
/* the same with bf16 feature maps: half the L2 traffic; identical results when the fp32 maps are widened copies of the bf16 ones */
int hd_roi_align_ml_fwd_bf16(const hd_roi_level* levels, int n_levels, const float* rois, const int64_t* level_of_roi, float* out,
                             int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio, hd_stream stream);
int hd_roi_align_ml_bwd(const hd_roi_level* levels, int n_levels, const float* grad_out, const float* rois,
                        const int64_t* level_of_roi, int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio,
                        hd_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* HALLUCIDET_B200_H */
