// roi_targets.cu -- target assignment and sample gathering of the RoI heads as two launches.
//
// The reference's detector is torchvision's Faster R-CNN; in training its RoI heads run, per image,
//   add_gt_proposals -> box_iou -> Matcher -> labels -> BalancedPositiveNegativeSampler -> gathers -> BoxCoder.encode
// (TV models/detection/roi_heads.py select_training_samples; reached from src/utils/eval_forward_fasterrcnn.py:112-128) and
// MultiScaleRoIAlign then converts the sampled boxes to RoI format and maps them to pyramid levels (TV ops/poolers.py).
// hallucidet_b200/detection.py restates that for the whole batch with ~80 element-wise PyTorch launches on small tensors: a
// few microseconds each, back to back on the critical path between the proposal filter and RoIAlign.  The two kernels here
// perform the same fp32 operations in the same order (no fused multiply-adds, IEEE division / sqrt / log), element for
// element, so their results are bit-identical to the PyTorch operator chain:
//   hd_roi_match_labels    box_iou(gt, [proposals | gt]) -> Matcher (no low-quality matches) -> class label per candidate
//   hd_roi_gather_samples  the drawn candidates (ascending flat positions) -> proposals, labels, matched gt, regression
//                          targets (encode_boxes), RoIs (image index, box) and FPN level (LevelMapper)
#include <math.h>
#include <string.h>

#include "hd_common.cuh"

namespace hd {

namespace {

constexpr int kMaxGt = 64;

__device__ __forceinline__ float box_area_rn(float x1, float y1, float x2, float y2) {
    return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
}

// candidates of image b: columns [0, T) = proposals (live below n_props[b]), [T, T + G) = the image's ground-truth boxes
__global__ void __launch_bounds__(256) roi_match_labels_kernel(const float4* __restrict__ props, const long long* __restrict__ n_props,
                                                             const float4* __restrict__ gt, const unsigned char* __restrict__ gt_present,
                                                             const long long* __restrict__ gt_labels, int B, int T, int G,
                                                             float low_thr, float high_thr, long long* __restrict__ labels,
                                                             long long* __restrict__ matched) {
    pdl_trigger();
    pdl_wait();
    __shared__ float4 s_gt[kMaxGt];
    __shared__ float s_area[kMaxGt];
    __shared__ unsigned char s_pres[kMaxGt];
    __shared__ long long s_lab[kMaxGt];
    const int b = blockIdx.y, N = T + G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float4 q = gt[b * G + g];
        s_gt[g] = q;
        s_area[g] = box_area_rn(q.x, q.y, q.z, q.w);
        s_pres[g] = gt_present[b * G + g];
        s_lab[g] = gt_labels[b * G + g];
    }
    __syncthreads();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    const float4 p = col < T ? props[static_cast<long>(b) * T + col] : s_gt[col - T];
    const bool present = col < T ? static_cast<long long>(col) < n_props[b] : s_pres[col - T] != 0;
    const float area2 = box_area_rn(p.x, p.y, p.z, p.w);
    float best = 0.f;
    int best_g = 0;
    for (int g = 0; g < G; ++g) {
        float v = -1.f;                                          // padded ground-truth rows never win
        if (s_pres[g]) {
            const float4 q = s_gt[g];
            const float w = fmaxf(__fsub_rn(fminf(q.z, p.z), fmaxf(q.x, p.x)), 0.f);
            const float h = fmaxf(__fsub_rn(fminf(q.w, p.w), fmaxf(q.y, p.y)), 0.f);
            const float inter = __fmul_rn(w, h);
            v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(s_area[g], area2), inter));
        }
        if (g == 0 || v > best) { best = v; best_g = g; }        // torch.max: the first maximal value
    }
    long long m = best_g;
    if (best < low_thr) m = -1;                                  // Matcher.BELOW_LOW_THRESHOLD
    else if (best < high_thr) m = -2;                            // Matcher.BETWEEN_THRESHOLDS
    const long long clamped = m < 0 ? 0 : m;
    long long lab = s_lab[clamped];
    if (m == -1) lab = 0;
    if (m == -2) lab = -1;
    if (!present) lab = -1;                                      // padding: ignored by the sampler
    const long idx = static_cast<long>(b) * N + col;
    labels[idx] = lab;
    matched[idx] = clamped;
}

struct GatherParams {
    const long long* flat;          // [S] ascending flat positions (image * N + column) of the drawn candidates, padded
    const int* counts;              // [B][4] sampler counts (.., .., drawn positives, drawn negatives)
    const float4* props;            // [B][T]
    const float4* gt;               // [B][G]
    const long long* labels;        // [B][N]
    const long long* matched;       // [B][N]
    int B, T, G, S;
    float wx, wy, ww, wh;           // BoxCoder weights
    float inv_s0, lvl0, eps, k_min, k_max;
    float4* out_props;              // [S]
    long long* out_labels;          // [S]   (-100 for padding rows)
    long long* out_matched;         // [S]
    long long* out_image;           // [S]
    float4* out_targets;            // [S]
    float* out_rois;                // [S][5]
    long long* out_levels;          // [S]
    long long* out_n_drawn;         // [1]
    long long* out_per_image;       // [B]
};

__global__ void __launch_bounds__(256) roi_gather_samples_kernel(const GatherParams P) {
    pdl_trigger();
    pdl_wait();
    const int N = P.T + P.G;
    long long n_drawn = 0;
    for (int b = 0; b < P.B; ++b) n_drawn += P.counts[b * 4 + 2] + P.counts[b * 4 + 3];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) *P.out_n_drawn = n_drawn;
    if (s < P.B) P.out_per_image[s] = P.counts[s * 4 + 2] + P.counts[s * 4 + 3];
    if (s >= P.S) return;
    const long long f = P.flat[s];
    const int img = static_cast<int>(f / N), col = static_cast<int>(f - static_cast<long long>(img) * N);
    const float4 p = col < P.T ? P.props[static_cast<long>(img) * P.T + col] : P.gt[img * P.G + (col - P.T)];
    const long long m = P.matched[f];
    const float4 r = P.gt[img * P.G + static_cast<int>(m)];
    P.out_props[s] = p;
    P.out_labels[s] = s < n_drawn ? P.labels[f] : -100;
    P.out_matched[s] = m;
    P.out_image[s] = img;
    // encode_boxes (TV models/detection/_utils.py:75-119)
    const float ex_w = __fsub_rn(p.z, p.x), ex_h = __fsub_rn(p.w, p.y);
    const float ex_cx = __fadd_rn(p.x, __fmul_rn(0.5f, ex_w)), ex_cy = __fadd_rn(p.y, __fmul_rn(0.5f, ex_h));
    const float gt_w = __fsub_rn(r.z, r.x), gt_h = __fsub_rn(r.w, r.y);
    const float gt_cx = __fadd_rn(r.x, __fmul_rn(0.5f, gt_w)), gt_cy = __fadd_rn(r.y, __fmul_rn(0.5f, gt_h));
    float4 t;
    t.x = __fdiv_rn(__fmul_rn(P.wx, __fsub_rn(gt_cx, ex_cx)), ex_w);
    t.y = __fdiv_rn(__fmul_rn(P.wy, __fsub_rn(gt_cy, ex_cy)), ex_h);
    t.z = __fmul_rn(P.ww, logf(__fdiv_rn(gt_w, ex_w)));
    t.w = __fmul_rn(P.wh, logf(__fdiv_rn(gt_h, ex_h)));
    P.out_targets[s] = t;
    // _convert_to_roi_format + LevelMapper (TV ops/poolers.py:47-84)
    float* roi = P.out_rois + static_cast<long>(s) * 5;
    roi[0] = static_cast<float>(img); roi[1] = p.x; roi[2] = p.y; roi[3] = p.z; roi[4] = p.w;
    const float sz = sqrtf(box_area_rn(p.x, p.y, p.z, p.w));
    float lvl = floorf(__fadd_rn(__fadd_rn(P.lvl0, log2f(__fmul_rn(sz, P.inv_s0))), P.eps));
    lvl = fminf(fmaxf(lvl, P.k_min), P.k_max);
    P.out_levels[s] = static_cast<long long>(lvl) - static_cast<long long>(P.k_min);
}

constexpr int kMaxPredLevels = 8;

struct PredLevels {
    float* pred[kMaxPredLevels];        // [B][hw][cp] channels-last predictor output (or its gradient) of the level
    int hw[kMaxPredLevels];             // pixels per image
    int cp[kMaxPredLevels];             // channel pitch (>= 5 * a)
    int first_pixel[kMaxPredLevels];    // pixels of the earlier levels (per image)
    int levels, a, pixels;              // anchors per pixel, pixels per image over all levels
};

// objectness [B][pixels * a] / deltas [B][pixels * a][4] <-> per-level channels-last predictor maps (channels: a objectness
// logits, then 4 * a box deltas): torchvision's concat_box_prediction_layers (TV rpn.py:81-110) and its adjoint.
template <bool kBackward>
__global__ void __launch_bounds__(256) rpn_concat_kernel(const PredLevels L, int B, float* __restrict__ obj, float* __restrict__ deltas) {
    pdl_trigger();
    pdl_wait();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B * L.pixels) return;
    const int b = q / L.pixels, px = q - b * L.pixels;
    int l = 0;
    while (l + 1 < L.levels && px >= L.first_pixel[l + 1]) ++l;
    const int hw = px - L.first_pixel[l];
    float* p = L.pred[l] + (static_cast<long>(b) * L.hw[l] + hw) * L.cp[l];
    const long o = (static_cast<long>(b) * L.pixels + px) * L.a;
    if (L.a == 3 && L.cp[l] == 16) {
        // the stock RPN head (3 anchors per pixel, 15 channels padded to 16): one pixel = four 16-byte words of the map, its
        // twelve box deltas = three aligned 16-byte words of the flat tensor (48 bytes per pixel)
        float4* p4 = reinterpret_cast<float4*>(p);
        float4* d4 = reinterpret_cast<float4*>(deltas + o * 4);
        if (!kBackward) {
            const float4 v0 = p4[0], v1 = p4[1], v2 = p4[2], v3 = p4[3];
            obj[o] = v0.x; obj[o + 1] = v0.y; obj[o + 2] = v0.z;
            d4[0] = make_float4(v0.w, v1.x, v1.y, v1.z);
            d4[1] = make_float4(v1.w, v2.x, v2.y, v2.z);
            d4[2] = make_float4(v2.w, v3.x, v3.y, v3.z);
        } else {
            const float4 e0 = d4[0], e1 = d4[1], e2 = d4[2];
            p4[0] = make_float4(obj[o], obj[o + 1], obj[o + 2], e0.x);
            p4[1] = make_float4(e0.y, e0.z, e0.w, e1.x);
            p4[2] = make_float4(e1.y, e1.z, e1.w, e2.x);
            p4[3] = make_float4(e2.y, e2.z, e2.w, 0.f);
        }
        return;
    }
    if (!kBackward) {
        for (int j = 0; j < L.a; ++j) obj[o + j] = p[j];
        for (int j = 0; j < 4 * L.a; ++j) deltas[o * 4 + j] = p[L.a + j];
    } else {
        for (int j = 0; j < L.a; ++j) p[j] = obj[o + j];
        for (int j = 0; j < 4 * L.a; ++j) p[L.a + j] = deltas[o * 4 + j];
        for (int j = 5 * L.a; j < L.cp[l]; ++j) p[j] = 0.f;
    }
}

// RegionProposalNetwork.assign_targets_to_anchors + BoxCoder.encode (TV rpn.py:190-229, _utils.py:139-160) for the whole
// batch: box_iou(gt, anchors), Matcher WITH low-quality matches (every anchor that attains a ground-truth box's best IoU gets
// its plain arg-max back), labels 1 / 0 / -1 and the regression target of every anchor -- two launches instead of ~85
// element-wise ones over [B, G, A] tensors (they run on a side stream next to the backbone's persistent kernels, where every
// small launch has to wait for a gap: the chain used to finish AFTER the proposal filter it is supposed to hide under).
// Pass 1: per ground-truth box the best IoU over all anchors (bit pattern maximum: IoUs are >= 0); pass 2: the assignment.
__device__ __forceinline__ float iou_rn(const float4 q, float area_q, const float4 p, float area_p) {
    const float w = fmaxf(__fsub_rn(fminf(q.z, p.z), fmaxf(q.x, p.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(q.w, p.w), fmaxf(q.y, p.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_q, area_p), inter));
}

template <int kPass>
__global__ void __launch_bounds__(256) rpn_assign_kernel(const float4* __restrict__ anchors, const float4* __restrict__ gt,
                                                         const unsigned char* __restrict__ gt_present, int B, int A, int G, float low_thr,
                                                         float high_thr, int allow_low_quality, float wx, float wy, float ww, float wh,
                                                         unsigned int* __restrict__ highest, float* __restrict__ labels,
                                                         float4* __restrict__ targets) {
    pdl_trigger();
    pdl_wait();
    __shared__ float4 s_gt[kMaxGt];
    __shared__ float s_area[kMaxGt];
    __shared__ unsigned char s_pres[kMaxGt];
    __shared__ unsigned int s_hi[kMaxGt];
    const int b = blockIdx.y;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float4 q = gt[b * G + g];
        s_gt[g] = q;
        s_area[g] = box_area_rn(q.x, q.y, q.z, q.w);
        s_pres[g] = gt_present[b * G + g];
        s_hi[g] = kPass == 1 ? 0u : highest[b * G + g];
    }
    __syncthreads();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < A;
    const float4 p = live ? anchors[a] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float area_p = box_area_rn(p.x, p.y, p.z, p.w);
    if (kPass == 1) {
        for (int g = 0; g < G; ++g) {
            if (!s_pres[g]) continue;                                    // (uniform per block)
            unsigned int v = live ? __float_as_uint(iou_rn(s_gt[g], s_area[g], p, area_p)) : 0u;
            for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0) atomicMax(&s_hi[g], v);
        }
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += blockDim.x)
            if (s_pres[g]) atomicMax(&highest[b * G + g], s_hi[g]);
        return;
    }
    if (!live) return;
    float best = 0.f;
    int best_g = 0;
    bool restore = false;
    for (int g = 0; g < G; ++g) {
        float v = -1.f;
        if (s_pres[g]) {
            v = iou_rn(s_gt[g], s_area[g], p, area_p);
            restore = restore || __float_as_uint(v) == s_hi[g];        // attains this box's best IoU (exact equality, as torch)
        }
        if (g == 0 || v > best) { best = v; best_g = g; }
    }
    int m = best_g;
    if (best < low_thr) m = -1;
    else if (best < high_thr) m = -2;
    if (allow_low_quality && restore) m = best_g;
    const long idx = static_cast<long>(b) * A + a;
    labels[idx] = m >= 0 ? 1.f : (m == -2 ? -1.f : 0.f);
    // regression target against the matched box (clamped index 0 for unmatched anchors, as torchvision gathers it)
    const float4 r = s_gt[m < 0 ? 0 : m];
    const float ex_w = __fsub_rn(p.z, p.x), ex_h = __fsub_rn(p.w, p.y);
    const float ex_cx = __fadd_rn(p.x, __fmul_rn(0.5f, ex_w)), ex_cy = __fadd_rn(p.y, __fmul_rn(0.5f, ex_h));
    const float gt_w = __fsub_rn(r.z, r.x), gt_h = __fsub_rn(r.w, r.y);
    const float gt_cx = __fadd_rn(r.x, __fmul_rn(0.5f, gt_w)), gt_cy = __fadd_rn(r.y, __fmul_rn(0.5f, gt_h));
    float4 t;
    t.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gt_cx, ex_cx)), ex_w);
    t.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gt_cy, ex_cy)), ex_h);
    t.z = __fmul_rn(ww, logf(__fdiv_rn(gt_w, ex_w)));
    t.w = __fmul_rn(wh, logf(__fdiv_rn(gt_h, ex_h)));
    targets[idx] = t;
}

// The RPN proposal filter's element-wise front end for the SELECTED candidates only: torchvision decodes all anchors
// (BoxCoder.decode_single, TV models/detection/_utils.py:188-226), then gathers the per-level top-k, applies sigmoid, clips to
// the image and tests size / score (TV rpn.py:263-276) -- ~45 launches, 25 of them over every anchor.  The operations are
// element-wise, so decoding only the gathered rows gives the same bits (same fp32 operations in the same order, no FMA).
__global__ void __launch_bounds__(256) rpn_decode_selected_kernel(const float* __restrict__ objectness, const float4* __restrict__ deltas,
                                                                  const float4* __restrict__ anchors, const long long* __restrict__ idx,
                                                                  int B, int A, int M, float inv_wx, float inv_wy, float inv_ww, float inv_wh,
                                                                  float xform_clip, float img_w, float img_h, float min_size, float score_thresh,
                                                                  float4* __restrict__ boxes, float* __restrict__ scores,
                                                                  unsigned char* __restrict__ valid) {
    pdl_trigger();
    pdl_wait();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B * M) return;
    const int b = q / M;
    const long long a = idx[q];
    const float4 an = anchors[a], d = deltas[static_cast<long>(b) * A + a];
    const float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
    const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
    const float dx = __fmul_rn(d.x, inv_wx), dy = __fmul_rn(d.y, inv_wy);
    const float dw = fminf(__fmul_rn(d.z, inv_ww), xform_clip), dh = fminf(__fmul_rn(d.w, inv_wh), xform_clip);
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
    float x1 = __fsub_rn(pcx, hw), y1 = __fsub_rn(pcy, hh), x2 = __fadd_rn(pcx, hw), y2 = __fadd_rn(pcy, hh);
    x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);          // clip_boxes_to_image
    y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
    const float sc = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-objectness[static_cast<long>(b) * A + a])));
    boxes[q] = make_float4(x1, y1, x2, y2);
    scores[q] = sc;
    valid[q] = (__fsub_rn(x2, x1) >= min_size && __fsub_rn(y2, y1) >= min_size && sc >= score_thresh) ? 1 : 0;
}

}  // namespace

}  // namespace hd

using namespace hd;

extern "C" int hd_rpn_assign_targets(const float* anchors, const float* gt, const uint8_t* gt_present, int batch, int n_anchors, int n_gt,
                                     float low_threshold, float high_threshold, int allow_low_quality, const float* coder_weights,
                                     void* highest_ws, float* labels, float* regression_targets, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(anchors != nullptr && gt != nullptr && gt_present != nullptr && coder_weights != nullptr && highest_ws != nullptr);
    HD_CHECK_ARG(labels != nullptr && regression_targets != nullptr && batch > 0 && n_anchors > 0 && n_gt >= 1 && n_gt <= kMaxGt);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(anchors) & 15) == 0 && (reinterpret_cast<uintptr_t>(gt) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(regression_targets) & 15) == 0);
    HD_CUDA_OK(cudaMemsetAsync(highest_ws, 0, static_cast<size_t>(batch) * n_gt * sizeof(unsigned int), stream));
    const dim3 grid((n_anchors + 255) / 256, batch);
    HD_CUDA_OK(hd::launch(rpn_assign_kernel<1>, grid, dim3(256), 0, stream, reinterpret_cast<const float4*>(anchors),
                          reinterpret_cast<const float4*>(gt), gt_present, batch, n_anchors, n_gt, low_threshold, high_threshold,
                          allow_low_quality, coder_weights[0], coder_weights[1], coder_weights[2], coder_weights[3],
                          static_cast<unsigned int*>(highest_ws), labels, reinterpret_cast<float4*>(regression_targets)));
    HD_CUDA_OK(hd::launch(rpn_assign_kernel<2>, grid, dim3(256), 0, stream, reinterpret_cast<const float4*>(anchors),
                          reinterpret_cast<const float4*>(gt), gt_present, batch, n_anchors, n_gt, low_threshold, high_threshold,
                          allow_low_quality, coder_weights[0], coder_weights[1], coder_weights[2], coder_weights[3],
                          static_cast<unsigned int*>(highest_ws), labels, reinterpret_cast<float4*>(regression_targets)));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_rpn_decode_selected(const float* objectness, const float* deltas, const float* anchors, const int64_t* idx, int batch,
                                      int anchors_per_image, int selected, const float* coder_weights, float xform_clip, float img_w,
                                      float img_h, float min_size, float score_thresh, float* boxes, float* scores, uint8_t* valid,
                                      hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(objectness != nullptr && deltas != nullptr && anchors != nullptr && idx != nullptr && coder_weights != nullptr);
    HD_CHECK_ARG(boxes != nullptr && scores != nullptr && valid != nullptr && batch > 0 && anchors_per_image > 0 && selected > 0);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(deltas) & 15) == 0 && (reinterpret_cast<uintptr_t>(anchors) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
    const long total = static_cast<long>(batch) * selected;
    // ATen divides a tensor by a host scalar as a multiplication by its fp32 reciprocal
    HD_CUDA_OK(hd::launch(rpn_decode_selected_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, objectness,
                          reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(anchors),
                          reinterpret_cast<const long long*>(idx), batch, anchors_per_image, selected, 1.0f / coder_weights[0],
                          1.0f / coder_weights[1], 1.0f / coder_weights[2], 1.0f / coder_weights[3], xform_clip, img_w, img_h, min_size,
                          score_thresh, reinterpret_cast<float4*>(boxes), scores, valid));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

// direction 0: preds -> (objectness, deltas); 1: (objectness, deltas) -> preds (every channel written, padding = 0)
extern "C" int hd_rpn_concat_preds(void* const* preds, const int* hw, const int* channel_pitch, int levels, int batch, int anchors_per_pixel,
                                   float* objectness, float* deltas, int direction, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(preds != nullptr && hw != nullptr && channel_pitch != nullptr && levels >= 1 && levels <= kMaxPredLevels);
    HD_CHECK_ARG(batch > 0 && anchors_per_pixel >= 1 && objectness != nullptr && deltas != nullptr && (direction == 0 || direction == 1));
    PredLevels L;
    memset(&L, 0, sizeof(L));
    int px = 0;
    for (int l = 0; l < levels; ++l) {
        HD_CHECK_ARG(preds[l] != nullptr && hw[l] > 0 && channel_pitch[l] >= 5 * anchors_per_pixel);
        L.pred[l] = static_cast<float*>(preds[l]); L.hw[l] = hw[l]; L.cp[l] = channel_pitch[l]; L.first_pixel[l] = px;
        px += hw[l];
    }
    L.levels = levels; L.a = anchors_per_pixel; L.pixels = px;
    const long total = static_cast<long>(batch) * px;
    const dim3 grid(static_cast<unsigned>((total + 255) / 256));
    if (direction == 0) HD_CUDA_OK(hd::launch(rpn_concat_kernel<false>, grid, dim3(256), 0, stream, L, batch, objectness, deltas));
    else HD_CUDA_OK(hd::launch(rpn_concat_kernel<true>, grid, dim3(256), 0, stream, L, batch, objectness, deltas));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

// See include/hallucidet_b200.h.
extern "C" int hd_roi_match_labels(const float* props, const int64_t* n_props, const float* gt, const uint8_t* gt_present,
                                   const int64_t* gt_labels, int batch, int slots, int n_gt, float low_threshold, float high_threshold,
                                   int64_t* labels, int64_t* matched, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(props != nullptr && n_props != nullptr && gt != nullptr && gt_present != nullptr && gt_labels != nullptr);
    HD_CHECK_ARG(labels != nullptr && matched != nullptr && batch > 0 && slots >= 0 && n_gt >= 1 && n_gt <= kMaxGt);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(props) & 15) == 0 && (reinterpret_cast<uintptr_t>(gt) & 15) == 0);
    const int n = slots + n_gt;
    HD_CUDA_OK(hd::launch(roi_match_labels_kernel, dim3((n + 255) / 256, batch), dim3(256), 0, stream, reinterpret_cast<const float4*>(props),
                          reinterpret_cast<const long long*>(n_props), reinterpret_cast<const float4*>(gt), gt_present,
                          reinterpret_cast<const long long*>(gt_labels), batch, slots, n_gt, low_threshold, high_threshold,
                          reinterpret_cast<long long*>(labels), reinterpret_cast<long long*>(matched)));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_roi_gather_samples(const hd_roi_gather_args* a, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(a != nullptr && a->flat != nullptr && a->counts != nullptr && a->props != nullptr && a->gt != nullptr);
    HD_CHECK_ARG(a->labels != nullptr && a->matched != nullptr && a->batch > 0 && a->slots >= 0 && a->n_gt >= 1 && a->rows > 0);
    HD_CHECK_ARG(a->out_props != nullptr && a->out_labels != nullptr && a->out_matched != nullptr && a->out_image != nullptr);
    HD_CHECK_ARG(a->out_targets != nullptr && a->out_rois != nullptr && a->out_levels != nullptr && a->out_n_drawn != nullptr);
    HD_CHECK_ARG(a->out_per_image != nullptr && a->rows >= a->batch);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(a->props) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->gt) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->out_props) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out_targets) & 15) == 0);
    GatherParams P;
    P.flat = reinterpret_cast<const long long*>(a->flat); P.counts = a->counts;
    P.props = reinterpret_cast<const float4*>(a->props); P.gt = reinterpret_cast<const float4*>(a->gt);
    P.labels = reinterpret_cast<const long long*>(a->labels); P.matched = reinterpret_cast<const long long*>(a->matched);
    P.B = a->batch; P.T = a->slots; P.G = a->n_gt; P.S = a->rows;
    P.wx = a->weights[0]; P.wy = a->weights[1]; P.ww = a->weights[2]; P.wh = a->weights[3];
    P.inv_s0 = 1.0f / a->canonical_scale;                 // ATen divides by a host scalar as a multiplication by its reciprocal
    P.lvl0 = a->canonical_level; P.eps = a->eps; P.k_min = a->k_min; P.k_max = a->k_max;
    P.out_props = reinterpret_cast<float4*>(a->out_props); P.out_labels = reinterpret_cast<long long*>(a->out_labels);
    P.out_matched = reinterpret_cast<long long*>(a->out_matched); P.out_image = reinterpret_cast<long long*>(a->out_image);
    P.out_targets = reinterpret_cast<float4*>(a->out_targets); P.out_rois = a->out_rois;
    P.out_levels = reinterpret_cast<long long*>(a->out_levels); P.out_n_drawn = reinterpret_cast<long long*>(a->out_n_drawn);
    P.out_per_image = reinterpret_cast<long long*>(a->out_per_image);
    HD_CUDA_OK(hd::launch(roi_gather_samples_kernel, dim3((a->rows + 255) / 256), dim3(256), 0, stream, P));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}
