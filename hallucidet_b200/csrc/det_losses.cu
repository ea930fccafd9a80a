// det_losses.cu -- the two loss pairs of the Faster R-CNN training tail, value and gradient in one launch each.
//
// The reference's detector is torchvision's Faster R-CNN (src/utils/eval_forward_fasterrcnn.py:62-99, :112-140):
//   RegionProposalNetwork.compute_loss (TV models/detection/rpn.py): binary cross-entropy with logits over the sampled anchors
//       (mean) + smooth-L1 (beta 1/9, summed) over the sampled positive anchors / number of sampled anchors;
//   fastrcnn_loss (TV models/detection/roi_heads.py): cross-entropy over the sampled proposals (mean) + smooth-L1 (beta 1/9,
//       summed) of the matched class's box deltas over the foreground proposals / number of sampled proposals.
// As PyTorch operators each pair is ~25 small launches forward and ~35 backward, all on the critical path of the step.  Here
// one CTA computes both losses of a pair AND their gradients w.r.t. the network outputs (the row lists are <= a few thousand
// rows), summing in a fixed order (deterministic); autograd then only scales the stored gradients by the incoming
// gradient of each loss (hallucidet_b200/detection.py: _FastRCNNLoss, _RPNLoss).  Per-row arithmetic as in PyTorch's
// kernels (log-sum-exp with the row maximum subtracted; max(x, 0) - x * t + log1p(exp(-|x|)) for the logits BCE).
#include <math.h>

#include "hd_common.cuh"

namespace hd {

namespace {

constexpr int kLossThreads = 1024;

__device__ __forceinline__ float block_sum(float v, float* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                      // s_red may still be read from the previous call
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) t += s_red[w];   // every thread adds the warp sums in the same order
    return t;
}

__device__ __forceinline__ void smooth_l1(float d, float beta, float& loss, float& grad) {
    const float a = fabsf(d);
    if (a < beta) { loss = 0.5f * d * d / beta; grad = d / beta; }
    else { loss = a - 0.5f * beta; grad = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
}

// rows: labels[s] >= 0 sampled proposal of class labels[s] (0 = background), -100 = padding row
__global__ void __launch_bounds__(kLossThreads) fastrcnn_loss_kernel(const float* __restrict__ logits, const float* __restrict__ box,
                                                                    const long long* __restrict__ labels, const float4* __restrict__ targets,
                                                                    int S, int C, float beta, float* __restrict__ losses,
                                                                    float* __restrict__ g_logits, float* __restrict__ g_box) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_red[kLossThreads / 32];
    float cnt = 0.f;
    for (int s = threadIdx.x; s < S; s += kLossThreads) cnt += labels[s] != -100 ? 1.f : 0.f;
    const float n = block_sum(cnt, s_red);
    const float inv_n = n > 0.f ? 1.f / n : 0.f;
    float ce = 0.f, bl = 0.f;
    for (int s = threadIdx.x; s < S; s += kLossThreads) {
        const long long lab = labels[s];
        const float* x = logits + static_cast<long>(s) * C;
        float* gx = g_logits + static_cast<long>(s) * C;
        float* gb = g_box + static_cast<long>(s) * 4 * C;
        for (int j = 0; j < 4 * C; ++j) gb[j] = 0.f;
        if (lab == -100) {
            for (int c = 0; c < C; ++c) gx[c] = 0.f;
            continue;
        }
        float m = x[0];
        for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
        float z = 0.f;
        for (int c = 0; c < C; ++c) z += expf(x[c] - m);
        const float lse = m + logf(z);
        ce += lse - x[lab];
        for (int c = 0; c < C; ++c) gx[c] = (expf(x[c] - lse) - (c == lab ? 1.f : 0.f)) * inv_n;
        if (lab > 0) {
            const float4 t = targets[s];
            const float tv[4] = {t.x, t.y, t.z, t.w};
            const float* p = box + (static_cast<long>(s) * C + lab) * 4;
            for (int j = 0; j < 4; ++j) {
                float l, g;
                smooth_l1(p[j] - tv[j], beta, l, g);
                bl += l;
                gb[lab * 4 + j] = g * inv_n;
            }
        }
    }
    const float ce_sum = block_sum(ce, s_red);
    const float bl_sum = block_sum(bl, s_red);
    if (threadIdx.x == 0) { losses[0] = ce_sum * inv_n; losses[1] = bl_sum * inv_n; }
}

// flat[r]: position (image * A + anchor) of the r-th sampled anchor; code[flat[r]] = 1 positive / 2 negative; rows r >= n_drawn are padding.
// g_obj / g_deltas: dense gradients, zeroed by the caller.
__global__ void __launch_bounds__(kLossThreads) rpn_loss_kernel(const float* __restrict__ objectness, const float4* __restrict__ deltas,
                                                               const float* __restrict__ labels, const float4* __restrict__ targets,
                                                               const long long* __restrict__ flat, const unsigned char* __restrict__ code,
                                                               const int* __restrict__ counts, int B, int R, float beta,
                                                               float* __restrict__ losses, float* __restrict__ g_obj, float4* __restrict__ g_deltas) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_red[kLossThreads / 32];
    long long n_drawn = 0;
    for (int b = 0; b < B; ++b) n_drawn += counts[b * 4 + 2] + counts[b * 4 + 3];
    const float inv_n = n_drawn > 0 ? 1.f / static_cast<float>(n_drawn) : 0.f;
    float ol = 0.f, bl = 0.f;
    for (int r = threadIdx.x; r < R; r += kLossThreads) {
        if (r >= n_drawn) continue;
        const long long f = flat[r];
        const float x = objectness[f], t = labels[f];
        // binary_cross_entropy_with_logits: (1 - t) * x + max(-x, 0) + log1p(exp(-|x|))
        ol += (1.f - t) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
        g_obj[f] = (1.f / (1.f + expf(-x)) - t) * inv_n;
        if (code[f] == 1) {
            const float4 p = deltas[f], q = targets[f];
            const float pv[4] = {p.x, p.y, p.z, p.w}, qv[4] = {q.x, q.y, q.z, q.w};
            float gv[4];
            for (int j = 0; j < 4; ++j) {
                float l;
                smooth_l1(pv[j] - qv[j], beta, l, gv[j]);
                bl += l;
                gv[j] *= inv_n;
            }
            g_deltas[f] = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
    }
    const float ol_sum = block_sum(ol, s_red);
    const float bl_sum = block_sum(bl, s_red);
    if (threadIdx.x == 0) { losses[0] = ol_sum * inv_n; losses[1] = bl_sum * inv_n; }
}

}  // namespace

}  // namespace hd

using namespace hd;

// See include/hallucidet_b200.h.
extern "C" int hd_fastrcnn_loss(const float* class_logits, const float* box_regression, const int64_t* labels, const float* regression_targets,
                                int rows, int num_classes, float beta, float* losses, float* grad_logits, float* grad_box, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(class_logits != nullptr && box_regression != nullptr && labels != nullptr && regression_targets != nullptr);
    HD_CHECK_ARG(losses != nullptr && grad_logits != nullptr && grad_box != nullptr && rows > 0 && num_classes >= 2 && beta > 0.f);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(regression_targets) & 15) == 0);
    HD_CUDA_OK(hd::launch(fastrcnn_loss_kernel, dim3(1), dim3(kLossThreads), 0, stream, class_logits, box_regression,
                          reinterpret_cast<const long long*>(labels), reinterpret_cast<const float4*>(regression_targets), rows, num_classes,
                          beta, losses, grad_logits, grad_box));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_rpn_loss(const float* objectness, const float* pred_bbox_deltas, const float* labels, const float* regression_targets,
                           const int64_t* flat, const uint8_t* sampled, const int32_t* counts, int batch, int rows, float beta, float* losses,
                           float* grad_objectness, float* grad_deltas, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(objectness != nullptr && pred_bbox_deltas != nullptr && labels != nullptr && regression_targets != nullptr);
    HD_CHECK_ARG(flat != nullptr && sampled != nullptr && counts != nullptr && losses != nullptr && grad_objectness != nullptr);
    HD_CHECK_ARG(grad_deltas != nullptr && batch > 0 && rows > 0 && beta > 0.f);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(pred_bbox_deltas) & 15) == 0 && (reinterpret_cast<uintptr_t>(regression_targets) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(grad_deltas) & 15) == 0);
    HD_CUDA_OK(hd::launch(rpn_loss_kernel, dim3(1), dim3(kLossThreads), 0, stream, objectness, reinterpret_cast<const float4*>(pred_bbox_deltas),
                          labels, reinterpret_cast<const float4*>(regression_targets), reinterpret_cast<const long long*>(flat), sampled, counts,
                          batch, rows, beta, losses, grad_objectness, reinterpret_cast<float4*>(grad_deltas)));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}
