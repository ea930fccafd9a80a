// roi_align.cu -- backward of RoIAlign (the box head's pooling over the FPN levels), the largest single kernel of the
// detection tail's backward pass.
//
// The reference's detector is torchvision's Faster R-CNN: its RoI heads pool 7x7 features with
// torchvision.ops.roi_align (sampling_ratio 2, aligned = False) -- a third-party op whose published CUDA backward
// (torchvision csrc/ops/cuda/roi_align_kernel.cu: roi_align_backward_kernel_impl) gives every (roi, channel, bin) element a
// thread that issues 16 scattered 4-byte atomics into the NCHW gradient: 822 M single-sector atomics, 1.75 ms per train
// step at config 2.  The bilinear sample positions and weights do not depend on the channel, so here the gradient is
// accumulated channels-last: one CTA per RoI stages its [C][PH*PW] output gradient in shared memory, computes the 196
// sample descriptors once, and every (sample, corner) contribution becomes one fully coalesced vector reduction
// (red.global.add.v4.f32, 64 lanes = the 256 channels of one pixel = 1 KB contiguous): 8x fewer L2 sector operations.
// hd_nhwc_to_nchw_f32 then hands the result to autograd in torchvision's layout.  Same arithmetic per contribution
// (grad * w / count); only the fp32 summation order differs (it is not deterministic in torchvision either).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hd_common.cuh"

namespace hd {

namespace {

constexpr int kRoiThreads = 256;
constexpr int kMaxLevels = 8;

// feature-pyramid levels a RoI can be pooled from (level_of_roi selects one; a single-level call passes level_of_roi = NULL)
struct RoiLevels {
    const void* feat[kMaxLevels];       // forward: channels-last feature maps [N][H][W][C], fp32 or bf16
    float* grad[kMaxLevels];            // backward: channels-last gradient scratch (zeroed by the caller)
    int h[kMaxLevels], w[kMaxLevels];
    float scale[kMaxLevels];
};
constexpr int kMaxBins = 7 * 7;
constexpr int kMaxSamples = kMaxBins * 4;       // sampling_ratio <= 2
constexpr int kMaxC = 256;

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// grad_in: [N][H][W][C] fp32 (channels last)
__global__ void __launch_bounds__(kRoiThreads) roi_align_bwd_nhwc_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois,
                                                                         const long long* __restrict__ level_of_roi, const RoiLevels L,
                                                                         int C, int PH, int PW, int sampling_ratio) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float gsm[];        // [nbins][C + 4]: one float4 per thread and item, conflict free
    __shared__ float s_w[kMaxSamples * 4];
    __shared__ int s_off[kMaxSamples * 4];              // pixel index y * W + x, or -1
    const int k = blockIdx.x;
    const float* roi = rois + static_cast<long>(k) * 5;
    const int n = static_cast<int>(roi[0]);
    const int lvl = level_of_roi != nullptr ? static_cast<int>(level_of_roi[k]) : 0;
    const int H = L.h[lvl], W = L.w[lvl];
    const float spatial_scale = L.scale[lvl];
    float* __restrict__ grad_in = L.grad[lvl];
    // torchvision, aligned = false
    const float roi_start_w = roi[1] * spatial_scale, roi_start_h = roi[2] * spatial_scale;
    const float roi_end_w = roi[3] * spatial_scale, roi_end_h = roi[4] * spatial_scale;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f), roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
    const float bin_size_h = roi_height / static_cast<float>(PH), bin_size_w = roi_width / static_cast<float>(PW);
    const int grid_h = sampling_ratio, grid_w = sampling_ratio;
    const float inv_count = 1.f / static_cast<float>(grid_h * grid_w);   // count = 1 or 4: (g * w) / count == (g * w) * inv_count exactly
    const int nbins = PH * PW, per_bin = grid_h * grid_w, nsamp = nbins * per_bin, pitch = C + 4;

    const float* gout = grad_out + static_cast<long>(k) * C * nbins;
    for (int i = threadIdx.x; i < C * nbins; i += kRoiThreads) {
        const int c = i / nbins, b = i - c * nbins;
        gsm[b * pitch + c] = gout[i];
    }
    for (int s = threadIdx.x; s < nsamp; s += kRoiThreads) {
        const int bin = s / per_bin, r = s - bin * per_bin;
        const int ph = bin / PW, pw = bin - ph * PW, iy = r / grid_w, ix = r - iy * grid_w;
        float y = roi_start_h + ph * bin_size_h + (iy + .5f) * bin_size_h / static_cast<float>(grid_h);
        float x = roi_start_w + pw * bin_size_w + (ix + .5f) * bin_size_w / static_cast<float>(grid_w);
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        int off[4] = {-1, -1, -1, -1};
        if (!(y < -1.0f || y > H || x < -1.0f || x > W)) {
            if (y <= 0) y = 0;
            if (x <= 0) x = 0;
            int y_low = static_cast<int>(y), x_low = static_cast<int>(x), y_high, x_high;
            if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else y_high = y_low + 1;
            if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else x_high = x_low + 1;
            const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
            w[0] = hy * hx; w[1] = hy * lx; w[2] = ly * hx; w[3] = ly * lx;
            off[0] = y_low * W + x_low; off[1] = y_low * W + x_high; off[2] = y_high * W + x_low; off[3] = y_high * W + x_high;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) { s_w[s * 4 + c4] = w[c4]; s_off[s * 4 + c4] = off[c4]; }
    }
    __syncthreads();
    // items = (sample, corner); a group of C/4 threads covers the channels of one item, kRoiThreads/(C/4) items at a time
    const int lanes = C / 4, groups = kRoiThreads / lanes;
    const int grp = threadIdx.x / lanes, c0 = (threadIdx.x - grp * lanes) * 4;
    float* gin = grad_in + static_cast<long>(n) * H * W * C + c0;
    if (grp < groups) {
        for (int it = grp; it < nsamp * 4; it += groups) {
            const int off = s_off[it];
            if (off < 0) continue;
            const float wgt = s_w[it];
            const int bin = (it >> 2) / per_bin;
            const float4 g = *reinterpret_cast<const float4*>(gsm + bin * pitch + c0);
            red_add_v4(gin + static_cast<long>(off) * C, g.x * wgt * inv_count, g.y * wgt * inv_count, g.z * wgt * inv_count,
                       g.w * wgt * inv_count);
        }
    }
}

// Backward, separable form.  The bilinear weight of sample (sy, sx) on pixel (y, x) factors into wy(sy, y) * wx(sx, x), the
// validity test factors the same way, and the gradient value depends on the BIN only -- so the sum over the 4 x 49 (sample,
// corner) contributions landing on one pixel is
//     (1 / count) * sum_ph sum_pw WY[ph][y] * WX[pw][x] * g[ph][pw],      WY[ph][y] = sum of wy over the samples of bin row ph.
// Every pixel of the RoI's footprint (F_y x F_x, about 16 x 16 at the level the RoI is assigned to) is therefore written by ONE
// vector reduction instead of the ~3 the per-sample scatter issues on average (784 items on ~256 pixels): the kernel is bound
// by L2 atomic throughput, so that is its cost.  Footprints wider than kSepMax pixels fall back to the scatter loop.
constexpr int kSepMax = 32;

__global__ void __launch_bounds__(kRoiThreads) roi_align_bwd_sep_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois,
                                                                        const long long* __restrict__ level_of_roi, const RoiLevels L,
                                                                        int C, int PH, int PW, int sampling_ratio) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float gsm[];        // [nbins][C + 4]
    __shared__ float s_wy[7 * kSepMax], s_wx[7 * kSepMax];
    __shared__ int s_lo[2 * 14], s_hi[2 * 14];          // per axis sample: low / high pixel (or -1)
    __shared__ float s_wl[2 * 14], s_wh[2 * 14];        // weights on the low / high pixel
    __shared__ int s_geo[4];                            // y0, F_y, x0, F_x
    __shared__ float s_cw[2 * kSepMax * 7];             // compacted weights: [axis][pixel][k]
    __shared__ int s_cb[2 * kSepMax * 7], s_cn[2 * kSepMax];
    __shared__ float s_w[kMaxSamples * 4];
    __shared__ int s_off[kMaxSamples * 4];
    const int k = blockIdx.x;
    const float* roi = rois + static_cast<long>(k) * 5;
    const int n = static_cast<int>(roi[0]);
    const int lvl = level_of_roi != nullptr ? static_cast<int>(level_of_roi[k]) : 0;
    const int H = L.h[lvl], W = L.w[lvl];
    const float spatial_scale = L.scale[lvl];
    float* __restrict__ grad_in = L.grad[lvl];
    const float roi_start_w = roi[1] * spatial_scale, roi_start_h = roi[2] * spatial_scale;
    const float roi_end_w = roi[3] * spatial_scale, roi_end_h = roi[4] * spatial_scale;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f), roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
    const float bin_size_h = roi_height / static_cast<float>(PH), bin_size_w = roi_width / static_cast<float>(PW);
    const int grid = sampling_ratio;                    // grid_h == grid_w
    const float inv_count = 1.f / static_cast<float>(grid * grid);
    const int nbins = PH * PW, pitch = C + 4;
    const int ny = PH * grid, nx = PW * grid;           // samples per axis (<= 14)

    const float* gout = grad_out + static_cast<long>(k) * C * nbins;
    for (int i = threadIdx.x; i < C * nbins; i += kRoiThreads) {
        const int c = i / nbins, b = i - c * nbins;
        gsm[b * pitch + c] = gout[i];
    }
    if (threadIdx.x < ny + nx) {
        // one axis sample: the per-axis half of torchvision's bilinear_interpolate_gradient
        const bool is_y = static_cast<int>(threadIdx.x) < ny;
        const int s = is_y ? threadIdx.x : threadIdx.x - ny;
        const int bin = s / grid, i = s - bin * grid;
        const float start = is_y ? roi_start_h : roi_start_w, bsz = is_y ? bin_size_h : bin_size_w;
        const int size = is_y ? H : W;
        float t = start + bin * bsz + (i + .5f) * bsz / static_cast<float>(grid);
        int lo = -1, hi = -1;
        float wl = 0.f, wh = 0.f;
        if (!(t < -1.0f || t > size)) {
            if (t <= 0) t = 0;
            lo = static_cast<int>(t);
            if (lo >= size - 1) { hi = lo = size - 1; t = static_cast<float>(lo); } else hi = lo + 1;
            const float l = t - lo;
            wl = 1.f - l; wh = l;
        }
        const int o = is_y ? s : 14 + s;
        s_lo[o] = lo; s_hi[o] = hi; s_wl[o] = wl; s_wh[o] = wh;
    }
    for (int i = threadIdx.x; i < 7 * kSepMax; i += kRoiThreads) { s_wy[i] = 0.f; s_wx[i] = 0.f; }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int ax = threadIdx.x, cnt = ax == 0 ? ny : nx;
        int mn = 1 << 30, mx = -1;
        for (int s = 0; s < cnt; ++s) {
            const int lo = s_lo[ax * 14 + s], hi = s_hi[ax * 14 + s];
            if (lo >= 0) { mn = min(mn, lo); mx = max(mx, hi); }
        }
        const int F = mx >= 0 ? mx - mn + 1 : 0;
        s_geo[ax * 2] = mn; s_geo[ax * 2 + 1] = F;
        if (F > 0 && F <= kSepMax) {
            float* wtab = ax == 0 ? s_wy : s_wx;
            for (int s = 0; s < cnt; ++s) {
                const int lo = s_lo[ax * 14 + s], hi = s_hi[ax * 14 + s];
                if (lo < 0) continue;
                const int bin = s / grid;
                wtab[bin * kSepMax + (lo - mn)] += s_wl[ax * 14 + s];
                wtab[bin * kSepMax + (hi - mn)] += s_wh[ax * 14 + s];
            }
        }
    }
    __syncthreads();
    const int y0 = s_geo[0], Fy = s_geo[1], x0 = s_geo[2], Fx = s_geo[3];
    if (Fy == 0 || Fx == 0) return;                      // every sample outside the map
    const int lanes = C / 4, groups = kRoiThreads / lanes;
    const int grp = threadIdx.x / lanes, c0 = (threadIdx.x - grp * lanes) * 4;
    float* gin = grad_in + static_cast<long>(n) * H * W * C + c0;
    if (Fy <= kSepMax && Fx <= kSepMax) {
        // compact the weight tables: per footprint row / column the bins that reach it (usually 1-2, all 7 for a tiny RoI)
        if (threadIdx.x < 2 * kSepMax) {
            const int ax = threadIdx.x / kSepMax, f = threadIdx.x - ax * kSepMax;
            const float* wtab = ax == 0 ? s_wy : s_wx;
            const int nb = ax == 0 ? PH : PW;
            int cnt = 0;
            if (f < (ax == 0 ? Fy : Fx))
                for (int b = 0; b < nb; ++b) {
                    const float w = wtab[b * kSepMax + f];
                    if (w != 0.f) { s_cw[(ax * kSepMax + f) * 7 + cnt] = w; s_cb[(ax * kSepMax + f) * 7 + cnt] = b; ++cnt; }
                }
            s_cn[ax * kSepMax + f] = cnt;
        }
        __syncthreads();
        if ((C & 7) == 0) {
            // eight channels per thread (sixteen: 0.48 ms, bank conflicts on the strided 16-byte loads; four: 0.51 ms): the index / weight bookkeeping of a pixel (most of the instructions: ncu shows the
            // kernel issue-bound at 62 % of the SM's issue slots, not atomics-bound) is shared by two vector reductions
            constexpr int V = 2;
            const int lanesV = C / (4 * V), groupsV = kRoiThreads / lanesV;
            const int gV = threadIdx.x / lanesV, cV = (threadIdx.x - gV * lanesV) * 4 * V;
            if (gV >= groupsV) return;
            float* ginV = grad_in + static_cast<long>(n) * H * W * C + cV;
            for (int q = gV; q < Fy * Fx; q += groupsV) {
                const int py = q / Fx, px = q - py * Fx;
                const int cy = s_cn[py], cx = s_cn[kSepMax + px];
                if (cy == 0 || cx == 0) continue;
                float4 acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = 0; i < cy; ++i) {
                    const float wy = s_cw[py * 7 + i];
                    const float* grow = gsm + s_cb[py * 7 + i] * PW * pitch + cV;
                    for (int j = 0; j < cx; ++j) {
                        const float w = wy * s_cw[(kSepMax + px) * 7 + j];
                        const float* gp = grow + s_cb[(kSepMax + px) * 7 + j] * pitch;
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            const float4 u = *reinterpret_cast<const float4*>(gp + 4 * v);
                            acc[v].x += w * u.x; acc[v].y += w * u.y; acc[v].z += w * u.z; acc[v].w += w * u.w;
                        }
                    }
                }
                float* dst = ginV + (static_cast<long>(y0 + py) * W + (x0 + px)) * C;
#pragma unroll
                for (int v = 0; v < V; ++v)
                    red_add_v4(dst + 4 * v, acc[v].x * inv_count, acc[v].y * inv_count, acc[v].z * inv_count, acc[v].w * inv_count);
            }
            return;
        }
        if (grp >= groups) return;
        for (int q = grp; q < Fy * Fx; q += groups) {
            const int py = q / Fx, px = q - py * Fx;
            const int cy = s_cn[py], cx = s_cn[kSepMax + px];
            if (cy == 0 || cx == 0) continue;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < cy; ++i) {
                const float wy = s_cw[py * 7 + i];
                const float* grow = gsm + s_cb[py * 7 + i] * PW * pitch + c0;
                for (int j = 0; j < cx; ++j) {
                    const float w = wy * s_cw[(kSepMax + px) * 7 + j];
                    const float4 g = *reinterpret_cast<const float4*>(grow + s_cb[(kSepMax + px) * 7 + j] * pitch);
                    acc.x += w * g.x; acc.y += w * g.y; acc.z += w * g.z; acc.w += w * g.w;
                }
            }
            red_add_v4(gin + (static_cast<long>(y0 + py) * W + (x0 + px)) * C, acc.x * inv_count, acc.y * inv_count, acc.z * inv_count,
                       acc.w * inv_count);
        }
        return;
    }
    // wide footprint: the per-sample scatter of roi_align_bwd_nhwc_kernel
    const int per_bin = grid * grid, nsamp = nbins * per_bin;
    for (int s = threadIdx.x; s < nsamp; s += kRoiThreads) {
        const int bin = s / per_bin, r = s - bin * per_bin;
        const int ph = bin / PW, pw = bin - ph * PW, iy = r / grid, ix = r - iy * grid;
        const int sy = ph * grid + iy, sx = 14 + pw * grid + ix;
        const bool ok = s_lo[sy] >= 0 && s_lo[sx] >= 0;
        const float wyl = s_wl[sy], wyh = s_wh[sy], wxl = s_wl[sx], wxh = s_wh[sx];
        s_w[s * 4 + 0] = wyl * wxl; s_w[s * 4 + 1] = wyl * wxh; s_w[s * 4 + 2] = wyh * wxl; s_w[s * 4 + 3] = wyh * wxh;
        s_off[s * 4 + 0] = ok ? s_lo[sy] * W + s_lo[sx] : -1; s_off[s * 4 + 1] = ok ? s_lo[sy] * W + s_hi[sx] : -1;
        s_off[s * 4 + 2] = ok ? s_hi[sy] * W + s_lo[sx] : -1; s_off[s * 4 + 3] = ok ? s_hi[sy] * W + s_hi[sx] : -1;
    }
    __syncthreads();
    if (grp < groups) {
        for (int it = grp; it < nsamp * 4; it += groups) {
            const int off = s_off[it];
            if (off < 0) continue;
            const float wgt = s_w[it];
            const int bin = (it >> 2) / per_bin;
            const float4 g = *reinterpret_cast<const float4*>(gsm + bin * pitch + c0);
            red_add_v4(gin + static_cast<long>(off) * C, g.x * wgt * inv_count, g.y * wgt * inv_count, g.z * wgt * inv_count,
                       g.w * wgt * inv_count);
        }
    }
}

// Forward on a channels-last copy of the feature map: every (sample, corner) read is 64 lanes x 16 bytes = the 256
// channels of one pixel (torchvision gathers 16 scattered 4-byte values per output element from NCHW).  Same expression
// per output element as torchvision's roi_align_forward_kernel_impl: val += w1*v1 + w2*v2 + w3*v3 + w4*v4 over the samples
// in (iy, ix) order, then val /= count.  The [C][PH*PW] result of a RoI is staged in shared memory and written coalesced.
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);                     // 4 channels: exact widening to fp32
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xffff0000u));
}

template <typename T>
__global__ void __launch_bounds__(kRoiThreads) roi_align_fwd_nhwc_kernel(const float* __restrict__ rois, const long long* __restrict__ level_of_roi,
                                                                         const RoiLevels L, float* __restrict__ out, int C, int PH, int PW,
                                                                         int sampling_ratio) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float gsm[];        // [C][nbins] output tile of this RoI
    __shared__ float s_w[kMaxSamples * 4];
    __shared__ int s_off[kMaxSamples * 4];
    const int k = blockIdx.x;
    const float* roi = rois + static_cast<long>(k) * 5;
    const int n = static_cast<int>(roi[0]);
    const int lvl = level_of_roi != nullptr ? static_cast<int>(level_of_roi[k]) : 0;
    const int H = L.h[lvl], W = L.w[lvl];
    const float spatial_scale = L.scale[lvl];
    const T* __restrict__ feat = static_cast<const T*>(L.feat[lvl]);
    const float roi_start_w = roi[1] * spatial_scale, roi_start_h = roi[2] * spatial_scale;
    const float roi_end_w = roi[3] * spatial_scale, roi_end_h = roi[4] * spatial_scale;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f), roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
    const float bin_size_h = roi_height / static_cast<float>(PH), bin_size_w = roi_width / static_cast<float>(PW);
    const int grid_h = sampling_ratio, grid_w = sampling_ratio;
    const float count = fmaxf(static_cast<float>(grid_h * grid_w), 1.f);
    const int nbins = PH * PW, per_bin = grid_h * grid_w, nsamp = nbins * per_bin;
    for (int s = threadIdx.x; s < nsamp; s += kRoiThreads) {
        const int bin = s / per_bin, r = s - bin * per_bin;
        const int ph = bin / PW, pw = bin - ph * PW, iy = r / grid_w, ix = r - iy * grid_w;
        float y = roi_start_h + ph * bin_size_h + (iy + .5f) * bin_size_h / static_cast<float>(grid_h);
        float x = roi_start_w + pw * bin_size_w + (ix + .5f) * bin_size_w / static_cast<float>(grid_w);
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        int off[4] = {-1, -1, -1, -1};
        if (!(y < -1.0f || y > H || x < -1.0f || x > W)) {     // torchvision bilinear_interpolate: out of range -> 0
            if (y <= 0) y = 0;
            if (x <= 0) x = 0;
            int y_low = static_cast<int>(y), x_low = static_cast<int>(x), y_high, x_high;
            if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else y_high = y_low + 1;
            if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else x_high = x_low + 1;
            const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
            w[0] = hy * hx; w[1] = hy * lx; w[2] = ly * hx; w[3] = ly * lx;
            off[0] = y_low * W + x_low; off[1] = y_low * W + x_high; off[2] = y_high * W + x_low; off[3] = y_high * W + x_high;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) { s_w[s * 4 + c4] = w[c4]; s_off[s * 4 + c4] = off[c4]; }
    }
    __syncthreads();
    const int lanes = C / 4, groups = kRoiThreads / lanes;
    const int grp = threadIdx.x / lanes, c0 = (threadIdx.x - grp * lanes) * 4;
    const T* fin = feat + static_cast<long>(n) * H * W * C + c0;
    if (grp < groups) {
        for (int bin = grp; bin < nbins; bin += groups) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < per_bin; ++r) {
                const int s4 = (bin * per_bin + r) * 4;
                if (s_off[s4] < 0) continue;                   // sample outside the map contributes 0
                const float w1 = s_w[s4], w2 = s_w[s4 + 1], w3 = s_w[s4 + 2], w4 = s_w[s4 + 3];
                const float4 v1 = load4<T>(fin + static_cast<long>(s_off[s4]) * C);
                const float4 v2 = load4<T>(fin + static_cast<long>(s_off[s4 + 1]) * C);
                const float4 v3 = load4<T>(fin + static_cast<long>(s_off[s4 + 2]) * C);
                const float4 v4 = load4<T>(fin + static_cast<long>(s_off[s4 + 3]) * C);
                acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
                acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
                acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
                acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
            }
            gsm[(c0 + 0) * nbins + bin] = acc.x / count;
            gsm[(c0 + 1) * nbins + bin] = acc.y / count;
            gsm[(c0 + 2) * nbins + bin] = acc.z / count;
            gsm[(c0 + 3) * nbins + bin] = acc.w / count;
        }
    }
    __syncthreads();
    float* o = out + static_cast<long>(k) * C * nbins;
    for (int i = threadIdx.x; i < C * nbins; i += kRoiThreads) o[i] = gsm[i];
}

// [N][C][H][W] fp32 -> [N][H][W][C] fp32 through a 32 x 32 shared-memory tile
__global__ void __launch_bounds__(256) nchw_to_nhwc_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
    pdl_trigger();
    pdl_wait();
    __shared__ float t[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xs = x + static_cast<long>(n) * C * HW;
    float* ys = y + static_cast<long>(n) * HW * C;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int c = c0 + ty + j, p = p0 + tx;
        t[ty + j][tx] = (p < HW && c < C) ? xs[static_cast<long>(c) * HW + p] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int p = p0 + ty + j, c = c0 + tx;
        if (p < HW && c < C) ys[static_cast<long>(p) * C + c] = t[tx][ty + j];
    }
}

// [N][H][W][C] fp32 -> [N][C][H][W] fp32 through a 32 x 32 shared-memory tile
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
    pdl_trigger();
    pdl_wait();
    __shared__ float t[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const float* xs = x + static_cast<long>(n) * HW * C;
    float* ys = y + static_cast<long>(n) * C * HW;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int p = p0 + ty + j, c = c0 + tx;
        t[ty + j][tx] = (p < HW && c < C) ? xs[static_cast<long>(p) * C + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int c = c0 + ty + j, p = p0 + tx;
        if (p < HW && c < C) ys[static_cast<long>(c) * HW + p] = t[tx][ty + j];
    }
}

}  // namespace

}  // namespace hd

using namespace hd;

static int roi_check(int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio) {
    HD_CHECK_ARG(num_rois >= 0 && channels > 0 && channels % 4 == 0 && channels <= kMaxC && kRoiThreads % (channels / 4) == 0);
    HD_CHECK_ARG(pooled_h > 0 && pooled_w > 0 && pooled_h * pooled_w <= kMaxBins && sampling_ratio >= 1 && sampling_ratio <= 2);
    return HD_OK;
}

static int roi_bwd_launch(const float* grad_out, const float* rois, const long long* level_of_roi, const RoiLevels& L, int num_rois,
                          int channels, int pooled_h, int pooled_w, int sampling_ratio, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(pooled_h * pooled_w) * (channels + 4) * sizeof(float);
    static SmemAttrOnce smem_attr;
    static int sep = -1;
    if (sep < 0) {
        const char* e = getenv("HD_ROI_BWD_SEPARABLE");
        sep = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (sep && pooled_h <= 7 && pooled_w <= 7 && pooled_h * sampling_ratio <= 14 && pooled_w * sampling_ratio <= 14) {
        static SmemAttrOnce smem_attr_sep;
        HD_CUDA_OK(ensure_dyn_smem(smem_attr_sep, roi_align_bwd_sep_kernel, static_cast<int>(kMaxBins * (kMaxC + 4) * sizeof(float))));
        HD_CUDA_OK(hd::launch(roi_align_bwd_sep_kernel, dim3(num_rois), dim3(kRoiThreads), smem, stream, grad_out, rois, level_of_roi, L, channels,
                              pooled_h, pooled_w, sampling_ratio));
        HD_CUDA_OK(cudaPeekAtLastError());
        return HD_OK;
    }
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, roi_align_bwd_nhwc_kernel, static_cast<int>(kMaxBins * (kMaxC + 4) * sizeof(float))));
    HD_CUDA_OK(hd::launch(roi_align_bwd_nhwc_kernel, dim3(num_rois), dim3(kRoiThreads), smem, stream, grad_out, rois, level_of_roi, L, channels, pooled_h, pooled_w,
                                                                       sampling_ratio));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

static int roi_fwd_launch(const float* rois, const long long* level_of_roi, const RoiLevels& L, float* out, int num_rois, int channels,
                          int pooled_h, int pooled_w, int sampling_ratio, cudaStream_t stream, bool bf16 = false) {
    const size_t smem = static_cast<size_t>(channels) * pooled_h * pooled_w * sizeof(float);
    static SmemAttrOnce smem_attr, smem_attr16;
    if (bf16) {
        HD_CUDA_OK(ensure_dyn_smem(smem_attr16, roi_align_fwd_nhwc_kernel<__nv_bfloat16>, static_cast<int>(kMaxC * kMaxBins * sizeof(float))));
        HD_CUDA_OK(hd::launch(roi_align_fwd_nhwc_kernel<__nv_bfloat16>, dim3(num_rois), dim3(kRoiThreads), smem, stream, rois, level_of_roi, L, out,
                              channels, pooled_h, pooled_w, sampling_ratio));
        HD_CUDA_OK(cudaPeekAtLastError());
        return HD_OK;
    }
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, roi_align_fwd_nhwc_kernel<float>, static_cast<int>(kMaxC * kMaxBins * sizeof(float))));
    HD_CUDA_OK(hd::launch(roi_align_fwd_nhwc_kernel<float>, dim3(num_rois), dim3(kRoiThreads), smem, stream, rois, level_of_roi, L, out, channels, pooled_h, pooled_w, sampling_ratio));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

// grad_in_nhwc ([N][H][W][C] fp32 channels-last scratch, zeroed by the caller) += d/d(input) of
// torchvision.ops.roi_align(input, rois, spatial_scale, PH, PW, sampling_ratio, aligned = False) for grad_out [K][C][PH][PW];
// rois [K][5] = (batch index, x1, y1, x2, y2).  C % 4 == 0, C <= 256, PH * PW <= 49, sampling_ratio in {1, 2}.
extern "C" int hd_roi_align_bwd_nhwc(const float* grad_out, const float* rois, float* grad_in_nhwc, int num_rois, int channels,
                                     int height, int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio,
                                     hd_stream stream_) {
    if (int e = roi_check(num_rois, channels, pooled_h, pooled_w, sampling_ratio)) return e;
    HD_CHECK_ARG(height > 0 && width > 0);
    if (num_rois == 0) return HD_OK;
    HD_CHECK_ARG(grad_out != nullptr && rois != nullptr && grad_in_nhwc != nullptr && (reinterpret_cast<uintptr_t>(grad_in_nhwc) & 15) == 0);
    RoiLevels L;
    memset(&L, 0, sizeof(L));
    L.grad[0] = grad_in_nhwc; L.h[0] = height; L.w[0] = width; L.scale[0] = spatial_scale;
    return roi_bwd_launch(grad_out, rois, nullptr, L, num_rois, channels, pooled_h, pooled_w, sampling_ratio, static_cast<cudaStream_t>(stream_));
}

// Multi-level variants (torchvision MultiScaleRoIAlign): RoI k is pooled from level level_of_roi[k] (device int64), so the
// whole pyramid is one launch and no per-level index lists (host syncs) are needed.  levels: HOST array of n_levels entries.
extern "C" int hd_roi_align_ml_fwd(const hd_roi_level* levels, int n_levels, const float* rois, const int64_t* level_of_roi, float* out,
                                   int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio, hd_stream stream_) {
    if (int e = roi_check(num_rois, channels, pooled_h, pooled_w, sampling_ratio)) return e;
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= kMaxLevels);
    if (num_rois == 0) return HD_OK;
    HD_CHECK_ARG(rois != nullptr && level_of_roi != nullptr && out != nullptr);
    RoiLevels L;
    memset(&L, 0, sizeof(L));
    for (int i = 0; i < n_levels; ++i) {
        HD_CHECK_ARG(levels[i].feat_nhwc != nullptr && (reinterpret_cast<uintptr_t>(levels[i].feat_nhwc) & 15) == 0 && levels[i].h > 0 && levels[i].w > 0);
        L.feat[i] = levels[i].feat_nhwc; L.h[i] = levels[i].h; L.w[i] = levels[i].w; L.scale[i] = levels[i].scale;
    }
    return roi_fwd_launch(rois, reinterpret_cast<const long long*>(level_of_roi), L, out, num_rois, channels, pooled_h, pooled_w,
                          sampling_ratio, static_cast<cudaStream_t>(stream_));
}

// the same with bf16 channels-last feature maps (levels[i].feat_nhwc points at __nv_bfloat16): half the L2 traffic, identical
// results when the fp32 maps are widened copies of these
extern "C" int hd_roi_align_ml_fwd_bf16(const hd_roi_level* levels, int n_levels, const float* rois, const int64_t* level_of_roi, float* out,
                                   int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio, hd_stream stream_) {
    if (int e = roi_check(num_rois, channels, pooled_h, pooled_w, sampling_ratio)) return e;
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= kMaxLevels);
    if (num_rois == 0) return HD_OK;
    HD_CHECK_ARG(rois != nullptr && level_of_roi != nullptr && out != nullptr);
    RoiLevels L;
    memset(&L, 0, sizeof(L));
    for (int i = 0; i < n_levels; ++i) {
        HD_CHECK_ARG(levels[i].feat_nhwc != nullptr && (reinterpret_cast<uintptr_t>(levels[i].feat_nhwc) & 7) == 0 && levels[i].h > 0 && levels[i].w > 0);
        L.feat[i] = levels[i].feat_nhwc; L.h[i] = levels[i].h; L.w[i] = levels[i].w; L.scale[i] = levels[i].scale;
    }
    return roi_fwd_launch(rois, reinterpret_cast<const long long*>(level_of_roi), L, out, num_rois, channels, pooled_h, pooled_w,
                          sampling_ratio, static_cast<cudaStream_t>(stream_), true);
}

extern "C" int hd_roi_align_ml_bwd(const hd_roi_level* levels, int n_levels, const float* grad_out, const float* rois,
                                   const int64_t* level_of_roi, int num_rois, int channels, int pooled_h, int pooled_w, int sampling_ratio,
                                   hd_stream stream_) {
    if (int e = roi_check(num_rois, channels, pooled_h, pooled_w, sampling_ratio)) return e;
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= kMaxLevels);
    if (num_rois == 0) return HD_OK;
    HD_CHECK_ARG(rois != nullptr && level_of_roi != nullptr && grad_out != nullptr);
    RoiLevels L;
    memset(&L, 0, sizeof(L));
    for (int i = 0; i < n_levels; ++i) {
        HD_CHECK_ARG(levels[i].grad_nhwc != nullptr && (reinterpret_cast<uintptr_t>(levels[i].grad_nhwc) & 15) == 0 && levels[i].h > 0 && levels[i].w > 0);
        L.grad[i] = levels[i].grad_nhwc; L.h[i] = levels[i].h; L.w[i] = levels[i].w; L.scale[i] = levels[i].scale;
    }
    return roi_bwd_launch(grad_out, rois, reinterpret_cast<const long long*>(level_of_roi), L, num_rois, channels, pooled_h, pooled_w,
                          sampling_ratio, static_cast<cudaStream_t>(stream_));
}

extern "C" int hd_nhwc_to_nchw_f32(const float* x_nhwc, float* y_nchw, int n, int channels, int height, int width, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(x_nhwc != nullptr && y_nchw != nullptr && n > 0 && channels > 0 && height > 0 && width > 0 && n < 65536);
    const int hw = height * width;
    dim3 grid((hw + 31) / 32, (channels + 31) / 32, n);
    HD_CUDA_OK(hd::launch(nhwc_to_nchw_f32_kernel, dim3(grid), dim3(256), 0, stream, x_nhwc, y_nchw, channels, hw));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

// out [K][C][PH][PW] = torchvision.ops.roi_align(input, rois, spatial_scale, PH, PW, sampling_ratio, aligned = False), with the
// input given channels-last (feat_nhwc [N][H][W][C] fp32, see hd_nchw_to_nhwc_f32).  Same constraints as the backward.
extern "C" int hd_roi_align_fwd_nhwc(const float* feat_nhwc, const float* rois, float* out, int num_rois, int channels, int height,
                                     int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, hd_stream stream_) {
    if (int e = roi_check(num_rois, channels, pooled_h, pooled_w, sampling_ratio)) return e;
    HD_CHECK_ARG(height > 0 && width > 0);
    if (num_rois == 0) return HD_OK;
    HD_CHECK_ARG(feat_nhwc != nullptr && rois != nullptr && out != nullptr && (reinterpret_cast<uintptr_t>(feat_nhwc) & 15) == 0);
    RoiLevels L;
    memset(&L, 0, sizeof(L));
    L.feat[0] = feat_nhwc; L.h[0] = height; L.w[0] = width; L.scale[0] = spatial_scale;
    return roi_fwd_launch(rois, nullptr, L, out, num_rois, channels, pooled_h, pooled_w, sampling_ratio, static_cast<cudaStream_t>(stream_));
}

extern "C" int hd_nchw_to_nhwc_f32(const float* x_nchw, float* y_nhwc, int n, int channels, int height, int width, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(x_nchw != nullptr && y_nhwc != nullptr && n > 0 && channels > 0 && height > 0 && width > 0 && n < 65536);
    const int hw = height * width;
    dim3 grid((hw + 31) / 32, (channels + 31) / 32, n);
    HD_CUDA_OK(hd::launch(nchw_to_nhwc_f32_kernel, dim3(grid), dim3(256), 0, stream, x_nchw, y_nhwc, channels, hw));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}
