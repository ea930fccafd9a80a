// elementwise.cu -- the memory-bound kernels of the HalluciDet hot path (HBM roofline): weight packing,
// stem im2col/col2im, train-mode BatchNorm finalize/apply/backward, max-pool, nearest up-sampling, FPN
// top-down add, layout converters, detector input transform, sigmoid-head backward and the pixel regulariser.
// All activation kernels move 16 bytes (8 bf16 channels) per thread per access, consecutive threads on
// consecutive channel groups / pixels (coalesced NHWC), grid sized from the element count.
#include "hd_common.cuh"

namespace hd {

constexpr int kEwThreads = 256;

static inline int ew_blocks(long work) {
    long b = (work + kEwThreads - 1) / kEwThreads;
    if (b < 1) b = 1;
    const long cap = 148L * 64;                      // grid-stride above ~64 blocks per SM
    return static_cast<int>(b > cap ? cap : b);
}

struct bf8 {                                          // 8 bf16 channels = one 16-byte access
    uint4 u;
    __device__ __forceinline__ void load(const void* p) { u = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(void* p) const { *reinterpret_cast<uint4*>(p) = u; }
    __device__ __forceinline__ void unpack(float* f) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { f[2 * i] = bf16_lo(w[i]); f[2 * i + 1] = bf16_hi(w[i]); }
    }
    __device__ __forceinline__ void pack(const float* f) {
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    }
};

// -------------------------------------------------------------------------------------------------
// weight packing
// -------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, int cout, int cin,
                                   int kh, int kw, __nv_bfloat16* w_fwd, int cout_pad, int k_pad,
                                   __nv_bfloat16* w_dgrad, int cin_pad, __nv_bfloat16* w_t) {
    pdl_trigger();
    pdl_wait();
    const int taps = kh * kw;
    const long n_fwd = static_cast<long>(cout_pad) * k_pad;
    const long n_dg = static_cast<long>(cin_pad) * taps * cout;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_fwd + n_dg;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        if (i < n_fwd) {
            const int co = static_cast<int>(i / k_pad), k = static_cast<int>(i % k_pad);
            float v = 0.f;
            if (co < cout && k < taps * cin) {
                const int tap = k / cin, ci = k % cin;
                v = w[(static_cast<long>(co) * cin + ci) * taps + tap];
                if (scale) v *= scale[co];
            }
            const __nv_bfloat16 b = __float2bfloat16_rn(v);
            if (w_fwd) w_fwd[i] = b;
            if (w_t) w_t[static_cast<long>(k) * cout_pad + co] = b;
        } else if (w_dgrad) {
            const long j = i - n_fwd;
            const int ci = static_cast<int>(j / (taps * cout));
            const int rem = static_cast<int>(j % (taps * cout));
            const int tap = rem / cout, co = rem % cout;
            float v = 0.f;
            if (ci < cin) {
                v = w[(static_cast<long>(co) * cin + ci) * taps + tap];
                if (scale) v *= scale[co];
            }
            w_dgrad[j] = __float2bfloat16_rn(v);
        }
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, float* __restrict__ g, int cout, int cin, int taps,
                                    int tap_stride, int row_stride, float scale) {
    pdl_trigger();
    pdl_wait();
    const long n = static_cast<long>(cout) * cin * taps;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % taps);
        const int ci = static_cast<int>((i / taps) % cin);
        const int co = static_cast<int>(i / (static_cast<long>(taps) * cin));
        g[i] = dw[static_cast<long>(co) * row_stride + static_cast<long>(tap) * tap_stride + ci] * scale;
    }
}

// All layers of a network in one launch (the per-layer kernels above are 5-10 us of launch + DRAM latency each, 47 of
// them per step): block -> layer through the prefix of block counts stored in the descriptor table.
constexpr int kMultiItems = 4;                           // elements per thread

__device__ __forceinline__ int find_desc(const int* first_block, int stride_ints, int n, int block) {
    int lo = 0, hi = n - 1;                              // last descriptor whose first_block <= block
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (first_block[static_cast<long>(mid) * stride_ints] <= block) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Tiled path of the multi-layer packer: a block stages 16 output channels x TCI input channels x taps of the OIHW master
// weight (contiguous 4 * TCI * taps bytes per output channel) in shared memory and writes both bf16 layouts from there --
// the forward rows as 2 * TCI contiguous bytes per (cout, tap), the dgrad rows as full 32-byte sectors (16 couts) per
// (cin, tap).  The element-per-thread path below reads the master weight with a stride of `taps` (forward) or
// `cin * taps` (dgrad) floats: one 32-byte sector per element, 8x the traffic.
constexpr int kPackCo = 16, kPackCiMax = 64, kPackTapsMax = 9;

__host__ __device__ __forceinline__ bool pack_tiled_ok(const hd_pack_desc& d, int taps) {
    return d.w_fwd != nullptr && d.w_dgrad != nullptr && d.w_t == nullptr && taps <= kPackTapsMax && d.cout % kPackCo == 0 &&
           d.cin % 16 == 0 && d.cout_pad == d.cout && d.cin_pad == d.cin && d.k_pad == taps * d.cin;
}

// One Adam step of one element (torch.optim.Adam: lerp of the moments, bias-corrected step) after scale + clip of the gradient.
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const hd_adam_args& A) {
    g *= A.grad_scale;
    if (A.clip > 0.f) g = fminf(fmaxf(g, -A.clip), A.clip);
    m = m + (g - m) * A.one_minus_beta1;
    v = A.beta2 * v + A.one_minus_beta2 * g * g;
    const float denom = sqrtf(v) / sqrtf(A.bias_correction2) + A.eps;
    return p - (A.lr / A.bias_correction1) * (m / denom);
}

template <bool kAdam>
__device__ void pack_tiled(const hd_pack_desc& d, int taps, int blk, int nblk, float* sm, const hd_adam_args& A) {
    const int tci = d.cin % 64 == 0 ? 64 : (d.cin % 32 == 0 ? 32 : 16);
    const int ci_tiles = d.cin / tci, tiles = (d.cout / kPackCo) * ci_tiles;
    const int row = tci * taps, pitch = row + 1;                     // +1: the dgrad pass reads a column of 16 rows
    __nv_bfloat16* w_fwd = static_cast<__nv_bfloat16*>(d.w_fwd);
    __nv_bfloat16* w_dgrad = static_cast<__nv_bfloat16*>(d.w_dgrad);
    for (int t = blk; t < tiles; t += nblk) {
        const int co0 = (t / ci_tiles) * kPackCo, ci0 = (t % ci_tiles) * tci;
        __syncthreads();                                             // previous tile fully consumed
        for (int i = threadIdx.x; i < kPackCo * row; i += blockDim.x) {
            const int r = i / row, c = i - r * row;
            const long gi = (static_cast<long>(co0 + r) * d.cin + ci0) * taps + c;
            float v = d.w[gi];
            if (kAdam && d.g != nullptr) {                       // the optimizer step happens on the one read of the master weight
                float mm = d.m[gi], vv = d.v[gi];
                v = adam_update(v, d.g[gi], mm, vv, A);
                const_cast<float*>(d.w)[gi] = v;
                d.m[gi] = mm;
                d.v[gi] = vv;
            }
            if (d.scale) v *= d.scale[co0 + r];
            sm[r * pitch + c] = v;
        }
        __syncthreads();
        // forward layout [cout][tap * cin + ci]: pairs of consecutive input channels
        const int half = tci / 2;
        for (int i = threadIdx.x; i < kPackCo * taps * half; i += blockDim.x) {
            const int cp = i % half, tap = (i / half) % taps, r = i / (half * taps);
            const float a = sm[r * pitch + (2 * cp) * taps + tap], b = sm[r * pitch + (2 * cp + 1) * taps + tap];
            *reinterpret_cast<__nv_bfloat162*>(w_fwd + static_cast<long>(co0 + r) * d.k_pad + tap * d.cin + ci0 + 2 * cp) =
                __floats2bfloat162_rn(a, b);
        }
        // dgrad layout [cin][tap * cout + co]: pairs of consecutive output channels
        for (int i = threadIdx.x; i < tci * taps * (kPackCo / 2); i += blockDim.x) {
            const int rp = i % (kPackCo / 2), tap = (i / (kPackCo / 2)) % taps, c = i / ((kPackCo / 2) * taps);
            const float a = sm[(2 * rp) * pitch + c * taps + tap], b = sm[(2 * rp + 1) * pitch + c * taps + tap];
            *reinterpret_cast<__nv_bfloat162*>(w_dgrad + (static_cast<long>(ci0 + c) * taps + tap) * d.cout + co0 + 2 * rp) =
                __floats2bfloat162_rn(a, b);
        }
    }
}

template <bool kAdam>
__global__ void pack_weights_multi_kernel(const hd_pack_desc* __restrict__ descs, int n, int total_blocks, const hd_adam_args A) {
    pdl_trigger();
    pdl_wait();
    __shared__ float pack_sm[kPackCo * (kPackCiMax * kPackTapsMax + 1)];
    const int li = find_desc(&descs[0].first_block, sizeof(hd_pack_desc) / sizeof(int), n, blockIdx.x);
    const hd_pack_desc d = descs[li];
    const int taps = d.kh * d.kw;
    if (pack_tiled_ok(d, taps)) {
        const int next = li + 1 < n ? descs[li + 1].first_block : total_blocks;
        pack_tiled<kAdam>(d, taps, blockIdx.x - d.first_block, next - d.first_block, pack_sm, A);
        return;
    }
    const long n_fwd = static_cast<long>(d.cout_pad) * d.k_pad;
    const long n_dg = d.w_dgrad ? static_cast<long>(d.cin_pad) * taps * d.cout : 0;
    __nv_bfloat16* w_fwd = static_cast<__nv_bfloat16*>(d.w_fwd);
    __nv_bfloat16* w_dgrad = static_cast<__nv_bfloat16*>(d.w_dgrad);
    __nv_bfloat16* w_t = static_cast<__nv_bfloat16*>(d.w_t);
    const long base = static_cast<long>(blockIdx.x - d.first_block) * blockDim.x * kMultiItems;
#pragma unroll
    for (int it = 0; it < kMultiItems; ++it) {
        const long i = base + it * blockDim.x + threadIdx.x;
        if (i >= n_fwd + n_dg) break;
        if (i < n_fwd) {
            const int co = static_cast<int>(i / d.k_pad), k = static_cast<int>(i % d.k_pad);
            float v = 0.f;
            if (co < d.cout && k < taps * d.cin) {
                const int tap = k / d.cin, ci = k % d.cin;
                v = d.w[(static_cast<long>(co) * d.cin + ci) * taps + tap];
                if (d.scale) v *= d.scale[co];
            }
            const __nv_bfloat16 b = __float2bfloat16_rn(v);
            if (w_fwd) w_fwd[i] = b;
            if (w_t) w_t[static_cast<long>(k) * d.cout_pad + co] = b;
        } else {
            const long j = i - n_fwd;
            const int ci = static_cast<int>(j / (taps * d.cout));
            const int rem = static_cast<int>(j % (taps * d.cout));
            const int tap = rem / d.cout, co = rem % d.cout;
            float v = 0.f;
            if (ci < d.cin) {
                v = d.w[(static_cast<long>(co) * d.cin + ci) * taps + tap];
                if (d.scale) v *= d.scale[co];
            }
            w_dgrad[j] = __float2bfloat16_rn(v);
        }
    }
}

// One block = whole output-channel rows of one layer (as many as fit 1024 elements, at least one): a row is read in the
// accumulator's order ([tap][ci], coalesced), transposed through shared memory and written in the gradient's order ([ci][tap],
// coalesced).  The element-wise form read 4-byte words `tap_stride` apart and spent 64-bit divisions on every element (124 us
// for the U-Net's 24.4 M gradients).  Rows longer than kUnpackRow elements are walked in ci chunks.
constexpr int kUnpackRow = 8192;

__global__ void __launch_bounds__(kEwThreads) unpack_wgrads_multi_kernel(const hd_unpack_desc* __restrict__ descs, int n) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[kUnpackRow];
    const int li = find_desc(&descs[0].first_block, sizeof(hd_unpack_desc) / sizeof(int), n, blockIdx.x);
    const hd_unpack_desc d = descs[li];
    const int row = d.cin * d.taps;
    const int rows_per_block = row >= kMultiItems * kEwThreads ? 1 : (kMultiItems * kEwThreads) / row;
    const int co0 = (blockIdx.x - d.first_block) * rows_per_block;
    const int ci_chunk = row <= kUnpackRow ? d.cin : kUnpackRow / d.taps;
    for (int r = 0; r < rows_per_block; ++r) {
        const int co = co0 + r;
        if (co >= d.cout) break;
        const float* src = d.dw + static_cast<long>(co) * d.row_stride;
        float* dst = d.g + static_cast<long>(co) * row;
        for (int c0 = 0; c0 < d.cin; c0 += ci_chunk) {
            const int nc = min(ci_chunk, d.cin - c0);
            __syncthreads();                                   // the tile of the previous chunk / row has been written out
            for (int idx = threadIdx.x; idx < nc * d.taps; idx += kEwThreads) {
                const int t = idx / nc, c = idx - t * nc;
                tile[c * d.taps + t] = src[static_cast<long>(t) * d.tap_stride + c0 + c] * d.scale;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < nc * d.taps; i += kEwThreads) dst[static_cast<long>(c0) * d.taps + i] = tile[i];
        }
    }
}

__global__ void adam_multi_kernel(const hd_adam_desc* __restrict__ descs, int n, const hd_adam_args A) {
    pdl_trigger();
    pdl_wait();
    const int li = find_desc(&descs[0].first_block, sizeof(hd_adam_desc) / sizeof(int), n, blockIdx.x);
    const hd_adam_desc d = descs[li];
    const long base = static_cast<long>(blockIdx.x - d.first_block) * blockDim.x * kMultiItems;
#pragma unroll
    for (int it = 0; it < kMultiItems; ++it) {
        const long i = base + it * blockDim.x + threadIdx.x;
        if (i >= d.n) break;
        float mm = d.m[i], vv = d.v[i];
        d.p[i] = adam_update(d.p[i], d.g[i], mm, vv, A);
        d.m[i] = mm;
        d.v[i] = vv;
    }
}

// -------------------------------------------------------------------------------------------------
// stem 7x7/2 patches
// -------------------------------------------------------------------------------------------------
// One CTA = 8 x 32 output pixels: the (2*8+5) x (2*32+5) x 3 input patch is staged once in shared memory as HWC bf16 (so the
// 21 values of a filter row are contiguous), then every thread assembles 8-value groups of the patch rows and writes them
// with coalesced 16-byte stores.  (A direct gather issued 8 scattered global loads per 16 output bytes.)
constexpr int kI2cTH = 8, kI2cTW = 32, kI2cPH = 2 * kI2cTH + 5, kI2cPW = 2 * kI2cTW + 5, kI2cPitch = kI2cPW * 3 + 1;

__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ patches, int n,
                                                          int h, int w, int k_pad) {
    pdl_trigger();
    pdl_wait();
    __shared__ __nv_bfloat16 sp[kI2cPH * kI2cPitch];
    const int ho = h / 2, wo = w / 2, groups = k_pad / 8;
    const int tiles_w = (wo + kI2cTW - 1) / kI2cTW, tiles_h = (ho + kI2cTH - 1) / kI2cTH;
    const int tile = blockIdx.x;
    const int b = tile / (tiles_w * tiles_h), t2 = tile - b * tiles_w * tiles_h;
    const int oh0 = (t2 / tiles_w) * kI2cTH, ow0 = (t2 % tiles_w) * kI2cTW;
    const int ih0 = 2 * oh0 - 3, iw0 = 2 * ow0 - 3;
    for (int q = threadIdx.x; q < 3 * kI2cPH * kI2cPW; q += blockDim.x) {
        const int col = q % kI2cPW, t3 = q / kI2cPW, row = t3 % kI2cPH, c = t3 / kI2cPH;
        const int ih = ih0 + row, iw = iw0 + col;
        float v = 0.f;
        if (ih >= 0 && ih < h && iw >= 0 && iw < w) v = __ldg(x + ((static_cast<long>(b) * 3 + c) * h + ih) * w + iw);
        sp[row * kI2cPitch + col * 3 + c] = __float2bfloat16(v);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kI2cTH * kI2cTW * groups; q += blockDim.x) {
        const int g = q % groups, p = q / groups;
        const int ohl = p / kI2cTW, owl = p - ohl * kI2cTW;
        const int oh = oh0 + ohl, ow = ow0 + owl;
        if (oh >= ho || ow >= wo) continue;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = g * 8 + j;                           // k = (r*7 + s)*3 + c = r*21 + (s*3 + c)
            const int r = k / 21, rem = k - r * 21;
            f[j] = k < 147 ? __bfloat162float(sp[(2 * ohl + r) * kI2cPitch + 2 * owl * 3 + rem]) : 0.f;
        }
        bf8 o;
        o.pack(f);
        o.store(patches + ((static_cast<long>(b) * ho + oh) * wo + ow) * k_pad + g * 8);
    }
}

// Single-channel variant for the hallucination U-Net: HalluciDet feeds the IR plane replicated three times
// (src/utils/utils.py:52-53), so conv(W, [x, x, x]) == conv(sum_c W[:, c], x) exactly; the stem runs with K = 49 (padded to 64)
// instead of 147 (padded to 160): 2.5x less patch traffic.  The input may be the uint8 camera plane itself (scale = 1/255:
// the dataloader's ToTensor division, src/dataloader/dataloader.py:13-73) -- one byte per pixel over PCIe instead of four.
template <typename T>
__global__ void __launch_bounds__(256) stem_im2col1_kernel(const T* __restrict__ x, float scale, __nv_bfloat16* __restrict__ patches,
                                                           int n, int h, int w, int k_pad) {
    pdl_trigger();
    pdl_wait();
    constexpr int pitch = kI2cPW + 1;
    __shared__ __nv_bfloat16 sp[kI2cPH * pitch];
    const int ho = h / 2, wo = w / 2, groups = k_pad / 8;
    const int tiles_w = (wo + kI2cTW - 1) / kI2cTW, tiles_h = (ho + kI2cTH - 1) / kI2cTH;
    const int tile = blockIdx.x;
    const int b = tile / (tiles_w * tiles_h), t2 = tile - b * tiles_w * tiles_h;
    const int oh0 = (t2 / tiles_w) * kI2cTH, ow0 = (t2 % tiles_w) * kI2cTW;
    const int ih0 = 2 * oh0 - 3, iw0 = 2 * ow0 - 3;
    for (int q = threadIdx.x; q < kI2cPH * kI2cPW; q += blockDim.x) {
        const int col = q % kI2cPW, row = q / kI2cPW;
        const int ih = ih0 + row, iw = iw0 + col;
        float v = 0.f;
        if (ih >= 0 && ih < h && iw >= 0 && iw < w) v = static_cast<float>(x[(static_cast<long>(b) * h + ih) * w + iw]) * scale;
        sp[row * pitch + col] = __float2bfloat16(v);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kI2cTH * kI2cTW * groups; q += blockDim.x) {
        const int g = q % groups, p = q / groups;
        const int ohl = p / kI2cTW, owl = p - ohl * kI2cTW;
        const int oh = oh0 + ohl, ow = ow0 + owl;
        if (oh >= ho || ow >= wo) continue;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = g * 8 + j;                           // k = r*7 + s
            const int r = k / 7, c = k - r * 7;
            f[j] = k < 49 ? __bfloat162float(sp[(2 * ohl + r) * pitch + 2 * owl + c]) : 0.f;
        }
        bf8 o;
        o.pack(f);
        o.store(patches + ((static_cast<long>(b) * ho + oh) * wo + ow) * k_pad + g * 8);
    }
}

__global__ void stem_col2im_kernel(const __nv_bfloat16* __restrict__ dp, float* __restrict__ dx, int n, int h, int w, int k_pad) {
    pdl_trigger();
    pdl_wait();
    const int ho = h / 2, wo = w / 2;
    const long total = static_cast<long>(n) * h * w;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int iw = static_cast<int>(i % w), ih = static_cast<int>((i / w) % h), b = static_cast<int>(i / (static_cast<long>(w) * h));
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int r = (ih + 3) & 1; r < 7; r += 2) {
            const int oh = (ih + 3 - r) / 2;
            if (ih + 3 - r < 0 || oh >= ho) continue;
            for (int s = (iw + 3) & 1; s < 7; s += 2) {
                const int ow = (iw + 3 - s) / 2;
                if (iw + 3 - s < 0 || ow >= wo) continue;
                const __nv_bfloat16* p = dp + ((static_cast<long>(b) * ho + oh) * wo + ow) * k_pad + (r * 7 + s) * 3;
                a0 += __bfloat162float(p[0]); a1 += __bfloat162float(p[1]); a2 += __bfloat162float(p[2]);
            }
        }
        const long plane = static_cast<long>(h) * w;
        float* o = dx + static_cast<long>(b) * 3 * plane + static_cast<long>(ih) * w + iw;
        o[0] = a0; o[plane] = a1; o[2 * plane] = a2;
    }
}

// -------------------------------------------------------------------------------------------------
// train-mode BatchNorm
// -------------------------------------------------------------------------------------------------
// one block per channel: fixed-order reduction over the per-tile partial rows (deterministic), fp64 mean / variance
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int rows, int C, double count, const float* gamma,
                                   const float* beta, float eps, float momentum, float* rm, float* rv, float* mean_out,
                                   float* invstd_out, float* scale_out, float* shift_out) {
    pdl_trigger();
    pdl_wait();
    __shared__ double ss[128], sq[128];
    const int c = blockIdx.x, t = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int r = t; r < rows; r += 128) {
        s += stats[(static_cast<long>(r) * 2) * C + c];
        q += stats[(static_cast<long>(r) * 2 + 1) * C + c];
    }
    ss[t] = s;
    sq[t] = q;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (t < o) { ss[t] += ss[t + o]; sq[t] += sq[t + o]; }
        __syncthreads();
    }
    if (t != 0) return;
    s = ss[0];
    q = sq[0];
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = gamma[c] * invstd;
    if (mean_out) mean_out[c] = static_cast<float>(mean);
    if (invstd_out) invstd_out[c] = invstd;
    scale_out[c] = sc;
    shift_out[c] = beta[c] - static_cast<float>(mean) * sc;
    if (rm) rm[c] = (1.f - momentum) * rm[c] + momentum * static_cast<float>(mean);
    if (rv) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rv[c] = (1.f - momentum) * rv[c] + momentum * static_cast<float>(unbiased);
    }
}

// The channel group of a thread is loop invariant (G = C/8 divides the 256-thread block, hence the grid stride), so the
// per-channel coefficients live in registers; two 16-byte groups are in flight per thread and iteration.
__global__ void __launch_bounds__(kEwThreads) bn_apply_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                                const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                                const float* __restrict__ rscale, const float* __restrict__ rshift, int relu,
                                __nv_bfloat16* __restrict__ y, long n_pix, int C) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8;
    const long total = n_pix * G;
    const int c0 = static_cast<int>(threadIdx.x % G) * 8;
    // per-channel coefficients: loaded once per block into shared memory (32 scalar global loads per thread were most of the
    // run time of the small layers, where a thread handles one or two 16-byte groups), then 16-byte shared loads
    extern __shared__ __align__(16) float coef[];          // [4][C]: scale, shift, res scale, res shift
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        coef[c] = __ldg(scale + c); coef[C + c] = __ldg(shift + c);
        coef[2 * C + c] = rscale ? __ldg(rscale + c) : 1.f; coef[3 * C + c] = rscale ? __ldg(rshift + c) : 0.f;
    }
    __syncthreads();
    float sc[8], sh[8], rs[8], rb[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(coef + c0 + 4 * h), b = *reinterpret_cast<const float4*>(coef + C + c0 + 4 * h);
        const float4 c = *reinterpret_cast<const float4*>(coef + 2 * C + c0 + 4 * h), d = *reinterpret_cast<const float4*>(coef + 3 * C + c0 + 4 * h);
        sc[4 * h] = a.x; sc[4 * h + 1] = a.y; sc[4 * h + 2] = a.z; sc[4 * h + 3] = a.w;
        sh[4 * h] = b.x; sh[4 * h + 1] = b.y; sh[4 * h + 2] = b.z; sh[4 * h + 3] = b.w;
        rs[4 * h] = c.x; rs[4 * h + 1] = c.y; rs[4 * h + 2] = c.z; rs[4 * h + 3] = c.w;
        rb[4 * h] = d.x; rb[4 * h + 1] = d.y; rb[4 * h + 2] = d.z; rb[4 * h + 3] = d.w;
    }
    const long stride = static_cast<long>(gridDim.x) * blockDim.x;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += 2 * stride) {
        const long i2 = i + stride;
        const bool two = i2 < total;
        bf8 v[2], rv[2];
        v[0].load(z + i * 8);
        if (two) v[1].load(z + i2 * 8);
        if (res) {
            rv[0].load(res + i * 8);
            if (two) rv[1].load(res + i2 * 8);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            float f[8];
            v[u].unpack(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
            if (res) {
                float r[8];
                rv[u].unpack(r);
                if (rscale) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fmaf(r[j], rs[j], rb[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += r[j];
            }
            if (relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            bf8 o;
            o.pack(f);
            o.store(y + (u ? i2 : i) * 8);
        }
    }
}

// sums[0][c] = sum g, sums[1][c] = sum g*xhat with g = dy*(y>0)
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ yrelu,
                                     const float* __restrict__ rscale, const float* __restrict__ rshift,
                                     const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, float* __restrict__ sums, long n_pix, int C,
                                     long pix_per_block) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float sacc[];      // [2][C] sums, then [4][C] coefficients (mean, invstd, relu scale, relu shift)
    float* coef = sacc + 2 * C;
    const int G = C / 8;
    const int L = blockDim.x / G;                      // pixel lanes
    const int g = threadIdx.x % G, l = threadIdx.x / G;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        coef[c] = __ldg(mean + c); coef[C + c] = __ldg(invstd + c);
        coef[2 * C + c] = rscale ? __ldg(rscale + c) : 0.f; coef[3 * C + c] = rscale ? __ldg(rshift + c) : 0.f;
    }
    __syncthreads();
    float a[8], b[8], mu[8], is[8], rs[8], rb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = 0.f; b[j] = 0.f; mu[j] = coef[g * 8 + j]; is[j] = coef[C + g * 8 + j];
        rs[j] = coef[2 * C + g * 8 + j]; rb[j] = coef[3 * C + g * 8 + j];
    }
    const long p0 = blockIdx.x * pix_per_block;
    long p1 = p0 + pix_per_block;
    if (p1 > n_pix) p1 = n_pix;
    if (l < L) {
        for (long p = p0 + l; p < p1; p += 2 * L) {   // two pixels (16-byte groups) in flight per thread
            const bool two = p + L < p1;
            const long offs[2] = {p * C + g * 8, (p + L) * C + g * 8};
            bf8 d[2], zz[2], yy[2];
            d[0].load(dy + offs[0]);
            zz[0].load(z + offs[0]);
            if (yrelu) yy[0].load(yrelu + offs[0]);
            if (two) {
                d[1].load(dy + offs[1]);
                zz[1].load(z + offs[1]);
                if (yrelu) yy[1].load(yrelu + offs[1]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                float df[8], zf[8];
                d[u].unpack(df);
                zz[u].unpack(zf);
                if (yrelu) {
                    float yf[8];
                    yy[u].unpack(yf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (!(yf[j] > 0.f)) df[j] = 0.f;
                } else if (rscale) {                  // ReLU directly after this BN: the mask is recomputed from z
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (!(fmaf(zf[j], rs[j], rb[j]) > 0.f)) df[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { a[j] += df[j]; b[j] += df[j] * (zf[j] - mu[j]) * is[j]; }
            }
        }
        if (G < 32 && (G & (G - 1)) == 0) {              // lanes sharing a channel group inside the warp
            for (int o = 16; o >= G; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
                    b[j] += __shfl_xor_sync(0xffffffffu, b[j], o);
                }
            }
            if ((threadIdx.x & 31) < G) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[g * 8 + j], a[j]); atomicAdd(&sacc[C + g * 8 + j], b[j]); }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[g * 8 + j], a[j]); atomicAdd(&sacc[C + g * 8 + j], b[j]); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], sacc[i]);
}

__global__ void __launch_bounds__(kEwThreads) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ yrelu,
                                    const float* __restrict__ rscale, const float* __restrict__ rshift,
                                    const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ sums, float inv_count, __nv_bfloat16* __restrict__ dz,
                                    __nv_bfloat16* __restrict__ gout, float* dgamma, float* dbeta, long n_pix, int C) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8;
    const long total = n_pix * G;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta) dbeta[c] = sums[c];
            if (dgamma) dgamma[c] = sums[C + c];
        }
    }
    // loop-invariant channel group (see bn_apply_kernel): coefficients in registers, computed once per block via shared memory
    const int c0 = static_cast<int>(threadIdx.x % G) * 8;
    // dz = k1 * (g - m1 - xhat * m2), xhat = (z - mu) * is  ==  A * g + B * z + D  (three coefficients per channel)
    extern __shared__ __align__(16) float coef[];          // [5][C]: A, B, D, relu scale, relu shift
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float is = __ldg(invstd + c), mu = __ldg(mean + c);
        const float k1 = __ldg(gamma + c) * is;
        const float m1 = __ldg(sums + c) * inv_count, m2 = __ldg(sums + C + c) * inv_count;
        coef[c] = k1; coef[C + c] = -k1 * is * m2; coef[2 * C + c] = k1 * (is * m2 * mu - m1);
        coef[3 * C + c] = rscale ? __ldg(rscale + c) : 0.f; coef[4 * C + c] = rscale ? __ldg(rshift + c) : 0.f;
    }
    __syncthreads();
    float ca[8], cb[8], cd[8], rs[8], rb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        ca[j] = coef[c0 + j]; cb[j] = coef[C + c0 + j]; cd[j] = coef[2 * C + c0 + j];
        rs[j] = coef[3 * C + c0 + j]; rb[j] = coef[4 * C + c0 + j];
    }
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        bf8 d, zz;
        d.load(dy + i * 8);
        zz.load(z + i * 8);
        float df[8], zf[8], o[8];
        d.unpack(df);
        zz.unpack(zf);
        if (yrelu) {
            bf8 yy;
            yy.load(yrelu + i * 8);
            float yf[8];
            yy.unpack(yf);
#pragma unroll
            for (int j = 0; j < 8; ++j) if (!(yf[j] > 0.f)) df[j] = 0.f;
        } else if (rscale) {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (!(fmaf(zf[j], rs[j], rb[j]) > 0.f)) df[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(ca[j], df[j], fmaf(cb[j], zf[j], cd[j]));
        bf8 ov;
        ov.pack(o);
        ov.store(dz + i * 8);
        if (gout) {
            bf8 gv;
            gv.pack(df);
            gv.store(gout + i * 8);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Fused train-mode BatchNorm backward: reduce + apply in ONE persistent kernel (one CTA per SM, all co-resident).
//   phase 1: every CTA reduces its contiguous pixel slice (sum g, sum g*xhat with the ReLU mask folded into g) and adds its
//            partial sums to sums[2][C]; when the slice fits, the masked g and z stay in shared memory;
//   grid barrier (sense-reversing counter, self-resetting: safe under stream order and CUDA-graph replay);
//   phase 2: dz = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) from shared memory (small layers: g and z are read from
//            HBM/L2 exactly once) or by re-reading the slice (large layers: second read served mostly by the 126 MB L2).
// One launch instead of two per BatchNorm layer, 3 instead of 5 tensor passes for the 33 layers whose slice fits on chip.
// -------------------------------------------------------------------------------------------------
constexpr int kBnFusedThreads = 512;
constexpr int kBnUnroll = 4;

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int* vgen = bar + 1;
        const unsigned int gen = *vgen;                    // generation before arriving: the release cannot be missed
        __threadfence();
        const unsigned int prev = atomicAdd(bar, 1u);
        if (prev == nblocks - 1u) {
            bar[0] = 0u;                                   // nobody arrives again before the release below
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*vgen == gen) __nanosleep(40);
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kBnFusedThreads, 1)
bn_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ yrelu, const float* __restrict__ rscale,
                    const float* __restrict__ rshift, const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, float* sums, float inv_count,
                    __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ gout, float* dgamma, float* dbeta, long n_pix, int C,
                    long pix_per_block, int cache, unsigned int* bar) {
    pdl_trigger();
    extern __shared__ __align__(16) uint8_t bsm[];
    float* sacc = reinterpret_cast<float*>(bsm);                       // [2][C]
    uint4* cache_g = reinterpret_cast<uint4*>(bsm + 2 * C * sizeof(float));
    const int G = C / 8;
    const int L = kBnFusedThreads / G;                                 // pixel lanes
    const int g = threadIdx.x % G, l = threadIdx.x / G;
    const long p0 = blockIdx.x * pix_per_block;
    long p1 = p0 + pix_per_block;
    if (p1 > n_pix) p1 = n_pix;
    const int iters = static_cast<int>((pix_per_block + L - 1) / L);   // slice-local steps of this thread (cache index)
    uint4* cache_z = cache_g + static_cast<long>(iters) * kBnFusedThreads;
    for (int i = threadIdx.x; i < 2 * C; i += kBnFusedThreads) sacc[i] = 0.f;
    pdl_wait();
    float a[8], b[8], mu[8], is[8], rs[8], rb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = 0.f; b[j] = 0.f; mu[j] = __ldg(mean + g * 8 + j); is[j] = __ldg(invstd + g * 8 + j);
        rs[j] = rscale ? __ldg(rscale + g * 8 + j) : 0.f; rb[j] = rscale ? __ldg(rshift + g * 8 + j) : 0.f;
    }
    __syncthreads();
    // ---- phase 1
    for (int it0 = 0; it0 < iters; it0 += kBnUnroll) {
        bf8 d[kBnUnroll], zz[kBnUnroll], yy[kBnUnroll];
        bool on[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const long p = p0 + l + static_cast<long>(it0 + u) * L;
            on[u] = (it0 + u) < iters && p < p1;
            if (on[u]) {
                const long off = p * C + g * 8;
                d[u].load(dy + off);
                zz[u].load(z + off);
                if (yrelu) yy[u].load(yrelu + off);
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (!on[u]) continue;
            float df[8], zf[8];
            d[u].unpack(df);
            zz[u].unpack(zf);
            if (yrelu) {
                float yf[8];
                yy[u].unpack(yf);
#pragma unroll
                for (int j = 0; j < 8; ++j) if (!(yf[j] > 0.f)) df[j] = 0.f;
            } else if (rscale) {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (!(fmaf(zf[j], rs[j], rb[j]) > 0.f)) df[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += df[j]; b[j] += df[j] * (zf[j] - mu[j]) * is[j]; }
            if (cache) {
                bf8 gm;
                gm.pack(df);                                           // masking only zeroes values: exact in bf16
                cache_g[static_cast<long>(it0 + u) * kBnFusedThreads + threadIdx.x] = gm.u;
                cache_z[static_cast<long>(it0 + u) * kBnFusedThreads + threadIdx.x] = zz[u].u;
            }
        }
    }
    if (G < 32) {                                                      // lanes sharing a channel group inside the warp
        for (int o = 16; o >= G; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
                b[j] += __shfl_xor_sync(0xffffffffu, b[j], o);
            }
        }
        if ((threadIdx.x & 31) < G) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[g * 8 + j], a[j]); atomicAdd(&sacc[C + g * 8 + j], b[j]); }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[g * 8 + j], a[j]); atomicAdd(&sacc[C + g * 8 + j], b[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kBnFusedThreads) atomicAdd(&sums[i], sacc[i]);
    grid_barrier(bar, gridDim.x);
    // ---- phase 2
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += kBnFusedThreads) {
            if (dbeta) dbeta[c] = __ldcg(sums + c);
            if (dgamma) dgamma[c] = __ldcg(sums + C + c);
        }
    }
    float ca[8], cb[8], cd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        const float k1 = __ldg(gamma + c) * is[j];
        const float m1 = __ldcg(sums + c) * inv_count, m2 = __ldcg(sums + C + c) * inv_count;
        ca[j] = k1; cb[j] = -k1 * is[j] * m2; cd[j] = k1 * (is[j] * m2 * mu[j] - m1);
    }
    for (int it0 = 0; it0 < iters; it0 += kBnUnroll) {
        bf8 d[kBnUnroll], zz[kBnUnroll], yy[kBnUnroll];
        bool on[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const long p = p0 + l + static_cast<long>(it0 + u) * L;
            on[u] = (it0 + u) < iters && p < p1;
            if (!on[u]) continue;
            if (cache) {
                d[u].u = cache_g[static_cast<long>(it0 + u) * kBnFusedThreads + threadIdx.x];
                zz[u].u = cache_z[static_cast<long>(it0 + u) * kBnFusedThreads + threadIdx.x];
            } else {
                const long off = p * C + g * 8;
                d[u].load(dy + off);
                zz[u].load(z + off);
                if (yrelu) yy[u].load(yrelu + off);
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (!on[u]) continue;
            const long off = (p0 + l + static_cast<long>(it0 + u) * L) * C + g * 8;
            float df[8], zf[8], o[8];
            d[u].unpack(df);
            zz[u].unpack(zf);
            if (!cache) {
                if (yrelu) {
                    float yf[8];
                    yy[u].unpack(yf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (!(yf[j] > 0.f)) df[j] = 0.f;
                } else if (rscale) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (!(fmaf(zf[j], rs[j], rb[j]) > 0.f)) df[j] = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(ca[j], df[j], fmaf(cb[j], zf[j], cd[j]));
            bf8 ov;
            ov.pack(o);
            ov.store(dz + off);
            if (gout) {
                bf8 gv;
                gv.pack(df);
                gv.store(gout + off);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 pad 1 (NHWC)
// -------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   unsigned char* __restrict__ idx, int mask_nonpositive, int n, int h, int w, int C) {
    pdl_trigger();
    pdl_wait();
    const int ho = h / 2, wo = w / 2, G = C / 8;
    // 32-bit index arithmetic (the host checks n*h*w*C/8 < 2^31)
    const unsigned total = static_cast<unsigned>(n) * ho * wo * G;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const unsigned pix = i / G;
        const int ow = static_cast<int>(pix % wo), oh = static_cast<int>((pix / wo) % ho), b = static_cast<int>(pix / (static_cast<unsigned>(wo) * ho));
        float m[8];
        unsigned am[8];                                    // window position (r*3+s) of the FIRST maximum, as ATen
#pragma unroll
        for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; am[j] = 15u; }
        for (int r = 0; r < 3; ++r) {
            const int ih = 2 * oh + r - 1;
            if (ih < 0 || ih >= h) continue;
            for (int s = 0; s < 3; ++s) {
                const int iw = 2 * ow + s - 1;
                if (iw < 0 || iw >= w) continue;
                bf8 v;
                v.load(x + ((static_cast<long>(b) * h + ih) * w + iw) * C + g * 8);
                float f[8];
                v.unpack(f);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (f[j] > m[j]) { m[j] = f[j]; am[j] = r * 3 + s; }
            }
        }
        bf8 o;
        o.pack(m);
        o.store(y + static_cast<size_t>(pix) * C + g * 8);
        if (idx != nullptr) {
            unsigned lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                lo |= ((mask_nonpositive && !(m[j] > 0.f)) ? 15u : am[j]) << (8 * j);
                hi |= ((mask_nonpositive && !(m[j + 4] > 0.f)) ? 15u : am[j + 4]) << (8 * j);
            }
            *reinterpret_cast<uint2*>(idx + static_cast<size_t>(pix) * C + g * 8) = make_uint2(lo, hi);
        }
    }
}

// backward from the stored arg-max positions: dx[ih,iw] = (add) + sum of dy over the (<= 4) windows whose arg-max is here
__global__ void maxpool_bwd_idx_kernel(const unsigned char* __restrict__ idx, const __nv_bfloat16* __restrict__ dy,
                                       const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx, int n, int h,
                                       int w, int C) {
    pdl_trigger();
    pdl_wait();
    // 32-bit index arithmetic (the host checks n*h*w*C/8 < 2^31); the arg-max bytes of all candidate windows are loaded
    // first, the dy vectors only for windows that hit
    const int ho = h / 2, wo = w / 2, G = C / 8;
    const unsigned total = static_cast<unsigned>(n) * h * w * G;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned g = i % G, pix = i / G;
        const unsigned iw = pix % w, t = pix / w, ih = t % h, b = t / h;
        float acc[8];
        if (add) {
            bf8 av;
            av.load(add + static_cast<size_t>(pix) * C + g * 8);
            av.unpack(acc);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        }
        // windows with 2*oh-1 <= ih <= 2*oh+1: oh = ih/2 and, for odd ih, also (ih+1)/2
        const unsigned oh0 = ih >> 1, ow0 = iw >> 1;
        uint2 am[4];
        unsigned hit[4];
        size_t off[4];
        bool on_q[4];
        bf8 dvq[4];
        // all loads of the candidate windows (arg-max bytes AND their dy vectors) are issued before anything is decoded: one
        // L2 round trip per element instead of two dependent ones (a vector of 8 channels is hit ~90 % of the time anyway)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned oh = oh0 + (q >> 1), ow = ow0 + (q & 1);
            on_q[q] = ((q >> 1) == 0 || (ih & 1u)) && ((q & 1) == 0 || (iw & 1u)) && oh < static_cast<unsigned>(ho) &&
                      ow < static_cast<unsigned>(wo);
            off[q] = (static_cast<size_t>(b * ho + oh) * wo + ow) * C + g * 8;
            am[q] = on_q[q] ? *reinterpret_cast<const uint2*>(idx + off[q]) : make_uint2(0xffffffffu, 0xffffffffu);
            if (on_q[q]) dvq[q].load(dy + off[q]);
            else dvq[q].u = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned oh = oh0 + (q >> 1), ow = ow0 + (q & 1);
            const bool on = on_q[q];
            const unsigned pos = (ih + 1 - 2 * oh) * 3 + (iw + 1 - 2 * ow);
            const unsigned pat = pos * 0x01010101u;                    // pos in every byte
            const unsigned x0 = am[q].x ^ pat, x1 = am[q].y ^ pat;     // a zero byte = arg-max here
            unsigned hbits = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hbits |= (((x0 >> (8 * j)) & 0xffu) == 0u) ? (1u << j) : 0u;
                hbits |= (((x1 >> (8 * j)) & 0xffu) == 0u) ? (1u << (j + 4)) : 0u;
            }
            hit[q] = on ? hbits : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!hit[q]) continue;
            float df[8];
            dvq[q].unpack(df);
#pragma unroll
            for (int j = 0; j < 8; ++j) if (hit[q] & (1u << j)) acc[j] += df[j];
        }
        bf8 o;
        o.pack(acc);
        o.store(dx + static_cast<size_t>(pix) * C + g * 8);
    }
}

// dx[ih,iw] = (add) + sum over windows whose FIRST maximum (scan order) is (ih,iw) of dy; optional x>0 mask
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                   const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ add,
                                   __nv_bfloat16* __restrict__ dx, int n, int h, int w, int C, int relu_mask) {
    pdl_trigger();
    pdl_wait();
    const int ho = h / 2, wo = w / 2, G = C / 8;
    const long total = static_cast<long>(n) * h * w * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const long pix = i / G;
        const int iw = static_cast<int>(pix % w), ih = static_cast<int>((pix / w) % h), b = static_cast<int>(pix / (static_cast<long>(w) * h));
        bf8 xv;
        xv.load(x + pix * C + g * 8);
        float xf[8], acc[8];
        xv.unpack(xf);
        if (add) {
            bf8 av;
            av.load(add + pix * C + g * 8);
            av.unpack(acc);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        }
        for (int oh = ih / 2; oh <= (ih + 1) / 2; ++oh) {      // windows with 2*oh-1 <= ih <= 2*oh+1
            if (oh < 0 || oh >= ho) continue;
            for (int ow = iw / 2; ow <= (iw + 1) / 2; ++ow) {
                if (ow < 0 || ow >= wo) continue;
                const long opix = (static_cast<long>(b) * ho + oh) * wo + ow;
                bf8 yv, dv;
                yv.load(y + opix * C + g * 8);
                float yf[8];
                yv.unpack(yf);
                unsigned hit = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) hit |= (xf[j] == yf[j]) ? (1u << j) : 0u;
                if (!hit) continue;
                // drop channels where an earlier window element already equals the max (ATen keeps the first)
                const int rr = ih - (2 * oh - 1), ss = iw - (2 * ow - 1);
                for (int r = 0; r <= rr && hit; ++r) {
                    const int jh = 2 * oh + r - 1;
                    if (jh < 0) continue;
                    const int s_end = (r == rr) ? ss : 3;
                    for (int s = 0; s < s_end; ++s) {
                        const int jw = 2 * ow + s - 1;
                        if (jw < 0 || jw >= w) continue;
                        bf8 ev;
                        ev.load(x + ((static_cast<long>(b) * h + jh) * w + jw) * C + g * 8);
                        float ef[8];
                        ev.unpack(ef);
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (ef[j] == yf[j]) hit &= ~(1u << j);
                    }
                }
                if (!hit) continue;
                dv.load(dy + opix * C + g * 8);
                float df[8];
                dv.unpack(df);
#pragma unroll
                for (int j = 0; j < 8; ++j) if (hit & (1u << j)) acc[j] += df[j];
            }
        }
        if (relu_mask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (!(xf[j] > 0.f)) acc[j] = 0.f;
        }
        bf8 o;
        o.pack(acc);
        o.store(dx + pix * C + g * 8);
    }
}

// -------------------------------------------------------------------------------------------------
// nearest up-sampling x2 and FPN top-down add
// -------------------------------------------------------------------------------------------------
__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w,
                                      int C) {
    pdl_trigger();
    pdl_wait();            // h,w = OUTPUT size
    // one thread per INPUT vector (8 channels): loaded once, stored to its 2 x 2 output pixels; 32-bit index arithmetic
    // (the host checks the element count) -- the per-output-vector form spent its time in 64-bit divisions (3.0 TB/s)
    const unsigned G = C / 8, hi = h / 2, wi = w / 2;
    const unsigned total = static_cast<unsigned>(n) * hi * wi * G;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned g = i % G, pix = i / G;
        const unsigned iw = pix % wi, t = pix / wi, ih = t % hi, b = t / hi;
        bf8 v;
        v.load(x + static_cast<size_t>(pix) * C + g * 8);
        __nv_bfloat16* o = y + ((static_cast<size_t>(b) * h + 2 * ih) * w + 2 * iw) * C + g * 8;
        v.store(o);
        v.store(o + C);
        v.store(o + static_cast<size_t>(w) * C);
        v.store(o + static_cast<size_t>(w) * C + C);
    }
}

__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int n, int h, int w,
                                      int C) {
    pdl_trigger();
    pdl_wait();            // h,w = INPUT (small) size
    const int G = C / 8;
    const long total = static_cast<long>(n) * h * w * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const long pix = i / G;
        const int iw = static_cast<int>(pix % w), ih = static_cast<int>((pix / w) % h), b = static_cast<int>(pix / (static_cast<long>(w) * h));
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int r = 0; r < 2; ++r)
            for (int s = 0; s < 2; ++s) {
                bf8 v;
                v.load(dy + ((static_cast<long>(b) * (2 * h) + 2 * ih + r) * (2 * w) + 2 * iw + s) * C + g * 8);
                float f[8];
                v.unpack(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        bf8 o;
        o.pack(acc);
        o.store(dx + pix * C + g * 8);
    }
}

__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    const int s = static_cast<int>(floorf(static_cast<float>(dst) * scale));
    return s < in_size - 1 ? s : in_size - 1;
}

__global__ void add_nearest_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int hi, int wi,
                                       int ho, int wo, int C, float sh, float sw) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8;
    const long total = static_cast<long>(n) * ho * wo * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const long pix = i / G;
        const int ow = static_cast<int>(pix % wo), oh = static_cast<int>((pix / wo) % ho), b = static_cast<int>(pix / (static_cast<long>(wo) * ho));
        const int ih = nearest_src(oh, sh, hi), iw = nearest_src(ow, sw, wi);
        bf8 a, c;
        a.load(x + ((static_cast<long>(b) * hi + ih) * wi + iw) * C + g * 8);
        c.load(y + pix * C + g * 8);
        float fa[8], fc[8];
        a.unpack(fa);
        c.unpack(fc);
#pragma unroll
        for (int j = 0; j < 8; ++j) fc[j] += fa[j];
        c.pack(fc);
        c.store(y + pix * C + g * 8);
    }
}

__global__ void add_nearest_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int n, int hi,
                                       int wi, int ho, int wo, int C, float sh, float sw, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8;
    const long total = static_cast<long>(n) * hi * wi * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const long pix = i / G;
        const int iw = static_cast<int>(pix % wi), ih = static_cast<int>((pix / wi) % hi), b = static_cast<int>(pix / (static_cast<long>(wi) * hi));
        float acc[8];
        if (accumulate) {
            bf8 v;
            v.load(dx + pix * C + g * 8);
            v.unpack(acc);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        }
        int h_lo = static_cast<int>(ih / sh) - 2, w_lo = static_cast<int>(iw / sw) - 2;
        if (h_lo < 0) h_lo = 0;
        if (w_lo < 0) w_lo = 0;
        for (int oh = h_lo; oh < ho; ++oh) {
            const int sh_i = nearest_src(oh, sh, hi);
            if (sh_i < ih) continue;
            if (sh_i > ih) break;
            for (int ow = w_lo; ow < wo; ++ow) {
                const int sw_i = nearest_src(ow, sw, wi);
                if (sw_i < iw) continue;
                if (sw_i > iw) break;
                bf8 v;
                v.load(dy + ((static_cast<long>(b) * ho + oh) * wo + ow) * C + g * 8);
                float f[8];
                v.unpack(f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        }
        bf8 o;
        o.pack(acc);
        o.store(dx + pix * C + g * 8);
    }
}

// -------------------------------------------------------------------------------------------------
// odd feature maps (detector size 300: 75 / 19 pixels): stride-2 convolutions run on an even, zero-padded copy
// -------------------------------------------------------------------------------------------------
__global__ void pad_hw_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w, int hp,
                              int wp, int C) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8;
    const long total = static_cast<long>(n) * hp * wp * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const long pix = i / G;
        const int ow = static_cast<int>(pix % wp), oh = static_cast<int>((pix / wp) % hp), b = static_cast<int>(pix / (static_cast<long>(wp) * hp));
        bf8 v;
        if (oh < h && ow < w) v.load(x + ((static_cast<long>(b) * h + oh) * w + ow) * C + g * 8);
        else v.u = make_uint4(0, 0, 0, 0);
        v.store(y + pix * C + g * 8);
    }
}

// dx = (crop(dxp) + add) * (mask > 0)
__global__ void crop_add_mask_kernel(const __nv_bfloat16* __restrict__ dxp, const __nv_bfloat16* __restrict__ add,
                                     const __nv_bfloat16* __restrict__ mask, __nv_bfloat16* __restrict__ dx, int n, int h, int w,
                                     int hp, int wp, int C) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8;
    const long total = static_cast<long>(n) * h * w * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const long pix = i / G;
        const int ow = static_cast<int>(pix % w), oh = static_cast<int>((pix / w) % h), b = static_cast<int>(pix / (static_cast<long>(w) * h));
        bf8 v;
        v.load(dxp + ((static_cast<long>(b) * hp + oh) * wp + ow) * C + g * 8);
        float f[8];
        v.unpack(f);
        if (add) {
            bf8 a;
            a.load(add + pix * C + g * 8);
            float af[8];
            a.unpack(af);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += af[j];
        }
        if (mask) {
            bf8 m;
            m.load(mask + pix * C + g * 8);
            float mf[8];
            m.unpack(mf);
#pragma unroll
            for (int j = 0; j < 8; ++j) if (!(mf[j] > 0.f)) f[j] = 0.f;
        }
        v.pack(f);
        v.store(dx + pix * C + g * 8);
    }
}

// -------------------------------------------------------------------------------------------------
// layout converters
// -------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w, int Cs,
                                    int Cd, int accumulate) {
    pdl_trigger();
    pdl_wait();
    // thread = (pixel, 8-channel group); consecutive threads -> consecutive pixels (coalesced fp32 reads)
    const int G = Cd / 8;
    const long plane = static_cast<long>(h) * w;
    const long total = static_cast<long>(n) * plane * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long p = i % plane;
        const int g = static_cast<int>((i / plane) % G);
        const int b = static_cast<int>(i / (plane * G));
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            f[j] = c < Cs ? __ldg(x + (static_cast<long>(b) * Cs + c) * plane + p) : 0.f;
        }
        bf8 o;
        if (accumulate) {
            o.load(y + (static_cast<long>(b) * plane + p) * Cd + g * 8);
            float old[8];
            o.unpack(old);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += old[j];
        }
        o.pack(f);
        o.store(y + (static_cast<long>(b) * plane + p) * Cd + g * 8);
    }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int n, int h, int w, int Cs,
                                    int Cd) {
    pdl_trigger();
    pdl_wait();
    const int G = (Cd + 7) / 8;
    const long plane = static_cast<long>(h) * w;
    const long total = static_cast<long>(n) * plane * G;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long p = i % plane;
        const int g = static_cast<int>((i / plane) % G);
        const int b = static_cast<int>(i / (plane * G));
        bf8 v;
        v.load(x + (static_cast<long>(b) * plane + p) * Cs + g * 8);
        float f[8];
        v.unpack(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            if (c < Cd) y[(static_cast<long>(b) * Cd + c) * plane + p] = f[j];
        }
    }
}

__global__ void sigmoid_bwd_pack_kernel(const float* __restrict__ dhal, const float* __restrict__ hal,
                                        __nv_bfloat16* __restrict__ dl, int n, int h, int w, int ch, int Cd, float* dbias) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sb[4];
    if (threadIdx.x < 4) sb[threadIdx.x] = 0.f;
    __syncthreads();
    const long plane = static_cast<long>(h) * w;
    const long total = static_cast<long>(n) * plane;
    float loc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long p = i % plane;
        const int b = static_cast<int>(i / plane);
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = 0.f;
        for (int c = 0; c < ch && c < 4; ++c) {
            const long idx = (static_cast<long>(b) * ch + c) * plane + p;
            const float s = hal[idx];
            const float g = dhal[idx] * s * (1.f - s);
            f[c] = g;
            loc[c] += g;
        }
        for (int g8 = 0; g8 < 2; ++g8) {
            bf8 o;
            o.pack(f + g8 * 8);
            o.store(dl + i * Cd + g8 * 8);
        }
    }
    if (dbias) {
        for (int c = 0; c < ch && c < 4; ++c) {
            float v = loc[c];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sb[c], v);
        }
        __syncthreads();
        if (threadIdx.x < ch && threadIdx.x < 4) atomicAdd(dbias + threadIdx.x, sb[threadIdx.x]);
    }
}

// -------------------------------------------------------------------------------------------------
// detector input transform (fp32 NCHW)
// -------------------------------------------------------------------------------------------------
// grid (ceil(wo / 1024), ceil(ho / 4), n * c): four output pixels per thread (one 16-byte store when the row allows), four
// rows per block; row / plane are uniform per block, no per-element divisions (the flat grid-stride form spent its time in
// 64-bit index arithmetic: 59 us for 39 MB)
__global__ void __launch_bounds__(kEwThreads) resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int c, int hi, int wi,
                                                                int ho, int wo, float sh, float sw, const float* mean, const float* stdv) {
    pdl_trigger();
    pdl_wait();
    const int ow0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, bc = blockIdx.z;
    if (ow0 >= wo) return;
    const int ch = bc % c;
    const float m = mean ? mean[ch] : 0.f, sd = mean ? stdv[ch] : 1.f;
    int iw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) iw[j] = nearest_src(min(ow0 + j, wo - 1), sw, wi);
    for (int r = 0; r < 4; ++r) {
        const int oh = blockIdx.y * 4 + r;
        if (oh >= ho) break;
        const float* src = x + (static_cast<size_t>(bc) * hi + nearest_src(oh, sh, hi)) * wi;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = __ldg(src + iw[j]);
            if (mean) v[j] = (v[j] - m) / sd;
        }
        float* dst = y + (static_cast<size_t>(bc) * ho + oh) * wo + ow0;
        if (ow0 + 3 < wo && (wo & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        else
            for (int j = 0; j < 4 && ow0 + j < wo; ++j) dst[j] = v[j];
    }
}

// grid (ceil(wi / 1024), ceil(hi / 4), n * c): gather-sum of the output pixels whose nearest source is (ih, iw); four input
// pixels per thread, four rows per block (same summation order as the scalar form: output rows, then columns)
__global__ void __launch_bounds__(kEwThreads) resize_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int n, int c, int hi, int wi,
                                                                int ho, int wo, float sh, float sw, const float* stdv, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int iw0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, bc = blockIdx.z;
    if (iw0 >= wi) return;
    const int ch = bc % c;
    int w0[4], w1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int iw = iw0 + j;
        int a = static_cast<int>(iw / sw) - 2;
        if (a < 0) a = 0;
        while (a < wo && nearest_src(a, sw, wi) < iw) ++a;
        int b = a;
        while (b < wo && nearest_src(b, sw, wi) == iw) ++b;
        w0[j] = a; w1[j] = iw < wi ? b : a;
    }
    const float* plane = dy + static_cast<size_t>(bc) * ho * wo;
    for (int r = 0; r < 4; ++r) {
        const int ih = blockIdx.y * 4 + r;
        if (ih >= hi) break;
        int h0 = static_cast<int>(ih / sh) - 2;
        if (h0 < 0) h0 = 0;
        while (h0 < ho && nearest_src(h0, sh, hi) < ih) ++h0;
        int h1 = h0;
        while (h1 < ho && nearest_src(h1, sh, hi) == ih) ++h1;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int oh = h0; oh < h1; ++oh) {
            const float* row = plane + static_cast<size_t>(oh) * wo;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                for (int ow = w0[j]; ow < w1[j]; ++ow) acc[j] += __ldg(row + ow);
        }
        float* o = dx + (static_cast<size_t>(bc) * hi + ih) * wi + iw0;
        for (int j = 0; j < 4 && iw0 + j < wi; ++j) {
            float a = acc[j];
            if (stdv) a /= stdv[ch];
            o[j] = accumulate ? o[j] + a : a;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// pixel regulariser (fused loss + gradient): warp-shuffle reduction, one atomic per block
// -------------------------------------------------------------------------------------------------
__global__ void regulariser_kernel(int kind, const float* __restrict__ hal, const float* __restrict__ rgb,
                                   const float* __restrict__ ir, float w_rgb, float w_ir, int n, long plane, float* loss,
                                   float* dhal, float grad_scale, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const long total = static_cast<long>(n) * 3 * plane;
    const float inv_n = 1.f / static_cast<float>(total);
    float l_rgb = 0.f, l_ir = 0.f;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long p = i % plane;
        const int b = static_cast<int>(i / (3 * plane));
        const float hv = hal[i];
        float g = 0.f;
        if (rgb) {
            const float d = hv - __ldg(rgb + i);
            if (kind == 0) { l_rgb += d * d; g += w_rgb * 2.f * d * inv_n; }
            else { l_rgb += fabsf(d); g += w_rgb * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n; }
        }
        if (ir) {
            const float d = hv - __ldg(ir + static_cast<long>(b) * plane + p);
            if (kind == 0) { l_ir += d * d; g += w_ir * 2.f * d * inv_n; }
            else { l_ir += fabsf(d); g += w_ir * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n; }
        }
        if (dhal) dhal[i] = accumulate ? dhal[i] + g * grad_scale : g * grad_scale;
    }
    __shared__ float s[2];
    if (threadIdx.x < 2) s[threadIdx.x] = 0.f;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
        l_rgb += __shfl_xor_sync(0xffffffffu, l_rgb, o);
        l_ir += __shfl_xor_sync(0xffffffffu, l_ir, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s[0], l_rgb); atomicAdd(&s[1], l_ir); }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(loss, s[0] * inv_n * w_rgb);
        atomicAdd(loss + 1, s[1] * inv_n * w_ir);
    }
}

}  // namespace hd

using namespace hd;

#define HD_LAUNCH_OK() HD_CUDA_OK(cudaPeekAtLastError())

extern "C" int hd_pack_conv_weight(const float* w, const float* scale, int cout, int cin, int kh, int kw, void* w_fwd,
                                   int cout_pad, int k_pad, void* w_dgrad, int cin_pad, void* w_t, hd_stream st) {
    HD_CHECK_ARG(w != nullptr && cout > 0 && cin > 0 && kh > 0 && kw > 0);
    HD_CHECK_ARG(cout_pad >= cout && k_pad >= kh * kw * cin && cin_pad >= cin);
    const long work = static_cast<long>(cout_pad) * k_pad + (w_dgrad ? static_cast<long>(cin_pad) * kh * kw * cout : 0);
    HD_CUDA_OK(hd::launch(pack_weight_kernel, dim3(ew_blocks(work)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), w, scale, cout, cin, kh, kw, static_cast<__nv_bfloat16*>(w_fwd), cout_pad, k_pad,
        static_cast<__nv_bfloat16*>(w_dgrad), cin_pad, static_cast<__nv_bfloat16*>(w_t)));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_unpack_wgrad(const float* dw, float* g, int cout, int cin, int kh, int kw, int tap_stride, int row_stride,
                               float scale, hd_stream st) {
    HD_CHECK_ARG(dw && g && cout > 0 && cin > 0);
    HD_CUDA_OK(hd::launch(unpack_wgrad_kernel, dim3(ew_blocks(static_cast<long>(cout) * cin * kh * kw)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), dw, g, cout, cin, kh * kw, tap_stride, row_stride, scale));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_multi_blocks(int64_t elements) {
    // blocks a layer with `elements` work items occupies in hd_pack_conv_weights / hd_unpack_wgrads (for first_block)
    return static_cast<int>((elements + static_cast<int64_t>(kEwThreads) * kMultiItems - 1) / (static_cast<int64_t>(kEwThreads) * kMultiItems));
}

extern "C" int hd_pack_blocks(const hd_pack_desc* d) {
    // blocks layer `d` (a HOST copy of its descriptor; first_block ignored) occupies in hd_pack_conv_weights
    if (d == nullptr) return 0;
    const int taps = d->kh * d->kw;
    if (pack_tiled_ok(*d, taps)) {
        const int tci = d->cin % 64 == 0 ? 64 : (d->cin % 32 == 0 ? 32 : 16);
        return (d->cout / kPackCo) * (d->cin / tci);                    // one block per (16 cout, tci cin) tile
    }
    const int64_t work = static_cast<int64_t>(d->cout_pad) * d->k_pad + (d->w_dgrad ? static_cast<int64_t>(d->cin_pad) * taps * d->cout : 0);
    return hd_multi_blocks(work);
}

extern "C" int hd_adam_pack_conv_weights(const hd_pack_desc* descs_dev, int n_layers, int total_blocks, const hd_adam_args* adam, hd_stream st) {
    HD_CHECK_ARG(descs_dev != nullptr && adam != nullptr && n_layers > 0 && total_blocks > 0);
    HD_CHECK_ARG(adam->bias_correction1 > 0.f && adam->bias_correction2 > 0.f);
    HD_CUDA_OK(hd::launch(pack_weights_multi_kernel<true>, dim3(total_blocks), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), descs_dev, n_layers, total_blocks, *adam));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_adam_multi(const hd_adam_desc* descs_dev, int n_tensors, int total_blocks, const hd_adam_args* adam, hd_stream st) {
    HD_CHECK_ARG(descs_dev != nullptr && adam != nullptr && n_tensors > 0 && total_blocks > 0);
    HD_CHECK_ARG(adam->bias_correction1 > 0.f && adam->bias_correction2 > 0.f);
    HD_CUDA_OK(hd::launch(adam_multi_kernel, dim3(total_blocks), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), descs_dev, n_tensors, *adam));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_pack_conv_weights(const hd_pack_desc* descs_dev, int n_layers, int total_blocks, hd_stream st) {
    HD_CHECK_ARG(descs_dev && n_layers > 0 && total_blocks > 0);
    HD_CUDA_OK(hd::launch(pack_weights_multi_kernel<false>, dim3(total_blocks), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), descs_dev, n_layers, total_blocks, hd_adam_args{}));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_unpack_blocks(int cout, int cin, int taps) {
    // blocks one layer occupies in hd_unpack_wgrads (for hd_unpack_desc.first_block): whole output-channel rows per block
    if (cout <= 0 || cin <= 0 || taps <= 0) return HD_ERR_BAD_ARG;
    const int row = cin * taps;
    const int rows_per_block = row >= kMultiItems * kEwThreads ? 1 : (kMultiItems * kEwThreads) / row;
    return (cout + rows_per_block - 1) / rows_per_block;
}

extern "C" int hd_unpack_wgrads(const hd_unpack_desc* descs_dev, int n_layers, int total_blocks, hd_stream st) {
    HD_CHECK_ARG(descs_dev && n_layers > 0 && total_blocks > 0);
    HD_CUDA_OK(hd::launch(unpack_wgrads_multi_kernel, dim3(total_blocks), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), descs_dev, n_layers));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_stem_im2col(const float* x, void* patches, int n, int h, int w, int k_pad, hd_stream st) {
    HD_CHECK_ARG(x && patches && h % 2 == 0 && w % 2 == 0 && k_pad >= 152 && k_pad % 8 == 0);
    const int ho = h / 2, wo = w / 2;
    const int tiles = n * ((ho + kI2cTH - 1) / kI2cTH) * ((wo + kI2cTW - 1) / kI2cTW);
    HD_CUDA_OK(hd::launch(stem_im2col_kernel, dim3(tiles), dim3(256), 0, static_cast<cudaStream_t>(st), x, static_cast<__nv_bfloat16*>(patches), n, h, w, k_pad));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_stem_im2col_1ch(const void* x, int x_dtype, float scale, void* patches, int n, int h, int w, int k_pad, hd_stream st) {
    HD_CHECK_ARG(x != nullptr && patches != nullptr && n > 0 && h % 2 == 0 && w % 2 == 0 && k_pad % 8 == 0 && k_pad >= 56);
    HD_CHECK_ARG(x_dtype == 0 || x_dtype == 1);
    const int tiles = n * ((h / 2 + kI2cTH - 1) / kI2cTH) * ((w / 2 + kI2cTW - 1) / kI2cTW);
    if (x_dtype == 0)
        HD_CUDA_OK(hd::launch(stem_im2col1_kernel<float>, dim3(tiles), dim3(256), 0, static_cast<cudaStream_t>(st), static_cast<const float*>(x), scale,
                              static_cast<__nv_bfloat16*>(patches), n, h, w, k_pad));
    else
        HD_CUDA_OK(hd::launch(stem_im2col1_kernel<unsigned char>, dim3(tiles), dim3(256), 0, static_cast<cudaStream_t>(st), static_cast<const unsigned char*>(x),
                              scale, static_cast<__nv_bfloat16*>(patches), n, h, w, k_pad));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_stem_col2im(const void* dp, float* dx, int n, int h, int w, int k_pad, hd_stream st) {
    HD_CHECK_ARG(dp && dx && h % 2 == 0 && w % 2 == 0);
    HD_CUDA_OK(hd::launch(stem_col2im_kernel, dim3(ew_blocks(static_cast<long>(n) * h * w)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(dp), dx, n, h, w, k_pad));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_bn_finalize(const float* stats, int reps, int C, double count, const float* gamma, const float* beta,
                              float eps, float momentum, float* rm, float* rv, float* mean_out, float* invstd_out,
                              float* scale_out, float* shift_out, hd_stream st) {
    HD_CHECK_ARG(stats && gamma && beta && scale_out && shift_out && C > 0 && reps > 0 && count > 0);
    HD_CUDA_OK(hd::launch(bn_finalize_kernel, dim3(C), dim3(128), 0, static_cast<cudaStream_t>(st), stats, reps, C, count, gamma, beta, eps,
                                                                                 momentum, rm, rv, mean_out, invstd_out,
                                                                                 scale_out, shift_out));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_bn_apply(const void* z, const float* scale, const float* shift, const void* res, const float* rscale,
                           const float* rshift, int relu, void* y, int64_t n_pix, int C, hd_stream st) {
    HD_CHECK_ARG(z && scale && shift && y && C % 8 == 0 && n_pix > 0 && kEwThreads % (C / 8) == 0);
    HD_CUDA_OK(hd::launch(bn_apply_kernel, dim3(ew_blocks(n_pix * (C / 8))), dim3(kEwThreads), 4 * C * sizeof(float), static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(z), scale, shift, static_cast<const __nv_bfloat16*>(res), rscale, rshift, relu,
        static_cast<__nv_bfloat16*>(y), n_pix, C));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_bn_bwd_reduce(const void* dy, const void* yrelu, const float* rscale, const float* rshift, const void* z,
                                const float* mean, const float* invstd, float* sums, int64_t n_pix, int C, hd_stream st) {
    HD_CHECK_ARG(dy && z && mean && invstd && sums && C % 8 == 0 && C / 8 <= 256 && n_pix > 0);
    const int G = C / 8;
    int L = 256 / G;
    if (L < 1) L = 1;
    const int threads = G * L;
    long blocks = 148L * 4;
    if (blocks > n_pix / 32) blocks = n_pix / 32 > 0 ? n_pix / 32 : 1;     // small maps: at least 32 pixels per block (2C atomics each)
    long ppb = (n_pix + blocks - 1) / blocks;
    if (ppb < L) ppb = L;
    blocks = (n_pix + ppb - 1) / ppb;
    HD_CUDA_OK(hd::launch(bn_bwd_reduce_kernel, dim3(static_cast<int>(blocks)), dim3(threads), 6 * C * sizeof(float), static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(yrelu), rscale, rshift,
        static_cast<const __nv_bfloat16*>(z), mean, invstd, sums, n_pix, C, ppb));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_bn_bwd_apply(const void* dy, const void* yrelu, const float* rscale, const float* rshift, const void* z,
                               const float* mean, const float* invstd, const float* gamma, const float* sums, double count,
                               void* dz, void* gout, float* dgamma, float* dbeta, int64_t n_pix, int C, hd_stream st) {
    HD_CHECK_ARG(dy && z && mean && invstd && gamma && sums && dz && C % 8 == 0 && n_pix > 0 && count > 0 && kEwThreads % (C / 8) == 0);
    HD_CUDA_OK(hd::launch(bn_bwd_apply_kernel, dim3(ew_blocks(n_pix * (C / 8))), dim3(kEwThreads), 5 * C * sizeof(float), static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(yrelu), rscale, rshift,
        static_cast<const __nv_bfloat16*>(z), mean, invstd, gamma, sums, static_cast<float>(1.0 / count),
        static_cast<__nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(gout), dgamma, dbeta, n_pix, C));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_bn_bwd_fused(const void* dy, const void* yrelu, const float* rscale, const float* rshift, const void* z,
                               const float* mean, const float* invstd, const float* gamma, float* sums, double count, void* dz,
                               void* g_out, float* dgamma, float* dbeta, int64_t n_pix, int C, uint32_t* barrier_words, hd_stream st) {
    HD_CHECK_ARG(dy != nullptr && z != nullptr && dz != nullptr && mean != nullptr && invstd != nullptr && gamma != nullptr &&
                 sums != nullptr && barrier_words != nullptr);
    HD_CHECK_ARG(C % 8 == 0 && kBnFusedThreads % (C / 8) == 0 && n_pix > 0);
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    const int L = kBnFusedThreads / (C / 8);
    long blocks = (n_pix + L - 1) / L;                       // at least one step of every pixel lane per CTA
    if (blocks > sms) blocks = sms;
    const long per = (n_pix + blocks - 1) / blocks;
    blocks = (n_pix + per - 1) / per;
    const long iters = (per + L - 1) / L;
    const size_t base = 2 * static_cast<size_t>(C) * sizeof(float);
    const size_t cache_bytes = static_cast<size_t>(iters) * kBnFusedThreads * 16 * 2;
    const size_t limit = 200 * 1024;
    const int cache = base + cache_bytes <= limit ? 1 : 0;
    const size_t smem = base + (cache ? cache_bytes : 0);
    static SmemAttrOnce smem_attr;
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, bn_bwd_fused_kernel, static_cast<int>(limit)));
    HD_CUDA_OK(hd::launch(bn_bwd_fused_kernel, dim3(static_cast<int>(blocks)), dim3(kBnFusedThreads), smem, static_cast<cudaStream_t>(st),
                          static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(yrelu), rscale, rshift,
                          static_cast<const __nv_bfloat16*>(z), mean, invstd, gamma, sums, static_cast<float>(1.0 / count),
                          static_cast<__nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(g_out), dgamma, dbeta, static_cast<long>(n_pix), C,
                          per, cache, barrier_words));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

extern "C" int hd_maxpool_fwd(const hd_act* x, const hd_act* y, void* idx, int mask_nonpositive, hd_stream st) {
    HD_CHECK_ARG(x && y && x->ptr && y->ptr && x->c % 8 == 0 && x->c == y->c && x->h % 2 == 0 && x->w % 2 == 0);
    HD_CHECK_ARG(y->h == x->h / 2 && y->w == x->w / 2 && y->n == x->n);
    HD_CHECK_ARG(idx == nullptr || (reinterpret_cast<uintptr_t>(idx) & 7) == 0);
    HD_CHECK_ARG(static_cast<long>(x->n) * x->h * x->w * (x->c / 8) < (1L << 31));         // 32-bit index arithmetic in the kernel
    HD_CUDA_OK(hd::launch(maxpool_fwd_kernel, dim3(ew_blocks(static_cast<long>(y->n) * y->h * y->w * (y->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(x->ptr),
                                                          static_cast<__nv_bfloat16*>(y->ptr), static_cast<unsigned char*>(idx),
                                                          mask_nonpositive, x->n, x->h, x->w, x->c));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_maxpool_bwd(const hd_act* x, const hd_act* y, const void* dy, const void* add, void* dx, int relu_mask,
                              const void* idx, hd_stream st) {
    HD_CHECK_ARG(x && y && dy && dx && x->c % 8 == 0 && x->c == y->c);
    HD_CHECK_ARG(y->h == x->h / 2 && y->w == x->w / 2 && y->n == x->n);
    if (idx != nullptr) {
        // arg-max positions stored by hd_maxpool_fwd (with mask_nonpositive when relu_mask semantics are wanted): x / y are
        // not read at all
        HD_CHECK_ARG(static_cast<long>(x->n) * x->h * x->w * (x->c / 8) < (1L << 31));     // 32-bit index arithmetic in the kernel
        HD_CUDA_OK(hd::launch(maxpool_bwd_idx_kernel, dim3(ew_blocks(static_cast<long>(x->n) * x->h * x->w * (x->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const unsigned char*>(idx), static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(add),
            static_cast<__nv_bfloat16*>(dx), x->n, x->h, x->w, x->c));
        HD_LAUNCH_OK();
        return HD_OK;
    }
    HD_CHECK_ARG(x->ptr && y->ptr);
    HD_CUDA_OK(hd::launch(maxpool_bwd_kernel, dim3(ew_blocks(static_cast<long>(x->n) * x->h * x->w * (x->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(x->ptr), static_cast<const __nv_bfloat16*>(y->ptr),
        static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dx),
        x->n, x->h, x->w, x->c, relu_mask));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_upsample2x_fwd(const hd_act* x, const hd_act* y, hd_stream st) {
    HD_CHECK_ARG(x && y && x->ptr && y->ptr && x->c == y->c && x->c % 8 == 0 && y->h == 2 * x->h && y->w == 2 * x->w && x->n == y->n);
    HD_CHECK_ARG(static_cast<long>(y->n) * y->h * y->w * (y->c / 8) < (1L << 31));        // 32-bit index arithmetic in the kernel
    HD_CUDA_OK(hd::launch(upsample2x_fwd_kernel, dim3(ew_blocks(static_cast<long>(x->n) * x->h * x->w * (x->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(x->ptr),
                                                             static_cast<__nv_bfloat16*>(y->ptr), y->n, y->h, y->w, y->c));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_upsample2x_bwd(const hd_act* dy, const hd_act* dx, hd_stream st) {
    HD_CHECK_ARG(dy && dx && dy->ptr && dx->ptr && dx->c == dy->c && dx->c % 8 == 0 && dy->h == 2 * dx->h && dy->w == 2 * dx->w && dx->n == dy->n);
    HD_CUDA_OK(hd::launch(upsample2x_bwd_kernel, dim3(ew_blocks(static_cast<long>(dx->n) * dx->h * dx->w * (dx->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(dy->ptr),
                                                             static_cast<__nv_bfloat16*>(dx->ptr), dx->n, dx->h, dx->w, dx->c));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_add_nearest_fwd(const hd_act* x, const hd_act* y, hd_stream st) {
    HD_CHECK_ARG(x && y && x->ptr && y->ptr && x->c == y->c && x->c % 8 == 0 && x->n == y->n);
    const float sh = static_cast<float>(x->h) / static_cast<float>(y->h), sw = static_cast<float>(x->w) / static_cast<float>(y->w);
    HD_CUDA_OK(hd::launch(add_nearest_fwd_kernel, dim3(ew_blocks(static_cast<long>(y->n) * y->h * y->w * (y->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(x->ptr),
                                                              static_cast<__nv_bfloat16*>(y->ptr), y->n, x->h, x->w, y->h,
                                                              y->w, y->c, sh, sw));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_add_nearest_bwd(const hd_act* dy, const hd_act* dx, int accumulate, hd_stream st) {
    HD_CHECK_ARG(dy && dx && dy->ptr && dx->ptr && dx->c == dy->c && dx->c % 8 == 0 && dx->n == dy->n);
    const float sh = static_cast<float>(dx->h) / static_cast<float>(dy->h), sw = static_cast<float>(dx->w) / static_cast<float>(dy->w);
    HD_CUDA_OK(hd::launch(add_nearest_bwd_kernel, dim3(ew_blocks(static_cast<long>(dx->n) * dx->h * dx->w * (dx->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(dy->ptr),
                                                              static_cast<__nv_bfloat16*>(dx->ptr), dx->n, dx->h, dx->w,
                                                              dy->h, dy->w, dx->c, sh, sw, accumulate));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_pad_hw(const hd_act* x, const hd_act* y, hd_stream st) {
    HD_CHECK_ARG(x && y && x->ptr && y->ptr && x->c == y->c && x->c % 8 == 0 && x->n == y->n && y->h >= x->h && y->w >= x->w);
    HD_CUDA_OK(hd::launch(pad_hw_kernel, dim3(ew_blocks(static_cast<long>(y->n) * y->h * y->w * (y->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(x->ptr), static_cast<__nv_bfloat16*>(y->ptr), x->n, x->h, x->w, y->h, y->w, x->c));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_crop_add_mask(const hd_act* dxp, const void* add, const void* mask, const hd_act* dx, hd_stream st) {
    HD_CHECK_ARG(dxp && dx && dxp->ptr && dx->ptr && dx->c == dxp->c && dx->c % 8 == 0 && dx->n == dxp->n && dxp->h >= dx->h && dxp->w >= dx->w);
    HD_CUDA_OK(hd::launch(crop_add_mask_kernel, dim3(ew_blocks(static_cast<long>(dx->n) * dx->h * dx->w * (dx->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(dxp->ptr), static_cast<const __nv_bfloat16*>(add), static_cast<const __nv_bfloat16*>(mask),
        static_cast<__nv_bfloat16*>(dx->ptr), dx->n, dx->h, dx->w, dxp->h, dxp->w, dx->c));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_nchw_f32_to_nhwc_bf16(const float* x, const hd_act* y, int channels, int accumulate, hd_stream st) {
    HD_CHECK_ARG(x && y && y->ptr && y->c % 8 == 0 && channels <= y->c && channels > 0);
    HD_CUDA_OK(hd::launch(nchw_to_nhwc_kernel, dim3(ew_blocks(static_cast<long>(y->n) * y->h * y->w * (y->c / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), x, static_cast<__nv_bfloat16*>(y->ptr), y->n, y->h, y->w, channels, y->c, accumulate));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_nhwc_bf16_to_nchw_f32(const hd_act* x, float* y, int channels, hd_stream st) {
    HD_CHECK_ARG(x && y && x->ptr && x->c % 8 == 0 && channels <= x->c && channels > 0);
    HD_CUDA_OK(hd::launch(nhwc_to_nchw_kernel, dim3(ew_blocks(static_cast<long>(x->n) * x->h * x->w * ((channels + 7) / 8))), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), static_cast<const __nv_bfloat16*>(x->ptr), y, x->n, x->h, x->w, x->c, channels));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_sigmoid_bwd_pack(const float* dhal, const float* hal, const hd_act* dl, int channels, float* dbias, hd_stream st) {
    HD_CHECK_ARG(dhal && hal && dl && dl->ptr && dl->c == 16 && channels >= 1 && channels <= 4);
    HD_CUDA_OK(hd::launch(sigmoid_bwd_pack_kernel, dim3(ew_blocks(static_cast<long>(dl->n) * dl->h * dl->w)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), dhal, hal, static_cast<__nv_bfloat16*>(dl->ptr), dl->n, dl->h, dl->w, channels, dl->c, dbias));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_resize_nearest_fwd(const float* x, float* y, int n, int c, int hi, int wi, int ho, int wo, const float* mean,
                                     const float* stdv, hd_stream st) {
    HD_CHECK_ARG(x && y && n > 0 && c > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && ((mean == nullptr) == (stdv == nullptr)));
    const float sh = static_cast<float>(hi) / static_cast<float>(ho), sw = static_cast<float>(wi) / static_cast<float>(wo);
    HD_CHECK_ARG(ho <= 65535 && n * c <= 65535);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    HD_CUDA_OK(hd::launch(resize_fwd_kernel, dim3((wo + 4 * kEwThreads - 1) / (4 * kEwThreads), (ho + 3) / 4, n * c), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), x, y, n, c, hi, wi, ho, wo, sh, sw, mean, stdv));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_resize_nearest_bwd(const float* dy, float* dx, int n, int c, int hi, int wi, int ho, int wo, const float* stdv,
                                     int accumulate, hd_stream st) {
    HD_CHECK_ARG(dy && dx && n > 0 && c > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0);
    const float sh = static_cast<float>(hi) / static_cast<float>(ho), sw = static_cast<float>(wi) / static_cast<float>(wo);
    HD_CHECK_ARG(hi <= 65535 && n * c <= 65535);
    HD_CUDA_OK(hd::launch(resize_bwd_kernel, dim3((wi + 4 * kEwThreads - 1) / (4 * kEwThreads), (hi + 3) / 4, n * c), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), dy, dx, n, c, hi, wi, ho, wo, sh, sw, stdv, accumulate));
    HD_LAUNCH_OK();
    return HD_OK;
}

extern "C" int hd_regulariser(int kind, const float* hal, const float* rgb, const float* ir, float w_rgb, float w_ir, int n,
                              int h, int w, float* loss, float* dhal, float grad_scale, int accumulate, hd_stream st) {
    HD_CHECK_ARG((kind == 0 || kind == 1) && hal && loss && n > 0 && h > 0 && w > 0);
    const long plane = static_cast<long>(h) * w;
    long blocks = (static_cast<long>(n) * 3 * plane + kEwThreads * 8 - 1) / (kEwThreads * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    HD_CUDA_OK(hd::launch(regulariser_kernel, dim3(static_cast<int>(blocks)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(st), kind, hal, rgb, ir, w_rgb, w_ir, n, plane, loss, dhal, grad_scale, accumulate));
    HD_LAUNCH_OK();
    return HD_OK;
}
