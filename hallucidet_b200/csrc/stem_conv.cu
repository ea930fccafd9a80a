// stem_conv.cu -- the 7x7 stride-2 pad-3 stem convolutions (3 or 1 input channels -> 64) as fused halo-patch kernels.
//
// Reference operators: encoder.conv1 of the hallucination U-Net (src/segmentation_models/encoders/resnet.py:50, input = the IR
// plane replicated x3 by src/utils/utils.py:52-53) and body.conv1 of the detector backbone (TV models/resnet.py:197), both
// followed by BatchNorm + ReLU.  Cin = 3 is not TMA / tcgen05 addressable (6-byte pixels), so round 1 went through an explicit
// bf16 patch matrix [pixels][160] in HBM (im2col 0.17 ms + GEMM 0.12 ms + col2im 0.17 ms per direction and network).
// Here a CTA stages the (2*8+5) x (2*32+5) input patch of an 8 x 32 output tile ONCE in shared memory (fp32 / uint8 -> bf16,
// pixel-interleaved so that the 7*Cin values of a filter row are one contiguous run), builds the mma.sync A fragments from
// 32-bit shared loads of that run (pixel stride 2*Cin elements = always 4-byte aligned), and keeps the 64 x K' weight
// matrix in shared memory for ldmatrix.  K' pads every filter row to a multiple of 8/32 values whose weights are zero, so
// the "extra" A values (whatever follows in the halo row) never matter:
//     Cin = 3:  K' = 7 rows x 32 (21 used)  = 224 -> 14 k16 steps      Cin = 1:  K' = 8 rows x 8 (7 rows x 7 used) = 64 -> 4 k16 steps
// The work is HBM-bound on the 64-channel output (105 MB for 8 x 320 x 320): tensor throughput of the legacy mma.sync path
// is ample (45 us of HMMA issue at Cin = 3).
//   forward:  optional per-channel bias (folded BatchNorm shift) + ReLU, bf16 NHWC store with full 128-byte pixel rows,
//             optional BatchNorm statistics (one deterministic row per CTA) + fused finalize (bn_tail.cuh).
//   dgrad:    (backbone only) d(input) from d(output): four output parity classes, each a small stride-1 convolution of the
//             64-channel gradient with 9 / 12 / 12 / 16 of the 49 taps; N = 3 channels padded to one n8 tile.
#include <string.h>

#include "hd_common.cuh"
#include "bn_tail.cuh"

namespace hd {

namespace {

constexpr int STH = 8, STW = 32;                       // output tile
constexpr int SPH = 2 * STH + 5 + 1;                   // halo rows (+1: the zero-weight 8th filter row of the 1-channel layout)
constexpr int SPW = 2 * STW + 5;                       // halo columns
constexpr int kStemThreads = 256;
constexpr int kStemWarps = 8;

struct StemParams {
    const void* x;               // [N][CIN][H][W] fp32 or uint8
    int x_u8;
    float x_scale;
    const float* w;              // fp32 OIHW master [64][3][7][7] (summed over the input channel when CIN == 1)
    const float* w_scale;        // optional per-cout scale (folded BatchNorm)
    const float* bias;           // optional per-cout bias
    int relu;
    __nv_bfloat16* y;            // [N][Ho][Wo][64]
    float* stats;                // [rows][2][64] or nullptr
    int stats_rows;
    BnFin fin;
    int N, H, W, Ho, Wo;
    int tiles_w, tiles_h, total_tiles;
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int CIN>
struct StemLayout {
    static constexpr int KROW = CIN == 3 ? 32 : 8;                 // padded K' values per filter row
    static constexpr int KP = CIN == 3 ? 224 : 64;                 // K'
    static constexpr int KSTEPS = KP / 16;
    static constexpr int HPITCH = SPW * CIN + (CIN == 3 ? 1 : 3);  // halo row pitch in elements (even; 208 / 72)
    static constexpr int HALO_ELEMS = SPH * HPITCH + 64;           // + slack for the reads past the last run (zero weights)
    static constexpr int WPITCH = KP + 8;                          // weight row pitch in elements (odd multiple of 16 bytes)
    static constexpr int OPITCH = 64 + 8;                          // output staging pitch in elements (144 bytes)
    static constexpr int SMEM = HALO_ELEMS * 2 + 64 * WPITCH * 2 + kStemWarps * 32 * OPITCH * 2 + 64 * 4 + kStemWarps * 128 * 4 + 64;
};

// All global loads of the patch are issued before the first shared-memory store (fully unrolled: 18 / 6 independent loads per
// thread in flight); one load -> convert -> store per iteration left the ~1 us global latency exposed 18 times per tile.
template <int CIN, typename T>
__device__ __forceinline__ void load_halo_tile(__nv_bfloat16* halo, const T* x, float scale, int b, int ih0, int iw0, int H, int W) {
    using L = StemLayout<CIN>;
    constexpr int TOTAL = CIN * SPH * SPW, NL = (TOTAL + kStemThreads - 1) / kStemThreads;
    float v[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int q = static_cast<int>(threadIdx.x) + i * kStemThreads;
        const int col = q % SPW, t3 = q / SPW, row = t3 % SPH, c = t3 / SPH;
        const int ih = ih0 + row, iw = iw0 + col;
        v[i] = 0.f;
        if (q < TOTAL && ih >= 0 && ih < H && iw >= 0 && iw < W) v[i] = static_cast<float>(__ldg(x + ((static_cast<long>(b) * CIN + c) * H + ih) * W + iw));
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int q = static_cast<int>(threadIdx.x) + i * kStemThreads;
        const int col = q % SPW, t3 = q / SPW, row = t3 % SPH, c = t3 / SPH;
        if (q < TOTAL) halo[row * L::HPITCH + col * CIN + c] = __float2bfloat16(v[i] * scale);
    }
}

template <int CIN>
__global__ void __launch_bounds__(kStemThreads, 2) stem7_fwd_kernel(const StemParams P) {
    using L = StemLayout<CIN>;
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t ssm[];
    __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(ssm);
    __nv_bfloat16* wsm = halo + L::HALO_ELEMS;
    __nv_bfloat16* ostage = wsm + 64 * L::WPITCH;
    float* bias_s = reinterpret_cast<float*>(ostage + kStemWarps * 32 * L::OPITCH);
    float* red = bias_s + 64;                                       // [warps][64][2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // ---- weights: fp32 OIHW master -> bf16 [64][K'] (zero in the padding of every filter row); constant during the launch's
    // lifetime but written by the optimizer / packer before it, so they are read after the dependency wait
    for (int i = threadIdx.x; i < 64 * L::WPITCH; i += kStemThreads) wsm[i] = __float2bfloat16(0.f);
    for (int i = threadIdx.x; i < L::HALO_ELEMS; i += kStemThreads) halo[i] = __float2bfloat16(0.f);
    pdl_wait();
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 49 * (CIN == 3 ? 3 : 1); i += kStemThreads) {
        int o, c, r, s;
        if (CIN == 3) { s = i % 7; r = (i / 7) % 7; c = (i / 49) % 3; o = i / 147; }
        else { s = i % 7; r = (i / 7) % 7; c = 0; o = i / 49; }
        float v;
        if (CIN == 3) v = P.w[i];
        else v = P.w[(o * 3 + 0) * 49 + r * 7 + s] + P.w[(o * 3 + 1) * 49 + r * 7 + s] + P.w[(o * 3 + 2) * 49 + r * 7 + s];
        if (P.w_scale) v *= P.w_scale[o];
        wsm[o * L::WPITCH + r * L::KROW + s * CIN + c] = __float2bfloat16(v);
    }
    if (threadIdx.x < 64) bias_s[threadIdx.x] = P.bias ? P.bias[threadIdx.x] : 0.f;
    float st_s[2] = {0.f, 0.f}, st_q[2] = {0.f, 0.f};                 // this lane's channels 2*lane, 2*lane+1 over all tiles of the warp
    const uint32_t halo_u = smem_u32(halo), wsm_u = smem_u32(wsm);
    __nv_bfloat16* my_stage = ostage + warp * 32 * L::OPITCH;
    const uint32_t stage_u = smem_u32(my_stage);
    // ldmatrix lane address inside a (2 n-tiles x 16 k) block of the weight matrix: matrices (n0, k0-7), (n0, k8-15), (n0+8, k0-7), (n0+8, k8-15)
    const uint32_t b_lane = static_cast<uint32_t>(((lane & 7) + ((lane >> 4) << 3)) * L::WPITCH + ((lane >> 3) & 1) * 8) * 2u;

    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int b = tile / (P.tiles_w * P.tiles_h), t2 = tile - b * P.tiles_w * P.tiles_h;
        const int oh0 = (t2 / P.tiles_w) * STH, ow0 = (t2 % P.tiles_w) * STW;
        __syncthreads();                                              // previous tile's halo fully consumed (and weights visible)
        if (P.x_u8) load_halo_tile<CIN, unsigned char>(halo, static_cast<const unsigned char*>(P.x), P.x_scale, b, 2 * oh0 - 3, 2 * ow0 - 3, P.H, P.W);
        else load_halo_tile<CIN, float>(halo, static_cast<const float*>(P.x), P.x_scale, b, 2 * oh0 - 3, 2 * ow0 - 3, P.H, P.W);
        __syncthreads();
        // ---- warp = output row `warp` of the tile, 32 pixels = two m16 groups sharing the B fragments
        float acc[2][8][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[m][n][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < L::KSTEPS; ++ks) {
            uint32_t a[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (CIN == 3) {
                    const int kh = ks >> 1, half = ks & 1;
                    const uint32_t base = halo_u + static_cast<uint32_t>(((2 * warp + kh) * L::HPITCH + 6 * (16 * m + g) + 16 * half + 2 * t) * 2);
                    a[m][0] = lds32(base);                            // pixel g,     k = 2t, 2t+1
                    a[m][1] = lds32(base + 6 * 8 * 2);                // pixel g + 8
                    a[m][2] = lds32(base + 8 * 2);                    // pixel g,     k + 8
                    a[m][3] = lds32(base + 6 * 8 * 2 + 8 * 2);
                } else {
                    const uint32_t base = halo_u + static_cast<uint32_t>(((2 * warp + 2 * ks) * L::HPITCH + 2 * (16 * m + g) + 2 * t) * 2);
                    a[m][0] = lds32(base);                            // filter row 2ks, taps 2t, 2t+1
                    a[m][1] = lds32(base + 2 * 8 * 2);
                    a[m][2] = lds32(base + L::HPITCH * 2);            // filter row 2ks + 1
                    a[m][3] = lds32(base + L::HPITCH * 2 + 2 * 8 * 2);
                }
            }
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                ldsm_x4(wsm_u + static_cast<uint32_t>((np * 16 * L::WPITCH + ks * 16) * 2) + b_lane, bfr);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma16816(acc[m][2 * np], a[m], bfr[0], bfr[1]);
                    mma16816(acc[m][2 * np + 1], a[m], bfr[2], bfr[3]);
                }
            }
        }
        // ---- epilogue: bias / ReLU, bf16, stage the warp's 32 pixels x 64 channels, then 16-byte stores of full pixel rows
        const int oh = oh0 + warp;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int ch = 8 * n + 2 * t;
                float v0 = acc[m][n][0] + bias_s[ch], v1 = acc[m][n][1] + bias_s[ch + 1];
                float v2 = acc[m][n][2] + bias_s[ch], v3 = acc[m][n][3] + bias_s[ch + 1];
                if (P.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                *reinterpret_cast<uint32_t*>(my_stage + (16 * m + g) * L::OPITCH + ch) = pack_bf16x2(v0, v1);
                *reinterpret_cast<uint32_t*>(my_stage + (16 * m + g + 8) * L::OPITCH + ch) = pack_bf16x2(v2, v3);
            }
        __syncwarp();
        if (oh < P.Ho) {
            const int npx = P.Wo - ow0 < STW ? P.Wo - ow0 : STW;      // valid pixels of this row segment
            __nv_bfloat16* dst = P.y + ((static_cast<long>(b) * P.Ho + oh) * P.Wo + ow0) * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int px = i * 4 + (lane >> 3), chunk = lane & 7;
                if (px < npx) {
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "r"(stage_u + static_cast<uint32_t>((px * L::OPITCH + chunk * 8) * 2)));
                    *reinterpret_cast<uint4*>(dst + px * 64 + chunk * 8) = v;
                }
            }
            if (P.stats != nullptr) {                                 // channels 2*lane, 2*lane+1 over the valid pixels (bf16-rounded values)
                for (int px = 0; px < npx; ++px) {
                    const uint32_t w2 = lds32(stage_u + static_cast<uint32_t>((px * L::OPITCH + 2 * lane) * 2));
                    const float a0 = bf16_lo(w2), a1 = bf16_hi(w2);
                    st_s[0] += a0; st_q[0] = fmaf(a0, a0, st_q[0]);
                    st_s[1] += a1; st_q[1] = fmaf(a1, a1, st_q[1]);
                }
            }
        }
        __syncwarp();
    }
    if (P.stats != nullptr) {
        // one statistics row per CTA: the warps' partials added in warp order (deterministic)
        red[(warp * 64 + 2 * lane) * 2] = st_s[0]; red[(warp * 64 + 2 * lane) * 2 + 1] = st_q[0];
        red[(warp * 64 + 2 * lane + 1) * 2] = st_s[1]; red[(warp * 64 + 2 * lane + 1) * 2 + 1] = st_q[1];
        __syncthreads();
        if (threadIdx.x < 64) {
            float s = 0.f, q = 0.f;
            for (int wv = 0; wv < kStemWarps; ++wv) { s += red[(wv * 64 + threadIdx.x) * 2]; q += red[(wv * 64 + threadIdx.x) * 2 + 1]; }
            float* row = P.stats + static_cast<long>(blockIdx.x) * 128;
            row[threadIdx.x] = s;
            row[64 + threadIdx.x] = q;
            for (int r = gridDim.x + blockIdx.x; r < P.stats_rows; r += gridDim.x) {
                P.stats[static_cast<long>(r) * 128 + threadIdx.x] = 0.f;
                P.stats[static_cast<long>(r) * 128 + 64 + threadIdx.x] = 0.f;
            }
        }
        if (P.fin.counter != nullptr) {
            __syncthreads();                                          // halo / weights are no longer needed: scratch for the reduction
            bn_finalize_tail(P.fin, P.stats, gridDim.x, 64, threadIdx.x, kStemThreads, 1, smem_u32(red), reinterpret_cast<double*>(ssm));
        }
    }
}

int stem_grid(int total_tiles) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return total_tiles < 2 * sms ? total_tiles : 2 * sms;
}

template <int CIN>
int launch_stem_fwd(const StemParams& P, cudaStream_t stream) {
    using L = StemLayout<CIN>;
    static SmemAttrOnce smem_attr;
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, stem7_fwd_kernel<CIN>, L::SMEM));
    HD_CUDA_OK(hd::launch(stem7_fwd_kernel<CIN>, dim3(stem_grid(P.total_tiles)), dim3(kStemThreads), L::SMEM, stream, P));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

}  // namespace

}  // namespace hd

using namespace hd;

extern "C" int hd_stem_fwd_rows(int n, int h, int w) {
    // rows of the statistics buffer hd_stem_fwd writes (one per CTA); host-only
    const int ho = h / 2, wo = w / 2;
    return stem_grid(n * ((ho + STH - 1) / STH) * ((wo + STW - 1) / STW));
}

extern "C" int hd_stem_fwd(const void* x, int x_dtype, float x_scale, int cin, const float* w_oihw, const float* w_scale, const float* bias,
                           int relu, void* y, int n, int h, int w, float* stats, int stats_rows, const hd_bn_fin* bn_fin, hd_stream st) {
    HD_CHECK_ARG(x != nullptr && w_oihw != nullptr && y != nullptr && n > 0 && h % 2 == 0 && w % 2 == 0 && h >= 2 && w >= 2);
    HD_CHECK_ARG((cin == 1 || cin == 3) && (x_dtype == 0 || x_dtype == 1));
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    StemParams P;
    memset(&P, 0, sizeof(P));
    P.x = x; P.x_u8 = x_dtype; P.x_scale = x_scale; P.w = w_oihw; P.w_scale = w_scale; P.bias = bias; P.relu = relu;
    P.y = static_cast<__nv_bfloat16*>(y);
    P.stats = stats; P.stats_rows = stats_rows;
    P.fin = make_bn_fin(stats != nullptr ? bn_fin : nullptr);
    P.N = n; P.H = h; P.W = w; P.Ho = h / 2; P.Wo = w / 2;
    P.tiles_w = (P.Wo + STW - 1) / STW; P.tiles_h = (P.Ho + STH - 1) / STH;
    P.total_tiles = n * P.tiles_w * P.tiles_h;
    HD_CHECK_ARG(static_cast<long>(n) * P.Ho * P.Wo * 64 < (1l << 40));
    if (stats != nullptr && stats_rows < stem_grid(P.total_tiles)) {
        set_last_error(__FILE__, __LINE__, "stats buffer has fewer rows than CTAs (size it with hd_stem_fwd_rows)");
        return HD_ERR_BAD_ARG;
    }
    return cin == 3 ? launch_stem_fwd<3>(P, static_cast<cudaStream_t>(st)) : launch_stem_fwd<1>(P, static_cast<cudaStream_t>(st));
}
