// narrow_conv.cu -- 3x3 stride-1 convolutions whose channel counts are 16 or 32 (the two full-resolution decoder
// blocks and the segmentation head of the hallucination U-Net: segmentation_models_pytorch decoder channels 32 / 16,
// reference src/models/encoder_decoder.py:22-30).
//
// These layers move ~170 MB per launch and need < 25 GFLOP: they are bound by HBM, not by the tensor pipe, and the
// 128 x N tcgen05 tiles of conv_gemm.cu / wgrad_gemm.cu spend their time re-fetching the same pixels once per filter
// tap in 32/64-byte rows.  Here a CTA fetches an (8+2) x (32+2) pixel halo patch ONCE (16-byte cp.async chunks, zero
// fill = the convolution padding, XOR-swizzled so that ldmatrix is bank-conflict free), and the nine taps are nine
// shifted ldmatrix views of that patch feeding warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  The tensor
// throughput of mma.sync is irrelevant at this arithmetic intensity; what matters is that every input byte crosses
// HBM -> SM once.
//
//   narrow_conv_kernel<K, NC>   out[p][n] = sum_t sum_k in[p + off_t][k] * Wm[n][t*K + k]      (forward, and -- with
//                               the taps mirrored and the dgrad-packed matrix -- the input gradient)
//   narrow_wgrad_kernel<KX, KY> dW[co][t][ci] += sum_p dy[p][co] * x[p + off_t][ci]              (fp32 red.global.add)
#include <stdlib.h>
#include <string.h>

#include "hd_common.cuh"
#include "bn_tail.cuh"

namespace hd {

namespace {

constexpr int NTH = 8, NTW = 32;                 // output tile: 8 rows x 32 pixels = 16 groups of 16 pixels
constexpr int NHH = NTH + 2, NHW = NTW + 2;      // halo patch
constexpr int kNThreads = 256;
constexpr int kNWarps = kNThreads / 32;

struct NarrowParams {
    const __nv_bfloat16* x;      // [N, H, W, K]
    const __nv_bfloat16* w;      // [round16(NC)][9 * K]
    __nv_bfloat16* y;            // [N, H, W, NC]
    const float* bias;
    float* stats;                // [stats_rows][2][NC] partial sum / sum of squares, one row per CTA (deterministic)
    int stats_rows;              // rows of the buffer (>= grid; the rows past the grid are written as zeros)
    BnFin fin;                   // fused BatchNorm finalize by the last CTA (fin.counter == nullptr: off)
    float* out_f32;              // [N, out_f32_c, H, W]
    int out_f32_c;
    int relu, sigmoid, store_bf16, flip;
    int N, H, W;
    int tiles_w, tiles_h, total_tiles;
    FastDiv fd_tiles_w, fd_tiles_h;
};

struct NarrowWgradParams {
    const __nv_bfloat16* x;      // [N, H, W, KX]
    const __nv_bfloat16* dy;     // [N, H, W, KY]
    float* dw;                   // [KY][9][KX]
    int N, H, W;
    int tiles_w, tiles_h, total_tiles;
    FastDiv fd_tiles_w, fd_tiles_h;
};

// 16-byte chunk swizzle of a pixel row: eight consecutive pixels' chunk c land in eight different 16-byte bank groups
template <int C>
__device__ __forceinline__ uint32_t chunk_swz(uint32_t ww) {
    return C == 16 ? ((ww >> 2) & 1u) : ((ww >> 1) & 3u);
}

// 16-byte global->shared async copy that bypasses L1 (streamed once); src_bytes = 0 writes zeros
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename PT>
__device__ __forceinline__ void decode(const PT& P, int tile, int& img, int& h0, int& w0) {
    const int t1 = static_cast<int>(fdiv(tile, P.fd_tiles_w));
    const int tw = tile - t1 * P.tiles_w;
    img = static_cast<int>(fdiv(t1, P.fd_tiles_h));
    h0 = (t1 - img * P.tiles_h) * NTH;
    w0 = tw * NTW;
}

// ROWS x (COLS pixels) patch of an NHWC tensor whose first pixel is (h0 + DH, w0 + DH) -> swizzled shared memory; chunks
// outside the image are zero filled (= the convolution padding).  A patch row is one contiguous run of COLS*C/8 16-byte
// chunks in global memory; all offsets fit 32 bits (checked on the host).
template <int C, int ROWS, int COLS, int DH>
__device__ __forceinline__ void load_patch(uint32_t dst, const __nv_bfloat16* x, int img, int h0, int w0, int H, int W) {
    constexpr int CH = C / 8, ROWCH = COLS * CH, total = ROWS * ROWCH;
    const int org = ((img * H + h0 + DH) * W + (w0 + DH)) * C;       // element offset of the patch origin (may be < 0)
    const int row_elems = W * C;
#pragma unroll
    for (int q0 = 0; q0 < total; q0 += kNThreads) {
        const int q = q0 + static_cast<int>(threadIdx.x);
        if (q0 + kNThreads <= total || q < total) {
            const int hh = q / ROWCH, i = q - hh * ROWCH;
            const int ww = i / CH, c = i - ww * CH;
            const bool ok = static_cast<unsigned>(h0 + DH + hh) < static_cast<unsigned>(H) &&
                            static_cast<unsigned>(w0 + DH + ww) < static_cast<unsigned>(W);
            const int off = ok ? org + hh * row_elems + i * 8 : 0;
            cp_async16_cg(dst + hh * (COLS * C * 2) + ww * (C * 2) + ((static_cast<uint32_t>(c) ^ chunk_swz<C>(ww)) << 4), x + off,
                          ok ? 16 : 0);
        }
    }
}
template <int C>
__device__ __forceinline__ void load_halo(uint32_t dst, const __nv_bfloat16* x, int img, int h0, int w0, int H, int W) {
    load_patch<C, NHH, NHW, -1>(dst, x, img, h0, w0, H, W);
}
template <int C>
__device__ __forceinline__ void load_tile(uint32_t dst, const __nv_bfloat16* x, int img, int h0, int w0, int H, int W) {
    load_patch<C, NTH, NTW, 0>(dst, x, img, h0, w0, H, W);
}

// halo patches (+ dy tiles) in flight per CTA: two CTAs per SM must keep >= ~64 KB outstanding to cover the HBM latency
__host__ __device__ constexpr int fwd_stages(int k) { return k == 16 ? 4 : 3; }
__host__ __device__ constexpr int wgrad_stages(int kx, int ky) { return (kx == 16 && ky == 16) ? 4 : ((kx == 32 && ky == 32) ? 2 : 3); }

enum { kModePlain = 0, kModeStats = 1, kModeF32 = 2 };

// ------------------------------------------------------------------------------------------------
// forward / input-gradient kernel.  K = GEMM-K channels per tap (16 / 32), NC = output channels (16 / 32).
// Each warp owns 16 output channels (role = warp % (NC/16)) and every (8 / roles)-th 16-pixel group of the tile;
// its slice of the weight matrix lives in registers for the whole (persistent) kernel.
// MODE: plain (bias / relu, bf16 store), + BatchNorm statistics, or the fp32 NCHW (sigmoid) output of the head.
// ------------------------------------------------------------------------------------------------
template <int K, int NC, int MODE>
__global__ void __launch_bounds__(kNThreads, 2) narrow_conv_kernel(const NarrowParams P) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t nsm[];
    constexpr int KS = K / 16;                       // k16 slices per tap
    constexpr int ROLES = NC / 16;
    constexpr int NST = fwd_stages(K);               // halo patches in flight per CTA (HBM latency x bandwidth)
    constexpr int HALO_BYTES = NHH * NHW * K * 2;
    constexpr int PITCH = NHW * K * 2;               // bytes per halo row
    const uint32_t buf0 = smem_u32(nsm);
    const uint32_t ostage = buf0 + NST * HALO_BYTES + (threadIdx.x >> 5) * 512;      // per-warp 16 pixels x 32 bytes
    float* red = reinterpret_cast<float*>(nsm + NST * HALO_BYTES + kNWarps * 512);  // [warps][16 channels][2]
    // K = 32: 72 weight registers per thread would spill; the weight matrix stays in shared memory instead (rows padded
    // to WPITCH bytes so that ldmatrix is conflict free) and B fragments are re-read per use
    constexpr bool WSM = K == 32;
    constexpr int WPITCH = 9 * K * 2 + 16;
    const uint32_t wsm = buf0 + NST * HALO_BYTES + kNWarps * 512 + kNWarps * 16 * 2 * 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp % ROLES, slice = warp / ROLES;
    const int g = lane >> 2, t4 = lane & 3;

    int tile = blockIdx.x;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
        const int tl = tile + p * gridDim.x;
        if (tl < P.total_tiles) {
            int img, h0, w0;
            decode(P, tl, img, h0, w0);
            load_halo<K>(buf0 + p * HALO_BYTES, P.x, img, h0, w0, P.H, P.W);
        }
        cp_async_commit();
    }

    // B fragments (weights): b0 = Wm[n = g][k = 2*t4, 2*t4+1], b1 = k + 8, per tap / k16 slice / 8-channel tile
    uint32_t wreg[9][KS][2][2];
    if (WSM) {
        constexpr int RCH = 9 * K / 8;               // 16-byte chunks per weight row
        for (int q = threadIdx.x; q < NC * RCH; q += kNThreads) {
            const int n = q / RCH, c = q - n * RCH;
            const uint4 v = *reinterpret_cast<const uint4*>(P.w + n * (9 * K) + c * 8);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wsm + n * WPITCH + c * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    } else {
        const __nv_bfloat16* wrow = P.w + (role * 16 + g) * (9 * K) + 2 * t4;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __nv_bfloat16* p = wrow + (j * 8) * (9 * K) + t * K + s * 16;
                    wreg[t][s][j][0] = *reinterpret_cast<const uint32_t*>(p);
                    wreg[t][s][j][1] = *reinterpret_cast<const uint32_t*>(p + 8);
                }
    }
    // ldmatrix lane addressing of the A operand (16 pixels x 16 channels): matrices = (pixels 0-7 | 8-15) x (k 0-7 | 8-15).
    // Filter tap (r, c) reads the patch at row offset r, column offset c (forward) or 2-r, 2-c (input gradient).
    const int mat = lane >> 3, mr = lane & 7;
    const int po = (mat & 1) * 8 + mr, kc = mat >> 1;
    uint32_t aoff[3][KS];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int d = P.flip ? 2 - c : c;
            aoff[c][s] = static_cast<uint32_t>((po + d) * K * 2) + ((static_cast<uint32_t>(s * 2 + kc) ^ chunk_swz<K>(po + d)) << 4);
        }
    const int row0 = P.flip ? 2 * PITCH : 0, row_step = P.flip ? -PITCH : PITCH;
    // B through ldmatrix (WSM): matrices (n 0-7 | 8-15) x (k 0-7 | 8-15) -> {b0, b1} of channel tile 0, then of tile 1
    const uint32_t wlane = wsm + (role * 16 + (mat >> 1) * 8 + mr) * WPITCH + (mat & 1) * 16;
    // output staging (stmatrix): matrix i = pixels (i>>1)*8.. x channels (i&1)*8.. of this warp's 16 x 16 block
    const uint32_t st_addr = ostage + ((mat >> 1) * 8 + mr) * 32 + (mat & 1) * 16;

    float bias_r[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (P.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            bias_r[j][0] = __ldg(P.bias + role * 16 + j * 8 + 2 * t4);
            bias_r[j][1] = __ldg(P.bias + role * 16 + j * 8 + 2 * t4 + 1);
        }
    }
    float st_s[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, st_q[2][2] = {{0.f, 0.f}, {0.f, 0.f}};     // MODE == kModeStats

    for (int it = 0; tile < P.total_tiles; tile += gridDim.x, ++it) {
        cp_async_wait<NST - 2>();                    // this tile's patch has landed (this thread's chunks)
        __syncthreads();                             // ... everyone's; and the patch of the previous tile is free
        {
            const int next = tile + (NST - 1) * gridDim.x;
            if (next < P.total_tiles) {
                int ni, nh, nw;
                decode(P, next, ni, nh, nw);
                load_halo<K>(buf0 + ((it + NST - 1) % NST) * HALO_BYTES, P.x, ni, nh, nw, P.H, P.W);
            }
            cp_async_commit();
        }
        const uint32_t buf = buf0 + (it % NST) * HALO_BYTES;

        int img, h0, w0;
        decode(P, tile, img, h0, w0);
        // U 16-pixel groups in flight per warp: independent accumulator chains hide the mma.sync latency
        constexpr int U = 2;
        constexpr int GSTEP = kNWarps / ROLES;
        for (int gi = slice; gi < 16; gi += U * GSTEP) {
            float acc[U][2][4];
            uint32_t gbase[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int gu = gi + u * GSTEP;
                gbase[u] = buf + (gu >> 1) * PITCH + (gu & 1) * 16 * (K * 2) + row0;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[u][j][i] = 0.f;
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int r = t / 3, c = t - r * 3;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    uint32_t a[U][4];
#pragma unroll
                    for (int u = 0; u < U; ++u) ldmatrix_x4(gbase[u] + r * row_step + aoff[c][s], a[u]);
                    uint32_t b[4];
                    if (WSM) {
                        ldmatrix_x4(wlane + (t * K + s * 16) * 2, b);
                    } else {
                        b[0] = wreg[t][s][0][0]; b[1] = wreg[t][s][0][1]; b[2] = wreg[t][s][1][0]; b[3] = wreg[t][s][1][1];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        mma_bf16_16816(acc[u][0], a[u], b[0], b[1]);
                        mma_bf16_16816(acc[u][1], a[u], b[2], b[3]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // epilogue: this thread holds pixels (g, g + 8) of the group x channels (j*8 + 2*t4, +1)
                const int gu = gi + u * GSTEP;
                const int gh = h0 + (gu >> 1);
                const int gw0 = w0 + (gu & 1) * 16;
                const int pix0 = (img * P.H + gh) * P.W + gw0;          // pixel index of the group's first pixel
                uint32_t pk[2][2];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int gw = gw0 + g + half * 8;
                    const bool valid = gh < P.H && gw < P.W;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float v0 = acc[u][j][half * 2] + bias_r[j][0], v1 = acc[u][j][half * 2 + 1] + bias_r[j][1];
                        if (P.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                        if (!valid) { v0 = 0.f; v1 = 0.f; }
                        if (MODE == kModeF32) {
                            const int ch = role * 16 + j * 8 + 2 * t4;
                            if (valid && ch < P.out_f32_c) {
                                float o0 = v0, o1 = v1;
                                if (P.sigmoid) { o0 = 1.f / (1.f + __expf(-o0)); o1 = 1.f / (1.f + __expf(-o1)); }
                                float* o = P.out_f32 + (static_cast<long>(img * P.out_f32_c + ch) * P.H + gh) * P.W + gw;
                                o[0] = o0;
                                if (ch + 1 < P.out_f32_c) o[static_cast<long>(P.H) * P.W] = o1;
                            }
                        }
                        pk[half][j] = pack_bf16x2(v0, v1);
                        if (MODE == kModeStats) {
                            const float a0 = bf16_lo(pk[half][j]), a1 = bf16_hi(pk[half][j]);
                            st_s[j][0] += a0; st_q[j][0] = fmaf(a0, a0, st_q[j][0]);
                            st_s[j][1] += a1; st_q[j][1] = fmaf(a1, a1, st_q[j][1]);
                        }
                    }
                }
                if (MODE != kModeF32 || P.store_bf16) {
                    // fragments -> (pixel, 16 channels) rows through a 512-byte per-warp staging block, then one coalesced
                    // 16-byte store per lane: lane = pixel * 2 + 8-channel half
                    __syncwarp();
                    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(st_addr), "r"(pk[0][0]), "r"(pk[0][1]),
                                 "r"(pk[1][0]), "r"(pk[1][1]) : "memory");
                    __syncwarp();
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ostage + lane * 16));
                    if (gh < P.H && gw0 + (lane >> 1) < P.W)
                        *reinterpret_cast<uint4*>(P.y + (pix0 + (lane >> 1)) * NC + role * 16 + (lane & 1) * 8) = v;
                }
            }
        }
    }
    cp_async_wait<0>();
    if (MODE == kModeStats) {
        // one statistics row per CTA, reduced in a fixed order: the 8 pixel lanes of a warp, then the warps of a role
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float s = st_s[j][e], q = st_q[j][e];
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                    q += __shfl_xor_sync(0xffffffffu, q, o);
                }
                if (g == 0) {
                    red[(warp * 16 + j * 8 + 2 * t4 + e) * 2] = s;
                    red[(warp * 16 + j * 8 + 2 * t4 + e) * 2 + 1] = q;
                }
            }
        __syncthreads();
        if (threadIdx.x < NC) {
            const int ch = threadIdx.x, rl = ch >> 4, cl = ch & 15;
            float s = 0.f, q = 0.f;
            for (int wv = rl; wv < kNWarps; wv += ROLES) {
                s += red[(wv * 16 + cl) * 2];
                q += red[(wv * 16 + cl) * 2 + 1];
            }
            float* row = P.stats + static_cast<long>(blockIdx.x) * 2 * NC;
            row[ch] = s;
            row[NC + ch] = q;
            // a buffer sized for more rows than CTAs (hd_conv_fwd_tiles without a channel hint): zero the surplus rows
            for (int r = gridDim.x + blockIdx.x; r < P.stats_rows; r += gridDim.x) {
                P.stats[static_cast<long>(r) * 2 * NC + ch] = 0.f;
                P.stats[static_cast<long>(r) * 2 * NC + NC + ch] = 0.f;
            }
        }
        if (P.fin.counter != nullptr) {
            __syncthreads();                                  // `red` has been consumed: its first word becomes the "last CTA" flag
            // (all cp.async groups were waited for above: the halo stages at the start of the dynamic shared memory are free)
            bn_finalize_tail(P.fin, P.stats, gridDim.x, NC, threadIdx.x, kNThreads, 1, smem_u32(red), reinterpret_cast<double*>(nsm));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight-gradient kernel.  KX = channels of x (16 / 32), KY = channels of dy (16 / 32).  A warp owns one 16 (co) x
// 16 (ci) block of all nine taps (72 fp32 accumulators) and every (8 / roles)-th 16-pixel group; the accumulators
// live in registers across all tiles of the persistent CTA and are reduced once at the end.
// ------------------------------------------------------------------------------------------------
template <int KX, int KY>
__global__ void __launch_bounds__(kNThreads, 2) narrow_wgrad_kernel(const NarrowWgradParams P) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t nsm[];
    constexpr int XR = KX / 16, YR = KY / 16, ROLES = XR * YR;
    constexpr int HALO_BYTES = NHH * NHW * KX * 2;
    constexpr int DY_BYTES = NTH * NTW * KY * 2;
    constexpr int STAGE = HALO_BYTES + DY_BYTES;
    constexpr int NST = wgrad_stages(KX, KY);
    constexpr int XPITCH = NHW * KX * 2;
    const uint32_t buf0 = smem_u32(nsm);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp % ROLES, slice = warp / ROLES;
    const int xi = role % XR, yi = role / XR;
    const int mat = lane >> 3, mr = lane & 7;

    // A = dy^T (co x pixels) through ldmatrix.trans: matrices (pixels 0-7 | 8-15) x (co 0-7 | 8-15) in a0..a3 order
    const int a_po = (mat >> 1) * 8 + mr, a_c = yi * 2 + (mat & 1);
    const uint32_t a_off = static_cast<uint32_t>(a_po * KY * 2) + ((static_cast<uint32_t>(a_c) ^ chunk_swz<KY>(a_po)) << 4);
    // B = x (pixels x ci) through ldmatrix.trans: {b0,b1} of the first 8 ci, then of the second 8 ci
    const int b_po = (mat & 1) * 8 + mr, b_c = xi * 2 + (mat >> 1);
    uint32_t b_off[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        b_off[d] = static_cast<uint32_t>((b_po + d) * KX * 2) + ((static_cast<uint32_t>(b_c) ^ chunk_swz<KX>(b_po + d)) << 4);

    float acc[9][2][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[t][j][i] = 0.f;

    int tile = blockIdx.x;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
        const int tl = tile + p * gridDim.x;
        if (tl < P.total_tiles) {
            int img, h0, w0;
            decode(P, tl, img, h0, w0);
            load_halo<KX>(buf0 + p * STAGE, P.x, img, h0, w0, P.H, P.W);
            load_tile<KY>(buf0 + p * STAGE + HALO_BYTES, P.dy, img, h0, w0, P.H, P.W);
        }
        cp_async_commit();
    }

    for (int it = 0; tile < P.total_tiles; tile += gridDim.x, ++it) {
        if (NST > 2) {
            cp_async_wait<(NST > 2 ? NST - 2 : 0)>();
            __syncthreads();                         // this tile landed; the previous tile's stage is free
        }
        {
            const int next = tile + (NST - 1) * gridDim.x;
            if (next < P.total_tiles) {
                int ni, nh, nw;
                decode(P, next, ni, nh, nw);
                const uint32_t nb = buf0 + ((it + NST - 1) % NST) * STAGE;
                load_halo<KX>(nb, P.x, ni, nh, nw, P.H, P.W);
                load_tile<KY>(nb + HALO_BYTES, P.dy, ni, nh, nw, P.H, P.W);
            }
            cp_async_commit();
        }
        if (NST == 2) {                              // two stages: the refill above targets the other stage
            cp_async_wait<1>();
            __syncthreads();
        }
        const uint32_t buf = buf0 + (it % NST) * STAGE;

        for (int gi = slice; gi < 16; gi += kNWarps / ROLES) {
            const int hl = gi >> 1, wl0 = (gi & 1) * 16;
            uint32_t a[4];
            ldmatrix_x4_trans(buf + HALO_BYTES + (hl * NTW + wl0) * (KY * 2) + a_off, a);
            const uint32_t xb = buf + hl * XPITCH + wl0 * (KX * 2);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int r = t / 3, c = t - r * 3;
                uint32_t b[4];
                ldmatrix_x4_trans(xb + r * XPITCH + b_off[c], b);
                mma_bf16_16816(acc[t][0], a, b[0], b[1]);
                mma_bf16_16816(acc[t][1], a, b[2], b[3]);
            }
        }
        if (NST == 2) __syncthreads();               // stage may be refilled at the top of the next iteration
    }
    cp_async_wait<0>();
    __syncthreads();

    // CTA reduction in a fixed order (slice by slice), then one fp32 red.global.add per weight and CTA
    float* red = reinterpret_cast<float*>(nsm);                    // [roles][9][2][4][32 lanes]
    constexpr int SLICES = kNWarps / ROLES;
    for (int s = 0; s < SLICES; ++s) {
        if (slice == s) {
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float* p = red + ((role * 9 + t) * 8 + j * 4 + i) * 32 + lane;
                        if (s == 0) *p = acc[t][j][i];
                        else *p += acc[t][j][i];
                    }
        }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < ROLES * 72 * 32; q += kNThreads) {
        const int ln = q & 31, e = q >> 5;
        const int i = e & 3, j = (e >> 2) & 1, t = (e >> 3) % 9, rl = (e >> 3) / 9;
        const int rxi = rl % XR, ryi = rl / XR;
        const int co = ryi * 16 + (ln >> 2) + (i >> 1) * 8;
        const int ci = rxi * 16 + j * 8 + 2 * (ln & 3) + (i & 1);
        atomicAdd(P.dw + (co * 9 + t) * KX + ci, red[q]);
    }
}

bool narrow_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("HD_NARROW");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

int narrow_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

bool narrow_c(int c) { return c == 16 || c == 32; }

int narrow_grid(int total_tiles) { return total_tiles < 2 * narrow_sms() ? total_tiles : 2 * narrow_sms(); }

template <int K, int NC, int MODE>
int launch_fwd_mode(const NarrowParams& P, cudaStream_t stream) {
    const size_t smem = fwd_stages(K) * NHH * NHW * K * 2 + kNWarps * 512 + kNWarps * 16 * 2 * sizeof(float) +
                        (K == 32 ? NC * (9 * K * 2 + 16) : 0);
    static SmemAttrOnce smem_attr;
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, narrow_conv_kernel<K, NC, MODE>, static_cast<int>(smem)));
    HD_CUDA_OK(hd::launch(narrow_conv_kernel<K, NC, MODE>, dim3(narrow_grid(P.total_tiles)), dim3(kNThreads), smem, stream, P));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

template <int K, int NC>
int launch_fwd(const NarrowParams& P, cudaStream_t stream) {
    if (P.out_f32 != nullptr) return launch_fwd_mode<K, NC, kModeF32>(P, stream);
    if (P.stats != nullptr) return launch_fwd_mode<K, NC, kModeStats>(P, stream);
    return launch_fwd_mode<K, NC, kModePlain>(P, stream);
}

template <int KX, int KY>
int launch_wgrad(const NarrowWgradParams& P, cudaStream_t stream) {
    size_t smem = static_cast<size_t>(wgrad_stages(KX, KY)) * (NHH * NHW * KX * 2 + NTH * NTW * KY * 2);
    const size_t red = static_cast<size_t>(KX / 16) * (KY / 16) * 72 * 32 * sizeof(float);
    if (smem < red) smem = red;
    static SmemAttrOnce smem_attr;
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, narrow_wgrad_kernel<KX, KY>, static_cast<int>(smem)));
    HD_CUDA_OK(hd::launch(narrow_wgrad_kernel<KX, KY>, dim3(narrow_grid(P.total_tiles)), dim3(kNThreads), smem, stream, P));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}

}  // namespace

// 3x3 stride-1, one source, 16/32 -> 16/32 channels, no fused residual / mask operands
bool narrow_conv_eligible(const hd_conv_args* a, bool dgrad) {
    if (!narrow_enabled() || a == nullptr) return false;
    if (a->kh != 3 || a->kw != 3 || a->stride != 1) return false;
    if (a->x1.ptr != nullptr || a->y1.ptr != nullptr || a->add != nullptr || a->mask != nullptr) return false;
    if (!narrow_c(a->x0.c) || !narrow_c(a->y0.c)) return false;
    if (dgrad && a->stats != nullptr) return false;
    if (a->stats != nullptr && a->out_f32_nchw != nullptr) return false;      // one specialised epilogue per launch
    if (a->out_f32_nhwc) return false;
    return true;
}

// rows of the statistics buffer the forward kernel writes: one per CTA
int narrow_conv_stats_rows(const hd_conv_args* a) {
    return narrow_grid(((a->x0.w + NTW - 1) / NTW) * ((a->x0.h + NTH - 1) / NTH) * a->x0.n);
}

int narrow_conv_launch(const hd_conv_args* a, bool dgrad, cudaStream_t stream) {
    HD_CHECK_ARG(a->w != nullptr && a->x0.ptr != nullptr && a->y0.ptr != nullptr);
    HD_CHECK_ARG(a->y0.n == a->x0.n && a->y0.h == a->x0.h && a->y0.w == a->x0.w);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x0.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->y0.ptr) & 15) == 0);
    HD_CHECK_ARG(a->store_bf16 || a->out_f32_nchw != nullptr);
    NarrowParams P;
    memset(&P, 0, sizeof(P));
    P.x = static_cast<const __nv_bfloat16*>(a->x0.ptr);
    P.w = static_cast<const __nv_bfloat16*>(a->w);
    P.y = static_cast<__nv_bfloat16*>(a->y0.ptr);
    P.bias = a->bias;
    P.stats = a->stats;
    P.stats_rows = a->stats_replicas;
    P.fin = make_bn_fin(a->stats != nullptr ? a->bn_fin : nullptr);
    P.out_f32 = a->out_f32_nchw;
    P.out_f32_c = a->out_f32_channels;
    P.relu = a->relu; P.sigmoid = a->sigmoid; P.store_bf16 = a->store_bf16;
    P.flip = dgrad ? 1 : 0;
    P.N = a->x0.n; P.H = a->x0.h; P.W = a->x0.w;
    P.tiles_w = (P.W + NTW - 1) / NTW;
    P.tiles_h = (P.H + NTH - 1) / NTH;
    P.total_tiles = P.tiles_w * P.tiles_h * P.N;
    P.fd_tiles_w = make_fastdiv(static_cast<uint32_t>(P.tiles_w));
    P.fd_tiles_h = make_fastdiv(static_cast<uint32_t>(P.tiles_h));
    HD_CHECK_ARG(static_cast<long>(P.N) * P.H * P.W * 32 < (1l << 31));          // 32-bit element offsets in the kernels
    if (P.stats != nullptr && a->stats_replicas < narrow_grid(P.total_tiles)) {
        set_last_error(__FILE__, __LINE__, "stats buffer has fewer rows than CTAs (size it with hd_conv_fwd_tiles)");
        return HD_ERR_BAD_ARG;
    }
    const int K = a->x0.c, NC = a->y0.c;
    if (K == 16 && NC == 16) return launch_fwd<16, 16>(P, stream);
    if (K == 32 && NC == 16) return launch_fwd<32, 16>(P, stream);
    if (K == 16 && NC == 32) return launch_fwd<16, 32>(P, stream);
    return launch_fwd<32, 32>(P, stream);
}

bool narrow_wgrad_eligible(const hd_conv_args* a) {
    if (!narrow_enabled() || a == nullptr) return false;
    if (a->kh != 3 || a->kw != 3 || a->stride != 1 || a->x1.ptr != nullptr) return false;
    return narrow_c(a->x0.c) && narrow_c(a->y0.c);
}

int narrow_wgrad_launch(const hd_conv_args* a, cudaStream_t stream) {
    HD_CHECK_ARG(a->dw != nullptr && a->x0.ptr != nullptr && a->y0.ptr != nullptr);
    HD_CHECK_ARG(a->y0.n == a->x0.n && a->y0.h == a->x0.h && a->y0.w == a->x0.w);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x0.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->y0.ptr) & 15) == 0);
    NarrowWgradParams P;
    memset(&P, 0, sizeof(P));
    P.x = static_cast<const __nv_bfloat16*>(a->x0.ptr);
    P.dy = static_cast<const __nv_bfloat16*>(a->y0.ptr);
    P.dw = a->dw;
    P.N = a->x0.n; P.H = a->x0.h; P.W = a->x0.w;
    P.tiles_w = (P.W + NTW - 1) / NTW;
    P.tiles_h = (P.H + NTH - 1) / NTH;
    P.total_tiles = P.tiles_w * P.tiles_h * P.N;
    P.fd_tiles_w = make_fastdiv(static_cast<uint32_t>(P.tiles_w));
    P.fd_tiles_h = make_fastdiv(static_cast<uint32_t>(P.tiles_h));
    HD_CHECK_ARG(static_cast<long>(P.N) * P.H * P.W * 32 < (1l << 31));
    const int KX = a->x0.c, KY = a->y0.c;
    if (KX == 16 && KY == 16) return launch_wgrad<16, 16>(P, stream);
    if (KX == 32 && KY == 16) return launch_wgrad<32, 16>(P, stream);
    if (KX == 16 && KY == 32) return launch_wgrad<16, 32>(P, stream);
    return launch_wgrad<32, 32>(P, stream);
}

}  // namespace hd
