// wgrad_gemm.cu -- convolution weight-gradient on the sm_100a tensor cores.
//
//   dW[co][tap][ci] = sum_{n,ho,wo}  dY[n,ho,wo,co] * X_tap[n,ho,wo,ci]
//
// GEMM with M = output channels, N = input channels, K = pixels.  Both operands are "MN-major" in shared
// memory exactly as TMA delivers an NHWC box ([pixel][channel], channel contiguous), so no transpose is
// needed: the UMMA descriptors carry the major-ness.  One CTA = (filter tap, co tile, ci tile, pixel
// range); the pixel range is the split-K dimension, partial sums are combined with vector fp32 reductions
// (red.global.add.v4.f32) into the caller-zeroed dW buffer.  X_tap is the input box shifted by the tap
// offset (TMA zero-fills the halo); stride-2 convolutions read X through the phase view used by the forward.
//
// warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = epilogue.
// Replaces cuDNN's convolution backward-weights as reached from nn.Conv2d in the U-Net (a6 in SURVEY.md 8a).
#include "hd_common.cuh"

#include <cstring>

namespace hd {

constexpr int kWMaxTaps = 9;
constexpr int kWCpThreads = 128;               // cp.async producers for narrow-channel operands (rows < 128 bytes)
constexpr int kWThreads = 192 + kWCpThreads;

struct WgradParams {
    CUtensorMap tmDY;
    CUtensorMap tmX[2];
    int TW, TH, tiles_w, tiles_h, total_tiles;
    int KP;                       // pixels per stage (multiple of 16)
    int M, ca, m_chunks;          // UMMA M (64/128), channel chunk of the dY boxes, chunks actually loaded
    int BNt, cb;                  // ci tile, channel chunk of the X boxes
    int Cout, Cin, C0;            // C0 = channels of source 0
    int ci_tiles, splits;
    int tap_dh[kWMaxTaps], tap_dw[kWMaxTaps], tap_p[kWMaxTaps], tap_q[kWMaxTaps];
    int x_qstride[2];
    int taps;
    float* dw;
    int stages, a_bytes, stage_bytes, tmem_cols;
    // narrow-channel operands are gathered with cp.async (TMA handles 32/64-byte rows one row at a time)
    int cp_a, cp_b;
    int n_acc;                    // rotating sub-accumulators (dependent MMAs on one accumulator are latency-bound)
    int group;                    // pixel tiles per pipeline stage (narrow-channel mode: fewer, fatter stages)
    int sub_bytes;                // bytes of one (A + B) tile pair inside a stage
    const __nv_bfloat16* dy_ptr;
    const __nv_bfloat16* x_ptr;
    int Ho, Wo, Hx, Wx, Cx;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWThreads) wgrad_gemm_kernel(const __grid_constant__ WgradParams P) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = P.stages;
    const uint32_t bar_base = smem_base + static_cast<uint32_t>(stages * P.stage_bytes);
    const uint32_t full0 = bar_base, empty0 = bar_base + 8u * stages, tfull = bar_base + 16u * stages;
    const uint32_t tmem_slot = tfull + 8u;

    const int tap = blockIdx.z;
    const int co_tile = blockIdx.y / P.ci_tiles, ci_tile = blockIdx.y % P.ci_tiles;
    const int co0 = co_tile * P.M, ci0 = ci_tile * P.BNt;
    const long t_lo = static_cast<long>(P.total_tiles) * blockIdx.x / P.splits;
    const long t_hi = static_cast<long>(P.total_tiles) * (blockIdx.x + 1) / P.splits;
    const int num_k = static_cast<int>(t_hi - t_lo);

    if (threadIdx.x == 0) {
        const int full_count = ((P.cp_a && P.cp_b) ? 0 : 1) + ((P.cp_a || P.cp_b) ? kWCpThreads / 32 : 0);   // one arrival per producer warp
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8u * s, full_count);
            mbar_init(empty0 + 8u * s, 1);
        }
        mbar_init(tfull, 1);
        mbar_fence_init();
        tma_prefetch_desc(&P.tmDY);
        tma_prefetch_desc(&P.tmX[0]);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, P.n_acc * P.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                    // prologue above overlaps the previous kernel's tail
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int n_chunks = P.BNt / P.cb;
    if (warp == 0) {
        if (!(P.cp_a && P.cp_b) && elect_one()) {
            int nb = 0;                                        // X chunks inside the tensor
            for (int j = 0; j < n_chunks; ++j) nb += (ci0 + j * P.cb < P.Cin) ? 1 : 0;
            const uint32_t tx_bytes = static_cast<uint32_t>(P.KP * 2 * ((P.cp_a ? 0 : P.m_chunks * P.ca) + (P.cp_b ? 0 : nb * P.cb)));
            int stage = 0;
            uint32_t phase = 0;
            // tile coordinates advance incrementally (64-bit divisions per tile would dominate the single producer thread)
            int tw_i = static_cast<int>(t_lo % P.tiles_w);
            int th_i = static_cast<int>((t_lo / P.tiles_w) % P.tiles_h);
            int img = static_cast<int>(t_lo / (P.tiles_w * P.tiles_h));
            const int n_x = P.cp_b ? 0 : nb, n_a = P.cp_a ? 0 : P.m_chunks;
            const int tdw = P.tap_dw[tap], tdh = P.tap_dh[tap], tp = P.tap_p[tap], tq = P.tap_q[tap];
            for (long t = t_lo; t < t_hi; ++t) {
                const int w0 = tw_i * P.TW, h0 = th_i * P.TH;
                mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                const uint32_t sa = smem_base + stage * P.stage_bytes;
                const uint32_t sb = sa + P.a_bytes;
                const uint32_t fb = full0 + 8u * stage;
                mbar_expect_tx(fb, tx_bytes);
                for (int i = 0; i < n_a; ++i)
                    tma_load_5d(sa + i * P.KP * P.ca * 2, &P.tmDY, fb, co0 + i * P.ca, w0, 0, h0, img);
                for (int j = 0; j < n_x; ++j) {
                    const int cc = ci0 + j * P.cb;
                    const int src = cc < P.C0 ? 0 : 1;
                    const int c = (src ? cc - P.C0 : cc) + tq * P.x_qstride[src];
                    tma_load_5d(sb + j * P.KP * P.cb * 2, &P.tmX[src], fb, c, w0 + tdw, tp, h0 + tdh, img);
                }
                if (++tw_i == P.tiles_w) { tw_i = 0; if (++th_i == P.tiles_h) { th_i = 0; ++img; } }
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(P.M, P.BNt, 1, 1);
            const uint32_t lta = swizzle_layout_type(P.ca * 2), ltb = swizzle_layout_type(P.cb * 2);
            // when fewer channel chunks were loaded than M=128 needs, the missing chunks alias chunk 0 (LBO = 0):
            // their accumulator rows are duplicates that the epilogue never reads
            const uint32_t lbo_a = (P.m_chunks * P.ca < P.M) ? 0u : P.KP * P.ca * 2;
            const uint32_t sbo_a = 8 * P.ca * 2, kstep_a = 16 * P.ca * 2;
            const uint32_t lbo_b = P.KP * P.cb * 2, sbo_b = 8 * P.cb * 2, kstep_b = 16 * P.cb * 2;
            // loop-invariant descriptor halves; only the start-address field of lo moves (16-byte units)
            const uint64_t da0 = make_smem_desc(0, lbo_a, sbo_a, lta), db0 = make_smem_desc(0, lbo_b, sbo_b, ltb);
            const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), b_hi = static_cast<uint32_t>(db0 >> 32);
            const uint32_t a_lo0 = static_cast<uint32_t>(da0) + (smem_base >> 4);
            const uint32_t b_lo0 = static_cast<uint32_t>(db0) + ((smem_base + P.a_bytes) >> 4);
            const uint32_t stage_u = P.stage_bytes >> 4, ka_u = kstep_a >> 4, kb_u = kstep_b >> 4;
            const int ksteps = P.KP / 16;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t u = 0;
            const uint32_t n_acc = P.n_acc, cols = P.tmem_cols;
            const int num_st = (num_k + P.group - 1) / P.group;
            const uint32_t sub_u = P.sub_bytes >> 4;
            for (int st = 0; st < num_st; ++st) {
                mbar_wait(full0 + 8u * stage, phase);
                if (P.cp_a || P.cp_b) fence_proxy_async_smem();    // cp.async wrote through the generic proxy
                tc_fence_after();
                for (int g = 0; g < P.group; ++g) {
                    uint32_t a_lo = a_lo0 + stage * stage_u + g * sub_u, b_lo = b_lo0 + stage * stage_u + g * sub_u;
                    for (int k = 0; k < ksteps; ++k) {
                        umma_bf16_lohi(tmem_base + (u % n_acc) * cols, a_lo, a_hi, b_lo, b_hi, idesc, u >= n_acc);
                        ++u;
                        a_lo += ka_u;
                        b_lo += kb_u;
                    }
                }
                umma_commit(empty0 + 8u * stage);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            umma_commit(tfull);
        }
    } else if (warp >= 6) {
        // ================= cp.async producers for narrow-channel operands =================
        if (P.cp_a || P.cp_b) {
            const int pt = threadIdx.x - 192;
            // all addressing state lives in registers: parameters are copied out of the constant bank once, the
            // (thread, slot) decomposition of the 16-byte pieces is computed once, per tile only adds / compares remain
            const int Ho = P.Ho, Wo = P.Wo, Hx = P.Hx, Wx = P.Wx, TW = P.TW, TH = P.TH, tiles_w = P.tiles_w, tiles_h = P.tiles_h;
            const int Cout = P.Cout, Cx = P.Cx, KP = P.KP, group = P.group;
            const uint32_t stage_bytes = P.stage_bytes, sub_bytes = P.sub_bytes, a_bytes = P.a_bytes;
            const bool cp_a = P.cp_a != 0, cp_b = P.cp_b != 0;
            const __nv_bfloat16* dy_ptr = P.dy_ptr;
            const __nv_bfloat16* x_ptr = P.x_ptr;
            const int dh = P.tap_dh[tap], dw = P.tap_dw[tap];
            const int cpr_a = P.ca / 8, cpr_b = P.cb / 8;
            int a_hl[4], a_wl[4], b_hh[4], b_ww[4], a_off[4], b_off[4];
            uint32_t a_dst[4], b_dst[4];
            bool a_on[4], b_on[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = pt + kWCpThreads * i;
                int r = q / cpr_a, ch = q - r * cpr_a;
                a_on[i] = cp_a && q < KP * cpr_a;
                a_hl[i] = r / TW; a_wl[i] = r - a_hl[i] * TW;
                a_dst[i] = swz(static_cast<uint32_t>(r * P.ca * 2 + ch * 16), cpr_a - 1);
                a_off[i] = (a_hl[i] * Wo + a_wl[i]) * Cout + co0 + ch * 8;
                r = q / cpr_b; ch = q - r * cpr_b;
                b_on[i] = cp_b && q < KP * cpr_b;
                b_hh[i] = r / TW + dh; b_ww[i] = r - (r / TW) * TW + dw;
                b_dst[i] = a_bytes + swz(static_cast<uint32_t>(r * P.cb * 2 + ch * 16), cpr_b - 1);
                b_off[i] = (b_hh[i] * Wx + b_ww[i]) * Cx + ci0 + ch * 8;
            }
            int stage = 0;
            uint32_t phase = 0;
            int tw_i = static_cast<int>(t_lo % tiles_w);
            int th_i = static_cast<int>((t_lo / tiles_w) % tiles_h);
            int img = static_cast<int>(t_lo / (tiles_w * tiles_h));
            // software pipeline inside each producer warp: copies of stage i are committed as one cp.async group; the warp
            // signals stage i-(kLag) once that group has landed (cp.async.wait_group), with ONE mbarrier arrival per warp
            // (128 per-thread arrivals on one shared-memory word serialise and cost more than the copies themselves)
            constexpr int kLag = 2;                      // requires >= 3 pipeline stages
            int sig_stage = 0, issued = 0;
            for (long t = t_lo; t < t_hi; t += group) {
                mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                for (int g = 0; g < group; ++g) {
                    const bool live = t + g < t_hi;                 // tiles past the range contribute zeros
                    const int w0 = tw_i * TW, h0 = th_i * TH;
                    const uint32_t sbase = smem_base + stage * stage_bytes + g * sub_bytes;
                    const __nv_bfloat16* dy_base = dy_ptr + static_cast<long>((img * Ho + h0) * Wo + w0) * Cout;
                    const __nv_bfloat16* x_base = x_ptr + static_cast<long>((img * Hx + h0) * Wx + w0) * Cx;
                    const int a_hmax = live ? Ho - h0 : 0, a_wmax = Wo - w0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (a_on[i]) {
                            const bool ok = a_hl[i] < a_hmax && a_wl[i] < a_wmax;
                            cp_async16(sbase + a_dst[i], ok ? dy_base + a_off[i] : dy_ptr, ok ? 16 : 0);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (b_on[i]) {
                            const bool ok = live && static_cast<unsigned>(h0 + b_hh[i]) < static_cast<unsigned>(Hx) &&
                                            static_cast<unsigned>(w0 + b_ww[i]) < static_cast<unsigned>(Wx);
                            cp_async16(sbase + b_dst[i], ok ? x_base + b_off[i] : x_ptr, ok ? 16 : 0);
                        }
                    }
                    if (live) { if (++tw_i == tiles_w) { tw_i = 0; if (++th_i == tiles_h) { th_i = 0; ++img; } } }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                ++issued;
                if (issued > kLag) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kLag) : "memory");
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive(full0 + 8u * sig_stage);
                    if (++sig_stage == stages) sig_stage = 0;
                }
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            // drain: signal the last (up to kLag) stages
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            {
                const int pending = issued < kLag ? issued : kLag;
                for (int i = 0; i < pending; ++i) {
                    if ((threadIdx.x & 31) == 0) mbar_arrive(full0 + 8u * sig_stage);
                    if (++sig_stage == stages) sig_stage = 0;
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
    } else if (num_k > 0) {
        const int quad = warp & 3;
        // M=128: row = TMEM lane; M=64: rows 16q..16q+15 live in lanes 0..15 of quadrant q
        const int row = P.M == 128 ? quad * 32 + lane : quad * 16 + lane;
        const bool row_ok = (P.M == 128 || lane < 16) && (co0 + row < P.Cout);
        mbar_wait(tfull, 0);
        tc_fence_after();
        float* out_row = P.dw + (static_cast<long>(co0 + row) * P.taps + tap) * P.Cin;
        for (int c16 = 0; c16 < P.BNt / 16; ++c16) {
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
            const long n_mma = static_cast<long>((num_k + P.group - 1) / P.group) * P.group * (P.KP / 16);
            const int n_used = n_mma < P.n_acc ? static_cast<int>(n_mma) : P.n_acc;
            for (int sa_i = 0; sa_i < n_used; ++sa_i) {
                uint32_t r[16];
                tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + sa_i * P.tmem_cols + c16 * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(r[j]);
            }
            const int ci = ci0 + c16 * 16;
            if (row_ok && ci < P.Cin) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) red_add_v4(out_row + ci + j, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, P.n_acc * P.tmem_cols);
}

static int wround_up(int a, int b) { return (a + b - 1) / b * b; }

static int chunk_of(int c0, int c1) {
    for (int c = 64; c >= 16; c >>= 1)
        if (c0 % c == 0 && c1 % c == 0) return c;
    return 0;
}

static int wact_map(CUtensorMap* m, const hd_act& t, bool phase_view, int box_c, int TW, int TH) {
    uint64_t dims[5], str[4];
    uint32_t box[5] = {static_cast<uint32_t>(box_c), static_cast<uint32_t>(TW), 1u, static_cast<uint32_t>(TH), 1u};
    const uint64_t C = t.c, W = t.w, H = t.h;
    if (!phase_view) {
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = t.n;
        str[0] = C * 2; str[1] = W * C * 2; str[2] = W * C * 2; str[3] = H * W * C * 2;
    } else {
        dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = t.n;
        str[0] = 2 * C * 2; str[1] = W * C * 2; str[2] = 2 * W * C * 2; str[3] = H * W * C * 2;
    }
    return make_tensor_map(m, t.ptr, 5, dims, str, box, box_c * 2);
}

}  // namespace hd

using namespace hd;

extern "C" int hd_conv_wgrad(const hd_conv_args* a, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(a != nullptr && a->dw != nullptr);
    const hd_act& x0 = a->x0;
    const hd_act& dy = a->y0;
    HD_CHECK_ARG(x0.ptr && dy.ptr && x0.c % 16 == 0 && dy.c % 16 == 0);
    const bool two = a->x1.ptr != nullptr;
    if (two) HD_CHECK_ARG(a->x1.c % 16 == 0 && a->x1.n == x0.n && a->x1.h == x0.h && a->x1.w == x0.w);
    const int k = a->kh, s = a->stride;
    HD_CHECK_ARG(a->kh == a->kw && (k == 1 || k == 3) && (s == 1 || s == 2));
    const int pad = k / 2;
    if (s == 2) HD_CHECK_ARG(x0.h % 2 == 0 && x0.w % 2 == 0);
    const int Ho = x0.h / s, Wo = x0.w / s, N = x0.n;
    HD_CHECK_ARG(dy.n == N && dy.h == Ho && dy.w == Wo);
    if (narrow_wgrad_eligible(a)) return narrow_wgrad_launch(a, stream);

    WgradParams P;
    memset(&P, 0, sizeof(P));
    P.Cout = dy.c;
    P.C0 = x0.c;
    P.Cin = x0.c + (two ? a->x1.c : 0);
    P.taps = k * k;
    P.M = 128;                                   // M=64 UMMAs (SS) expose the smem A-read latency; always issue M=128
    P.ca = chunk_of(P.Cout, P.Cout);
    P.cb = chunk_of(x0.c, two ? a->x1.c : x0.c);
    HD_CHECK_ARG(P.ca && P.cb);
    {
        const int want = P.M / P.ca, have = (P.Cout + P.ca - 1) / P.ca;
        P.m_chunks = want < have ? want : have;
    }
    P.BNt = P.Cin >= 128 ? 128 : wround_up(P.Cin, P.cb);
    P.ci_tiles = (P.Cin + P.BNt - 1) / P.BNt;
    const int co_tiles = (P.Cout + P.M - 1) / P.M;

    // pixel box: TW*TH a multiple of 16, <= 128, maximise useful pixels
    {
        double best = -1.0;
        int bw = 128, bh = 1;
        for (int tw = 1; tw <= 128; ++tw)
            for (int th = 1; tw * th <= 128; ++th) {
                if ((tw * th) % 16 != 0 || tw * th < 64) continue;
                const long tiles = static_cast<long>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
                const double eff = static_cast<double>(Ho) * Wo / (static_cast<double>(tiles) * tw * th);
                const double score = eff + 1e-4 * tw * th / 128.0 + 1e-6 * tw;
                if (score > best) { best = score; bw = tw; bh = th; }
            }
        P.TW = bw; P.TH = bh;
    }
    P.KP = P.TW * P.TH;
    P.tiles_w = (Wo + P.TW - 1) / P.TW;
    P.tiles_h = (Ho + P.TH - 1) / P.TH;
    P.total_tiles = P.tiles_w * P.tiles_h * N;

    int t = 0;
    for (int r = 0; r < k; ++r)
        for (int c = 0; c < k; ++c, ++t) {
            const int eh = r - pad, ew = c - pad;
            if (s == 1) {
                P.tap_dh[t] = eh; P.tap_dw[t] = ew;
            } else {
                const int p = eh & 1, q = ew & 1;
                P.tap_p[t] = p; P.tap_q[t] = q;
                P.tap_dh[t] = (eh - p) / 2; P.tap_dw[t] = (ew - q) / 2;
            }
        }
    P.x_qstride[0] = x0.c; P.x_qstride[1] = two ? a->x1.c : 0;
    P.dw = a->dw;

    // cp.async gather for operands whose rows are narrower than 128 bytes (single chunk group, plain stride-1 view)
    P.cp_a = (P.ca < 64 && P.Cout <= P.ca) ? 1 : 0;
    P.cp_b = (P.cb < 64 && s == 1 && !two && P.Cin <= P.cb) ? 1 : 0;
    P.a_bytes = wround_up(P.KP * P.m_chunks * P.ca * 2, 1024);
    P.sub_bytes = P.a_bytes + wround_up(P.KP * P.BNt * 2, 1024);
    P.group = 1;
    if (P.cp_a && P.cp_b) {                      // both operands narrow: ~32-48 KB per stage instead of 8-12 KB
        P.group = (24 * 1024) / P.sub_bytes;
        if (P.group > 4) P.group = 4;
        if (P.group < 1) P.group = 1;
    }
    P.stage_bytes = P.sub_bytes * P.group;
    int stages = (190 * 1024) / P.stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) stages = 2;
    P.stages = stages;
    int cols = 32;
    while (cols < P.BNt) cols *= 2;
    P.tmem_cols = cols;
    P.n_acc = 1;   // measured: no gain from rotating sub-accumulators

    const int units = P.taps * co_tiles * P.ci_tiles;
    int splits = a->split_k;
    if (splits <= 0) {
        // one CTA per SM (190 KB of stages), so aim for a single wave: measured on B200, 144 CTAs beat 297 by 1.4-1.5x
        // (64->64 @128x160: 65 vs 94 us; 128->128 @64x80: 28 vs 43 us; 256->256 @32x40: 27 vs 40 us)
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
        splits = sms / units;
        const int max_by_work = P.total_tiles / 4 > 0 ? P.total_tiles / 4 : 1;
        if (splits > max_by_work) splits = max_by_work;
    }
    if (splits > P.total_tiles) splits = P.total_tiles;
    if (splits < 1) splits = 1;
    P.splits = splits;

    P.dy_ptr = static_cast<const __nv_bfloat16*>(dy.ptr);
    P.x_ptr = static_cast<const __nv_bfloat16*>(x0.ptr);
    P.Ho = Ho; P.Wo = Wo; P.Hx = x0.h; P.Wx = x0.w; P.Cx = x0.c;
    if (wact_map(&P.tmDY, dy, false, P.ca, P.TW, P.TH)) return HD_ERR_CUDA;
    if (wact_map(&P.tmX[0], x0, s == 2, P.cb, P.TW, P.TH)) return HD_ERR_CUDA;
    if (two) { if (wact_map(&P.tmX[1], a->x1, s == 2, P.cb, P.TW, P.TH)) return HD_ERR_CUDA; }
    else P.tmX[1] = P.tmX[0];

    const size_t smem = 1024 + static_cast<size_t>(stages) * P.stage_bytes + 16 * stages + 16;
    static SmemAttrOnce smem_attr;
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, wgrad_gemm_kernel, 225 * 1024));
    dim3 grid(splits, co_tiles * P.ci_tiles, P.taps);
    HD_CUDA_OK(hd::launch(wgrad_gemm_kernel, dim3(grid), dim3(kWThreads), smem, stream, P));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}
