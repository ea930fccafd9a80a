// bn_tail.cuh -- train-mode BatchNorm finalize executed by the LAST CTA of the producing convolution.
//
// Every CTA of a convolution that feeds a BatchNorm writes ONE row of partial per-channel sums (its own tiles, added in
// a fixed order).  The CTA that arrives last on a per-layer counter then adds up the rows in row order (fp64) and writes
// mean / invstd / scale / shift and the running statistics, i.e. what the separate bn_finalize_kernel launch did
// (46 latency-bound launches per step).  Fixed row order + fixed tile->CTA assignment keep it run-to-run deterministic.
#pragma once

#include "hd_common.cuh"

namespace hd {

struct BnFin {                 // device copy of hd_bn_fin
    double count;
    const float* gamma;
    const float* beta;
    float* rm;
    float* rv;
    float* mean;
    float* invstd;
    float* scale;
    float* shift;
    unsigned int* counter;     // nullptr = finalize not fused
    float eps, momentum;
};

inline BnFin make_bn_fin(const hd_bn_fin* f) {
    BnFin b;
    if (f == nullptr) {
        b = BnFin{};
        return b;
    }
    b.count = f->count; b.gamma = f->gamma; b.beta = f->beta; b.rm = f->running_mean; b.rv = f->running_var;
    b.mean = f->mean; b.invstd = f->invstd; b.scale = f->scale; b.shift = f->shift;
    b.counter = f->counter; b.eps = f->eps; b.momentum = f->momentum;
    return b;
}

#ifdef __CUDACC__
// Called by the `nthreads` (a power of two >= 64) threads (tid = 0 .. nthreads-1) of a CTA that have just written the CTA's
// statistics row; they meet on named barrier `bar_id`.  `flag_smem`: 4 bytes of shared memory (shared-space address);
// `scratch`: >= 8 KB of 8-byte aligned shared memory that nobody else uses any more.
//
// The last CTA sums rows x 2C floats.  Done by one thread per channel that is `rows` dependent L2 round trips (26 us at
// C = 512, measured as +17 us on the layer); instead every thread owns one float4 column group and every S-th row
// (S = nthreads / (2C/4) row slices), keeps 4 fp64 partial sums over 8-way unrolled independent loads, and the S partials
// of a channel are then added in slice order -- fixed order, hence still bit-identical run to run.
__device__ __forceinline__ void bn_finalize_tail(const BnFin& F, const float* stats, int rows, int C, int tid, int nthreads,
                                                 int bar_id, uint32_t flag_smem, double* scratch) {
    __threadfence();                                        // this thread's row entries are visible device-wide
    named_bar_sync(bar_id, nthreads);
    if (tid == 0) {
        const unsigned int prev = atomicAdd(F.counter, 1u);
        const uint32_t last = (prev + 1u == gridDim.x) ? 1u : 0u;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag_smem), "r"(last) : "memory");
    }
    named_bar_sync(bar_id, nthreads);
    uint32_t last;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(flag_smem) : "memory");
    if (!last) return;
    __threadfence();                                        // acquire: the other CTAs' rows
    const int groups = 2 * C / 4;                           // float4 column groups of one row ([sum | sum of squares])
    int S = nthreads / groups;                              // row slices (C <= 512, nthreads >= 256  =>  S >= 1 ... )
    if (S < 1) S = 1;
    if (S * 2 * C * 8 > 8192) S = 8192 / (2 * C * 8);       // scratch budget
    for (int q = tid; q < groups * S; q += nthreads) {
        const int grp = q % groups, sl = q / groups;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const float4* src = reinterpret_cast<const float4*>(stats) + grp;
        int r = sl;
        for (; r + 7 * S < rows; r += 8 * S) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + static_cast<long>(r + u * S) * groups);
#pragma unroll
            for (int u = 0; u < 8; ++u) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
        }
        for (; r < rows; r += S) {
            const float4 v = __ldcg(src + static_cast<long>(r) * groups);
            a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
        }
        double* dst = scratch + static_cast<long>(sl) * 2 * C + grp * 4;
        dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
    }
    named_bar_sync(bar_id, nthreads);
    for (int c = tid; c < C; c += nthreads) {
        double s = 0.0, q = 0.0;
        for (int sl = 0; sl < S; ++sl) {
            s += scratch[static_cast<long>(sl) * 2 * C + c];
            q += scratch[static_cast<long>(sl) * 2 * C + C + c];
        }
        const double mean = s / F.count;
        double var = q / F.count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(F.eps)));
        const float sc = F.gamma[c] * invstd;
        if (F.mean) F.mean[c] = static_cast<float>(mean);
        if (F.invstd) F.invstd[c] = invstd;
        F.scale[c] = sc;
        F.shift[c] = F.beta[c] - static_cast<float>(mean) * sc;
        if (F.rm) F.rm[c] = (1.f - F.momentum) * F.rm[c] + F.momentum * static_cast<float>(mean);
        if (F.rv) {
            const double unbiased = F.count > 1.0 ? var * F.count / (F.count - 1.0) : var;
            F.rv[c] = (1.f - F.momentum) * F.rv[c] + F.momentum * static_cast<float>(unbiased);
        }
    }
    if (tid == 0) *F.counter = 0u;                          // ready for the next launch (stream order / graph replay)
}
#endif

}  // namespace hd
