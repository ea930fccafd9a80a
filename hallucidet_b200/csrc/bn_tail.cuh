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
__device__ __forceinline__ void bn_finalize_channel(const BnFin& F, const float* stats, int rows, int C, int c) {
    double s = 0.0, q = 0.0;
    for (int r = 0; r < rows; ++r) {
        s += static_cast<double>(__ldcg(stats + (static_cast<long>(r) * 2) * C + c));
        q += static_cast<double>(__ldcg(stats + (static_cast<long>(r) * 2 + 1) * C + c));
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

// Called by the `nthreads` threads (tid = 0 .. nthreads-1) of a CTA that have just written the CTA's statistics row;
// they meet on named barrier `bar_id`.  `flag_smem`: 4 bytes of shared memory (shared-space address).
__device__ __forceinline__ void bn_finalize_tail(const BnFin& F, const float* stats, int rows, int C, int tid, int nthreads,
                                                 int bar_id, uint32_t flag_smem) {
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
    for (int c = tid; c < C; c += nthreads) bn_finalize_channel(F, stats, rows, C, c);
    if (tid == 0) *F.counter = 0u;                          // ready for the next launch (stream order / graph replay)
}
#endif

}  // namespace hd
