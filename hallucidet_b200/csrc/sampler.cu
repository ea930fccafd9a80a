// sampler.cu -- torchvision's BalancedPositiveNegativeSampler for a whole batch WITHOUT a device->host round trip.
//
// The frozen detector of the reference is torchvision's Faster R-CNN.  Its RPN loss and its RoI heads subsample anchors /
// proposals with det_utils.BalancedPositiveNegativeSampler (TV models/detection/_utils.py): per image
//     positive = where(labels >= 1);  negative = where(labels == 0)
//     num_pos  = min(positive.numel(), int(batch_size_per_image * positive_fraction))
//     num_neg  = min(negative.numel(), batch_size_per_image - num_pos)
//     pos_idx  = positive[randperm(positive.numel())[:num_pos]];  neg_idx = negative[randperm(negative.numel())[:num_neg]]
// The two randperm sizes are data dependent, so every image costs a device->host read before the draw can be issued --
// the host then sits on the critical path of the train step (2 x 16 randperm calls, ~45 radix-sort launches).
//
// This file reproduces the draw on the device, bit for bit, from the same Philox stream torch.randperm would consume
// (ATen/native/cuda/Randperm.cu, Randperm.cuh, DistributionTemplates.h -- third-party, restated from the shipped headers):
//   * randperm(n) draws one key per element i: curand_init(seed, i, offset); (x, y, ., .) = curand4(); v = ((x << 32) | y) % R
//     with R = 2^32 - 1 for 32-bit keys (bits <= 32) and 2^64 - 1 otherwise, bits = ceil(log2(n - (6 n^2 + 1) / (12 ln 0.9)));
//   * the permutation is the stable ascending radix sort of the indices by (v & (2^bits - 1));
//   * runs of equal keys ("islands") are re-shuffled by the thread at the island's first sorted position t:
//     curand_init(seed, t, offset + 4); Fisher-Yates with curand() % (i + 1);
//   * the generator advances by 4 + ceil4(n) per call (nothing for n = 0).
// Only the first num_pos / num_neg entries of each permutation are used, i.e. the elements with the SMALLEST keys: instead of
// sorting n keys (n ~ 100 k anchors) the kernels keep the candidates below a key threshold (expected ~k + 4 sqrt(k) + 32 of
// them; an exact re-scan with a moved threshold in the rare case of a miss), sort those <= 2048 in shared memory, resolve
// the islands and mark the selected columns.  Offsets chain through device memory (image 0 pos, image 0 neg, image 1 pos ...
// as in torchvision's loop); the caller mirrors the final offset back into torch's generator at its next host sync.
#include <curand_kernel.h>
#include <math.h>

#include "hd_common.cuh"

namespace hd {

namespace {

constexpr int kSlices = 16;
constexpr int kGenThreads = 512;
constexpr int kSelThreads = 1024;
constexpr int kCap = 2048;
constexpr int kMaxCalls = 128;              // 2 x images per launch

struct Cand {
    unsigned long long key;
    unsigned int rank;                      // index into the positive / negative list (= value of the permutation entry)
    unsigned int col;
};

struct SampParams {
    const void* labels;
    int dtype;                              // 0 = float32, 1 = int64
    int B, N, bs, pos_max;
    unsigned long long seed;
    const unsigned long long* off_in;
    unsigned long long* off_out;
    unsigned char* mask;                    // [B][N]: 1 = sampled positive, 2 = sampled negative
    int* counts;                            // [B][4]: n_pos, n_neg, num_pos, num_neg
    int* slice_counts;                      // [2B][kSlices]
    Cand* cand;                             // [2B][kCap]
    int* cand_n;                            // [2B]
    double log_thr12;                       // std::log(0.9) * 12, evaluated on the host as ATen does
};

__device__ __forceinline__ bool in_class(const SampParams& P, long idx, int c) {
    if (P.dtype == 0) {
        const float v = static_cast<const float*>(P.labels)[idx];
        return c == 0 ? v >= 1.f : v == 0.f;
    }
    const long long v = static_cast<const long long*>(P.labels)[idx];
    return c == 0 ? v >= 1 : v == 0;
}

__device__ __forceinline__ unsigned long long call_increment(int n) {
    return n == 0 ? 0ull : 4ull + static_cast<unsigned long long>((n + 3) / 4) * 4ull;
}

struct CallInfo {
    int n, k, bits;
    unsigned long long offset, mask, range, thr;
};

// everything a CTA of call q needs, from the per-slice counts (s_cnt: shared copy of slice_counts, 2B x kSlices)
__device__ void call_info(const SampParams& P, const int* s_cnt, int q, CallInfo& ci) {
    const int b = q >> 1, c = q & 1;
    unsigned long long off = *P.off_in;
    int n_self = 0, n_pos_b = 0, n_neg_b = 0;
    for (int j = 0; j < 2 * P.B; ++j) {
        int t = 0;
        for (int s = 0; s < kSlices; ++s) t += s_cnt[j * kSlices + s];
        if (j < q) off += call_increment(t);
        if (j == q) n_self = t;
        if (j == 2 * b) n_pos_b = t;
        if (j == 2 * b + 1) n_neg_b = t;
    }
    const int num_pos = min(n_pos_b, P.pos_max);
    const int num_neg = min(n_neg_b, P.bs - num_pos);
    ci.n = n_self;
    ci.k = c == 0 ? num_pos : num_neg;
    ci.offset = off;
    const double nd = static_cast<double>(n_self);
    const double x = nd - (6.0 * nd * nd + 1.0) / P.log_thr12;
    int bits = 0;
    while (bits < 64 && ldexp(1.0, bits) < x) ++bits;
    ci.bits = bits;
    ci.mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    ci.range = bits <= 32 ? 0xFFFFFFFFull : ~0ull;
    const double want = static_cast<double>(ci.k) + 4.0 * sqrt(static_cast<double>(ci.k)) + 32.0;
    if (n_self <= kCap && static_cast<double>(n_self) <= 1.5 * want) {
        ci.thr = ci.mask;                        // short row, or most of it is needed anyway: every key is a candidate
    } else {
        const double t = ldexp(want / nd, bits);
        ci.thr = t >= 18446744073709551615.0 ? ci.mask : static_cast<unsigned long long>(t);
        if (ci.thr > ci.mask) ci.thr = ci.mask;
    }
}

__device__ __forceinline__ unsigned long long draw_key(unsigned long long seed, unsigned int rank, const CallInfo& ci) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, rank, ci.offset, &st);
    const uint4 r = curand4(&st);
    const unsigned long long v = (static_cast<unsigned long long>(r.x) << 32) | r.y;
    return (v % ci.range) & ci.mask;
}

// Scan columns [c0, c1) of image b in blockDim-sized chunks; every element of class c gets its rank (rank_base + number of
// class members before it) and key; keys <= ci.thr are appended to cand (slots from *counter; entries past kCap are counted
// but not stored).  s_warp: blockDim/32 ints of shared scratch.
__device__ void scan_collect(const SampParams& P, int b, int c, int c0, int c1, int rank_base, const CallInfo& ci, Cand* cand,
                             int* counter, int* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int running = rank_base;
    for (int base = c0; base < c1; base += blockDim.x) {
        const int col = base + threadIdx.x;
        const bool f = col < c1 && in_class(P, static_cast<long>(b) * P.N + col, c);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < nwarps; ++w) {
            const int t = s_warp[w];
            if (w < warp) before += t;
            total += t;
        }
        if (f) {
            const unsigned int rank = static_cast<unsigned int>(running + before + __popc(bal & ((1u << lane) - 1u)));
            const unsigned long long key = draw_key(P.seed, rank, ci);
            if (key <= ci.thr) {
                const int slot = atomicAdd(counter, 1);
                if (slot < kCap) cand[slot] = Cand{key, rank, static_cast<unsigned int>(col)};
            }
        }
        running += total;
        __syncthreads();
    }
}

__device__ __forceinline__ void slice_range(int N, int s, int& c0, int& c1) {
    const int L = (N + kSlices - 1) / kSlices;
    c0 = min(N, s * L);
    c1 = min(N, c0 + L);
}

// grid (kSlices, B): class counts of a column slice; clears the slice of the output mask
__global__ void __launch_bounds__(kGenThreads) samp_count_kernel(const SampParams P) {
    pdl_trigger();
    pdl_wait();
    const int s = blockIdx.x, b = blockIdx.y;
    int c0, c1;
    slice_range(P.N, s, c0, c1);
    int np = 0, nn = 0;
    for (int col = c0 + threadIdx.x; col < c1; col += blockDim.x) {
        const long idx = static_cast<long>(b) * P.N + col;
        np += in_class(P, idx, 0) ? 1 : 0;
        nn += in_class(P, idx, 1) ? 1 : 0;
        P.mask[idx] = 0;
    }
    __shared__ int s_p[kGenThreads / 32], s_n[kGenThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        np += __shfl_xor_sync(0xffffffffu, np, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if ((threadIdx.x & 31) == 0) { s_p[threadIdx.x >> 5] = np; s_n[threadIdx.x >> 5] = nn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tp = 0, tn = 0;
        for (int w = 0; w < kGenThreads / 32; ++w) { tp += s_p[w]; tn += s_n[w]; }
        P.slice_counts[(2 * b) * kSlices + s] = tp;
        P.slice_counts[(2 * b + 1) * kSlices + s] = tn;
        if (s == 0) { P.cand_n[2 * b] = 0; P.cand_n[2 * b + 1] = 0; }
    }
}

// grid (kSlices, 2B): keys of one column slice of one call; candidates below the threshold go to the call's list
__global__ void __launch_bounds__(kGenThreads) samp_generate_kernel(const SampParams P) {
    pdl_trigger();
    pdl_wait();
    __shared__ int s_cnt[kMaxCalls * kSlices];
    __shared__ int s_warp[kGenThreads / 32];
    __shared__ CallInfo s_ci;
    const int s = blockIdx.x, q = blockIdx.y, b = q >> 1, c = q & 1;
    for (int i = threadIdx.x; i < 2 * P.B * kSlices; i += blockDim.x) s_cnt[i] = P.slice_counts[i];
    __syncthreads();
    if (threadIdx.x == 0) call_info(P, s_cnt, q, s_ci);
    __syncthreads();
    const CallInfo ci = s_ci;
    if (ci.n == 0 || ci.k == 0) return;
    int rank_base = 0;
    for (int t = 0; t < s; ++t) rank_base += s_cnt[q * kSlices + t];
    int c0, c1;
    slice_range(P.N, s, c0, c1);
    scan_collect(P, b, c, c0, c1, rank_base, ci, P.cand + static_cast<long>(q) * kCap, P.cand_n + q, s_warp);
}

__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) {
    return a.key < b.key || (a.key == b.key && a.rank < b.rank);
}

// grid (2B): sort the call's candidates, resolve islands of equal keys, mark the first k columns
__global__ void __launch_bounds__(kSelThreads) samp_select_kernel(const SampParams P) {
    pdl_trigger();
    pdl_wait();
    __shared__ int s_cnt[kMaxCalls * kSlices];
    __shared__ int s_warp[kSelThreads / 32];
    __shared__ CallInfo s_ci;
    __shared__ int s_m;
    __shared__ Cand s_c[kCap];
    const int q = blockIdx.x, b = q >> 1, c = q & 1;
    for (int i = threadIdx.x; i < 2 * P.B * kSlices; i += blockDim.x) s_cnt[i] = P.slice_counts[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        call_info(P, s_cnt, q, s_ci);
        s_m = P.cand_n[q];
        if (c == 0) {
            int n_pos = 0, n_neg = 0;
            for (int s = 0; s < kSlices; ++s) { n_pos += s_cnt[q * kSlices + s]; n_neg += s_cnt[(q + 1) * kSlices + s]; }
            const int num_pos = min(n_pos, P.pos_max), num_neg = min(n_neg, P.bs - num_pos);
            P.counts[b * 4 + 0] = n_pos; P.counts[b * 4 + 1] = n_neg; P.counts[b * 4 + 2] = num_pos; P.counts[b * 4 + 3] = num_neg;
        }
        if (q == 0) {
            unsigned long long off = *P.off_in;
            for (int j = 0; j < 2 * P.B; ++j) {
                int t = 0;
                for (int s = 0; s < kSlices; ++s) t += s_cnt[j * kSlices + s];
                off += call_increment(t);
            }
            *P.off_out = off;
        }
    }
    __syncthreads();
    CallInfo ci = s_ci;
    if (ci.n == 0 || ci.k == 0) return;
    Cand* gc = P.cand + static_cast<long>(q) * kCap;
    // the threshold missed (fewer than k candidates, or more than the list holds): move it and re-scan the whole row here
    int m = s_m;
    while (ci.thr < ci.mask && (m < ci.k || m > kCap)) {
        if (m < ci.k) ci.thr = ci.thr > (ci.mask >> 1) ? (ci.n <= kCap ? ci.mask : ci.mask - 1ull) : ci.thr * 2ull + 1ull;
        else ci.thr = ci.thr / 2ull;
        __syncthreads();                       // everyone has read s_m
        if (threadIdx.x == 0) s_m = 0;
        __syncthreads();
        scan_collect(P, b, c, 0, P.N, 0, ci, gc, &s_m, s_warp);
        __syncthreads();
        m = s_m;
    }
    int sort_n = 64;                               // power of two >= m: the padding entries sort to the end
    while (sort_n < m) sort_n <<= 1;
    for (int i = threadIdx.x; i < sort_n; i += blockDim.x) s_c[i] = i < m ? gc[i] : Cand{~0ull, ~0u, 0u};
    __syncthreads();
    // bitonic sort, ascending by (key, rank): the stable radix sort of the keys restricted to the candidates
    for (int size = 2; size <= sort_n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < sort_n / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const Cand a = s_c[lo], d = s_c[hi];
                if (cand_less(d, a) == up) { s_c[lo] = d; s_c[hi] = a; }
            }
            __syncthreads();
        }
    }
    // islands of equal keys: re-shuffled by the thread at their first position (randperm_handle_duplicate_keys_kernel)
    for (int tid = threadIdx.x; tid < m - 1; tid += blockDim.x) {
        if (s_c[tid].key != s_c[tid + 1].key) continue;
        if (tid != 0 && s_c[tid].key == s_c[tid - 1].key) continue;
        int island = 0;
        do { ++island; } while (tid + island < m && s_c[tid + island].key == s_c[tid].key);
        curandStatePhilox4_32_10_t st;
        curand_init(P.seed, static_cast<unsigned long long>(tid), ci.offset + 4ull, &st);
        for (int i = island - 1; i > 0; --i) {
            const unsigned int r = curand(&st) % static_cast<unsigned int>(i + 1);
            if (static_cast<unsigned int>(i) != r) {
                const unsigned int tr = s_c[tid + i].rank, tc = s_c[tid + i].col;
                s_c[tid + i].rank = s_c[tid + r].rank; s_c[tid + i].col = s_c[tid + r].col;
                s_c[tid + r].rank = tr; s_c[tid + r].col = tc;
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ci.k; j += blockDim.x) P.mask[static_cast<long>(b) * P.N + s_c[j].col] = static_cast<unsigned char>(1 + c);
}

}  // namespace

}  // namespace hd

using namespace hd;

extern "C" int64_t hd_sample_balanced_workspace_bytes(int batch) {
    if (batch <= 0 || 2 * batch > kMaxCalls) return -1;
    return static_cast<int64_t>(2 * batch) * (kSlices * 4 + 4 + static_cast<int64_t>(kCap) * sizeof(Cand)) + 256;
}

// See include/hallucidet_b200.h.
extern "C" int hd_sample_balanced(const void* labels, int labels_dtype, int batch, int n, int batch_size_per_image, int num_pos_max,
                                  uint64_t seed, const uint64_t* offset_in, uint64_t* offset_out, uint8_t* sampled, int32_t* counts,
                                  void* workspace, int64_t workspace_bytes, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(labels != nullptr && (labels_dtype == 0 || labels_dtype == 1) && batch > 0 && 2 * batch <= kMaxCalls && n > 0);
    HD_CHECK_ARG(batch_size_per_image > 0 && batch_size_per_image <= kCap / 2 && num_pos_max >= 0 && num_pos_max <= batch_size_per_image);
    HD_CHECK_ARG(offset_in != nullptr && offset_out != nullptr && offset_in != offset_out && sampled != nullptr && counts != nullptr);
    HD_CHECK_ARG(workspace != nullptr && workspace_bytes >= hd_sample_balanced_workspace_bytes(batch) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
    SampParams P;
    P.labels = labels; P.dtype = labels_dtype; P.B = batch; P.N = n; P.bs = batch_size_per_image; P.pos_max = num_pos_max;
    P.seed = seed; P.off_in = reinterpret_cast<const unsigned long long*>(offset_in);
    P.off_out = reinterpret_cast<unsigned long long*>(offset_out);
    P.mask = sampled; P.counts = counts;
    uint8_t* w = static_cast<uint8_t*>(workspace);
    P.cand = reinterpret_cast<Cand*>(w);
    w += static_cast<size_t>(2 * batch) * kCap * sizeof(Cand);
    P.slice_counts = reinterpret_cast<int*>(w);
    w += static_cast<size_t>(2 * batch) * kSlices * 4;
    P.cand_n = reinterpret_cast<int*>(w);
    P.log_thr12 = std::log(0.9) * 12;
    HD_CUDA_OK(hd::launch(samp_count_kernel, dim3(kSlices, batch), dim3(kGenThreads), 0, stream, P));
    HD_CUDA_OK(hd::launch(samp_generate_kernel, dim3(kSlices, 2 * batch), dim3(kGenThreads), 0, stream, P));
    HD_CUDA_OK(hd::launch(samp_select_kernel, dim3(2 * batch), dim3(kSelThreads), 0, stream, P));
    HD_CUDA_OK(cudaPeekAtLastError());
    return HD_OK;
}
