// nms.cu -- greedy non-maximum suppression for the detector's proposal filter and final detections.
//
// The frozen detector of the reference is torchvision's Faster R-CNN / RetinaNet; its RPN (TV models/detection/rpn.py
// filter_proposals) and RoI heads (roi_heads.py postprocess_detections) call torchvision.ops.batched_nms -> nms, a
// third-party op that is not part of /root/reference.  Its published algorithm (torchvision csrc/ops/cuda/nms_kernel.cu):
// sort by score (stable, descending); bit (i, j) of a 64-column blocked mask says "box j > i overlaps box i with
// IoU > threshold"; a box is kept iff no earlier KEPT box has its bit set.  The IoU predicate below performs the same
// fp32 operations in the same order, so the keep set is identical.
//
// torchvision resolves the mask with one thread block that walks the boxes one by one, with a global-memory load and two
// block barriers per kept box (~0.5 us per box: 0.5 ms for the 4.3 k proposals of one image, 16 such launches per train
// step).  Here the 64 x 64 diagonal block of each 64-box chunk is resolved with register bit operations, the rows of the
// chunk are prefetched into shared memory (cp.async, double buffered) while the previous chunk is resolved, and the
// "removed" words of the later chunks are updated from shared memory: ~1 k cycles per chunk of 64 boxes instead of per box.
#include <string.h>

#include "hd_common.cuh"

namespace hd {

namespace {

constexpr int kNmsBox = 64;                 // boxes per mask word
constexpr int kScanThreads = 256;
constexpr int kMaxProblems = 64;            // per launch
constexpr int kMaxColBlocks = 128;          // shared-memory budget of the scan kernel: n <= 8192 boxes per problem

struct NmsBatch {
    int count;
    int first;                              // index of the first problem of this launch (into counts)
    const int* counts;                      // optional device array: actual box count per problem (<= capacity)
    const unsigned char* valid;             // optional device bytes, one per box slot: 0 = the box takes no part (never kept, suppresses nothing)
    int box_off[kMaxProblems];              // first box of the problem in the concatenated (sorted) box array
    int n[kMaxProblems];                    // capacity (boxes reserved for the problem)
    long mask_off[kMaxProblems];            // first mask word of the problem
};

__device__ __forceinline__ int problem_size(const NmsBatch& B, int pb) {
    const int cap = B.n[pb];
    if (B.counts == nullptr) return cap;
    const int c = B.counts[B.first + pb];
    return c < 0 ? 0 : (c < cap ? c : cap);
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, const float threshold) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float inter = __fmul_rn(width, height);
    const float sa = __fmul_rn(a.z - a.x, a.w - a.y);
    const float sb = __fmul_rn(b.z - b.x, b.w - b.y);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > threshold;
}

// grid (col_blocks, col_blocks, problems), 64 threads: word (row box, column block) of the upper triangle
__global__ void __launch_bounds__(kNmsBox) nms_mask_kernel(const float4* __restrict__ boxes, unsigned long long* __restrict__ mask,
                                                          const NmsBatch B, float threshold) {
    pdl_trigger();
    pdl_wait();
    const int pb = blockIdx.z;
    const int n = problem_size(B, pb);
    const int col_blocks = (n + kNmsBox - 1) / kNmsBox;
    const int pitch = (B.n[pb] + kNmsBox - 1) / kNmsBox;          // mask row pitch: from the capacity (host-known)
    const int row_start = blockIdx.y, col_start = blockIdx.x;
    if (row_start > col_start || col_start >= col_blocks) return;
    const float4* bx = boxes + B.box_off[pb];
    const int row_size = min(n - row_start * kNmsBox, kNmsBox), col_size = min(n - col_start * kNmsBox, kNmsBox);
    __shared__ float4 cb[kNmsBox];
    if (static_cast<int>(threadIdx.x) < col_size) cb[threadIdx.x] = bx[col_start * kNmsBox + threadIdx.x];
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < row_size) {
        const int cur = row_start * kNmsBox + threadIdx.x;
        const float4 me = bx[cur];
        unsigned long long t = 0;
        const int start = row_start == col_start ? threadIdx.x + 1 : 0;
        for (int i = start; i < col_size; ++i)
            if (iou_gt(me, cb[i], threshold)) t |= 1ULL << i;
        mask[B.mask_off[pb] + static_cast<long>(cur) * pitch + col_start] = t;
    }
}

__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// one CTA per problem: keep[i] = 1 iff sorted box i survives
__global__ void __launch_bounds__(kScanThreads, 1) nms_scan_kernel(const unsigned long long* __restrict__ mask_all,
                                                                  unsigned char* __restrict__ keep_all, const NmsBatch B) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned long long nsm[];
    const int pb = blockIdx.x;
    const int n = problem_size(B, pb);
    const int col_blocks = (n + kNmsBox - 1) / kNmsBox;
    const int pitch = (B.n[pb] + kNmsBox - 1) / kNmsBox;
    const unsigned long long* mask = mask_all + B.mask_off[pb];
    unsigned char* keep = keep_all + B.box_off[pb];
    unsigned long long* removed = nsm;                                   // [col_blocks]
    unsigned long long* rows = nsm + kMaxColBlocks;                      // [2][64][col_blocks]: words c.. of the rows of chunk c
    const int tid = threadIdx.x;
    for (int i = tid; i < col_blocks; i += kScanThreads) {
        unsigned long long r = 0;
        if (B.valid != nullptr) {                                        // boxes filtered out by the caller start out "removed"
            const unsigned char* v = B.valid + B.box_off[pb] + i * kNmsBox;
            for (int b = 0; b < kNmsBox && i * kNmsBox + b < n; ++b)
                if (!v[b]) r |= 1ULL << b;
        }
        removed[i] = r;
    }
    for (int i = n + tid; i < B.n[pb]; i += kScanThreads) keep[i] = 0;   // reserved but unused slots

    // rows of chunk c, words [c, col_blocks) -> rows[buf][i][0 .. col_blocks - c)
    auto prefetch = [&](int c, int buf) {
        const int wpr = col_blocks - c;
        const int nrow = min(n - c * kNmsBox, kNmsBox);
        const uint32_t dst0 = smem_u32(rows + static_cast<long>(buf) * kNmsBox * col_blocks);
        for (int q = tid; q < nrow * wpr; q += kScanThreads) {
            const int i = q / wpr, j = q - i * wpr;
            cp_async8(dst0 + (i * col_blocks + j) * 8, mask + static_cast<long>(c * kNmsBox + i) * pitch + c + j);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0, 0);
    for (int c = 0; c < col_blocks; ++c) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // chunk c landed; removed[] updates of chunk c-1 visible
        if (c + 1 < col_blocks) prefetch(c + 1, (c + 1) & 1);
        const unsigned long long* rb = rows + static_cast<long>(c & 1) * kNmsBox * col_blocks;
        const int nrow = min(n - c * kNmsBox, kNmsBox);
        // resolve the chunk against itself (every thread redundantly: shared-memory broadcasts, no barrier needed after)
        unsigned long long rem = removed[c];
        if (nrow < kNmsBox) rem |= ~0ULL << nrow;          // boxes past n: never kept
        unsigned long long kept = 0;
#pragma unroll
        for (int i0 = 0; i0 < kNmsBox; i0 += 16) {
            unsigned long long d[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = (i0 + i < nrow) ? rb[(i0 + i) * col_blocks] : 0ULL;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const bool k = !((rem >> (i0 + i)) & 1ULL);
                if (k) { kept |= 1ULL << (i0 + i); rem |= d[i]; }
            }
        }
        if (tid < nrow) keep[c * kNmsBox + tid] = static_cast<unsigned char>((kept >> tid) & 1ULL);
        // later chunks: removed[j] |= OR of the kept rows (thread <-> word; each word has one writer)
        for (int j = c + 1 + tid; j < col_blocks; j += kScanThreads) {
            unsigned long long acc = removed[j], kb = kept;
            while (kb) {
                const int i = __ffsll(static_cast<long long>(kb)) - 1;
                kb &= kb - 1;
                acc |= rb[i * col_blocks + (j - c)];
            }
            removed[j] = acc;
        }
    }
}

}  // namespace

}  // namespace hd

using namespace hd;

// boxes_sorted: [total][4] fp32 (x1, y1, x2, y2), the problems back to back, each sorted by descending score.
// offsets (host): problems + 1 box offsets (the capacity of each problem); counts_dev (optional, device): the number of
// boxes actually present in each problem (the rest of its slots get keep = 0) -- lets the caller stay sync-free.  mask_ws: sum over problems of n * ceil(n / 64) 64-bit words (device scratch).
// keep: [total] bytes, 1 = the box survives.  Replaces torchvision.ops.nms's kernels (see the header of this file).
extern "C" int hd_nms_valid(const float* boxes_sorted, const int* offsets, const int* counts_dev, const unsigned char* valid_dev,
                            int problems, float iou_threshold, void* mask_ws, unsigned char* keep, hd_stream stream_);

extern "C" int hd_nms(const float* boxes_sorted, const int* offsets, const int* counts_dev, int problems, float iou_threshold,
                      void* mask_ws, unsigned char* keep, hd_stream stream_) {
    return hd_nms_valid(boxes_sorted, offsets, counts_dev, nullptr, problems, iou_threshold, mask_ws, keep, stream_);
}

// hd_nms with an optional per-box-slot validity mask (device bytes): a box with valid = 0 is never kept and suppresses nothing,
// wherever it sits in its problem -- the caller's score / size filters need no compaction (and no host sync) before the NMS.
extern "C" int hd_nms_valid(const float* boxes_sorted, const int* offsets, const int* counts_dev, const unsigned char* valid_dev,
                            int problems, float iou_threshold, void* mask_ws, unsigned char* keep, hd_stream stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HD_CHECK_ARG(offsets != nullptr && problems >= 0);
    if (problems == 0) return HD_OK;
    HD_CHECK_ARG(boxes_sorted != nullptr && mask_ws != nullptr && keep != nullptr);
    HD_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes_sorted) & 15) == 0);
    const size_t smem = (kMaxColBlocks + 2 * kNmsBox * kMaxColBlocks) * sizeof(unsigned long long);
    static SmemAttrOnce smem_attr;
    HD_CUDA_OK(ensure_dyn_smem(smem_attr, nms_scan_kernel, static_cast<int>(smem)));
    long mask_off = 0;
    for (int p0 = 0; p0 < problems; p0 += kMaxProblems) {
        NmsBatch B;
        memset(&B, 0, sizeof(B));
        B.first = p0;
        B.counts = counts_dev;
        B.valid = valid_dev;
        int max_cb = 0;
        for (int p = p0; p < problems && p < p0 + kMaxProblems; ++p) {
            const int n = offsets[p + 1] - offsets[p];
            HD_CHECK_ARG(n >= 0);
            const int cb = (n + kNmsBox - 1) / kNmsBox;
            HD_CHECK_ARG(cb <= kMaxColBlocks);             // n <= 8192 boxes per problem
            B.box_off[B.count] = offsets[p];
            B.n[B.count] = n;
            B.mask_off[B.count] = mask_off;
            mask_off += static_cast<long>(n) * cb;
            if (cb > max_cb) max_cb = cb;
            ++B.count;
        }
        if (max_cb == 0) continue;
        HD_CUDA_OK(hd::launch(nms_mask_kernel, dim3(dim3(max_cb, max_cb, B.count)), dim3(kNmsBox), 0, stream, reinterpret_cast<const float4*>(boxes_sorted),
                                                                            static_cast<unsigned long long*>(mask_ws), B, iou_threshold));
        HD_CUDA_OK(cudaPeekAtLastError());
        const size_t sm = (kMaxColBlocks + 2 * static_cast<size_t>(kNmsBox) * max_cb) * sizeof(unsigned long long);
        HD_CUDA_OK(hd::launch(nms_scan_kernel, dim3(B.count), dim3(kScanThreads), sm, stream, static_cast<const unsigned long long*>(mask_ws), keep, B));
        HD_CUDA_OK(cudaPeekAtLastError());
    }
    return HD_OK;
}
