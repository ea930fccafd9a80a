// Micro-benchmark: issue rate of tcgen05.mma for K-major vs MN-major shared-memory operands (no loads; smem garbage).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../hallucidet_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include "hd_common.cuh"
using namespace hd;

__global__ void __launch_bounds__(128, 1) k(int n, int a_mn, int b_mn, int swz_bytes, int iters, int dep, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n, a_mn, b_mn);
        const uint32_t lt = swizzle_layout_type(swz_bytes);
        // K-major: SBO = 8 rows * swz; MN-major: LBO = chunk stride (16 KB), SBO = 8 k-rows * swz
        const uint64_t da = a_mn ? make_smem_desc(base, 16384, 8 * swz_bytes, lt) : make_smem_desc(base, 16, 8 * swz_bytes, lt);
        const uint64_t db = b_mn ? make_smem_desc(base + 65536, 16384, 8 * swz_bytes, lt) : make_smem_desc(base + 65536, 16, 8 * swz_bytes, lt);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = dep ? tm : tm + (i & 3) * 128;       // dependent chain vs 4 rotating accumulators
            umma_bf16(d, da, db, idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("%-8s %-5s %-5s %-5s %-4s %s\n", "N", "A_mn", "B_mn", "swz", "dep", "cycles/MMA");
    for (int n : {16, 64, 128, 256})
        for (int mode = 0; mode < 3; ++mode)            // 0: K-major both, 1: MN-major both, 2: A K-major / B MN-major
            for (int swz : {128, 32})
                for (int dep : {1, 0}) {
                    if (n == 256 && !dep) continue;
                    const int a_mn = mode == 1, b_mn = mode >= 1;
                    k<<<1, 128, 200 * 1024>>>(n, a_mn, b_mn, swz, iters, dep, d);
                    long long c = 0; cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("%-8d %-5d %-5d %-5d %-4d %.1f %s\n", n, a_mn, b_mn, swz, dep, (double)c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    return 0;
}
