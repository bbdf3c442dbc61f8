// Microbenchmark: cost of publishing an operand chunk into TMEM on sm_100a: per-iteration cycles of
// {2 x tcgen05.st.32x32b.x8, tcgen05.wait::st, tcgen05.fence::before_thread_sync} for 1..16 warps doing it at once,
// and of tcgen05.ld.32x32b.x32 + wait::ld for 1..8 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_st_latency tmem_st_latency.cu && ./tmem_st_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: st x8 x2 + wait ; 1: ld x32 + wait ; 2: st x8 x2, wait only every 4th iteration
__global__ void __launch_bounds__(512, 1) k_tmem(long long* out, int iters, int active_warps) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp < active_warps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
        uint32_t v = lane;
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (MODE == 1) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(base)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r[0] + r[31];
            } else {
                for (int h = 0; h < 2; ++h)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(base + 8u * h), "r"(v),
                                 "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7)
                                 : "memory");
                if (MODE == 0 || (it & 3) == 3) {
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
                v += 3;
            }
        }
        const long long t1 = clock64();
        if (lane == 0) out[blockIdx.x * 16 + warp] = (t1 - t0) + (acc == 0xdeadbeef);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* what, int active) {
    long long* d;
    cudaMalloc(&d, 148 * 16 * sizeof(long long));
    const int iters = 1000;
    for (int rep = 0; rep < 2; ++rep) k_tmem<MODE><<<148, 512>>>(d, iters, active);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int b = 0; b < 148; ++b)
        for (int w = 0; w < active; ++w) s += (double)h[b * 16 + w];
    printf("%-44s warps=%2d : %7.1f cycles / iteration %s\n", what, active, s / (148.0 * active) / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int a : {1, 4, 8, 16}) run<0>("2 x tcgen05.st.x8 + wait::st + fence", a);
    for (int a : {1, 16}) run<2>("2 x tcgen05.st.x8, wait::st every 4th", a);
    for (int a : {1, 4, 8}) run<1>("tcgen05.ld.x32 + wait::ld", a);
    return 0;
}
