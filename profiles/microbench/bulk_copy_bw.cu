// Microbenchmark: global(L2-resident) -> shared bandwidth of 1-D cp.async.bulk on sm_100a, per SM and chip-wide, as a
// function of the piece size, the bytes in flight per SM and the number of SMs streaming at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_bw bulk_copy_bw.cu && ./bulk_copy_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread per CTA streams `total` bytes from src (wrapping inside a `window`-byte region, all CTAs read the SAME region
// when shared != 0) into a ring of `stages` buffers of `stage_bytes`, each filled by pieces of `piece` bytes
__global__ void __launch_bounds__(128, 1) k_stream(const uint8_t* src, size_t window, int shared, uint32_t stage_bytes, int stages,
                                                   uint32_t piece, int iters, long long* cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t* base = src + (shared ? 0 : ((size_t)blockIdx.x * 7919 * 128) % (window / 2));
        size_t off = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters + stages; ++it) {
            const int s = it % stages;
            if (it >= stages) {          // wait for the fill issued `stages` iterations ago
                const uint32_t ph = ((it / stages) - 1) & 1;
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(ph) : "memory");
            }
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(stage_bytes) : "memory");
                for (uint32_t o = 0; o < stage_bytes; o += piece) {
                    const uint32_t nb = min(piece, stage_bytes - o);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(smem + (size_t)s * stage_bytes + o)),
                                 "l"(base + off + o), "r"(nb), "r"(smem_u32(&bar[s]))
                                 : "memory");
                }
                off += stage_bytes;
                if (off + stage_bytes > window / 2) off = 0;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const size_t window = 32u << 20;          // 32 MB: L2 resident
    uint8_t* src;
    cudaMalloc(&src, window);
    cudaMemset(src, 1, window);
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    struct Cfg { int ctas, shared; uint32_t stage, piece; int stages; };
    const Cfg cfgs[] = {
        {1, 1, 53248, 8192, 4},   {148, 1, 53248, 8192, 4},  {148, 0, 53248, 8192, 4},  {128, 1, 53248, 8192, 4},
        {148, 1, 53248, 53248, 4}, {148, 1, 53248, 2048, 4}, {148, 1, 26624, 8192, 8}, {148, 1, 28672, 28672, 4},
        {74, 1, 53248, 8192, 4},  {37, 1, 53248, 8192, 4},   {1, 1, 53248, 53248, 4},
    };
    for (const Cfg& c : cfgs) {
        const int iters = 256;
        for (int rep = 0; rep < 2; ++rep) k_stream<<<c.ctas, 128, (size_t)c.stage * c.stages>>>(src, window, c.shared, c.stage, c.stages, c.piece, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, d, sizeof(long long) * c.ctas, cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < c.ctas; ++i) s += (double)h[i];
        const double cyc = s / c.ctas, bytes = (double)iters * c.stage;
        printf("CTAs=%3d %s stage=%5u piece=%5u stages=%d : %6.1f B/cycle/SM  -> %6.2f KB/cycle chip  (%.2f TB/s at 1.965 GHz) %s\n", c.ctas,
               c.shared ? "same data " : "distinct  ", c.stage, c.piece, c.stages, bytes / cyc, bytes / cyc * c.ctas / 1024.0,
               bytes / cyc * c.ctas * 1.965e9 / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
