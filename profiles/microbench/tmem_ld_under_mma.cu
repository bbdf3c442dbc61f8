// Microbenchmark: does tcgen05.ld (the epilogue's accumulator drain) keep its rate while tcgen05.mma runs?
// One thread issues kind::f16 MMAs (M = 128, N = 128, TS mode) back to back into TMEM columns [0, 256); 8 warps drain
// columns [256, 384) in a loop (2 x tcgen05.ld.x32 + wait per iteration, as the forward kernel's epilogue does) until the
// MMA thread is done.  Printed: cycles per MMA, drain iterations per warp, drained bytes per cycle and SM - with the MMAs
// on and off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_under_mma tmem_ld_under_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
                 "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(id) : "memory");
}

template <int MODE>      // 0: drain only, 1: MMAs (TS) + drain, 2: MMAs (SS) + drain, 3: MMAs (TS) only
__global__ void __launch_bounds__(384, 1) k(long long* out, int rounds) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003800u;
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t id = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bdesc = make_desc(smem_u32(smem), 128 * 16u, 128u);
        const uint64_t adesc = make_desc(smem_u32(smem) + 32768u, 128 * 16u, 128u);
        const long long t0 = clock64();
        if (MODE != 0) {
            for (int r = 0; r < rounds; ++r) {
                if (MODE == 2) { mma_f16_ss(tmem, adesc, bdesc, id); mma_f16_ss(tmem + 128, adesc, bdesc, id); }
                else { mma_f16_ts(tmem, tmem + 480, bdesc, id); mma_f16_ts(tmem + 128, tmem + 480, bdesc, id); }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        } else {
            while (clock64() - t0 < 64LL * 2 * rounds) { }
        }
        out[3 * blockIdx.x] = clock64() - t0;
        stop = 1;
    } else if (warp >= 4 && MODE != 3) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u + 64u * ((warp - 4) >> 2);
        long long iters = 0;
        float acc = 0.f;
        while (!stop) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(base + 32u * h)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int x = 0; x < 32; ++x) acc = fmaf(__uint_as_float(r[x]), 1.0001f, acc);
            }
            ++iters;
        }
        if (lane == 0) out[3 * blockIdx.x + 1 + ((warp - 4) >> 2)] = iters + (acc == 123.f ? 1 : 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int rounds) {
    long long* d;
    cudaMalloc(&d, 3 * 148 * sizeof(long long));
    cudaMemset(d, 0, 3 * 148 * sizeof(long long));
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 384, 64 * 1024>>>(d, rounds);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3 * 148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0, it = 0;
    for (int i = 0; i < 148; ++i) { cyc += (double)h[3 * i]; it += 0.5 * (double)(h[3 * i + 1] + h[3 * i + 2]); }
    cyc /= 148; it /= 148;
    // per iteration every one of the 8 warps drains 2 x 4 KB
    printf("%-28s %8.1f cycles / MMA ; drain: %7.1f iterations per warp = %6.1f cycles per iteration, %6.1f B / cycle / SM  %s\n", name,
           cyc / (2.0 * rounds), it, it > 0 ? cyc / it : 0.0, it * 8 * 8192.0 / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<0>("drain only", 4000);
    run<3>("MMAs (TS) only", 4000);
    run<1>("MMAs (TS) + drain", 4000);
    run<2>("MMAs (SS) + drain", 4000);
    return 0;
}
