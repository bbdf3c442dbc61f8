// Microbenchmark: issue cadence of tcgen05.mma kind::tf32 (M=128, K=8) on sm_100a as a function of N, of the number of
// independent accumulators the MMAs are interleaved over, and of where the A operand lives (TMEM = TS mode, smem = SS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_cadence mma_cadence.cu && ./mma_cadence
// Prints cycles per MMA (SM clock) measured on one CTA per SM; the operands are zeros (timing does not depend on data).
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
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

template <int N, int NACC, bool TS, int NOISE = 0>     // NOISE 1: the other warps stream tcgen05.st, 2: tcgen05.ld, 3: smem stores
__global__ void __launch_bounds__(640, 1) k_cadence(long long* out, int rounds) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) {
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
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_tf32(128, N);
        const uint64_t bdesc = make_desc(smem_u32(smem), N * 16u, 128u);            // B: [2 k-halves][N rows][16 B]
        const uint64_t adesc = make_desc(smem_u32(smem) + 32768u, 128 * 16u, 128u);   // A (SS): [2][128 rows][16 B]
        const uint32_t a_tmem = tmem + 480;                                          // A (TS): 8 columns
        t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                const uint32_t d = tmem + (uint32_t)(j * N);
                if (TS) mma_ts(d, a_tmem, bdesc, idesc, 1u); else mma_ss(d, adesc, bdesc, idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        *reinterpret_cast<volatile int*>(&tmem_slot) = -1;      // tells the noise warps to stop
    } else if (NOISE != 0 && warp >= 4) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 464u + 8u * ((warp >> 2) & 1);
        uint32_t v = threadIdx.x, acc = 0;
        while (*reinterpret_cast<volatile int*>(&tmem_slot) != -1) {
            if (NOISE == 1) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(base), "r"(v), "r"(v + 1),
                             "r"(v + 2), "r"(v + 3), "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else if (NOISE == 2) {
                uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(base) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r0 + r7;
            } else {
                reinterpret_cast<volatile uint32_t*>(smem + 49152)[threadIdx.x] = v;
            }
            v += 3;
        }
        if (acc == 0xdeadbeef) out[0] = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int NACC, bool TS, int NOISE = 0>
void run(const char* name) {
    static_assert(N * NACC <= 448, "accumulators must fit beside the A columns");
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int rounds = 2048 / NACC;
    cudaFuncSetAttribute(k_cadence<N, NACC, TS, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) k_cadence<N, NACC, TS, NOISE><<<148, NOISE ? 640 : 128, 64 * 1024>>>(d, rounds);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; ++i) s += (double)h[i];
    printf("%-13s N=%3d accumulators=%d : %6.1f cycles / MMA   (floor 128*N/256 = %d)  %s\n", name, N, NACC, s / 148 / (rounds * NACC), N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<112, 1, true, 1>("TS+st noise"); run<112, 1, true, 2>("TS+ld noise"); run<112, 1, true, 3>("TS+sts noise");
    run<64, 2, true, 2>("TS+ld noise"); run<64, 2, true, 1>("TS+st noise");
    run<64, 1, true>("TS"); run<64, 2, true>("TS"); run<64, 4, true>("TS"); run<32, 4, true>("TS"); run<32, 8, true>("TS");
    run<128, 1, true>("TS"); run<128, 2, true>("TS"); run<112, 1, true>("TS"); run<112, 2, true>("TS"); run<256, 1, true>("TS");
    run<64, 1, false>("SS"); run<64, 2, false>("SS"); run<64, 4, false>("SS"); run<128, 2, false>("SS"); run<256, 1, false>("SS");
    return 0;
}
