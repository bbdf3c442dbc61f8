// Kernel 2 (tensor-core contraction path, sm_100a): the bilinear contractions of the decoder on tcgen05 with TMEM
// accumulators, fp32-accurate through an error-compensated 3xTF32 split.
//
//   reference: weighted_R = T.tensordot(relation_probs, R, axes=[[1],[2]])   learning/models/decoders/Bilinear.py:33
//              weightedC  (same)                                             learning/models/decoders/BilinearPlusSP.py:37
//              weightedC1/C2 = T.dot(relation_probs, C1.T / C2.T)            BilinearPlusSP.py:35-36, SelectionalPreferences.py:31-32
//              batched_tensordot / batched_dot consumers                     Bilinear.py:58-59,68-69,78-79
//              and T.grad through them                                       learning/Optimizers.py:27
//
// Three GEMM-shaped contractions per step, none of which may materialise [B, d*d] in HBM (64 KiB per example at d=128):
//   forward / backward-recompute  M_b = sum_k q_bk C[:,:,k]      D[b, n]  = sum_k P[b,k]  Cf[n,k]     -> v = M R, w = M^T L
//   backward dq                   dq_bk = <dM_b, C[:,:,k]>        D[b, k]  = sum_n G[b,n]  Cf[n,k]     (G = a R^T + L Y2^T, rank 2)
//   backward dC                   dC[n,k] = sum_b dM_b[n] q_bk    D[n, k]  = sum_b G[b,n]  q[b,k]
// fp32 parity: every operand x is split x = hi + lo (both TF32-exact) and each k-step issues hi.hi + hi.lo + lo.hi
// (3 tcgen05.mma kind::tf32, fp32 accumulation in TMEM; dropped lo.lo ~ 2^-22 relative).
//
// Operand placement is dictated by the measured shared-memory -> tensor-core feed (~64 B/cycle/SM): with M = 128 an SS
// MMA re-reads 4 KB of A per 8-deep k-step and runs at a third of the math rate.  So the M-side (A) operand lives in
// TMEM (TS mode): the threads that own a TMEM lane (= an example row b, or an output row n) write their row with
// tcgen05.st - q rows once per CTA for the forward, the generated operand G chunk by chunk for the backward - and only the
// small N-side (B) operand streams through shared memory.  B operands are pre-split and pre-arranged in HBM as the exact
// shared-memory image of the no-swizzle K-major canonical layout (float4 planes T[kq][row]; LBO = rows*16 B, SBO = 128 B),
// so a stage is filled by 1-D bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier.
//
// Warp roles (all kernels): warp 0 = bulk-copy producer, 1 = MMA issuer (warp-uniform loop, elect.sync lane issues),
// 2 = TMEM allocator, 3 = idle, 4.. = row-owning workers (epilogue / operand generators); worker warp w touches TMEM
// lanes 32*(w%4)..+31 as the hardware requires.
#include <stdlib.h>

#include <algorithm>

#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

namespace {

// optional in-kernel timeline (build with -DRAE_TRACE): SM cycle counter of CTA-local milestones, [CTA][64] slots
#ifdef RAE_TRACE
__device__ unsigned long long* g_tc_trace = nullptr;
// TC_TRACE_INIT() once per thread (reads the buffer pointer into a register: a timestamp then costs one clock read and
// one asynchronous store, not a dependent global load)
#define TC_TRACE_INIT() unsigned long long* const tc_trace_ptr_ = g_tc_trace
#define TC_TRACE(slot)                                                                                         \
    do {                                                                                                       \
        if (tc_trace_ptr_ != nullptr && (slot) < 64) tc_trace_ptr_[(size_t)blockIdx.x * 64 + (slot)] = clock64(); \
    } while (0)
#else
#define TC_TRACE_INIT() do { } while (0)
#define TC_TRACE(slot) do { } while (0)
#endif

__device__ int g_tc_bwd_dbg = 0;   // measurement knobs of the backward kernels (RAE_TC_DEBUG bits 8: no MMAs, 16: generators do not store)

constexpr int TC_M = 128;          // rows per CTA (TMEM lanes)
constexpr int TC_N = 64;           // forward: B-operand rows per chunk (TMEM columns per accumulator stage)
constexpr int TC_TSTAGES = 4;      // forward: accumulator stages in TMEM
constexpr int TC_BSTAGES = 4;      // B-operand smem stages
constexpr int TC_NC = 32;          // backward: reduction rows per chunk (4 k-steps of 8)
constexpr int TC_FWD_THREADS = 384;    // 4 control + 8 epilogue warps
constexpr int TC_BWD_THREADS = 640;    // 4 control + 16 generator warps
constexpr uint32_t TC_FWD_ACOL = 256;  // forward: first TMEM column of the resident P operand (hi, then lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_pieces(uint8_t* dst_smem, const uint8_t* src_gmem, uint32_t bytes, uint64_t* bar) {
    constexpr uint32_t PIECE = 1u << 16;    // one instruction per stage: each bulk copy costs ~85 cycles + bytes / 130 B per cycle (profiles/microbench)
    for (uint32_t off = 0; off < bytes; off += PIECE) bulk_g2s(dst_smem + off, src_gmem + off, min(PIECE, bytes - off), bar);
}
// ---- thread-block clusters: the B operand of a chunk is fetched ONCE per cluster (each CTA loads 1/cs of it and multicasts
// the slice into every CTA's stage) instead of once per CTA: the contractions are bound by L2 -> SM operand traffic
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// slice `crank` of a chunk -> the same stage offset of every CTA of the cluster (cs == 1: plain copy of the whole chunk)
__device__ __forceinline__ void bulk_g2s_chunk(uint8_t* stage, const uint8_t* chunk, uint32_t bytes, uint64_t* bar, uint32_t cs,
                                               uint32_t crank) {
    constexpr uint32_t PIECE = 1u << 16;    // one instruction per stage: each bulk copy costs ~85 cycles + bytes / 130 B per cycle (profiles/microbench)
    const uint32_t slice = bytes / cs, base = slice * crank;
    const uint16_t mask = (uint16_t)((1u << cs) - 1u);
    for (uint32_t off = 0; off < slice; off += PIECE) {
        const uint32_t nb = min(PIECE, slice - off);
        if (cs == 1) bulk_g2s(stage + base + off, chunk + base + off, nb, bar);
        else bulk_g2s_mc(stage + base + off, chunk + base + off, nb, bar, mask);
    }
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// one lane of a converged warp (the loops around it stay warp-uniform, so descriptor math lives in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32 (TS mode: A rows = TMEM lanes, one 32-bit element per column)
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread t = lane base + t)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 registers per thread -> 32 lanes x 8 columns of TMEM (thread t writes lane base + t)
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 consecutive floats of one accumulator row -> global memory; 16-byte stores when the row is 16-byte aligned
__device__ __forceinline__ void store_row32(float* o, const float (&t)[32], int nvalid, bool vec) {
    if (vec) {
#pragma unroll
        for (int x = 0; x < 32; x += 4)
            if (x < nvalid) *reinterpret_cast<float4*>(o + x) = make_float4(t[x], t[x + 1], t[x + 2], t[x + 3]);
    } else {
#pragma unroll
        for (int x = 0; x < 32; ++x)
            if (x < nvalid) o[x] = t[x];
    }
}

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous);
// LBO = byte distance between the two 16-byte K-halves of one MMA k-step, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    return d;                   // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
// descriptor with the start address advanced by `bytes` (the address field is the low 14 bits, in 16-byte units)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense, no negate
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// hot-loop split (operand generators, 16 warps per SM, instruction-bound): hi = round-to-nearest-away TF32 of a FINITE x as
// two integer ops (cvt.rna.tf32 compiles to three: it also guards inf / nan), lo = x - hi (exact) with its low 13 bits
// cleared (truncation: |error| < 2^-23 |x|, below the dropped lo*lo term)
__device__ __forceinline__ void split8(const float (&x)[8], float (&hi)[8], float (&lo)[8]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        hi[u] = __uint_as_float((__float_as_uint(x[u]) + 0x1000u) & 0xffffe000u);
        lo[u] = __uint_as_float(__float_as_uint(x[u] - hi[u]) & 0xffffe000u);
    }
}
__device__ __forceinline__ void split4(const float (&x)[4], float4& hi, float4& lo) {
    hi.x = tf32_hi(x[0]); hi.y = tf32_hi(x[1]); hi.z = tf32_hi(x[2]); hi.w = tf32_hi(x[3]);
    lo.x = tf32_hi(x[0] - hi.x); lo.y = tf32_hi(x[1] - hi.y); lo.z = tf32_hi(x[2] - hi.z); lo.w = tf32_hi(x[3] - hi.w);
}

// row n of the dense operand Cf -> source [K]-vector (nullptr = zero padding row)
//   n <  n_bil_rows : C[i, j, :] with i = n / DP, j = n % DP
//   then DP rows of C1[j,:] and DP rows of C2[j,:]
__device__ __forceinline__ const float* cf_row(const float* C, const float* C1, const float* C2, int d, int K, int DP,
                                               int n_bil_rows, int n) {
    if (n < n_bil_rows) {
        const int i = n / DP, j = n - i * DP;
        return (i < d && j < d && C != nullptr) ? C + ((size_t)i * d + j) * K : nullptr;
    }
    const int m = n - n_bil_rows;
    const int which = m / DP, j = m - which * DP;
    if (j >= d) return nullptr;
    const float* src = which == 0 ? C1 : C2;
    return src != nullptr ? src + (size_t)j * K : nullptr;
}

// ------------------------------------------------------------------------------------------------------------
// operand preparation (per step; the dense parameters change every step)
// ------------------------------------------------------------------------------------------------------------
// forward B operand: [chunk of 64 rows n][hi/lo][kq][row] float4, 4 consecutive relations per float4
__global__ void __launch_bounds__(256) k_tc_prep_c(const float* __restrict__ C, const float* __restrict__ C1,
                                                   const float* __restrict__ C2, int d, int K, int KQ, int DP, int n_bil_rows,
                                                   int n_rows_total, float4* __restrict__ out, int NK, float4* __restrict__ out2) {
    if (blockIdx.y == 1) {
        // second half of the grid: the transposed operand of the dq contraction
        // out2[c32][hi/lo][nq 0..7][krow 0..NK-1] = (Cf[32c+4nq+0..3][krow])
        const size_t total2 = (size_t)(n_rows_total / 4) * NK;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (size_t)gridDim.x * blockDim.x) {
            const int krow = (int)(idx % NK);
            const int nq_g = (int)(idx / NK);
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* src = cf_row(C, C1, C2, d, K, DP, n_bil_rows, 4 * nq_g + u);
                x[u] = (src != nullptr && krow < K) ? src[krow] : 0.f;
            }
            float4 hi, lo;
            split4(x, hi, lo);
            const int c32 = nq_g / 8, nq = nq_g - c32 * 8;
            float4* base = out2 + (size_t)c32 * 2 * 8 * NK;
            base[(size_t)nq * NK + krow] = hi;
            base[(size_t)(8 + nq) * NK + krow] = lo;
        }
        return;
    }
    const size_t total = (size_t)n_rows_total * KQ;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int kq = (int)(idx % KQ);
        const int n = (int)(idx / KQ);
        const float* src = cf_row(C, C1, C2, d, K, DP, n_bil_rows, n);
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * kq + u;
            x[u] = (src != nullptr && k < K) ? src[k] : 0.f;
        }
        float4 hi, lo;
        split4(x, hi, lo);
        const int chunk = n / TC_N, r = n - chunk * TC_N;
        float4* base = out + (size_t)chunk * 2 * KQ * TC_N;
        base[(size_t)kq * TC_N + r] = hi;
        base[(size_t)(KQ + kq) * TC_N + r] = lo;
    }
}

// dq B operand: Cf transposed, [chunk of 32 rows n][hi/lo][nq 0..7][krow 0..NK-1] float4 = (Cf[32c+4nq+0..3][krow])
__global__ void __launch_bounds__(256) k_tc_prep_ct(const float* __restrict__ C, const float* __restrict__ C1,
                                                    const float* __restrict__ C2, int d, int K, int NK, int DP, int n_bil_rows,
                                                    int n_rows_total, float4* __restrict__ out) {
    const size_t total = (size_t)(n_rows_total / 4) * NK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int krow = (int)(idx % NK);
        const int nq_g = (int)(idx / NK);
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* src = cf_row(C, C1, C2, d, K, DP, n_bil_rows, 4 * nq_g + u);
            x[u] = (src != nullptr && krow < K) ? src[krow] : 0.f;
        }
        float4 hi, lo;
        split4(x, hi, lo);
        const int c32 = nq_g / 8, nq = nq_g - c32 * 8;
        float4* base = out + (size_t)c32 * 2 * 8 * NK;
        base[(size_t)nq * NK + krow] = hi;
        base[(size_t)(8 + nq) * NK + krow] = lo;
    }
}

// dC B operand: q transposed, [chunk of 32 examples][hi/lo][bq 0..7][krow 0..NK-1] float4 = (q[32c+4bq+0..3][krow])
__global__ void __launch_bounds__(256) k_tc_prep_qt(const float* __restrict__ q, int B, int K, int NK, float4* __restrict__ out,
                                                    const float* __restrict__ A, const int32_t* __restrict__ a1,
                                                    const int32_t* __restrict__ a2, int d, int dp, int quirk, float* __restrict__ ev) {
    if (blockIdx.y == 1) {
        // second half of the grid: L = A[a1], R = A[a2] (A[a1] with the model-C quirk) -> ev, one warp per example
        const int lane = threadIdx.x & 31;
        for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += (gridDim.x * blockDim.x) >> 5) {
            const int r1 = a1[b], r2 = quirk ? r1 : a2[b];
            float* o = ev + (size_t)b * E_NV * dp;
            for (int j = lane; j < d; j += 32) {
                o[E_L * dp + j] = ld_nc(A + (size_t)r1 * d + j);
                o[E_R * dp + j] = ld_nc(A + (size_t)r2 * d + j);
            }
        }
        return;
    }
    const int nbc = (B + TC_NC - 1) / TC_NC;
    const size_t total = (size_t)nbc * 8 * NK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int krow = (int)(idx % NK);
        const int bq = (int)((idx / NK) % 8);
        const int bc = (int)(idx / ((size_t)8 * NK));
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b = bc * TC_NC + 4 * bq + u;
            x[u] = (b < B && krow < K) ? q[(size_t)b * K + krow] : 0.f;
        }
        float4 hi, lo;
        split4(x, hi, lo);
        float4* base = out + (size_t)bc * 2 * 8 * NK;
        base[(size_t)bq * NK + krow] = hi;
        base[(size_t)(8 + bq) * NK + krow] = lo;
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward contraction (also used for the backward recompute with L := a, R := c)
// TMEM map: accumulator stages [0,256) (4 x 64 columns), P operand hi at [256, 256+Kp), lo at [256+Kp, 256+2Kp)
// ------------------------------------------------------------------------------------------------------------
struct TcArgs {
    const float* q;         // [B,K]
    const float4* bop;      // B operand chunks
    const float* ev;        // per-example vectors (L at slotL, R at slotR), row stride E_NV*dp
    float* ev_out;          // SP chunks write c1 / c2 here (E_C1 / E_C2)
    float* vg;              // [2][B][dp]   v partial per epilogue group
    float* wp;              // [NS][2][B][dp] w partial per (split, group)
    int B, K, d, dp, KQ;
    int slotL, slotR;
    int n_bil_chunks;       // chunks holding bilinear rows
    int n_sp_chunks;        // chunks holding C1/C2 rows (forward only)
    int NS;                 // splits of the bilinear chunk range per tile
    int cs;                 // cluster size (1, 2 or 4): CTAs of a cluster = consecutive tiles of the SAME split
    int ntile;
    int dbg;                // measurement knobs (RAE_TC_DEBUG): 1 no operand copies after the first fills, 2 no MMAs, 4 no epilogue math
};

template <int DP>
__global__ void __launch_bounds__(TC_FWD_THREADS, 1) k_tc_bilinear(TcArgs p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TC_TRACE_INIT();
    const int split = blockIdx.x / p.ntile, tile = blockIdx.x - split * p.ntile;     // a cluster = cs consecutive tiles
    const uint32_t cs = (uint32_t)p.cs, crank = cs > 1 ? cluster_ctarank() : 0u;
    const uint32_t B_BYTES = 2u * p.KQ * TC_N * 16u;
    const uint32_t Kp = 4u * p.KQ;                          // relations padded to a multiple of 8
    uint8_t* smB = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TC_BSTAGES * B_BYTES);
    uint64_t* a_full = bars;                                // 8 epilogue-warp arrivals: P operand is in TMEM
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + TC_BSTAGES;
    uint64_t* t_full = b_empty + TC_BSTAGES;
    uint64_t* t_empty = t_full + TC_TSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + TC_TSTAGES);

    // chunk range of this CTA: the bilinear chunks are split NS ways at even boundaries, the last split also takes SP
    int c_begin, c_end;
    {
        const int pairs = (p.n_bil_chunks + 1) / 2;
        const int per = (pairs + p.NS - 1) / p.NS;
        c_begin = min(2 * per * split, p.n_bil_chunks);
        c_end = min(2 * per * (split + 1), p.n_bil_chunks);
        if (split == p.NS - 1) c_end = p.n_bil_chunks + p.n_sp_chunks;
    }
    const int nit = c_end - c_begin;
    if (threadIdx.x == 0) TC_TRACE(0);

    if (threadIdx.x == 0) {
        mbar_init(a_full, 8);
        // a stage is free once EVERY CTA of the cluster has consumed it (its next fill is multicast into all of them)
        for (int s = 0; s < TC_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], cs); }
        for (int s = 0; s < TC_TSTAGES; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();       // every CTA's barriers exist before a peer multicasts into them / arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_TRACE(1);

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    }
    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            for (int it = 0; it < nit; ++it) {
                const int s = it % TC_BSTAGES;
                const uint32_t ph = (it / TC_BSTAGES) & 1;
                mbar_wait(&b_empty[s], ph ^ 1);
                if ((p.dbg & 1) && it >= TC_BSTAGES) {      // measurement only: stage keeps its old contents
                    mbar_arrive(&b_full[s]);
                    continue;
                }
                mbar_expect_tx(&b_full[s], B_BYTES);
                bulk_g2s_chunk(smB + (size_t)s * B_BYTES, reinterpret_cast<const uint8_t*>(p.bop) + (size_t)(c_begin + it) * B_BYTES,
                               B_BYTES, &b_full[s], cs, crank);
                if (it == 0) TC_TRACE(2);
            }
            TC_TRACE(3);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
        if (nit > 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, TC_N);
            const uint32_t a_hi = tmem_base + TC_FWD_ACOL, a_lo = a_hi + Kp;
            uint64_t dbh0[TC_BSTAGES], dbl0[TC_BSTAGES];
#pragma unroll
            for (int s = 0; s < TC_BSTAGES; ++s) {
                const uint32_t b_hi = smem_u32(smB + (size_t)s * B_BYTES);
                dbh0[s] = make_desc(b_hi, TC_N * 16u, 128u);
                dbl0[s] = make_desc(b_hi + p.KQ * TC_N * 16u, TC_N * 16u, 128u);
            }
            const int ksteps = p.KQ / 2;
            mbar_wait(a_full, 0);
            tc_fence_after();
            if (lane == 0) TC_TRACE(4);
            // Two chunks are issued INTERLEAVED (different accumulator stages): consecutive MMAs into one accumulator are
            // a dependent chain whose latency (~60 cycles, measured) exceeds the 32-cycle issue slot of a 128x64x8 MMA.
            for (int it = 0; it < nit; it += 2) {
                const bool two = it + 1 < nit;
                const int s0 = it % TC_BSTAGES, ts0 = it % TC_TSTAGES;
                const int s1 = (it + 1) % TC_BSTAGES, ts1 = (it + 1) % TC_TSTAGES;
                mbar_wait(&t_empty[ts0], ((it / TC_TSTAGES) & 1) ^ 1);
                if (lane == 0 && it < 4) TC_TRACE(8 + 2 * it);
                mbar_wait(&b_full[s0], (it / TC_BSTAGES) & 1);
                if (two) {
                    mbar_wait(&t_empty[ts1], (((it + 1) / TC_TSTAGES) & 1) ^ 1);
                    mbar_wait(&b_full[s1], ((it + 1) / TC_BSTAGES) & 1);
                }
                if (lane == 0 && it < 4) TC_TRACE(9 + 2 * it);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(ts0 * TC_N), d1 = tmem_base + (uint32_t)(ts1 * TC_N);
                if (elect_one()) {
                    uint64_t h0 = dbh0[s0], l0 = dbl0[s0], h1 = dbh0[s1], l1 = dbl0[s1];
                    for (int ks = 0; ks < ((p.dbg & 2) ? 0 : ksteps); ++ks) {
                        const uint32_t acc = ks > 0 ? 1u : 0u;
                        tc_mma_tf32_ts(d0, a_hi + 8u * ks, h0, idesc, acc);                   // hi * hi
                        if (two) tc_mma_tf32_ts(d1, a_hi + 8u * ks, h1, idesc, acc);
                        tc_mma_tf32_ts(d0, a_hi + 8u * ks, l0, idesc, 1u);                    // hi * lo
                        if (two) tc_mma_tf32_ts(d1, a_hi + 8u * ks, l1, idesc, 1u);
                        tc_mma_tf32_ts(d0, a_lo + 8u * ks, h0, idesc, 1u);                    // lo * hi
                        if (two) tc_mma_tf32_ts(d1, a_lo + 8u * ks, h1, idesc, 1u);
                        h0 = desc_advance(h0, 2u * TC_N * 16u);
                        l0 = desc_advance(l0, 2u * TC_N * 16u);
                        h1 = desc_advance(h1, 2u * TC_N * 16u);
                        l1 = desc_advance(l1, 2u * TC_N * 16u);
                    }
                    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
                    // smem stages reusable once these MMAs have read them (signalled to every CTA of the cluster)
                    if (cs > 1) tc_commit_mc(&b_empty[s0], cmask); else tc_commit(&b_empty[s0]);
                    tc_commit(&t_full[ts0]);     // accumulators complete
                    if (two) {
                        if (cs > 1) tc_commit_mc(&b_empty[s1], cmask); else tc_commit(&b_empty[s1]);
                        tc_commit(&t_full[ts1]);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== row-owning warps: P operand -> TMEM, then epilogue =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
        const int ew = warp - 4, g = ew >> 2, q4 = ew & 3;
        const int row = q4 * 32 + lane;
        const int b = tile * TC_M + row;
        const bool ok = b < p.B;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        {
            // q row -> TF32 hi / lo planes of the A operand in TMEM.  Group g converts relations [56 g, 56 g + 56): all of a
            // thread's loads are issued before the first conversion (one memory latency) and the code stays small
            // (straight-line code that runs once is paid in instruction-cache misses).
            const float* qr = p.q + (size_t)(ok ? b : 0) * p.K;
            const int kb = 56 * g;
            float qh[56];
            if ((p.K & 3) == 0) {
                const float4* q4 = reinterpret_cast<const float4*>(qr + kb);
#pragma unroll
                for (int i = 0; i < 14; ++i) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && kb + 4 * i < p.K) v = q4[i];
                    qh[4 * i] = v.x; qh[4 * i + 1] = v.y; qh[4 * i + 2] = v.z; qh[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 56; ++i) qh[i] = (ok && kb + i < p.K) ? qr[kb + i] : 0.f;
            }
            const uint32_t a_hi_col = lane_base + TC_FWD_ACOL + (uint32_t)kb, a_lo_col = a_hi_col + Kp;
#pragma unroll
            for (int c8 = 0; c8 < 7; ++c8) {
                if ((uint32_t)(kb + 8 * c8) < Kp) {
                    float x[8], hi[8], lo[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = qh[8 * c8 + u];
                    split8(x, hi, lo);
                    tc_st8(a_hi_col + 8u * c8, hi);
                    tc_st8(a_lo_col + 8u * c8, lo);
                }
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
            if (warp == 4 && lane == 0) TC_TRACE(5);
        }
        const float* evb = p.ev + (size_t)(ok ? b : 0) * E_NV * p.dp;
        constexpr int RW = (DP >= 64) ? 64 : 32;      // columns of R / w held per thread
        const int jbase = (DP == 128) ? 64 * g : 0;
        float Rr[RW], Wr[RW];
#pragma unroll
        for (int c = 0; c < RW; ++c) {
            const int j = jbase + c;
            Rr[c] = (ok && j < p.d) ? evb[p.slotR * p.dp + j] : 0.f;
            Wr[c] = 0.f;
        }
        for (int it = g; it < nit; it += 2) {
            const int ts = it % TC_TSTAGES;
            const uint32_t tph = (it / TC_TSTAGES) & 1;
            const int c = c_begin + it;
            // L values this chunk needs (issued before the wait so the loads overlap it)
            float L0 = 0.f, L1 = 0.f;
            int i0 = 0;
            if (c < p.n_bil_chunks) {
                i0 = (DP == 128) ? (c >> 1) : (DP == 64 ? c : 2 * c);
                if (ok && i0 < p.d) L0 = evb[p.slotL * p.dp + i0];
                if (DP == 32 && ok && i0 + 1 < p.d) L1 = evb[p.slotL * p.dp + i0 + 1];
            }
            if (ew == 0 && lane == 0 && it < 8) TC_TRACE(24 + it);
            mbar_wait(&t_full[ts], tph);
            if (ew == 0 && lane == 0 && it < 8) TC_TRACE(25 + it);
            tc_fence_after();
            const bool bil = c < p.n_bil_chunks;
            const int sc = c - p.n_bil_chunks;
            float* o = p.ev_out + (size_t)(ok ? b : 0) * E_NV * p.dp;
            float vsum = 0.f;
            // the 64 accumulator columns are consumed in two halves of 32 to bound register pressure
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float t[32];
                tc_ld32(lane_base + (uint32_t)(ts * TC_N + 32 * hf), t);
                if (ew == 0 && lane == 0 && (it == 4 || it == 6)) TC_TRACE(16 + 2 * (it - 4) + hf);
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[ts]);     // accumulator stage free for the next MMA
                }
                if (p.dbg & 4) continue;
                if (bil) {
                    if (DP >= 64) {
                        float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            v4[x & 3] = fmaf(t[x], Rr[(32 * hf + x) % RW], v4[x & 3]);
                            Wr[(32 * hf + x) % RW] = fmaf(t[x], L0, Wr[(32 * hf + x) % RW]);
                        }
                        vsum += (v4[0] + v4[1]) + (v4[2] + v4[3]);
                    } else {
                        const float Lh = hf == 0 ? L0 : L1;
                        float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            v4[x & 3] = fmaf(t[x], Rr[x % RW], v4[x & 3]);
                            Wr[x % RW] = fmaf(t[x], Lh, Wr[x % RW]);
                        }
                        if (ok && i0 + hf < p.d) p.vg[((size_t)g * p.B + b) * p.dp + i0 + hf] = (v4[0] + v4[1]) + (v4[2] + v4[3]);
                    }
                } else if (ok) {
                    // selectional-preference rows: the accumulator row IS c1 / c2
                    if (DP == 128) {
                        const int slot = (sc < 2) ? E_C1 : E_C2, jb = 64 * (sc & 1) + 32 * hf;
#pragma unroll
                        for (int x = 0; x < 32; ++x)
                            if (jb + x < p.d) o[slot * p.dp + jb + x] = t[x];
                    } else if (DP == 64) {
                        const int slot = (sc == 0) ? E_C1 : E_C2;
#pragma unroll
                        for (int x = 0; x < 32; ++x)
                            if (32 * hf + x < p.d) o[slot * p.dp + 32 * hf + x] = t[x];
                    } else {
                        const int slot = hf == 0 ? E_C1 : E_C2;
#pragma unroll
                        for (int x = 0; x < 32; ++x)
                            if (x < p.d) o[slot * p.dp + x] = t[x];
                    }
                }
            }
            if (DP >= 64 && bil && ok && i0 < p.d) p.vg[((size_t)g * p.B + b) * p.dp + i0] = vsum;
            if (ew == 0 && lane == 0 && (it == 4 || it == 6)) TC_TRACE(18 + 2 * (it - 4));
        }
        if (ew == 0 && lane == 0) TC_TRACE(6);
        if (ok) {
            float* o = p.wp + (((size_t)split * 2 + g) * p.B + b) * p.dp;      // dp % 4 == 0: 16-byte stores
#pragma unroll
            for (int c = 0; c < RW; c += 4) {
                const int j = jbase + c;
                if (j < p.dp) *reinterpret_cast<float4*>(o + j) = make_float4(Wr[c], Wr[c + 1], Wr[c + 2], Wr[c + 3]);
            }
        }
        if (ew == 0 && lane == 0) TC_TRACE(7);
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();       // nobody leaves while a peer may still write into its stages / barriers
    if (threadIdx.x == 0) TC_TRACE(62);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        if (lane == 0) TC_TRACE(63);
    }
}

// v[b,i]: DP=128 both epilogue groups hold a half-row partial; DP=64 group i%2 produced it; DP=32 group (i/2)%2.
// w[b,j]: sum over splits of the group partials (DP=128: only group j/64 holds column j).  Nothing needs pre-zeroing.
__global__ void __launch_bounds__(256) k_tc_combine(const float* __restrict__ vg, const float* __restrict__ wp, float* __restrict__ ev,
                                                    int B, int d, int dp, int DP, int NS, int slotV, int slotW) {
    const size_t total = (size_t)B * dp;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx % dp);
        const size_t b = idx / dp;
        if (j >= d) continue;
        float v;
        if (DP == 128) v = vg[idx] + vg[total + idx];
        else if (DP == 64) v = vg[(size_t)(j & 1) * total + idx];
        else v = vg[(size_t)((j >> 1) & 1) * total + idx];
        float w = 0.f;
        if (DP == 128) {
            const int g = j >> 6;
            for (int s = 0; s < NS; ++s) w += wp[(size_t)(2 * s + g) * total + idx];
        } else {
            for (int s = 0; s < 2 * NS; ++s) w += wp[(size_t)s * total + idx];
        }
        ev[(b * E_NV + slotV) * dp + j] = v;
        ev[(b * E_NV + slotW) * dp + j] = w;
    }
}

// L = A[a1], R = A[a2] (A[a1] with the model-C quirk) -> ev
__global__ void __launch_bounds__(256) k_tc_gather_lr(const float* __restrict__ A, const int32_t* __restrict__ a1,
                                                      const int32_t* __restrict__ a2, int B, int d, int dp, int quirk,
                                                      float* __restrict__ ev) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int r1 = a1[b], r2 = quirk ? r1 : a2[b];
    float* o = ev + (size_t)b * E_NV * dp;
    for (int j = lane; j < d; j += 32) {
        o[E_L * dp + j] = ld_nc(A + (size_t)r1 * d + j);
        o[E_R * dp + j] = ld_nc(A + (size_t)r2 * d + j);
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward kernels: generated A operand in TMEM.
// TMEM map: accumulator [0,128) (NK <= 128 columns used); A stage s: hi at [128 + 64 s, +32), lo at [128 + 64 s + 32, +32)
// generated operand  G[b,n]:  bilinear row n=(i,j): a_bi R_bj + L_bi Y2_bj ;  C1 row j: a_bj + G2_b L_bj ;  C2 row j: c_bj + G1_b R_bj
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t TC_BWD_ACOL = 128;
constexpr int TC_ASTAGES = 4;          // generated-operand stages in TMEM (64 columns each: hi 32 + lo 32); 128 + 4*64 = 384 <= 512
// Second accumulator of the dC contraction at [384, 512).  tcgen05.mma adds into the fp32 accumulator with truncation, so
// a long chain of MMAs into ONE accumulator drifts towards zero by ~2e-8 of the sum per MMA (measured at the target shape,
// profiles/r01_accum_chain.md: 1536 MMAs -> 3.4e-5 of ||dC||_inf, 180 MMAs -> 3.2e-6).  The reduction over examples is
// therefore dealt over two accumulators (even / odd chunks, summed with a rounded fp32 add in the epilogue) and over
// enough batch splits (tc_init) that no accumulator takes more than TC_DC_MAX_CHAIN chunks of 12 MMAs.
constexpr uint32_t TC_BWD_ACC2 = 384;
constexpr int TC_DC_MAX_CHAIN = 32;    // chunks (of 32 examples, 12 MMAs each) per accumulator
constexpr int TC_DQ_MAX_CHAIN = 130;   // dq: chunks (of 32 reduction rows, 12 MMAs each) per CTA; random-sign terms drift less

// transposed copies aT[i][b], LT[i][b] so that lane = example reads of a_bi / L_bi are coalesced
__global__ void __launch_bounds__(256) k_tc_transpose_al(const float* __restrict__ ev, int B, int d, int dp, float* __restrict__ aT,
                                                         float* __restrict__ LT) {
    __shared__ float ta[32][33], tl[32][33];
    const int b0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows of 32
    for (int r = ty; r < 32; r += 8) {
        const int b = b0 + r, i = i0 + tx;
        const bool in = b < B && i < d;
        ta[r][tx] = in ? ev[((size_t)b * E_NV + E_A) * dp + i] : 0.f;
        tl[r][tx] = in ? ev[((size_t)b * E_NV + E_L) * dp + i] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, b = b0 + tx;
        if (i < dp && b < B) {
            aT[(size_t)i * B + b] = ta[tx][r];
            LT[(size_t)i * B + b] = tl[tx][r];
        }
    }
}

struct TcDqArgs {
    const float4* bop2;     // Cf^T chunks [c32][hi/lo][8][NK]
    const float* ev; const float* sc; const float* aT; const float* LT;
    float* dqp;             // [NS][B][NK]
    int B, d, dp, K, NK;
    int n_bil_rows, n_chunks32, NS;
    int cs, ntile;          // cluster size; CTAs of a cluster = consecutive example tiles of the SAME split (same streamed chunks)
};

// shared skeleton pieces of the two backward kernels ------------------------------------------------------------
struct BwdBars {
    uint64_t* a_full; uint64_t* a_empty; uint64_t* b_full; uint64_t* b_empty; uint64_t* acc_full; uint32_t* tmem_slot;
};

__device__ __forceinline__ BwdBars bwd_setup(uint8_t* smem_raw, uint32_t B_BYTES, int warp, uint32_t& tmem_base, uint32_t cs) {
    BwdBars br;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TC_BSTAGES * B_BYTES);
    br.a_full = bars;                         // [TC_ASTAGES] 16 generator-warp arrivals
    br.a_empty = bars + TC_ASTAGES;           // [TC_ASTAGES] tcgen05.commit
    br.b_full = bars + 2 * TC_ASTAGES;        // [TC_BSTAGES] bulk-copy tx
    br.b_empty = br.b_full + TC_BSTAGES;      // [TC_BSTAGES] tcgen05.commit
    br.acc_full = br.b_empty + TC_BSTAGES;
    br.tmem_slot = reinterpret_cast<uint32_t*>(br.acc_full + 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_ASTAGES; ++s) { mbar_init(&br.a_full[s], 16); mbar_init(&br.a_empty[s], 1); }
        // a stage is free once EVERY CTA of the cluster has consumed it (its next fill is multicast into all of them)
        for (int s = 0; s < TC_BSTAGES; ++s) { mbar_init(&br.b_full[s], 1); mbar_init(&br.b_empty[s], cs); }
        mbar_init(br.acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(br.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();
    tc_fence_after();
    tmem_base = *br.tmem_slot;
    return br;
}

__device__ __forceinline__ void bwd_producer(const BwdBars& br, uint8_t* smB, const uint8_t* src, uint32_t B_BYTES, int c_begin, int nit,
                                             uint32_t cs, uint32_t crank) {
    for (int it = 0; it < nit; ++it) {
        const int s = it % TC_BSTAGES;
        const uint32_t ph = (it / TC_BSTAGES) & 1;
        mbar_wait(&br.b_empty[s], ph ^ 1);
        mbar_expect_tx(&br.b_full[s], B_BYTES);
        bulk_g2s_chunk(smB + (size_t)s * B_BYTES, src + (size_t)(c_begin + it) * B_BYTES, B_BYTES, &br.b_full[s], cs, crank);
    }
}

__device__ __forceinline__ void bwd_mma(const BwdBars& br, uint8_t* smB, uint32_t B_BYTES, int NK, uint32_t tmem_base, int nit, int trace_base,
                                        uint32_t cs, int nacc = 1) {
    TC_TRACE_INIT();
    const uint32_t idesc = make_idesc_tf32(TC_M, NK);
    uint64_t dbh0[TC_BSTAGES], dbl0[TC_BSTAGES];
#pragma unroll
    for (int s = 0; s < TC_BSTAGES; ++s) {
        const uint32_t b_hi = smem_u32(smB + (size_t)s * B_BYTES);
        dbh0[s] = make_desc(b_hi, (uint32_t)NK * 16u, 128u);
        dbl0[s] = make_desc(b_hi + 8u * (uint32_t)NK * 16u, (uint32_t)NK * 16u, 128u);
    }
    for (int it = 0; it < nit; ++it) {
        const int s = it % TC_BSTAGES, as = it % TC_ASTAGES;
        const uint32_t ph = (it / TC_BSTAGES) & 1, aph = (it / TC_ASTAGES) & 1;
        mbar_wait(&br.a_full[as], aph);
        if ((threadIdx.x & 31) == 0 && it >= 40 && it < 43) TC_TRACE(trace_base + 2 * (it - 40));
        mbar_wait(&br.b_full[s], ph);
        if ((threadIdx.x & 31) == 0 && it >= 40 && it < 43) TC_TRACE(trace_base + 2 * (it - 40) + 1);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t a_hi = tmem_base + TC_BWD_ACOL + 64u * as, a_lo = a_hi + 32u;
            // chunk `it` accumulates into accumulator it % nacc; the first chunk of each accumulator overwrites it
            const uint32_t acc = tmem_base + ((nacc == 2 && (it & 1)) ? TC_BWD_ACC2 : 0u);
            uint64_t dbh = dbh0[s], dbl = dbl0[s];
#pragma unroll
            for (int ks = 0; ks < ((g_tc_bwd_dbg & 8) ? 0 : 4); ++ks) {
                tc_mma_tf32_ts(acc, a_hi + 8u * ks, dbh, idesc, (it >= nacc || ks > 0) ? 1u : 0u);
                tc_mma_tf32_ts(acc, a_hi + 8u * ks, dbl, idesc, 1u);
                tc_mma_tf32_ts(acc, a_lo + 8u * ks, dbh, idesc, 1u);
                dbh = desc_advance(dbh, 2u * (uint32_t)NK * 16u);
                dbl = desc_advance(dbl, 2u * (uint32_t)NK * 16u);
            }
            if (it >= 40 && it < 43) TC_TRACE(trace_base + 6 + (it - 40));     // MMAs of chunk `it` issued
            tc_commit(&br.a_empty[as]);
            if (cs > 1) tc_commit_mc(&br.b_empty[s], (uint16_t)((1u << cs) - 1u)); else tc_commit(&br.b_empty[s]);
            if (it == nit - 1) tc_commit(br.acc_full);
        }
        __syncwarp();
    }
}

// generator warp publishes its 8 columns of A stage `as`
__device__ __forceinline__ void bwd_publish(const BwdBars& br, uint32_t lane_base, int as, int cg, const float (&g)[8], int lane) {
    float hi[8], lo[8];
    split8(g, hi, lo);
    const uint32_t col = lane_base + TC_BWD_ACOL + 64u * as + 8u * cg;
    if (!(g_tc_bwd_dbg & 16)) {
        tc_st8(col, hi);
        tc_st8(col + 32u, lo);
        tc_wait_st();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&br.a_full[as]);
}

// dq: rows = examples.  Generator thread = (row b, octet cg of the chunk's 32 reduction rows); the R / Y2 values it needs
// for every bilinear chunk are cached in registers (X, Y), a_bi / L_bi come coalesced from the transposed copies.
template <int DP>
__global__ void __launch_bounds__(TC_BWD_THREADS, 1) k_tc_dq(TcDqArgs p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TC_TRACE_INIT();
    const int split = blockIdx.x / p.ntile, tile = blockIdx.x - split * p.ntile;
    const uint32_t cs = (uint32_t)p.cs, crank = cs > 1 ? cluster_ctarank() : 0u;
    const uint32_t B_BYTES = 2u * 8u * (uint32_t)p.NK * 16u;
    uint8_t* smB = smem_raw;
    uint32_t tmem_base;
    if (threadIdx.x == 0) TC_TRACE(32);
    const BwdBars br = bwd_setup(smem_raw, B_BYTES, warp, tmem_base, cs);
    if (threadIdx.x == 0) TC_TRACE(33);
    constexpr int JQ = DP / 32;                 // chunks per bilinear row i
    // split the chunk range at row-i boundaries
    const int per = ((p.n_chunks32 + p.NS - 1) / p.NS + JQ - 1) / JQ * JQ;
    const int c_begin = min(per * split, p.n_chunks32), c_end = min(per * (split + 1), p.n_chunks32);
    const int nit = c_end - c_begin;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    }
    if (warp == 0) {
        if (lane == 0) bwd_producer(br, smB, reinterpret_cast<const uint8_t*>(p.bop2), B_BYTES, c_begin, nit, cs, crank);
    } else if (warp == 1) {
        if (nit > 0) bwd_mma(br, smB, B_BYTES, p.NK, tmem_base, nit, 8, cs);
    } else if (warp >= 4) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
        const int gw = warp - 4, q4 = gw & 3, cg = gw >> 2;          // lane quarter, column octet
        const int row = q4 * 32 + lane;
        const int b = tile * TC_M + row;
        const bool ok = b < p.B;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const float* evb = p.ev + (size_t)(ok ? b : 0) * E_NV * p.dp;
        const float* scb = p.sc + (size_t)(ok ? b : 0) * SC_N;
        // register cache: for chunk jq of a row i this thread needs j = 32 jq + 8 cg + 0..7 (two 16-byte loads each)
        float X[JQ][8], Y[JQ][8];
#pragma unroll
        for (int jq = 0; jq < JQ; ++jq) {
            const int j = 32 * jq + 8 * cg;
            float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0, y0 = x0, y1 = x0;
            if (ok && j < p.dp) {              // dp is a multiple of 4 and rows of ev are 16-byte aligned
                x0 = *reinterpret_cast<const float4*>(evb + E_R * p.dp + j);
                y0 = *reinterpret_cast<const float4*>(evb + E_Y2 * p.dp + j);
            }
            if (ok && j + 4 < p.dp) {
                x1 = *reinterpret_cast<const float4*>(evb + E_R * p.dp + j + 4);
                y1 = *reinterpret_cast<const float4*>(evb + E_Y2 * p.dp + j + 4);
            }
            X[jq][0] = x0.x; X[jq][1] = x0.y; X[jq][2] = x0.z; X[jq][3] = x0.w; X[jq][4] = x1.x; X[jq][5] = x1.y; X[jq][6] = x1.z; X[jq][7] = x1.w;
            Y[jq][0] = y0.x; Y[jq][1] = y0.y; Y[jq][2] = y0.z; Y[jq][3] = y0.w; Y[jq][4] = y1.x; Y[jq][5] = y1.y; Y[jq][6] = y1.z; Y[jq][7] = y1.w;
        }
        const int n_bil_chunks = p.n_bil_rows / TC_NC;
        auto emit = [&](const float (&g)[8], int it) {
            const int as = it % TC_ASTAGES;
            const uint32_t aph = (it / TC_ASTAGES) & 1;
            const bool tr = gw == 0 && lane == 0 && it >= 40 && it < 44;      // steady state (not the first fills)
            if (tr) TC_TRACE(36 + 3 * (it - 40));
            mbar_wait(&br.a_empty[as], aph ^ 1);
            if (tr) TC_TRACE(37 + 3 * (it - 40));
            tc_fence_after();
            bwd_publish(br, lane_base, as, cg, g, lane);
            if (tr) TC_TRACE(38 + 3 * (it - 40));
        };
        // bilinear rows: the CTA's range starts and ends at row boundaries, so every row i contributes its JQ chunks in
        // order (static indexing of the register cache).  a_bi / L_bi are loaded one row AHEAD of their use.
        const int nbil = max(0, min(c_end, n_bil_chunks) - c_begin);
        int it = 0;
        float ai_n = 0.f, li_n = 0.f;
        if (nbil > 0 && ok && c_begin / JQ < p.dp) {
            ai_n = p.aT[(size_t)(c_begin / JQ) * p.B + b];
            li_n = p.LT[(size_t)(c_begin / JQ) * p.B + b];
        }
        for (int i = c_begin / JQ; it < nbil; ++i) {
            const float ai = ai_n, li = li_n;
            const bool more = ok && i + 1 < p.dp && it + JQ < nbil;
            ai_n = more ? p.aT[(size_t)(i + 1) * p.B + b] : 0.f;
            li_n = more ? p.LT[(size_t)(i + 1) * p.B + b] : 0.f;
#pragma unroll
            for (int jq = 0; jq < JQ; ++jq, ++it) {
                float g[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) g[u] = fmaf(ai, X[jq][u], li * Y[jq][u]);
                emit(g, it);
            }
        }
        // selectional-preference rows (a few chunks per tile): direct loads
        for (; it < nit; ++it) {
            const int c = c_begin + it;
            const int m = c * TC_NC - p.n_bil_rows;
            const int which = m / DP, j0 = m - which * DP + 8 * cg;
            const int sx = which == 0 ? E_A : E_CV, sy = which == 0 ? E_L : E_R;
            const float s2 = ok ? (which == 0 ? scb[SC_G2] : scb[SC_G1]) : 0.f;
            float g[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u;
                g[u] = (ok && j < p.dp) ? fmaf(s2, evb[sy * p.dp + j], evb[sx * p.dp + j]) : 0.f;
            }
            emit(g, it);
        }
        if (gw == 0 && lane == 0) TC_TRACE(34);
        if (gw < 4) {
            // ===== epilogue: accumulator row -> dq partial =====
            if (nit > 0) {
                mbar_wait(br.acc_full, 0);
                tc_fence_after();
            }
            float* o = p.dqp + ((size_t)split * p.B + (ok ? b : 0)) * p.NK;
            for (int c0 = 0; c0 < p.NK; c0 += 32) {
                float t[32];
                if (nit > 0) {
                    tc_ld32(lane_base + (uint32_t)c0, t);
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) t[x] = 0.f;
                }
                if (ok) store_row32(o + c0, t, p.NK - c0, true);       // NK is a multiple of 16
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();
    if (threadIdx.x == 0) TC_TRACE(35);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// dC: rows = operand rows n (one CTA per 128-row tile, batch range split NSb ways), reduction over examples.
struct TcDcArgs {
    const float4* pop3;     // q^T chunks [bc][hi/lo][8][NK]
    const float* ev; const float* sc; const float* aT; const float* LT;
    float* out;             // gC_part [NSb][units*d*K]
    int B, d, dp, K, NK, DP;
    int n_bil_rows, n_rows_total, n_bchunks, NSb, hasM;
    int nacc;               // TMEM accumulators the example chunks are dealt over (1 or 2)
    size_t split_stride;    // units*d*K
    int cs, n_ntiles_pad;   // cluster size; CTAs of a cluster = consecutive ROW tiles of the same batch split; the row-tile count
                            // is padded to a multiple of cs (tiles past n_rows_total hold padding rows only)
};

__global__ void __launch_bounds__(TC_BWD_THREADS, 1) k_tc_dc(TcDcArgs p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TC_TRACE_INIT();
    const int split = blockIdx.x / p.n_ntiles_pad, ntile = blockIdx.x - split * p.n_ntiles_pad;
    const uint32_t cs = (uint32_t)p.cs, crank = cs > 1 ? cluster_ctarank() : 0u;
    const uint32_t B_BYTES = 2u * 8u * (uint32_t)p.NK * 16u;
    uint8_t* smB = smem_raw;
    const int per = (p.n_bchunks + p.NSb - 1) / p.NSb;
    const int c_begin = min(per * split, p.n_bchunks), c_end = min(per * (split + 1), p.n_bchunks);
    const int nit = c_end - c_begin;
    // Per-example scalars of the generated operand, staged once in shared memory.  A 32-row quarter of the tile is one
    // bilinear row i (P1 = a_bi, P2 = L_bi) or one selectional-preference table (P1 = 1, P2 = G2_b | G1_b); the tile
    // holds TC_M / DP such sources.  Coalesced reads of the transposed copies aT / LT.
    const int nsrc = TC_M / p.DP;
    const int nbc = per * TC_NC;
    float* sP1 = reinterpret_cast<float*>(smem_raw + TC_BSTAGES * B_BYTES + 256);
    float* sP2 = sP1 + (size_t)nsrc * nbc;
    for (int idx = threadIdx.x; idx < nsrc * nbc; idx += blockDim.x) {
        const int src = idx / nbc, bl = idx - src * nbc;
        const int b = c_begin * TC_NC + bl;
        const int n0 = ntile * TC_M + src * p.DP;
        float v1 = 0.f, v2 = 0.f;
        if (b < p.B && bl < nit * TC_NC) {
            if (n0 < p.n_bil_rows) {
                const int i = n0 / p.DP;
                if (i < p.d) { v1 = p.aT[(size_t)i * p.B + b]; v2 = p.LT[(size_t)i * p.B + b]; }
            } else if (n0 < p.n_rows_total) {
                v1 = 1.f;
                v2 = p.sc[(size_t)b * SC_N + ((n0 - p.n_bil_rows) / p.DP == 0 ? SC_G2 : SC_G1)];
            }
        }
        sP1[idx] = v1;
        sP2[idx] = v2;
    }
    if (threadIdx.x == 0) TC_TRACE(52);
    uint32_t tmem_base;
    const BwdBars br = bwd_setup(smem_raw, B_BYTES, warp, tmem_base, cs);  // __syncthreads inside: staging visible
    if (threadIdx.x == 0) TC_TRACE(53);

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    }
    if (warp == 0) {
        if (lane == 0) bwd_producer(br, smB, reinterpret_cast<const uint8_t*>(p.pop3), B_BYTES, c_begin, nit, cs, crank);
    } else if (warp == 1) {
        if (nit > 0) bwd_mma(br, smB, B_BYTES, p.NK, tmem_base, nit, 100, cs, p.nacc);
    } else if (warp >= 4) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
        const int gw = warp - 4, q4 = gw & 3, cg = gw >> 2;
        const int row = q4 * 32 + lane;
        const int n = ntile * TC_M + row;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        // decode the row once: g(b) = P1(b) * X(b) + P2(b) * Y(b)
        int type = -1, i = 0, j = 0;          // -1: padding row (zero)
        if (n < p.n_bil_rows) {
            i = n / p.DP; j = n - i * p.DP;
            if (i < p.d && j < p.d) type = 0;
        } else if (n < p.n_rows_total) {
            const int m = n - p.n_bil_rows;
            const int which = m / p.DP;
            j = m - which * p.DP;
            if (j < p.d) type = 1 + which;
        }
        const size_t estride = (size_t)E_NV * p.dp;
        int oX = 0, oY = 0;
        if (type == 0) { oX = E_R * p.dp + j; oY = E_Y2 * p.dp + j; }
        else if (type == 1) { oX = E_A * p.dp + j; oY = E_L * p.dp + j; }
        else if (type == 2) { oX = E_CV * p.dp + j; oY = E_R * p.dp + j; }
        const int src = q4 / (p.DP >> 5);
        const float* s1 = sP1 + (size_t)src * nbc + 8 * cg;
        const float* s2 = sP2 + (size_t)src * nbc + 8 * cg;
        // the 16 per-lane loads of chunk it+1 are issued before chunk it is generated and published (register double
        // buffer): the generator loop no longer pays a global-memory latency per chunk
        auto load = [&](int it, float (&x)[8], float (&y)[8]) {
            const int b0 = (c_begin + it) * TC_NC + 8 * cg;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* evb = p.ev + (size_t)min(b0 + u, p.B - 1) * estride;
                x[u] = evb[oX];
                y[u] = evb[oY];
            }
        };
        auto process = [&](int it, const float (&x)[8], const float (&y)[8]) {
            const int as = it % TC_ASTAGES;
            const uint32_t aph = (it / TC_ASTAGES) & 1;
            const int b0 = (c_begin + it) * TC_NC + 8 * cg;
            const float4 pa = *reinterpret_cast<const float4*>(s1 + it * TC_NC), pb = *reinterpret_cast<const float4*>(s1 + it * TC_NC + 4);
            const float4 qa = *reinterpret_cast<const float4*>(s2 + it * TC_NC), qb = *reinterpret_cast<const float4*>(s2 + it * TC_NC + 4);
            const float p1[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
            const float p2[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            float g[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) g[u] = (type >= 0 && b0 + u < p.B) ? fmaf(p1[u], x[u], p2[u] * y[u]) : 0.f;
            if (gw == 0 && lane == 0 && it < 4) TC_TRACE(100 + 2 * it);      // (slots >= 64 are dropped: dq owns 36..47 now)
            mbar_wait(&br.a_empty[as], aph ^ 1);
            tc_fence_after();
            bwd_publish(br, lane_base, as, cg, g, lane);
            if (gw == 0 && lane == 0 && it < 4) TC_TRACE(101 + 2 * it);
        };
        float xa[8], ya[8], xb[8], yb[8];
        if (nit > 0) load(0, xa, ya);
        for (int it = 0; it < nit; it += 2) {
            if (it + 1 < nit) load(it + 1, xb, yb);
            process(it, xa, ya);
            if (it + 1 < nit) {
                if (it + 2 < nit) load(it + 2, xa, ya);
                process(it + 1, xb, yb);
            }
        }
        if (gw == 0 && lane == 0) TC_TRACE(55);
        if (gw < 4) {
            if (nit > 0) {
                mbar_wait(br.acc_full, 0);
                tc_fence_after();
            }
            // destination inside the split's block: units are [bilinear rows i][C1][C2], each [d][K]
            size_t off = 0;
            if (type == 0) off = ((size_t)i * p.d + j) * p.K;
            else if (type > 0) off = ((size_t)((p.hasM ? p.d : 0) + (type - 1)) * p.d + j) * p.K;
            float* o = p.out + (size_t)split * p.split_stride + off;
            for (int c0 = 0; c0 < p.NK; c0 += 32) {
                float t[32];
                if (nit > 0) {
                    tc_ld32(lane_base + (uint32_t)c0, t);
                    if (p.nacc == 2 && nit > 1) {            // odd chunks went to the second accumulator
                        float t2[32];
                        tc_ld32(lane_base + TC_BWD_ACC2 + (uint32_t)c0, t2);
#pragma unroll
                        for (int x = 0; x < 32; ++x) t[x] += t2[x];
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) t[x] = 0.f;
                }
                if (type >= 0) store_row32(o + c0, t, p.K - c0, (p.K & 3) == 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();
    if (threadIdx.x == 0) TC_TRACE(54);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// per-example finishing of the backward (one warp per example): SP terms of dL/dR, dq += entropy term, softmax backward
__global__ void __launch_bounds__(256) k_tc_bwd_finish(float* __restrict__ ev, const float* __restrict__ sc, const float* __restrict__ q,
                                                       const float* __restrict__ logq, const float* __restrict__ dqp, float* __restrict__ dz,
                                                       float* __restrict__ dzsum_part, int B, int K, int NK, int NS, int d, int dp,
                                                       int hasSP, float ent_coef, const float* __restrict__ vg, const float* __restrict__ wp,
                                                       int tcDP, int tcNS) {
    extern __shared__ float dzs[];     // [8][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * 8 + warp;
    if (b < B) {
        float* evb = ev + (size_t)b * E_NV * dp;
        {
            // d cost / d L = M c (+ SP term), d cost / d R = M^T a (+ SP term): the recompute pass left M c and M^T a in
            // the contraction's partial buffers
            const float gp = sc[(size_t)b * SC_N + SC_GP], g1 = sc[(size_t)b * SC_N + SC_G1], g2 = sc[(size_t)b * SC_N + SC_G2];
            for (int j = lane; j < d; j += 32) {
                float ga1, ga2;
                tc_combined_vw(vg, wp, B, dp, tcDP, tcNS, b, j, ga1, ga2);
                if (hasSP) {
                    ga1 = fmaf(gp + g2, evb[E_C1 * dp + j], ga1);
                    ga2 = fmaf(gp + g1, evb[E_C2 * dp + j], ga2);
                }
                evb[E_GA1 * dp + j] = ga1;
                evb[E_GA2 * dp + j] = ga2;
            }
        }
        float dot = 0.f;
        for (int k = lane; k < K; k += 32) {
            float v = 0.f;
            for (int s = 0; s < NS; ++s) v += dqp[((size_t)s * B + b) * NK + k];
            v = fmaf(ent_coef, logq[(size_t)b * K + k] + 1.f, v);
            dzs[warp * K + k] = v;
            dot = fmaf(q[(size_t)b * K + k], v, dot);
        }
        dot = warp_sum(dot);
        for (int k = lane; k < K; k += 32) {
            const float v = q[(size_t)b * K + k] * (dzs[warp * K + k] - dot);
            dzs[warp * K + k] = v;
            dz[(size_t)b * K + k] = v;
        }
    } else {
        for (int k = lane; k < K; k += 32) dzs[warp * K + k] = 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += dzs[w * K + k];
        dzsum_part[(size_t)blockIdx.x * K + k] = s;
    }
}

int tc_dp(int d) { return d <= 32 ? 32 : d <= 64 ? 64 : 128; }

#ifdef RAE_TRACE
}  // namespace
extern "C" int rae_debug_set_trace(unsigned long long* dev_buf) {
    return (int)cudaMemcpyToSymbol(g_tc_trace, &dev_buf, sizeof(dev_buf));
}
namespace {
#endif

}  // namespace

// shapes the tensor path handles: 16 < d <= 128 (padded to 32/64/128 columns per row) and K <= 104 (P operand resident
// in TMEM next to the accumulators)
int tc_supported(const rae_engine* h) {
    if (!h->hasM) return 0;
    if (h->d <= 16 || h->d > 128) return 0;
    if (h->K > 104) return 0;
    return 1;
}

int tc_init(rae_engine* h) {
    const int DP = tc_dp(h->d);
    TcState& t = h->tc;
    t.DP = DP;
    t.KQ = 2 * ((h->K + 7) / 8);
    const int di = (DP == 32) ? ((h->d + 1) & ~1) : h->d;
    t.n_bil_rows = di * DP;
    t.n_bil_chunks = t.n_bil_rows / TC_N;
    t.n_sp_chunks = h->hasSP ? (2 * DP) / TC_N : 0;
    t.n_rows_total = (t.n_bil_chunks + t.n_sp_chunks) * TC_N;
    t.ntile = (h->B + TC_M - 1) / TC_M;
    int ns = h->num_sms / t.ntile;
    const int pairs = (t.n_bil_chunks + 1) / 2;
    if (ns > pairs) ns = pairs;
    if (ns < 1) ns = 1;
    t.NS = ns;
    // cluster of consecutive example tiles sharing the streamed operand (multicast): 4, 2 or none
    t.cs = (t.ntile % 4 == 0) ? 4 : (t.ntile % 2 == 0 ? 2 : 1);
    if (!(h->cfg.flags & RAE_FLAG_CLUSTER_MULTICAST)) t.cs = 1;
    if ((2 * t.KQ * TC_N * 16) % (16 * t.cs) != 0) t.cs = 1;
    t.smem = (size_t)TC_BSTAGES * 2 * t.KQ * TC_N * 16 + 256;
    // backward operands: NK = relations padded to a multiple of 16, reduction chunks of 32 rows
    t.NK = (h->K + 15) & ~15;
    t.n_chunks32 = t.n_rows_total / TC_NC;
    t.NS2 = std::max(1, std::min(t.n_chunks32 / (DP / 32), h->num_sms / t.ntile));
    // accuracy bound for large batches (many example tiles leave few splits per tile): the measured-safe chain of the dq
    // reduction is 130 chunks = 1560 MMAs into one accumulator (4.5e-6 of ||dW||_inf at the target shape)
    t.NS2 = std::max(t.NS2, std::min(t.n_chunks32 / (DP / 32), (t.n_chunks32 + TC_DQ_MAX_CHAIN - 1) / TC_DQ_MAX_CHAIN));
    if (const char* e = getenv("RAE_TC_DQ_SPLITS")) {          // A/B knob: reduction splits of the dq contraction
        const int v = atoi(e);
        if (v > 0) t.NS2 = std::max(1, std::min(t.n_chunks32 / (DP / 32), v));
    }
    t.smem_dq = (size_t)TC_BSTAGES * (2 * 8 * t.NK * 16) + 256;
    t.n_ntiles = (t.n_rows_total + TC_M - 1) / TC_M;
    t.n_bchunks = (h->B + TC_NC - 1) / TC_NC;
    t.NSb = std::max(1, std::min(t.n_bchunks, h->num_sms / t.n_ntiles));
    // accuracy bound (see TC_DC_MAX_CHAIN): at most TC_DC_MAX_CHAIN chunks per TMEM accumulator
    t.dc_nacc = 2;
    if (const char* e = getenv("RAE_TC_DC_ACCS")) t.dc_nacc = atoi(e) == 1 ? 1 : 2;     // A/B knob
    t.NSb = std::max(t.NSb, (t.n_bchunks + t.dc_nacc * TC_DC_MAX_CHAIN - 1) / (t.dc_nacc * TC_DC_MAX_CHAIN));
    if (const char* e = getenv("RAE_TC_DC_SPLITS")) {          // A/B knob: batch splits of the dC contraction
        const int v = atoi(e);
        if (v > 0) t.NSb = std::max(1, std::min(t.n_bchunks, v));
    }
    for (;;) {
        const int per = (t.n_bchunks + t.NSb - 1) / t.NSb;       // batch chunks per CTA of the dC kernel
        t.smem_dc = t.smem_dq + (size_t)(TC_M / DP) * 2 * per * TC_NC * sizeof(float);
        if (t.smem_dc <= (size_t)h->max_smem_optin || t.NSb >= t.n_bchunks) break;
        t.NSb = std::min(t.n_bchunks, t.NSb * 2);               // large batches: more batch splits, smaller staging area
    }
    cudaError_t e;
    if ((e = cudaMalloc((void**)&t.bop, (size_t)(t.n_bil_chunks + t.n_sp_chunks) * 2 * t.KQ * TC_N * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.vg, (size_t)2 * h->B * h->dp * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.wp, (size_t)2 * t.NS * h->B * h->dp * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.bop2, (size_t)t.n_chunks32 * 2 * 8 * t.NK * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.dqp, (size_t)t.NS2 * h->B * t.NK * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.pop3, (size_t)t.n_bchunks * 2 * 8 * t.NK * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.aT, (size_t)h->dp * h->B * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.LT, (size_t)h->dp * h->B * sizeof(float))) != cudaSuccess)
        return fail(h, RAE_ENOMEM, "tensor-path workspace: %s", cudaGetErrorString(e));
#define RAE_TC_ATTR(KERN, BYTES)                                                                                     \
    if ((e = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES))) != cudaSuccess) \
        return fail(h, RAE_ECUDA, "cudaFuncSetAttribute(" #KERN "): %s", cudaGetErrorString(e));
    RAE_TC_ATTR(k_tc_bilinear<32>, t.smem) RAE_TC_ATTR(k_tc_bilinear<64>, t.smem) RAE_TC_ATTR(k_tc_bilinear<128>, t.smem)
    RAE_TC_ATTR(k_tc_dq<32>, t.smem_dq) RAE_TC_ATTR(k_tc_dq<64>, t.smem_dq) RAE_TC_ATTR(k_tc_dq<128>, t.smem_dq)
    RAE_TC_ATTR(k_tc_dc, t.smem_dc)
#undef RAE_TC_ATTR
    t.ready = true;
    return RAE_OK;
}

void tc_free(rae_engine* h) {
    TcState& t = h->tc;
    cudaFree(t.bop); cudaFree(t.vg); cudaFree(t.wp); cudaFree(t.bop2); cudaFree(t.dqp); cudaFree(t.pop3);
    cudaFree(t.aT); cudaFree(t.LT);
    t = TcState{};
}

// pre-split / pre-arrange the dense operands (call whenever C, C1, C2 changed, i.e. once per step)
int tc_prepare_c(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total = (size_t)t.n_rows_total * t.KQ;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_c<<<dim3(blocks, 2), 256, 0, st>>>(h->P[RAE_P_C], h->P[RAE_P_C1], h->P[RAE_P_C2], h->d, h->K, t.KQ, t.DP, t.n_bil_rows,
                                                 t.n_rows_total, t.bop, t.NK, t.bop2);
    h->launches += 1;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// q-dependent operand of the dC contraction (after the encoder)
int tc_prepare_p(rae_engine* h, const int32_t* a1, const int32_t* a2, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total3 = (size_t)t.n_bchunks * 8 * t.NK;
    const int blocks3 = (int)std::min<size_t>((total3 + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_qt<<<dim3(blocks3, 2), 256, 0, st>>>(h->q, h->B, h->K, t.NK, t.pop3, h->P[RAE_P_A], a1, a2, h->d, h->dp,
                                                   h->quirk ? 1 : 0, h->ev);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// the q-dependent operand alone (only the dC contraction at the end of the step needs it: prepared off the critical path)
int tc_prepare_qt(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total3 = (size_t)t.n_bchunks * 8 * t.NK;
    const int blocks3 = (int)std::min<size_t>((total3 + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_qt<<<dim3(blocks3, 1), 256, 0, st>>>(h->q, h->B, h->K, t.NK, t.pop3, h->P[RAE_P_A], nullptr, nullptr, h->d, h->dp,
                                                   h->quirk ? 1 : 0, h->ev);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// one contraction pass: (slotL, slotR) in -> (slotV = M R [+SP rows to E_C1/E_C2 when with_sp], slotW = M^T L) out
int tc_contract(rae_engine* h, int slotL, int slotR, int slotV, int slotW, bool with_sp, cudaStream_t st) {
    TcState& t = h->tc;
    TcArgs p{};
    p.q = h->q; p.bop = t.bop; p.ev = h->ev; p.ev_out = h->ev; p.vg = t.vg; p.wp = t.wp;
    p.B = h->B; p.K = h->K; p.d = h->d; p.dp = h->dp; p.KQ = t.KQ; p.slotL = slotL; p.slotR = slotR;
    p.n_bil_chunks = t.n_bil_chunks; p.n_sp_chunks = with_sp ? t.n_sp_chunks : 0; p.NS = t.NS;
    const int grid = t.ntile * t.NS;
    p.ntile = t.ntile;
    p.cs = t.cs;
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char* e = getenv("RAE_TC_DEBUG");
            dbg = e ? atoi(e) : 0;
            cudaMemcpyToSymbol(g_tc_bwd_dbg, &dbg, sizeof(int));
        }
        p.dbg = dbg;
    }
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TC_FWD_THREADS);
        cfg.dynamicSmemBytes = t.smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)t.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = t.cs > 1 ? 1 : 0;
        cudaError_t e;
        if (t.DP == 32) e = cudaLaunchKernelEx(&cfg, k_tc_bilinear<32>, p);
        else if (t.DP == 64) e = cudaLaunchKernelEx(&cfg, k_tc_bilinear<64>, p);
        else e = cudaLaunchKernelEx(&cfg, k_tc_bilinear<128>, p);
        if (e != cudaSuccess) return fail(h, RAE_ECUDA, "k_tc_bilinear launch (cluster %d): %s", t.cs, cudaGetErrorString(e));
    }
    // no combine pass: k_score (forward) and k_tc_bwd_finish (backward) read the partial buffers directly
    (void)slotV; (void)slotW;
    h->launches += 1;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// backward on the tensor path: dL, dR (through the forward kernel with L := a, R := c), dq, softmax backward -> dz
// backward on the tensor path, three parts (separately timed phases): (1) M c and M^T a through the forward kernel with
// L := a, R := c; (2) dq contraction; (3) per-example finish: dL, dR, entropy term, softmax backward -> dz
int tc_backward_recompute(rae_engine* h, cudaStream_t st) { return tc_contract(h, E_A, E_CV, E_GA1, E_GA2, false, st); }

int tc_backward_dq(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    k_tc_transpose_al<<<dim3((h->B + 31) / 32, (h->dp + 31) / 32), 256, 0, st>>>(h->ev, h->B, h->d, h->dp, t.aT, t.LT);
    TcDqArgs p{};
    p.bop2 = t.bop2; p.ev = h->ev; p.sc = h->sc; p.aT = t.aT; p.LT = t.LT; p.dqp = t.dqp;
    p.B = h->B; p.d = h->d; p.dp = h->dp; p.K = h->K; p.NK = t.NK;
    p.n_bil_rows = t.n_bil_rows; p.n_chunks32 = t.n_chunks32; p.NS = t.NS2;
    const int grid = t.ntile * t.NS2;
    p.cs = ((2 * 8 * t.NK * 16) % (16 * t.cs) == 0) ? t.cs : 1;
    p.ntile = t.ntile;
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TC_BWD_THREADS);
        cfg.dynamicSmemBytes = t.smem_dq;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = p.cs > 1 ? 1 : 0;
        cudaError_t e;
        if (t.DP == 32) e = cudaLaunchKernelEx(&cfg, k_tc_dq<32>, p);
        else if (t.DP == 64) e = cudaLaunchKernelEx(&cfg, k_tc_dq<64>, p);
        else e = cudaLaunchKernelEx(&cfg, k_tc_dq<128>, p);
        if (e != cudaSuccess) return fail(h, RAE_ECUDA, "k_tc_dq launch (cluster %d): %s", p.cs, cudaGetErrorString(e));
    }
    h->launches += 2;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int tc_backward_finish(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const int blocks = (h->B + 7) / 8;
    if (blocks > h->n_dz_part) return fail(h, RAE_EINVAL, "internal: dzsum_part too small");
    h->dz_part_used = blocks;
    k_tc_bwd_finish<<<blocks, 256, sizeof(float) * 8 * h->K, st>>>(h->ev, h->sc, h->q, h->logq, t.dqp, h->dz, h->dzsum_part, h->B, h->K,
                                                                  t.NK, t.NS2, h->d, h->dp, h->hasSP ? 1 : 0,
                                                                  (float)(2.0 * h->cfg.alpha / h->Z), t.vg, t.wp, t.DP, t.NS);
    h->launches += 1;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// dC, dC1, dC2 partials on the tensor path (k_dense_finalize sums the NSb batch splits)
int tc_grad_dense(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    TcDcArgs p{};
    p.pop3 = t.pop3; p.ev = h->ev; p.sc = h->sc; p.aT = t.aT; p.LT = t.LT; p.out = h->gC_part;
    p.B = h->B; p.d = h->d; p.dp = h->dp; p.K = h->K; p.NK = t.NK; p.DP = t.DP;
    p.n_bil_rows = t.n_bil_rows; p.n_rows_total = t.n_rows_total; p.n_bchunks = t.n_bchunks; p.NSb = t.NSb; p.hasM = h->hasM ? 1 : 0;
    p.nacc = t.dc_nacc;
    p.split_stride = (size_t)h->off_gWb;
    {
        int cs = (h->cfg.flags & RAE_FLAG_CLUSTER_MULTICAST) ? 4 : 1;
        while (cs > 1 && (2 * 8 * t.NK * 16) % (16 * cs) != 0) cs >>= 1;
        p.cs = cs;
        p.n_ntiles_pad = (t.n_ntiles + cs - 1) / cs * cs;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(p.n_ntiles_pad * t.NSb);
        cfg.blockDim = dim3(TC_BWD_THREADS);
        cfg.dynamicSmemBytes = t.smem_dc;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = cs > 1 ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_tc_dc, p);
        if (e != cudaSuccess) return fail(h, RAE_ECUDA, "k_tc_dc launch (cluster %d): %s", cs, cudaGetErrorString(e));
    }
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int tc_gather_lr(rae_engine* h, const int32_t* a1, const int32_t* a2, cudaStream_t st) {
    const int blocks = (h->B * 32 + 255) / 256;
    k_tc_gather_lr<<<blocks, 256, 0, st>>>(h->P[RAE_P_A], a1, a2, h->B, h->d, h->dp, h->quirk ? 1 : 0, h->ev);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

}  // namespace rae
