// Kernel 2 (tensor-core contraction path, sm_100a): the [B,K] x [K, d*d] bilinear contraction on tcgen05 with TMEM
// accumulators, fp32-accurate through an error-compensated 3xTF32 split.
//
//   reference: weighted_R = T.tensordot(relation_probs, R, axes=[[1],[2]])   learning/models/decoders/Bilinear.py:33
//              weightedC  (same)                                             learning/models/decoders/BilinearPlusSP.py:37
//              weightedC1/C2 = T.dot(relation_probs, C1.T / C2.T)            BilinearPlusSP.py:35-36, SelectionalPreferences.py:31-32
//              batched_tensordot / batched_dot consumers                     Bilinear.py:58-59,68-69,78-79
//
// M_b = sum_k q_bk C[:,:,k] is a GEMM  P[B,K] . Cf^T[K, d*d]  whose [B, d*d] result must never reach HBM (64 KiB per
// example at d = 128).  A CTA owns 128 examples (= the 128 TMEM lanes).  The B operand (rows n = (i,j) of Cf, K-major) is
// streamed in chunks of 64 rows; each chunk is one 128x64 accumulator (64 TMEM columns, 4 stages) produced by 13 k-steps
// x 3 MMAs (hi.hi + hi.lo + lo.hi of the TF32 hi/lo split: products carry ~2^-22 relative error, fp32-class).  Epilogue
// warps read the accumulator rows back with tcgen05.ld and immediately fold them into v = M R and w = M^T L, so only
// [B,d] vectors leave the SM.  The selectional-preference tensors C1, C2 are extra rows of the same B operand.
//
// Operands are pre-split and pre-arranged in global memory in the exact shared-memory image the MMA descriptors expect
// (no-swizzle K-major canonical layout: float4 planes T[kq][row]; leading-dim byte offset = rows*16, stride-dim byte
// offset = 128), so a stage is filled by ONE 1-D bulk copy (cp.async.bulk -> UBLKCP) completing on an mbarrier.
//
// Warp roles: 0 = bulk-copy producer, 1 = MMA issuer (one elected lane), 2 = TMEM allocator, 4..11 = epilogue
// (two groups of 4 warps; group g takes the chunks with index parity g, warp w reads TMEM lanes 32*(w%4)..+31).
#include <algorithm>

#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

namespace {

constexpr int TC_M = 128;          // examples per CTA (TMEM lanes)
constexpr int TC_N = 64;           // B-operand rows per chunk (TMEM columns per accumulator stage)
constexpr int TC_TSTAGES = 4;      // accumulator stages in TMEM
constexpr int TC_BSTAGES = 2;      // B-operand smem stages
constexpr int TC_THREADS = 384;    // 12 warps

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense, no negate
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------------------
// P operand: [tile][split hi/lo][kq][row 0..127] float4, rows = examples, 4 consecutive relations per float4
__global__ void __launch_bounds__(256) k_tc_prep_p(const float* __restrict__ q, int B, int K, int KQ, float4* __restrict__ out) {
    const int ntile = (B + TC_M - 1) / TC_M;
    const size_t total = (size_t)ntile * KQ * TC_M;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx % TC_M);
        const int kq = (int)((idx / TC_M) % KQ);
        const int tile = (int)(idx / ((size_t)TC_M * KQ));
        const int b = tile * TC_M + r;
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * kq + u;
            x[u] = (b < B && k < K) ? q[(size_t)b * K + k] : 0.f;
        }
        float4 hi, lo;
        hi.x = tf32_hi(x[0]); hi.y = tf32_hi(x[1]); hi.z = tf32_hi(x[2]); hi.w = tf32_hi(x[3]);
        lo.x = tf32_hi(x[0] - hi.x); lo.y = tf32_hi(x[1] - hi.y); lo.z = tf32_hi(x[2] - hi.z); lo.w = tf32_hi(x[3] - hi.w);
        float4* base = out + (size_t)tile * 2 * KQ * TC_M;
        base[(size_t)kq * TC_M + r] = hi;
        base[(size_t)(KQ + kq) * TC_M + r] = lo;
    }
}

// B operand: [chunk][split][kq][row 0..63] float4; row n of the operand = a [K]-vector of C / C1 / C2:
//   n <  di*DP            : C[i, j, :] with i = n / DP, j = n % DP (zero row when i >= d or j >= d)
//   then DP rows of C1[j,:] and DP rows of C2[j,:] (when the model has them)
__global__ void __launch_bounds__(256) k_tc_prep_c(const float* __restrict__ C, const float* __restrict__ C1,
                                                   const float* __restrict__ C2, int d, int K, int KQ, int DP, int n_bil_rows,
                                                   int n_rows_total, float4* __restrict__ out) {
    const size_t total = (size_t)n_rows_total * KQ;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int kq = (int)(idx % KQ);
        const int n = (int)(idx / KQ);
        const float* src = nullptr;
        if (n < n_bil_rows) {
            const int i = n / DP, j = n - i * DP;
            if (i < d && j < d && C != nullptr) src = C + ((size_t)i * d + j) * K;
        } else {
            const int m = n - n_bil_rows;
            const int which = m / DP, j = m - which * DP;
            if (j < d) src = (which == 0 ? C1 : C2);
            if (src != nullptr) src += (size_t)j * K;
        }
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * kq + u;
            x[u] = (src != nullptr && k < K) ? src[k] : 0.f;
        }
        float4 hi, lo;
        hi.x = tf32_hi(x[0]); hi.y = tf32_hi(x[1]); hi.z = tf32_hi(x[2]); hi.w = tf32_hi(x[3]);
        lo.x = tf32_hi(x[0] - hi.x); lo.y = tf32_hi(x[1] - hi.y); lo.z = tf32_hi(x[2] - hi.z); lo.w = tf32_hi(x[3] - hi.w);
        const int chunk = n / TC_N, r = n - chunk * TC_N;
        float4* base = out + (size_t)chunk * 2 * KQ * TC_N;
        base[(size_t)kq * TC_N + r] = hi;
        base[(size_t)(KQ + kq) * TC_N + r] = lo;
    }
}

// ------------------------------------------------------------------------------------------------------------
// the contraction kernel
// ------------------------------------------------------------------------------------------------------------
struct TcArgs {
    const float4* pop;      // P operand tiles
    const float4* bop;      // B operand chunks
    const float* ev;        // per-example vectors (L at slotL, R at slotR), row stride E_NV*dp
    float* ev_out;          // SP chunks write c1 / c2 here (E_C1 / E_C2)
    float* vg;              // [2][B][dp]   v partial per epilogue group
    float* wp;              // [NS][2][B][dp] w partial per (split, group)
    int B, d, dp, KQ;
    int slotL, slotR;
    int n_bil_chunks;       // chunks holding bilinear rows
    int n_sp_chunks;        // chunks holding C1/C2 rows (forward only)
    int NS;                 // splits of the bilinear chunk range per tile
};

template <int DP>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_bilinear(TcArgs p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / p.NS, split = blockIdx.x - tile * p.NS;
    const uint32_t A_BYTES = 2u * p.KQ * TC_M * 16u;
    const uint32_t B_BYTES = 2u * p.KQ * TC_N * 16u;
    uint8_t* smA = smem_raw;
    uint8_t* smB = smem_raw + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + A_BYTES + TC_BSTAGES * B_BYTES);
    uint64_t* a_full = bars;
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

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < TC_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < TC_TSTAGES; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(TC_TSTAGES * TC_N))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    }
    if (warp == 0) {
        // ===== producer: one bulk copy for the resident P tile, one per B chunk =====
        if (lane == 0 && nit > 0) {
            mbar_expect_tx(a_full, A_BYTES);
            bulk_g2s(smA, reinterpret_cast<const uint8_t*>(p.pop) + (size_t)tile * A_BYTES, A_BYTES, a_full);
            for (int it = 0; it < nit; ++it) {
                const int s = it % TC_BSTAGES;
                const uint32_t ph = (it / TC_BSTAGES) & 1;
                mbar_wait(&b_empty[s], ph ^ 1);
                mbar_expect_tx(&b_full[s], B_BYTES);
                bulk_g2s(smB + (size_t)s * B_BYTES, reinterpret_cast<const uint8_t*>(p.bop) + (size_t)(c_begin + it) * B_BYTES,
                         B_BYTES, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && nit > 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, TC_N);
            const uint32_t a_hi = smem_u32(smA), a_lo = a_hi + p.KQ * TC_M * 16u;
            const int ksteps = p.KQ / 2;
            mbar_wait(a_full, 0);
            for (int it = 0; it < nit; ++it) {
                const int s = it % TC_BSTAGES, ts = it % TC_TSTAGES;
                const uint32_t ph = (it / TC_BSTAGES) & 1, tph = (it / TC_TSTAGES) & 1;
                mbar_wait(&t_empty[ts], tph ^ 1);
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(smB + (size_t)s * B_BYTES), b_lo = b_hi + p.KQ * TC_N * 16u;
                const uint32_t dcol = tmem_base + (uint32_t)(ts * TC_N);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t ao = (uint32_t)(2 * ks) * TC_M * 16u, bo = (uint32_t)(2 * ks) * TC_N * 16u;
                    const uint64_t dah = make_desc(a_hi + ao, TC_M * 16u, 128u), dal = make_desc(a_lo + ao, TC_M * 16u, 128u);
                    const uint64_t dbh = make_desc(b_hi + bo, TC_N * 16u, 128u), dbl = make_desc(b_lo + bo, TC_N * 16u, 128u);
                    tc_mma_tf32(dcol, dah, dbh, idesc, ks > 0 ? 1u : 0u);   // hi * hi
                    tc_mma_tf32(dcol, dah, dbl, idesc, 1u);                 // hi * lo
                    tc_mma_tf32(dcol, dal, dbh, idesc, 1u);                 // lo * hi
                }
                tc_commit(&b_empty[s]);     // smem stage reusable once these MMAs have read it
                tc_commit(&t_full[ts]);     // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
        const int ew = warp - 4, g = ew >> 2, q4 = ew & 3;
        const int row = q4 * 32 + lane;
        const int b = tile * TC_M + row;
        const bool ok = b < p.B;
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
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
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
            mbar_wait(&t_full[ts], tph);
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
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[ts]);     // accumulator stage free for the next MMA
                }
                if (bil) {
                    if (DP >= 64) {
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            vsum = fmaf(t[x], Rr[(32 * hf + x) % RW], vsum);
                            Wr[(32 * hf + x) % RW] = fmaf(t[x], L0, Wr[(32 * hf + x) % RW]);
                        }
                    } else {
                        const float Lh = hf == 0 ? L0 : L1;
                        float vh = 0.f;
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            vh = fmaf(t[x], Rr[x % RW], vh);
                            Wr[x % RW] = fmaf(t[x], Lh, Wr[x % RW]);
                        }
                        if (ok && i0 + hf < p.d) p.vg[((size_t)g * p.B + b) * p.dp + i0 + hf] = vh;
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
        }
        if (ok) {
            float* o = p.wp + (((size_t)split * 2 + g) * p.B + b) * p.dp;
#pragma unroll
            for (int c = 0; c < RW; ++c) {
                const int j = jbase + c;
                if (j < p.dp) o[j] = Wr[c];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(TC_TSTAGES * TC_N))
                     : "memory");
    }
}

// v[b,:] = vg[0] + vg[1] ; w[b,:] = sum over (split, group) of wp  ->  ev slots (fixed order)
__global__ void __launch_bounds__(256) k_tc_combine(const float* __restrict__ vg, const float* __restrict__ wp, float* __restrict__ ev,
                                                    int B, int d, int dp, int NS, int slotV, int slotW) {
    const size_t total = (size_t)B * dp;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx % dp);
        const size_t b = idx / dp;
        if (j >= d) continue;
        const float v = vg[idx] + vg[total + idx];
        float w = 0.f;
        for (int s = 0; s < 2 * NS; ++s) w += wp[(size_t)s * total + idx];
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
// backward, dq[b,k] = sum_n G[b,n] Cf[n,k]  with the generated operand
//   bilinear row n=(i,j): G = a_bi R_bj + L_bi Y2_bj   (dM_b = a R^T + L Y2^T, rank 2)
//   C1 row j            : G = a_bj + G2_b L_bj         (d cost / d c1)
//   C2 row j            : G = c_bj + G1_b R_bj         (d cost / d c2)
// UMMA: M = 128 examples, N = NK (relations padded to 16), reduction over n in chunks of 32 rows.  The A operand is
// produced by 8 generator warps straight into the canonical smem image (hi/lo split, fence.proxy.async, mbarrier), the B
// operand (Cf transposed: K-major along n) is pre-arranged in HBM and bulk-copied.  One TMEM accumulator [128 x NK].
// ------------------------------------------------------------------------------------------------------------
constexpr int TC_NC = 32;          // reduction rows per chunk (8 float4 planes)

__global__ void __launch_bounds__(256) k_tc_prep_ct(const float* __restrict__ C, const float* __restrict__ C1,
                                                    const float* __restrict__ C2, int d, int K, int NK, int DP, int n_bil_rows,
                                                    int n_rows_total, float4* __restrict__ out) {
    // out[c32][split][nq 0..7][krow 0..NK-1] = (Cf[32c+4nq+0..3][krow])
    const size_t total = (size_t)(n_rows_total / 4) * NK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int krow = (int)(idx % NK);
        const int nq_g = (int)(idx / NK);            // global quad index along n
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = 4 * nq_g + u;
            const float* src = nullptr;
            if (n < n_bil_rows) {
                const int i = n / DP, j = n - i * DP;
                if (i < d && j < d && C != nullptr) src = C + ((size_t)i * d + j) * K;
            } else {
                const int m = n - n_bil_rows;
                const int which = m / DP, j = m - which * DP;
                if (j < d) src = (which == 0 ? C1 : C2);
                if (src != nullptr) src += (size_t)j * K;
            }
            x[u] = (src != nullptr && krow < K) ? src[krow] : 0.f;
        }
        float4 hi, lo;
        hi.x = tf32_hi(x[0]); hi.y = tf32_hi(x[1]); hi.z = tf32_hi(x[2]); hi.w = tf32_hi(x[3]);
        lo.x = tf32_hi(x[0] - hi.x); lo.y = tf32_hi(x[1] - hi.y); lo.z = tf32_hi(x[2] - hi.z); lo.w = tf32_hi(x[3] - hi.w);
        const int c32 = nq_g / 8, nq = nq_g - c32 * 8;
        float4* base = out + (size_t)c32 * 2 * 8 * NK;
        base[(size_t)nq * NK + krow] = hi;
        base[(size_t)(8 + nq) * NK + krow] = lo;
    }
}

struct TcDqArgs {
    const float4* bop2;     // Cf^T chunks [c32][hi/lo][8][NK]
    const float* ev; const float* sc;
    float* dqp;             // [NS][B][NK]
    int B, d, dp, K, NK, DP;
    int n_bil_rows, n_chunks32, NS;
};

// the generated operand for 4 consecutive rows n = base..base+3 of chunk c32 and example row evb (zero beyond dp)
__device__ __forceinline__ float4 gen_g4(const float* __restrict__ evb, const float* __restrict__ scb, int dp, int DP,
                                         int n_bil_rows, int n, bool ok) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ok) return g;
    float s1, s2;
    int sx, sy, j;
    if (n < n_bil_rows) {
        const int i = n / DP;
        j = n - i * DP;
        if (i >= dp) return g;
        s1 = evb[E_A * dp + i]; s2 = evb[E_L * dp + i]; sx = E_R; sy = E_Y2;
    } else {
        const int m = n - n_bil_rows;
        const int which = m / DP;
        j = m - which * DP;
        s1 = 1.f;
        if (which == 0) { sx = E_A; s2 = scb[SC_G2]; sy = E_L; }
        else { sx = E_CV; s2 = scb[SC_G1]; sy = E_R; }
    }
    if (j >= dp) return g;
    const float4 x = *reinterpret_cast<const float4*>(evb + sx * dp + j);
    const float4 y = *reinterpret_cast<const float4*>(evb + sy * dp + j);
    g.x = fmaf(s1, x.x, s2 * y.x); g.y = fmaf(s1, x.y, s2 * y.y); g.z = fmaf(s1, x.z, s2 * y.z); g.w = fmaf(s1, x.w, s2 * y.w);
    return g;
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_dq(TcDqArgs p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / p.NS, split = blockIdx.x - tile * p.NS;
    const uint32_t A_BYTES = 2u * 8u * TC_M * 16u;          // hi/lo x 8 planes x 128 rows
    const uint32_t B_BYTES = 2u * 8u * (uint32_t)p.NK * 16u;
    uint8_t* smA = smem_raw;
    uint8_t* smB = smem_raw + 2 * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * A_BYTES + 2 * B_BYTES);
    uint64_t* a_full = bars;            // [2] 256 generator arrivals
    uint64_t* a_empty = bars + 2;       // [2] tcgen05.commit
    uint64_t* b_full = bars + 4;        // [2] bulk-copy tx
    uint64_t* b_empty = bars + 6;       // [2] tcgen05.commit
    uint64_t* acc_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const int per = (p.n_chunks32 + p.NS - 1) / p.NS;
    const int c_begin = min(per * split, p.n_chunks32), c_end = min(per * (split + 1), p.n_chunks32);
    const int nit = c_end - c_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 256); mbar_init(&a_empty[s], 1); mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nit; ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&b_empty[s], ph ^ 1);
                mbar_expect_tx(&b_full[s], B_BYTES);
                bulk_g2s(smB + (size_t)s * B_BYTES, reinterpret_cast<const uint8_t*>(p.bop2) + (size_t)(c_begin + it) * B_BYTES,
                         B_BYTES, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nit > 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, p.NK);
            for (int it = 0; it < nit; ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&a_full[s], ph);
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smA + (size_t)s * A_BYTES), a_lo = a_hi + 8u * TC_M * 16u;
                const uint32_t b_hi = smem_u32(smB + (size_t)s * B_BYTES), b_lo = b_hi + 8u * (uint32_t)p.NK * 16u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ao = (uint32_t)(2 * ks) * TC_M * 16u, bo = (uint32_t)(2 * ks) * (uint32_t)p.NK * 16u;
                    const uint64_t dah = make_desc(a_hi + ao, TC_M * 16u, 128u), dal = make_desc(a_lo + ao, TC_M * 16u, 128u);
                    const uint64_t dbh = make_desc(b_hi + bo, (uint32_t)p.NK * 16u, 128u), dbl = make_desc(b_lo + bo, (uint32_t)p.NK * 16u, 128u);
                    tc_mma_tf32(tmem_base, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    tc_mma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    tc_mma_tf32(tmem_base, dal, dbh, idesc, 1u);
                }
                tc_commit(&a_empty[s]);
                tc_commit(&b_empty[s]);
            }
            tc_commit(acc_full);
        }
    } else if (warp >= 4) {
        // ===== generators (8 warps): thread = (row, half of the chunk's 8 planes) =====
        const int gt = threadIdx.x - 128;
        const int row = gt & 127, half = gt >> 7;
        const int b = tile * TC_M + row;
        const bool ok = b < p.B;
        const float* evb = p.ev + (size_t)(ok ? b : 0) * E_NV * p.dp;
        const float* scb = p.sc + (size_t)(ok ? b : 0) * SC_N;
        for (int it = 0; it < nit; ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int n0 = (c_begin + it) * TC_NC + 16 * half;
            float4 g[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) g[qd] = gen_g4(evb, scb, p.dp, p.DP, p.n_bil_rows, n0 + 4 * qd, ok);
            mbar_wait(&a_empty[s], ph ^ 1);
            float4* ah = reinterpret_cast<float4*>(smA + (size_t)s * A_BYTES);
            float4* al = ah + 8 * TC_M;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float4 hi, lo;
                hi.x = tf32_hi(g[qd].x); hi.y = tf32_hi(g[qd].y); hi.z = tf32_hi(g[qd].z); hi.w = tf32_hi(g[qd].w);
                lo.x = tf32_hi(g[qd].x - hi.x); lo.y = tf32_hi(g[qd].y - hi.y); lo.z = tf32_hi(g[qd].z - hi.z); lo.w = tf32_hi(g[qd].w - hi.w);
                ah[(4 * half + qd) * TC_M + row] = hi;
                al[(4 * half + qd) * TC_M + row] = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the MMA (async proxy)
            mbar_arrive(&a_full[s]);
        }
        if (warp < 8) {
            // ===== epilogue: accumulator row -> dq partial =====
            if (nit > 0) {
                mbar_wait(acc_full, 0);
                tc_fence_after();
            }
            const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            float* o = p.dqp + ((size_t)split * p.B + (ok ? b : 0)) * p.NK;
            for (int c0 = 0; c0 < p.NK; c0 += 32) {
                float t[32];
                if (nit > 0) {
                    tc_ld32(lane_base + (uint32_t)c0, t);
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) t[x] = 0.f;
                }
                if (ok) {
#pragma unroll
                    for (int x = 0; x < 32; ++x)
                        if (c0 + x < p.NK) o[c0 + x] = t[x];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward, dense-parameter gradients: dCf[n,k] = sum_b G[b,n] q[b,k]   (dC, dC1, dC2 in one operand)
// UMMA: M = 128 rows n (one CTA owns an n-tile), N = NK, reduction over the examples in chunks of 32.  The A operand is
// G^T generated in shared memory (planes over 4 consecutive examples), the B operand q^T is pre-arranged and bulk-copied.
// The batch range can be split over CTAs (split-K); partial tiles land in gC_part[split] and are summed in fixed order
// by k_dense_finalize.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tc_prep_qt(const float* __restrict__ q, int B, int K, int NK, float4* __restrict__ out) {
    // out[bc][split][bq 0..7][krow 0..NK-1] = (q[32bc+4bq+0..3][krow])
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
        hi.x = tf32_hi(x[0]); hi.y = tf32_hi(x[1]); hi.z = tf32_hi(x[2]); hi.w = tf32_hi(x[3]);
        lo.x = tf32_hi(x[0] - hi.x); lo.y = tf32_hi(x[1] - hi.y); lo.z = tf32_hi(x[2] - hi.z); lo.w = tf32_hi(x[3] - hi.w);
        float4* base = out + (size_t)bc * 2 * 8 * NK;
        base[(size_t)bq * NK + krow] = hi;
        base[(size_t)(8 + bq) * NK + krow] = lo;
    }
}

struct TcDcArgs {
    const float4* pop3;     // q^T chunks [bc][hi/lo][8][NK]
    const float* ev; const float* sc;
    float* out;             // gC_part [NSb][units*d*K]
    int B, d, dp, K, NK, DP;
    int n_bil_rows, n_rows_total, n_bchunks, NSb, hasM;
    size_t split_stride;    // units*d*K
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_dc(TcDcArgs p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntile = blockIdx.x / p.NSb, split = blockIdx.x - ntile * p.NSb;
    const uint32_t A_BYTES = 2u * 8u * TC_M * 16u;
    const uint32_t B_BYTES = 2u * 8u * (uint32_t)p.NK * 16u;
    uint8_t* smA = smem_raw;
    uint8_t* smB = smem_raw + 2 * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * A_BYTES + 2 * B_BYTES);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + 2;
    uint64_t* b_full = bars + 4;
    uint64_t* b_empty = bars + 6;
    uint64_t* acc_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const int per = (p.n_bchunks + p.NSb - 1) / p.NSb;
    const int c_begin = min(per * split, p.n_bchunks), c_end = min(per * (split + 1), p.n_bchunks);
    const int nit = c_end - c_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 256); mbar_init(&a_empty[s], 1); mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nit; ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&b_empty[s], ph ^ 1);
                mbar_expect_tx(&b_full[s], B_BYTES);
                bulk_g2s(smB + (size_t)s * B_BYTES, reinterpret_cast<const uint8_t*>(p.pop3) + (size_t)(c_begin + it) * B_BYTES,
                         B_BYTES, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nit > 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, p.NK);
            for (int it = 0; it < nit; ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&a_full[s], ph);
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smA + (size_t)s * A_BYTES), a_lo = a_hi + 8u * TC_M * 16u;
                const uint32_t b_hi = smem_u32(smB + (size_t)s * B_BYTES), b_lo = b_hi + 8u * (uint32_t)p.NK * 16u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ao = (uint32_t)(2 * ks) * TC_M * 16u, bo = (uint32_t)(2 * ks) * (uint32_t)p.NK * 16u;
                    const uint64_t dah = make_desc(a_hi + ao, TC_M * 16u, 128u), dal = make_desc(a_lo + ao, TC_M * 16u, 128u);
                    const uint64_t dbh = make_desc(b_hi + bo, (uint32_t)p.NK * 16u, 128u), dbl = make_desc(b_lo + bo, (uint32_t)p.NK * 16u, 128u);
                    tc_mma_tf32(tmem_base, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    tc_mma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    tc_mma_tf32(tmem_base, dal, dbh, idesc, 1u);
                }
                tc_commit(&a_empty[s]);
                tc_commit(&b_empty[s]);
            }
            tc_commit(acc_full);
        }
    } else if (warp >= 4) {
        // ===== generators: thread = (row n of the tile, half of the chunk's examples) =====
        const int gt = threadIdx.x - 128;
        const int row = gt & 127, half = gt >> 7;
        const int n = ntile * TC_M + row;
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
        int oP1, oX, oP2, oY;                 // float offsets inside an example's ev block (type 0) ...
        if (type == 0) { oP1 = E_A * p.dp + i; oX = E_R * p.dp + j; oP2 = E_L * p.dp + i; oY = E_Y2 * p.dp + j; }
        else if (type == 1) { oP1 = -1; oX = E_A * p.dp + j; oP2 = SC_G2; oY = E_L * p.dp + j; }
        else { oP1 = -1; oX = E_CV * p.dp + j; oP2 = SC_G1; oY = E_R * p.dp + j; }
        for (int it = 0; it < nit; ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int b0 = (c_begin + it) * TC_NC + 16 * half;
            float g[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int b = b0 + e;
                float v = 0.f;
                if (type >= 0 && b < p.B) {
                    const float* evb = p.ev + (size_t)b * estride;
                    const float p1 = (type == 0) ? evb[oP1] : 1.f;
                    const float p2 = (type == 0) ? evb[oP2] : p.sc[(size_t)b * SC_N + oP2];
                    v = fmaf(p1, evb[oX], p2 * evb[oY]);
                }
                g[e] = v;
            }
            mbar_wait(&a_empty[s], ph ^ 1);
            float4* ah = reinterpret_cast<float4*>(smA + (size_t)s * A_BYTES);
            float4* al = ah + 8 * TC_M;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float4 hi, lo;
                hi.x = tf32_hi(g[4 * qd]); hi.y = tf32_hi(g[4 * qd + 1]); hi.z = tf32_hi(g[4 * qd + 2]); hi.w = tf32_hi(g[4 * qd + 3]);
                lo.x = tf32_hi(g[4 * qd] - hi.x); lo.y = tf32_hi(g[4 * qd + 1] - hi.y);
                lo.z = tf32_hi(g[4 * qd + 2] - hi.z); lo.w = tf32_hi(g[4 * qd + 3] - hi.w);
                ah[(4 * half + qd) * TC_M + row] = hi;
                al[(4 * half + qd) * TC_M + row] = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&a_full[s]);
        }
        if (warp < 8) {
            if (nit > 0) {
                mbar_wait(acc_full, 0);
                tc_fence_after();
            }
            const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            // destination inside the split's block: units are [bilinear rows i][C1][C2], each [d][K]
            size_t off = 0;
            if (type == 0) off = ((size_t)i * p.d + j) * p.K;
            else if (type > 0) off = ((size_t)((p.hasM ? p.d : 0) + (type - 1)) * p.d + j) * p.K;
            float* o = p.out + (size_t)split * p.split_stride + off;
            for (int c0 = 0; c0 < p.NK; c0 += 32) {
                float t[32];
                if (nit > 0) {
                    tc_ld32(lane_base + (uint32_t)c0, t);
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) t[x] = 0.f;
                }
                if (type >= 0) {
#pragma unroll
                    for (int x = 0; x < 32; ++x)
                        if (c0 + x < p.K) o[c0 + x] = t[x];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// per-example finishing of the backward (one warp per example): SP terms of dL/dR, dq += entropy term, softmax backward
__global__ void __launch_bounds__(256) k_tc_bwd_finish(float* __restrict__ ev, const float* __restrict__ sc, const float* __restrict__ q,
                                                       const float* __restrict__ logq, const float* __restrict__ dqp, float* __restrict__ dz,
                                                       float* __restrict__ dzsum_part, int B, int K, int NK, int NS, int d, int dp,
                                                       int hasSP, float ent_coef) {
    extern __shared__ float dzs[];     // [8][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * 8 + warp;
    if (b < B) {
        float* evb = ev + (size_t)b * E_NV * dp;
        if (hasSP) {
            const float gp = sc[(size_t)b * SC_N + SC_GP], g1 = sc[(size_t)b * SC_N + SC_G1], g2 = sc[(size_t)b * SC_N + SC_G2];
            for (int j = lane; j < d; j += 32) {
                evb[E_GA1 * dp + j] = fmaf(gp + g2, evb[E_C1 * dp + j], evb[E_GA1 * dp + j]);
                evb[E_GA2 * dp + j] = fmaf(gp + g1, evb[E_C2 * dp + j], evb[E_GA2 * dp + j]);
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

}  // namespace

// shapes the tensor path handles: 16 < d <= 128 (padded to 32/64/128 columns per row) and K <= 104 (one resident P tile)
int tc_supported(const rae_engine* h) {
    if (!h->hasM) return 0;
    if (h->d <= 16 || h->d > 128) return 0;
    if (h->K > 104) return 0;
    const int KQ = 2 * ((h->K + 7) / 8);
    const size_t smem = (size_t)2 * KQ * TC_M * 16 + (size_t)TC_BSTAGES * 2 * KQ * TC_N * 16 + 256;
    return smem <= (size_t)h->max_smem_optin;
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
    t.smem = (size_t)2 * t.KQ * TC_M * 16 + (size_t)TC_BSTAGES * 2 * t.KQ * TC_N * 16 + 256;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&t.pop, (size_t)t.ntile * 2 * t.KQ * TC_M * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.bop, (size_t)(t.n_bil_chunks + t.n_sp_chunks) * 2 * t.KQ * TC_N * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.vg, (size_t)2 * h->B * h->dp * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.wp, (size_t)2 * t.NS * h->B * h->dp * sizeof(float))) != cudaSuccess)
        return fail(h, RAE_ENOMEM, "tensor-path workspace: %s", cudaGetErrorString(e));
#define RAE_TC_ATTR(DPV)                                                                                                   \
    if ((e = cudaFuncSetAttribute(k_tc_bilinear<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem)) != cudaSuccess) \
        return fail(h, RAE_ECUDA, "cudaFuncSetAttribute(k_tc_bilinear): %s", cudaGetErrorString(e));
    RAE_TC_ATTR(32) RAE_TC_ATTR(64) RAE_TC_ATTR(128)
#undef RAE_TC_ATTR
    // backward (dq) operand: Cf^T chunks of 32 reduction rows, NK = relations padded to a multiple of 16
    t.NK = (h->K + 15) & ~15;
    t.n_chunks32 = t.n_rows_total / TC_NC;
    t.NS2 = std::max(1, std::min(t.n_chunks32, h->num_sms / t.ntile));
    t.smem_dq = (size_t)2 * (2 * 8 * TC_M * 16) + (size_t)2 * (2 * 8 * t.NK * 16) + 256;
    if ((e = cudaMalloc((void**)&t.bop2, (size_t)t.n_chunks32 * 2 * 8 * t.NK * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&t.dqp, (size_t)t.NS2 * h->B * t.NK * sizeof(float))) != cudaSuccess)
        return fail(h, RAE_ENOMEM, "tensor-path workspace: %s", cudaGetErrorString(e));
    if ((e = cudaFuncSetAttribute(k_tc_dq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_dq)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tc_dc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_dq)) != cudaSuccess)
        return fail(h, RAE_ECUDA, "cudaFuncSetAttribute(k_tc_dq/dc): %s", cudaGetErrorString(e));
    // dC: n-tiles of 128 rows, batch split so that the grid fills the SMs
    t.n_ntiles = (t.n_rows_total + TC_M - 1) / TC_M;
    t.n_bchunks = (h->B + TC_NC - 1) / TC_NC;
    t.NSb = std::max(1, std::min(t.n_bchunks, h->num_sms / t.n_ntiles));
    if ((e = cudaMalloc((void**)&t.pop3, (size_t)t.n_bchunks * 2 * 8 * t.NK * 16)) != cudaSuccess)
        return fail(h, RAE_ENOMEM, "tensor-path workspace: %s", cudaGetErrorString(e));
    t.ready = true;
    return RAE_OK;
}

void tc_free(rae_engine* h) {
    TcState& t = h->tc;
    cudaFree(t.pop); cudaFree(t.bop); cudaFree(t.vg); cudaFree(t.wp); cudaFree(t.bop2); cudaFree(t.dqp); cudaFree(t.pop3);
    t = TcState{};
}

// pre-split / pre-arrange the dense operand (call whenever C, C1, C2 changed, i.e. once per step)
int tc_prepare_c(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total = (size_t)t.n_rows_total * t.KQ;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_c<<<blocks, 256, 0, st>>>(h->P[RAE_P_C], h->P[RAE_P_C1], h->P[RAE_P_C2], h->d, h->K, t.KQ, t.DP, t.n_bil_rows,
                                        t.n_rows_total, t.bop);
    {
        const size_t total2 = (size_t)(t.n_rows_total / 4) * t.NK;
        const int blocks2 = (int)std::min<size_t>((total2 + 255) / 256, (size_t)h->num_sms * 8);
        k_tc_prep_ct<<<blocks2, 256, 0, st>>>(h->P[RAE_P_C], h->P[RAE_P_C1], h->P[RAE_P_C2], h->d, h->K, t.NK, t.DP, t.n_bil_rows,
                                              t.n_rows_total, t.bop2);
    }
    h->launches += 2;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int tc_prepare_p(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total = (size_t)t.ntile * t.KQ * TC_M;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_p<<<blocks, 256, 0, st>>>(h->q, h->B, h->K, t.KQ, t.pop);
    {
        const size_t total3 = (size_t)t.n_bchunks * 8 * t.NK;
        const int blocks3 = (int)std::min<size_t>((total3 + 255) / 256, (size_t)h->num_sms * 8);
        k_tc_prep_qt<<<blocks3, 256, 0, st>>>(h->q, h->B, h->K, t.NK, t.pop3);
        h->launches++;
    }
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// one contraction pass: (slotL, slotR) in -> (slotV = M R [+SP rows to E_C1/E_C2 when with_sp], slotW = M^T L) out
int tc_contract(rae_engine* h, int slotL, int slotR, int slotV, int slotW, bool with_sp, cudaStream_t st) {
    TcState& t = h->tc;
    TcArgs p{};
    p.pop = t.pop; p.bop = t.bop; p.ev = h->ev; p.ev_out = h->ev; p.vg = t.vg; p.wp = t.wp;
    p.B = h->B; p.d = h->d; p.dp = h->dp; p.KQ = t.KQ; p.slotL = slotL; p.slotR = slotR;
    p.n_bil_chunks = t.n_bil_chunks; p.n_sp_chunks = with_sp ? t.n_sp_chunks : 0; p.NS = t.NS;
    RAE_CUDA(h, cudaMemsetAsync(t.vg, 0, (size_t)2 * h->B * h->dp * sizeof(float), st));
    RAE_CUDA(h, cudaMemsetAsync(t.wp, 0, (size_t)2 * t.NS * h->B * h->dp * sizeof(float), st));
    const int grid = t.ntile * t.NS;
    if (t.DP == 32) k_tc_bilinear<32><<<grid, TC_THREADS, t.smem, st>>>(p);
    else if (t.DP == 64) k_tc_bilinear<64><<<grid, TC_THREADS, t.smem, st>>>(p);
    else k_tc_bilinear<128><<<grid, TC_THREADS, t.smem, st>>>(p);
    const size_t total = (size_t)h->B * h->dp;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_combine<<<blocks, 256, 0, st>>>(t.vg, t.wp, h->ev, h->B, h->d, h->dp, t.NS, slotV, slotW);
    h->launches += 4;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// backward on the tensor path: dL, dR (through the forward kernel with L := a, R := c), dq, softmax backward -> dz
int tc_backward(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    int rc = tc_contract(h, E_A, E_CV, E_GA1, E_GA2, false, st);
    if (rc) return rc;
    TcDqArgs p{};
    p.bop2 = t.bop2; p.ev = h->ev; p.sc = h->sc; p.dqp = t.dqp;
    p.B = h->B; p.d = h->d; p.dp = h->dp; p.K = h->K; p.NK = t.NK; p.DP = t.DP;
    p.n_bil_rows = t.n_bil_rows; p.n_chunks32 = t.n_chunks32; p.NS = t.NS2;
    k_tc_dq<<<t.ntile * t.NS2, TC_THREADS, t.smem_dq, st>>>(p);
    const int blocks = (h->B + 7) / 8;
    if (blocks > h->n_dz_part) return fail(h, RAE_EINVAL, "internal: dzsum_part too small");
    h->dz_part_used = blocks;
    k_tc_bwd_finish<<<blocks, 256, sizeof(float) * 8 * h->K, st>>>(h->ev, h->sc, h->q, h->logq, t.dqp, h->dz, h->dzsum_part, h->B, h->K,
                                                                  t.NK, t.NS2, h->d, h->dp, h->hasSP ? 1 : 0,
                                                                  (float)(2.0 * h->cfg.alpha / h->Z));
    h->launches += 2;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// dC, dC1, dC2 partials on the tensor path (k_dense_finalize sums the NSb batch splits)
int tc_grad_dense(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    TcDcArgs p{};
    p.pop3 = t.pop3; p.ev = h->ev; p.sc = h->sc; p.out = h->gC_part;
    p.B = h->B; p.d = h->d; p.dp = h->dp; p.K = h->K; p.NK = t.NK; p.DP = t.DP;
    p.n_bil_rows = t.n_bil_rows; p.n_rows_total = t.n_rows_total; p.n_bchunks = t.n_bchunks; p.NSb = t.NSb; p.hasM = h->hasM ? 1 : 0;
    p.split_stride = (size_t)h->off_gWb;
    k_tc_dc<<<t.n_ntiles * t.NSb, TC_THREADS, t.smem_dq, st>>>(p);
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
