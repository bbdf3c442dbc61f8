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
    t.ready = true;
    return RAE_OK;
}

void tc_free(rae_engine* h) {
    TcState& t = h->tc;
    cudaFree(t.pop); cudaFree(t.bop); cudaFree(t.vg); cudaFree(t.wp);
    t = TcState{};
}

// pre-split / pre-arrange the dense operand (call whenever C, C1, C2 changed, i.e. once per step)
int tc_prepare_c(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total = (size_t)t.n_rows_total * t.KQ;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_c<<<blocks, 256, 0, st>>>(h->P[RAE_P_C], h->P[RAE_P_C1], h->P[RAE_P_C2], h->d, h->K, t.KQ, t.DP, t.n_bil_rows,
                                        t.n_rows_total, t.bop);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int tc_prepare_p(rae_engine* h, cudaStream_t st) {
    TcState& t = h->tc;
    const size_t total = (size_t)t.ntile * t.KQ * TC_M;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
    k_tc_prep_p<<<blocks, 256, 0, st>>>(h->q, h->B, h->K, t.KQ, t.pop);
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

int tc_gather_lr(rae_engine* h, const int32_t* a1, const int32_t* a2, cudaStream_t st) {
    const int blocks = (h->B * 32 + 255) / 256;
    k_tc_gather_lr<<<blocks, 256, 0, st>>>(h->P[RAE_P_A], a1, a2, h->B, h->d, h->dp, h->quirk ? 1 : 0, h->ev);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

}  // namespace rae
