// Kernel 2 (SIMT contraction path): decoder forward, scoring/loss, decoder backward, dense-parameter gradients.
//
//   reference: Bilinear.get_scores           learning/models/decoders/Bilinear.py:28-79
//              SelectionalPreferences        learning/models/decoders/SelectionalPreferences.py:30-51
//              BilinearPlusSP.get_scores     learning/models/decoders/BilinearPlusSP.py:34-102
//              -T.mean(all_scores)           learning/OieModel.py:90 ; T.grad  learning/Optimizers.py:27
//
// The score is  L^T M_b R + c1_b.L + c2_b.R  with  M_b = sum_k q_bk C[:,:,k]  (the expectation sits INSIDE the
// sigmoid, Bilinear.py:33,38), c1_b = C1 q_b, c2_b = C2 q_b.  [B,d,d] is never materialised: a CTA owns TB examples,
// streams C one row-block C[i,:,:] ([d,K], contiguous) at a time through shared memory (cp.async double buffer) and
// folds each row of M_b into v = M R and w = M^T L on the fly.  The selectional-preference tensors C1, C2 have the
// shape of one such row-block and ride the same pipeline as two extra "units".
// Backward re-forms the rows of M_b (needed for M c and M^T a) and contracts the rank-2 dM_b = a R^T + L Y2^T with the
// same staged block for dq; dC is a separate kernel that owns output blocks and loops over all examples, so no atomics
// are used anywhere and every sum has a fixed order.
#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

namespace {

struct SimtShape {
    int JT;      // j values per lane (template): rows j = lane + 32 t
    int WB;      // examples per warp (template)
    int TB;      // examples per CTA = 8 * WB
    int drows;   // 32 * JT staged rows (rows >= d are zero)
    int Kc;      // relations per staged chunk (multiple of 4, <= 128)
    int Kcp;     // smem row stride of a staged chunk (== 4 mod 8 -> conflict-free float4 across rows)
    int nkc;     // chunks over K
    int Kp;      // q tile row stride = nkc * Kc
    size_t smem_fwd, smem_bwd;
    bool ok;
};

SimtShape make_shape(int K, int d, int dp, int max_smem) {
    SimtShape s{};
    int jt = (d + 31) / 32;
    s.JT = jt <= 1 ? 1 : jt <= 2 ? 2 : jt <= 4 ? 4 : 8;
    s.ok = jt <= 8;
    s.drows = 32 * s.JT;
    const int Kr = (K + 3) & ~3;
    const size_t stage_budget = 112 * 1024;
    int maxKcp = (int)(stage_budget / (2 * (size_t)s.drows * 4));
    int kcp = maxKcp - ((maxKcp - 4) % 8 + 8) % 8;   // largest value <= maxKcp that is == 4 (mod 8)
    int kc = Kr < 128 ? Kr : 128;
    if (kc > kcp) kc = kcp;          // kcp % 4 == 0
    s.Kc = kc;
    s.Kcp = kc + ((4 - kc % 8) % 8 + 8) % 8;
    s.nkc = (K + kc - 1) / kc;
    s.Kp = s.nkc * kc;
    for (int wb = (s.JT == 8 ? 2 : 4); wb >= 1; wb >>= 1) {
        s.WB = wb;
        s.TB = 8 * wb;
        size_t stage = 2 * (size_t)s.drows * s.Kcp;
        s.smem_fwd = 4 * ((size_t)s.TB * s.Kp + 2 * (size_t)s.TB * dp + stage);
        s.smem_bwd = 4 * (2 * (size_t)s.TB * s.Kp + 3 * (size_t)s.TB * dp + 8 * (size_t)wb * s.drows + stage);
        if (s.smem_bwd <= (size_t)max_smem) return s;
    }
    s.ok = false;
    return s;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct UnitSrc {
    const float* base;  // [d rows][K] row-major block
    int type;           // 0 bilinear row, 1 C1, 2 C2
    int i;
};

// stage rows [0,d) x relations [k0, k0+kn) of a [d,K] block into smem [drows][Kcp]; pad columns zero-filled
__device__ __forceinline__ void stage_block(float* dst, const float* __restrict__ src, int d, int K, int k0, int kn,
                                            int Kcp, bool vec_ok) {
    const int nq = (kn + 3) >> 2;
    const int total = d * nq;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int j = idx / nq, qd = idx - j * nq;
        const int k = k0 + 4 * qd;
        float* o = dst + (size_t)j * Kcp + 4 * qd;
        const float* s = src + (size_t)j * K + k;
        if (vec_ok && 4 * qd + 4 <= kn) {
            cp_async16(o, s);
        } else {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * qd + 0 < kn) v.x = s[0];
            if (4 * qd + 1 < kn) v.y = s[1];
            if (4 * qd + 2 < kn) v.z = s[2];
            if (4 * qd + 3 < kn) v.w = s[3];
            *reinterpret_cast<float4*>(o) = v;
        }
    }
}

struct BilArgs {
    const float* q;       // [B,K]
    const float* logq;    // [B,K]
    const float* C;       // [d,d,K] or null
    const float* C1;      // [d,K] or null
    const float* C2;
    const float* A;       // [N,d]
    const int32_t* a1;
    const int32_t* a2;
    float* ev;            // [B,E_NV,dp]
    const float* sc;      // [B,SC_N]
    float* dz;            // [B,K]
    float* dzsum_part;    // [gridDim.x,K]
    int B, K, d, dp;
    int Kc, Kcp, nkc, Kp, drows;
    int hasM, hasSP, quirk;
    float ent_coef;       // 2*alpha/Z
};

__device__ __forceinline__ UnitSrc unit_of(const BilArgs& p, int u) {
    UnitSrc r;
    const int nM = p.hasM ? p.d : 0;
    if (u < nM) { r.type = 0; r.i = u; r.base = p.C + (size_t)u * p.d * p.K; }
    else if (u == nM) { r.type = 1; r.i = 0; r.base = p.C1; }
    else { r.type = 2; r.i = 0; r.base = p.C2; }
    return r;
}

// acc[e][t] += sum over the staged chunk of q[b_e, k] * Cs[j_t, k]
template <int WB, int JT>
__device__ __forceinline__ void gemm_rows(float (&acc)[WB][JT], const float* __restrict__ qs_warp, int Kp, int k0,
                                          const float* __restrict__ stage, int Kcp, int nqv, int lane) {
    for (int qd = 0; qd < nqv; ++qd) {
        float4 qv[WB];
#pragma unroll
        for (int e = 0; e < WB; ++e) qv[e] = *reinterpret_cast<const float4*>(qs_warp + (size_t)e * Kp + k0 + 4 * qd);
#pragma unroll
        for (int t = 0; t < JT; ++t) {
            const float4 cv = *reinterpret_cast<const float4*>(stage + (size_t)(lane + 32 * t) * Kcp + 4 * qd);
#pragma unroll
            for (int e = 0; e < WB; ++e) {
                acc[e][t] = fmaf(qv[e].x, cv.x, acc[e][t]);
                acc[e][t] = fmaf(qv[e].y, cv.y, acc[e][t]);
                acc[e][t] = fmaf(qv[e].z, cv.z, acc[e][t]);
                acc[e][t] = fmaf(qv[e].w, cv.w, acc[e][t]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward: gathers L = A[a1], R = A[a2] (quirk: A[a1]), writes ev slots L, R, V1 (= v), V2 (= w), C1, C2
// ------------------------------------------------------------------------------------------------------------
template <int WB, int JT>
__global__ void __launch_bounds__(256, 1) k_bilinear_forward(BilArgs p) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    constexpr int TB = 8 * WB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.x * TB;
    float* qs = smem;                                 // [TB][Kp]
    float* Ls = qs + (size_t)TB * p.Kp;               // [TB][dp]
    float* vs = Ls + (size_t)TB * p.dp;               // [TB][dp]
    float* stage = vs + (size_t)TB * p.dp;            // [2][drows][Kcp]
    const size_t stage_elems = (size_t)p.drows * p.Kcp;

    // zero the staging buffers once (rows >= d and pad columns stay zero)
    for (size_t i = threadIdx.x; i < 2 * stage_elems; i += blockDim.x) stage[i] = 0.f;
    // q tile, zero padded
    for (int idx = threadIdx.x; idx < TB * p.Kp; idx += blockDim.x) {
        const int bl = idx / p.Kp, k = idx - bl * p.Kp;
        const int b = b0 + bl;
        qs[idx] = (b < p.B && k < p.K) ? p.q[(size_t)b * p.K + k] : 0.f;
    }
    float Rr[WB][JT], Wr[WB][JT];
#pragma unroll
    for (int e = 0; e < WB; ++e) {
        const int bl = warp * WB + e, b = b0 + bl;
        const bool ok = b < p.B;
        const int r1 = ok ? p.a1[b] : 0;
        const int r2 = ok ? (p.quirk ? r1 : p.a2[b]) : 0;
#pragma unroll
        for (int t = 0; t < JT; ++t) {
            const int j = lane + 32 * t;
            float l = 0.f, r = 0.f;
            if (ok && j < p.d) {
                l = ld_nc(p.A + (size_t)r1 * p.d + j);
                r = ld_nc(p.A + (size_t)r2 * p.d + j);
                float* evb = p.ev + (size_t)b * E_NV * p.dp;
                evb[E_L * p.dp + j] = l;
                evb[E_R * p.dp + j] = r;
            }
            if (j < p.dp) { Ls[bl * p.dp + j] = l; vs[bl * p.dp + j] = 0.f; }
            Rr[e][t] = r;
            Wr[e][t] = 0.f;
        }
    }
    __syncthreads();

    const int nunits = (p.hasM ? p.d : 0) + (p.hasSP ? 2 : 0);
    const int nstages = nunits * p.nkc;
    const bool vec_ok = (p.K % 4) == 0;
    auto issue = [&](int s) {
        const int u = s / p.nkc, c = s - u * p.nkc;
        const UnitSrc us = unit_of(p, u);
        const int k0 = c * p.Kc;
        const int kn = min(p.Kc, p.K - k0);
        stage_block(stage + (size_t)(s & 1) * stage_elems, us.base, p.d, p.K, k0, kn, p.Kcp, vec_ok);
        cp_async_commit();
    };
    if (nstages > 0) issue(0);
    float acc[WB][JT];
    const float* qs_warp = qs + (size_t)warp * WB * p.Kp;
    for (int s = 0; s < nstages; ++s) {
        const int u = s / p.nkc, c = s - u * p.nkc;
        if (s + 1 < nstages) { issue(s + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        if (c == 0) {
#pragma unroll
            for (int e = 0; e < WB; ++e)
#pragma unroll
                for (int t = 0; t < JT; ++t) acc[e][t] = 0.f;
        }
        const int k0 = c * p.Kc;
        const int nqv = (min(p.Kc, p.K - k0) + 3) >> 2;
        gemm_rows<WB, JT>(acc, qs_warp, p.Kp, k0, stage + (size_t)(s & 1) * stage_elems, p.Kcp, nqv, lane);
        if (c == p.nkc - 1) {
            const UnitSrc us = unit_of(p, u);
            if (us.type == 0) {
#pragma unroll
                for (int e = 0; e < WB; ++e) {
                    const int bl = warp * WB + e;
                    float part = 0.f;
#pragma unroll
                    for (int t = 0; t < JT; ++t) part = fmaf(acc[e][t], Rr[e][t], part);
                    part = warp_sum(part);
                    if (lane == 0) vs[bl * p.dp + us.i] = part;
                    const float li = Ls[bl * p.dp + us.i];
#pragma unroll
                    for (int t = 0; t < JT; ++t) Wr[e][t] = fmaf(acc[e][t], li, Wr[e][t]);
                }
            } else {
                const int slot = us.type == 1 ? E_C1 : E_C2;
#pragma unroll
                for (int e = 0; e < WB; ++e) {
                    const int b = b0 + warp * WB + e;
#pragma unroll
                    for (int t = 0; t < JT; ++t) {
                        const int j = lane + 32 * t;
                        if (b < p.B && j < p.d) p.ev[((size_t)b * E_NV + slot) * p.dp + j] = acc[e][t];
                    }
                }
            }
        }
        __syncthreads();   // everyone done with buffer (s&1) before it is refilled at s+2
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < WB; ++e) {
        const int bl = warp * WB + e, b = b0 + bl;
#pragma unroll
        for (int t = 0; t < JT; ++t) {
            const int j = lane + 32 * t;
            if (b < p.B && j < p.d) {
                float* evb = p.ev + (size_t)b * E_NV * p.dp;
                evb[E_V1 * p.dp + j] = vs[bl * p.dp + j];
                evb[E_V2 * p.dp + j] = Wr[e][t];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward: GA1 = M c + (gp+G2) c1, GA2 = M^T a + (gp+G1) c2, dq, softmax/entropy backward -> dz, per-CTA sums of dz
// ------------------------------------------------------------------------------------------------------------
template <int WB, int JT>
__global__ void __launch_bounds__(256, 1) k_bilinear_backward(BilArgs p) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    constexpr int TB = 8 * WB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.x * TB;
    float* qs = smem;                                  // [TB][Kp]
    float* dqs = qs + (size_t)TB * p.Kp;               // [TB][Kp]
    float* Ls = dqs + (size_t)TB * p.Kp;               // [TB][dp]
    float* As = Ls + (size_t)TB * p.dp;                // [TB][dp]
    float* G1s = As + (size_t)TB * p.dp;               // [TB][dp]   GA1 tile
    float* gs = G1s + (size_t)TB * p.dp;               // [8 warps][WB][drows]
    float* stage = gs + (size_t)8 * WB * p.drows;      // [2][drows][Kcp]
    const size_t stage_elems = (size_t)p.drows * p.Kcp;
    float* gs_warp = gs + (size_t)warp * WB * p.drows;

    for (size_t i = threadIdx.x; i < 2 * stage_elems; i += blockDim.x) stage[i] = 0.f;
    for (int idx = threadIdx.x; idx < TB * p.Kp; idx += blockDim.x) {
        const int bl = idx / p.Kp, k = idx - bl * p.Kp;
        const int b = b0 + bl;
        qs[idx] = (b < p.B && k < p.K) ? p.q[(size_t)b * p.K + k] : 0.f;
        dqs[idx] = 0.f;
    }
    float Rr[WB][JT], Yr[WB][JT], Cr[WB][JT], G2r[WB][JT];
    float gG1[WB], gG2[WB], gP[WB];
#pragma unroll
    for (int e = 0; e < WB; ++e) {
        const int bl = warp * WB + e, b = b0 + bl;
        const bool ok = b < p.B;
        const float* evb = p.ev + (size_t)(ok ? b : 0) * E_NV * p.dp;
        gG1[e] = ok ? p.sc[(size_t)b * SC_N + SC_G1] : 0.f;
        gG2[e] = ok ? p.sc[(size_t)b * SC_N + SC_G2] : 0.f;
        gP[e] = ok ? p.sc[(size_t)b * SC_N + SC_GP] : 0.f;
#pragma unroll
        for (int t = 0; t < JT; ++t) {
            const int j = lane + 32 * t;
            const bool in = ok && j < p.d;
            Rr[e][t] = in ? evb[E_R * p.dp + j] : 0.f;
            Yr[e][t] = in ? evb[E_Y2 * p.dp + j] : 0.f;
            Cr[e][t] = in ? evb[E_CV * p.dp + j] : 0.f;
            G2r[e][t] = 0.f;
            if (j < p.dp) {
                Ls[bl * p.dp + j] = in ? evb[E_L * p.dp + j] : 0.f;
                As[bl * p.dp + j] = in ? evb[E_A * p.dp + j] : 0.f;
                G1s[bl * p.dp + j] = 0.f;
            }
        }
    }
    __syncthreads();

    const int nunits = (p.hasM ? p.d : 0) + (p.hasSP ? 2 : 0);
    const int nstages = nunits * p.nkc;
    const bool vec_ok = (p.K % 4) == 0;
    auto issue = [&](int s) {
        const int u = s / p.nkc, c = s - u * p.nkc;
        const UnitSrc us = unit_of(p, u);
        const int k0 = c * p.Kc;
        const int kn = min(p.Kc, p.K - k0);
        stage_block(stage + (size_t)(s & 1) * stage_elems, us.base, p.d, p.K, k0, kn, p.Kcp, vec_ok);
        cp_async_commit();
    };
    if (nstages > 0) issue(0);
    float acc[WB][JT];
    const float* qs_warp = qs + (size_t)warp * WB * p.Kp;
    const int nd4 = (p.d + 3) >> 2;
    for (int s = 0; s < nstages; ++s) {
        const int u = s / p.nkc, c = s - u * p.nkc;
        const UnitSrc us = unit_of(p, u);
        if (s + 1 < nstages) { issue(s + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* st = stage + (size_t)(s & 1) * stage_elems;
        const int k0 = c * p.Kc;
        const int kn = min(p.Kc, p.K - k0);
        const int nqv = (kn + 3) >> 2;
        if (c == 0) {
            // generated operand g[e][j] of this unit -> per-warp smem (broadcast source for the dq contraction)
#pragma unroll
            for (int e = 0; e < WB; ++e) {
                const int bl = warp * WB + e;
#pragma unroll
                for (int t = 0; t < JT; ++t) {
                    const int j = lane + 32 * t;
                    float g;
                    if (us.type == 0) {
                        const float ai = As[bl * p.dp + us.i], li = Ls[bl * p.dp + us.i];
                        g = fmaf(ai, Rr[e][t], li * Yr[e][t]);                       // dM_b[i,j] = a_i R_j + L_i Y2_j
                    } else if (us.type == 1) {
                        g = (j < p.dp) ? fmaf(gG2[e], Ls[bl * p.dp + j], As[bl * p.dp + j]) : 0.f;   // dc1 = a + G2 L
                    } else {
                        g = fmaf(gG1[e], Rr[e][t], Cr[e][t]);                        // dc2 = c + G1 R
                    }
                    gs_warp[e * p.drows + j] = g;
                    acc[e][t] = 0.f;
                }
            }
            __syncwarp();
        }
        if (us.type == 0) gemm_rows<WB, JT>(acc, qs_warp, p.Kp, k0, st, p.Kcp, nqv, lane);
        // dq[b, k0 + 4 lane .. +3] += sum_j g[b,j] * Cs[j, 4 lane ..]
        if (lane < nqv) {
            float4 dqa[WB];
#pragma unroll
            for (int e = 0; e < WB; ++e) dqa[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j4 = 0; j4 < nd4; ++j4) {
                float4 gv[WB];
#pragma unroll
                for (int e = 0; e < WB; ++e) gv[e] = *reinterpret_cast<const float4*>(gs_warp + e * p.drows + 4 * j4);
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const float4 cv = *reinterpret_cast<const float4*>(st + (size_t)(4 * j4 + uu) * p.Kcp + 4 * lane);
#pragma unroll
                    for (int e = 0; e < WB; ++e) {
                        const float g = uu == 0 ? gv[e].x : uu == 1 ? gv[e].y : uu == 2 ? gv[e].z : gv[e].w;
                        dqa[e].x = fmaf(g, cv.x, dqa[e].x);
                        dqa[e].y = fmaf(g, cv.y, dqa[e].y);
                        dqa[e].z = fmaf(g, cv.z, dqa[e].z);
                        dqa[e].w = fmaf(g, cv.w, dqa[e].w);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < WB; ++e) {
                float4* o = reinterpret_cast<float4*>(dqs + (size_t)(warp * WB + e) * p.Kp + k0 + 4 * lane);
                float4 cur = *o;
                cur.x += dqa[e].x; cur.y += dqa[e].y; cur.z += dqa[e].z; cur.w += dqa[e].w;
                *o = cur;
            }
        }
        if (us.type == 0 && c == p.nkc - 1) {
#pragma unroll
            for (int e = 0; e < WB; ++e) {
                const int bl = warp * WB + e;
                float part = 0.f;
#pragma unroll
                for (int t = 0; t < JT; ++t) part = fmaf(acc[e][t], Cr[e][t], part);   // (M c)_i
                part = warp_sum(part);
                if (lane == 0) G1s[bl * p.dp + us.i] = part;
                const float ai = As[bl * p.dp + us.i];
#pragma unroll
                for (int t = 0; t < JT; ++t) G2r[e][t] = fmaf(acc[e][t], ai, G2r[e][t]);   // (M^T a)_j
            }
        }
        __syncthreads();
    }
    __syncwarp();
    // entity-row gradients of the positive pair
#pragma unroll
    for (int e = 0; e < WB; ++e) {
        const int bl = warp * WB + e, b = b0 + bl;
        if (b >= p.B) continue;
        float* evb = p.ev + (size_t)b * E_NV * p.dp;
#pragma unroll
        for (int t = 0; t < JT; ++t) {
            const int j = lane + 32 * t;
            if (j < p.d) {
                float g1 = G1s[bl * p.dp + j], g2 = G2r[e][t];
                if (p.hasSP) {
                    g1 = fmaf(gP[e] + gG2[e], evb[E_C1 * p.dp + j], g1);
                    g2 = fmaf(gP[e] + gG1[e], evb[E_C2 * p.dp + j], g2);
                }
                if (p.quirk) { g1 += g2; g2 = 0.f; }    // R was A[a1]: its gradient lands on row a1
                evb[E_GA1 * p.dp + j] = g1;
                evb[E_GA2 * p.dp + j] = g2;
            }
        }
    }
    // dq += (2 alpha / Z)(log q + 1) ; dz = q * (dq - sum_k q dq)      (OieModel.py:81, RelationClassifier.py:36)
#pragma unroll
    for (int e = 0; e < WB; ++e) {
        const int bl = warp * WB + e, b = b0 + bl;
        float* dqr = dqs + (size_t)bl * p.Kp;
        const float* qr = qs + (size_t)bl * p.Kp;
        float dot = 0.f;
        for (int k = lane; k < p.K; k += 32) {
            float v = dqr[k];
            if (b < p.B) v = fmaf(p.ent_coef, p.logq[(size_t)b * p.K + k] + 1.f, v);
            dqr[k] = v;
            dot = fmaf(qr[k], v, dot);
        }
        dot = warp_sum(dot);
        for (int k = lane; k < p.K; k += 32) {
            const float v = (b < p.B) ? qr[k] * (dqr[k] - dot) : 0.f;
            dqr[k] = v;
            if (b < p.B) p.dz[(size_t)b * p.K + k] = v;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
        float s = 0.f;
        for (int bl = 0; bl < TB; ++bl) s += dqs[(size_t)bl * p.Kp + k];
        p.dzsum_part[(size_t)blockIdx.x * p.K + k] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// scoring: negatives, loss, d cost / d score, a, c, Y2     (one warp per example)
// ------------------------------------------------------------------------------------------------------------
struct ScoreArgs {
    float* ev; float* sc; float* gn1; float* gn2;
    const float* A; const float* Ab;
    const int32_t* a1; const int32_t* a2; const int32_t* neg1; const int32_t* neg2;
    int64_t neg_ld;
    double* loss_part;
    int B, S, d, dp, hasM, hasSP;
    float invZ;
    const float* vg; const float* wp; int tcDP; TcSched tcSch;   // tensor path: v, w come from the contraction's partial buffers
};

// Two warps per example, one per side (side 0: neg1 / e1 slot, side 1: neg2 / e2 slot); 4 examples per CTA.  Both warps
// recompute the cheap positive score; each gathers only its own S negative rows, UN rows in flight per lane.
template <int DT>
__global__ void __launch_bounds__(256) k_score(ScoreArgs p) {
    pdl_enter();
    __shared__ double wl[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int side = warp & 1;
    const int b = blockIdx.x * 4 + (warp >> 1);
    double loss = 0.0;
    if (b < p.B) {
        float* evb = p.ev + (size_t)b * E_NV * p.dp;
        float L[DT], R[DT], V[DT];      // V: direction the negatives of this side are scored against (v + c1 | w + c2)
        float sp1 = 0.f, sp2 = 0.f, pos = 0.f;
        const int tcNS = p.vg != nullptr ? tcs_nslots(p.tcSch, b >> 7) : 0;     // partial slots of this example's tile
#pragma unroll
        for (int t = 0; t < DT; ++t) {
            const int j = lane + 32 * t;
            const bool in = j < p.d;
            L[t] = in ? evb[E_L * p.dp + j] : 0.f;
            R[t] = in ? evb[E_R * p.dp + j] : 0.f;
            float v = 0.f, w = 0.f;
            if (in && p.hasM) {
                if (p.vg != nullptr) tc_combined_vw(p.vg, p.wp, p.B, p.dp, p.tcDP, tcNS, b, j, v, w);
                else { v = evb[E_V1 * p.dp + j]; w = evb[E_V2 * p.dp + j]; }
            }
            const float c1 = (in && p.hasSP) ? evb[E_C1 * p.dp + j] : 0.f;
            const float c2 = (in && p.hasSP) ? evb[E_C2 * p.dp + j] : 0.f;
            const float V1 = v + c1, V2 = w + c2;
            V[t] = side == 0 ? V1 : V2;
            sp1 = fmaf(c1, L[t], sp1);
            sp2 = fmaf(c2, R[t], sp2);
            pos = fmaf(L[t], V1, pos);
        }
        sp1 = warp_sum(sp1);
        sp2 = warp_sum(sp2);
        pos = warp_sum(pos) + sp2;       // L.v + c1.L + c2.R   (Bilinear.py:58-59 / BilinearPlusSP.py:68-72)
        // both warps have read v, w before either overwrites its slot with the gradient direction
        __syncthreads();
#pragma unroll
        for (int t = 0; t < DT; ++t) {
            const int j = lane + 32 * t;
            if (j < p.d) evb[(side == 0 ? E_V1 : E_V2) * p.dp + j] = V[t];
        }
        const int r = side == 0 ? p.a1[b] : p.a2[b];
        const float u = pos + ld_nc(p.Ab + r);                                           // Bilinear.py:36
        // lane 0 carries the positive term and the entropy; lanes 0..UN-1 carry one negative each (summed at the end)
        float lsum = lane == 0 ? log_sigmoid(u) + p.sc[(size_t)b * SC_N + SC_ENT] : 0.f;  // :38-39 (entropy counted twice)
        const float gu = -sigmoidf(-u) * p.invZ;
        float X[DT];
#pragma unroll
        for (int t = 0; t < DT; ++t) X[t] = 0.f;
        float G = 0.f;
        constexpr int UN = (DT <= 2) ? 10 : (DT <= 4 ? 8 : 4);
        const int32_t* neg = side == 0 ? p.neg1 : p.neg2;
        float* gn = side == 0 ? p.gn1 : p.gn2;
        const float add = side == 0 ? sp2 : sp1;         // BilinearPlusSP.py:86-87 / :100,102
        for (int s0 = 0; s0 < p.S; s0 += UN) {
            float x[UN][DT];
            float dots[UN];
            // lane uu owns negative s0 + uu: its id, its bias, later its sigmoid / log-sigmoid
            const bool mine = lane < UN && s0 + lane < p.S;
            const int myrow = mine ? neg[(size_t)(s0 + lane) * p.neg_ld + b] : 0;
            const float myab = mine ? ld_nc(p.Ab + myrow) : 0.f;
#pragma unroll
            for (int uu = 0; uu < UN; ++uu) {
                const int row = __shfl_sync(kFull, myrow, uu);
                const bool ok = s0 + uu < p.S;
#pragma unroll
                for (int t = 0; t < DT; ++t) {
                    const int j = lane + 32 * t;
                    x[uu][t] = (ok && j < p.d) ? ld_nc(p.A + (size_t)row * p.d + j) : 0.f;
                }
            }
#pragma unroll
            for (int uu = 0; uu < UN; ++uu) {
                float dot = 0.f;
#pragma unroll
                for (int t = 0; t < DT; ++t) dot = fmaf(x[uu][t], V[t], dot);
                dots[uu] = dot;
            }
            // UN independent butterfly reductions (no control flow in between: the shuffles pipeline)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) dots[uu] += __shfl_xor_sync(kFull, dots[uu], o);
            float mydot = 0.f;
#pragma unroll
            for (int uu = 0; uu < UN; ++uu)
                if (lane == uu) mydot = dots[uu];
            float myg = 0.f;
            if (mine) {
                mydot += add + myab;
                lsum += log_sigmoid(-mydot);                        // Bilinear.py:47
                myg = sigmoidf(mydot) * p.invZ;
                gn[(size_t)(s0 + lane) * p.B + b] = myg;
            }
#pragma unroll
            for (int uu = 0; uu < UN; ++uu) {
                const float g = __shfl_sync(kFull, myg, uu);        // 0 for the padding rows
                G += g;
#pragma unroll
                for (int t = 0; t < DT; ++t) X[t] = fmaf(g, x[uu][t], X[t]);
            }
        }
        lsum = warp_sum(lsum);
        // the other warp's gu is needed for gp = gu1 + gu2: recomputed here (one bias load + one sigmoid)
        const int ro = side == 0 ? p.a2[b] : p.a1[b];
        const float guo = -sigmoidf(-(pos + ld_nc(p.Ab + ro))) * p.invZ;
        const float gp = side == 0 ? gu + guo : guo + gu;
#pragma unroll
        for (int t = 0; t < DT; ++t) {
            const int j = lane + 32 * t;
            if (j < p.d) {
                if (side == 0) {
                    evb[E_A * p.dp + j] = fmaf(gp, L[t], X[t]);          // a = gp L + X1
                } else {
                    evb[E_CV * p.dp + j] = fmaf(gp, R[t], X[t]);         // c = gp R + Y2
                    evb[E_Y2 * p.dp + j] = X[t];
                }
            }
        }
        if (lane == 0) {
            float* sc = p.sc + (size_t)b * SC_N;
            if (side == 0) { sc[SC_GU1] = gu; sc[SC_GP] = gp; sc[SC_G1] = G; }
            else { sc[SC_GU2] = gu; sc[SC_G2] = G; }
        }
        loss = (double)lsum;
    } else {
        __syncthreads();
    }
    if (lane == 0) wl[warp] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += wl[w];
        p.loss_part[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------------------
// dense-parameter gradients: dC[i,j,k] = sum_b (a_bi R_bj + L_bi Y2_bj) q_bk ; dC1[j,k] = sum_b (a_bj + G2_b L_bj) q_bk ;
// dC2[j,k] = sum_b (c_bj + G1_b R_bj) q_bk.   CTA = (unit, j tile, k tile, batch split); 16x16 threads, RJ x RK each.
// ------------------------------------------------------------------------------------------------------------
struct GradCArgs {
    const float* q; const float* ev; const float* sc;
    float* out;   // [nsplit][units][d][K]
    int B, K, d, dp, hasM, hasSP, nsplit, njt, nkt;
};

template <int RJ, int RK>
__global__ void __launch_bounds__(256) k_grad_dense(GradCArgs p) {
    constexpr int BCH = 32;                      // examples staged per iteration
    __shared__ float gsm[BCH][16 * RJ + 4];
    __shared__ float qsm[BCH][16 * RK + 4];
    const int tj = threadIdx.x >> 4, tk = threadIdx.x & 15;
    int bid = blockIdx.x;
    const int kt = bid % p.nkt; bid /= p.nkt;
    const int jt = bid % p.njt; bid /= p.njt;
    const int split = bid % p.nsplit; bid /= p.nsplit;
    const int unit = bid;
    const int nM = p.hasM ? p.d : 0;
    const int type = unit < nM ? 0 : (unit == nM ? 1 : 2);
    const int j0 = jt * 16 * RJ, k0 = kt * 16 * RK;
    const int per = (p.B + p.nsplit - 1) / p.nsplit;
    const int bb = split * per, be = min(p.B, bb + per);
    float acc[RJ][RK];
#pragma unroll
    for (int a = 0; a < RJ; ++a)
#pragma unroll
        for (int c = 0; c < RK; ++c) acc[a][c] = 0.f;
    for (int bc = bb; bc < be; bc += BCH) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < BCH * 16 * RJ; idx += 256) {
            const int bl = idx / (16 * RJ), jl = idx - bl * (16 * RJ);
            const int b = bc + bl, j = j0 + jl;
            float g = 0.f;
            if (b < be && j < p.d) {
                const float* evb = p.ev + (size_t)b * E_NV * p.dp;
                if (type == 0) g = fmaf(evb[E_A * p.dp + unit], evb[E_R * p.dp + j], evb[E_L * p.dp + unit] * evb[E_Y2 * p.dp + j]);
                else if (type == 1) g = fmaf(p.sc[(size_t)b * SC_N + SC_G2], evb[E_L * p.dp + j], evb[E_A * p.dp + j]);
                else g = fmaf(p.sc[(size_t)b * SC_N + SC_G1], evb[E_R * p.dp + j], evb[E_CV * p.dp + j]);
            }
            gsm[bl][jl] = g;
        }
        for (int idx = threadIdx.x; idx < BCH * 16 * RK; idx += 256) {
            const int bl = idx / (16 * RK), kl = idx - bl * (16 * RK);
            const int b = bc + bl, k = k0 + kl;
            qsm[bl][kl] = (b < be && k < p.K) ? p.q[(size_t)b * p.K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int bl = 0; bl < BCH; ++bl) {
            float gv[RJ], qv[RK];
#pragma unroll
            for (int a = 0; a < RJ; ++a) gv[a] = gsm[bl][tj + 16 * a];
#pragma unroll
            for (int c = 0; c < RK; ++c) qv[c] = qsm[bl][tk + 16 * c];
#pragma unroll
            for (int a = 0; a < RJ; ++a)
#pragma unroll
                for (int c = 0; c < RK; ++c) acc[a][c] = fmaf(gv[a], qv[c], acc[a][c]);
        }
    }
    const int nunits = nM + (p.hasSP ? 2 : 0);
    float* o = p.out + ((size_t)split * nunits + unit) * p.d * p.K;
#pragma unroll
    for (int a = 0; a < RJ; ++a) {
        const int j = j0 + tj + 16 * a;
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            const int k = k0 + tk + 16 * c;
            if (j < p.d && k < p.K) o[(size_t)j * p.K + k] = acc[a][c];
        }
    }
}

BilArgs make_bil_args(rae_engine* h, const SimtShape& s, const int32_t* a1, const int32_t* a2) {
    BilArgs p{};
    p.q = h->q; p.logq = h->logq;
    p.C = h->P[RAE_P_C]; p.C1 = h->P[RAE_P_C1]; p.C2 = h->P[RAE_P_C2];
    p.A = h->P[RAE_P_A]; p.a1 = a1; p.a2 = a2;
    p.ev = h->ev; p.sc = h->sc; p.dz = h->dz; p.dzsum_part = h->dzsum_part;
    p.B = h->B; p.K = h->K; p.d = h->d; p.dp = h->dp;
    p.Kc = s.Kc; p.Kcp = s.Kcp; p.nkc = s.nkc; p.Kp = s.Kp; p.drows = s.drows;
    p.hasM = h->hasM; p.hasSP = h->hasSP; p.quirk = h->quirk;
    p.ent_coef = (float)(2.0 * h->cfg.alpha / h->Z);
    return p;
}

template <typename KernelT>
int set_smem(rae_engine* h, KernelT kern, size_t bytes) {
    RAE_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return RAE_OK;
}

}  // namespace

int simt_supported(const rae_engine* h, char* why, size_t n) {
    SimtShape s = make_shape(h->K, h->d, h->dp, h->max_smem_optin);
    if (!s.ok) {
        snprintf(why, n, "SIMT decoder path supports d <= 256 and shapes whose tiles fit %d B of shared memory (K=%d d=%d)",
                 h->max_smem_optin, h->K, h->d);
        return 0;
    }
    return 1;
}

int simt_grid_blocks(const rae_engine* h) {
    SimtShape s = make_shape(h->K, h->d, h->dp, h->max_smem_optin);
    return (h->B + s.TB - 1) / s.TB;
}

#define RAE_DISPATCH_WBJT(KERN, s, ...)                                          \
    do {                                                                         \
        if (s.WB == 4 && s.JT == 1) { auto k = KERN<4, 1>; __VA_ARGS__ }         \
        else if (s.WB == 4 && s.JT == 2) { auto k = KERN<4, 2>; __VA_ARGS__ }    \
        else if (s.WB == 4 && s.JT == 4) { auto k = KERN<4, 4>; __VA_ARGS__ }    \
        else if (s.WB == 2 && s.JT == 1) { auto k = KERN<2, 1>; __VA_ARGS__ }    \
        else if (s.WB == 2 && s.JT == 2) { auto k = KERN<2, 2>; __VA_ARGS__ }    \
        else if (s.WB == 2 && s.JT == 4) { auto k = KERN<2, 4>; __VA_ARGS__ }    \
        else if (s.WB == 2 && s.JT == 8) { auto k = KERN<2, 8>; __VA_ARGS__ }    \
        else if (s.WB == 1 && s.JT == 1) { auto k = KERN<1, 1>; __VA_ARGS__ }    \
        else if (s.WB == 1 && s.JT == 2) { auto k = KERN<1, 2>; __VA_ARGS__ }    \
        else if (s.WB == 1 && s.JT == 4) { auto k = KERN<1, 4>; __VA_ARGS__ }    \
        else { auto k = KERN<1, 8>; __VA_ARGS__ }                                \
    } while (0)

int launch_bilinear_forward_simt(rae_engine* h, const int32_t* a1, const int32_t* a2, cudaStream_t st) {
    SimtShape s = make_shape(h->K, h->d, h->dp, h->max_smem_optin);
    if (!s.ok) return fail(h, RAE_EINVAL, "shape K=%d d=%d unsupported by the SIMT decoder path", h->K, h->d);
    BilArgs p = make_bil_args(h, s, a1, a2);
    const int blocks = (h->B + s.TB - 1) / s.TB;
    RAE_DISPATCH_WBJT(k_bilinear_forward, s, {
        int rc = set_smem(h, k, s.smem_fwd);
        if (rc) return rc;
        k<<<blocks, 256, s.smem_fwd, st>>>(p);
    });
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_bilinear_backward_simt(rae_engine* h, cudaStream_t st) {
    SimtShape s = make_shape(h->K, h->d, h->dp, h->max_smem_optin);
    if (!s.ok) return fail(h, RAE_EINVAL, "shape K=%d d=%d unsupported by the SIMT decoder path", h->K, h->d);
    BilArgs p = make_bil_args(h, s, nullptr, nullptr);
    const int blocks = (h->B + s.TB - 1) / s.TB;
    if (blocks > h->n_dz_part) return fail(h, RAE_EINVAL, "internal: dzsum_part too small (%d > %d)", blocks, h->n_dz_part);
    h->dz_part_used = blocks;
    RAE_DISPATCH_WBJT(k_bilinear_backward, s, {
        int rc = set_smem(h, k, s.smem_bwd);
        if (rc) return rc;
        k<<<blocks, 256, s.smem_bwd, st>>>(p);
    });
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_score(rae_engine* h, const int32_t* a1, const int32_t* a2, const int32_t* neg1, const int32_t* neg2,
                 int64_t neg_ld, cudaStream_t st) {
    ScoreArgs p{};
    p.ev = h->ev; p.sc = h->sc; p.gn1 = h->gn1; p.gn2 = h->gn2;
    p.A = h->P[RAE_P_A]; p.Ab = h->P[RAE_P_AB];
    p.a1 = a1; p.a2 = a2; p.neg1 = neg1; p.neg2 = neg2; p.neg_ld = neg_ld;
    p.loss_part = h->loss_part;
    p.B = h->B; p.S = h->S; p.d = h->d; p.dp = h->dp; p.hasM = h->hasM; p.hasSP = h->hasSP;
    p.invZ = (float)(1.0 / h->Z);
    const int blocks = (h->B + 3) / 4;
    if (blocks > h->n_loss_part) return fail(h, RAE_EINVAL, "internal: loss_part too small");
    const int dt = (h->d + 31) / 32;
    if (dt <= 1) launch_pdl(k_score<1>, dim3(blocks), dim3(256), 0, st, p);
    else if (dt <= 2) launch_pdl(k_score<2>, dim3(blocks), dim3(256), 0, st, p);
    else if (dt <= 4) launch_pdl(k_score<4>, dim3(blocks), dim3(256), 0, st, p);
    else if (dt <= 8) launch_pdl(k_score<8>, dim3(blocks), dim3(256), 0, st, p);
    else return fail(h, RAE_EINVAL, "d=%d > 256 is not supported", h->d);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_grad_dense_simt(rae_engine* h, cudaStream_t st) {
    GradCArgs p{};
    p.q = h->q; p.ev = h->ev; p.sc = h->sc; p.out = h->gC_part;
    p.B = h->B; p.K = h->K; p.d = h->d; p.dp = h->dp; p.hasM = h->hasM; p.hasSP = h->hasSP;
    p.nsplit = h->gC_nsplit;
    const int nunits = (h->hasM ? h->d : 0) + (h->hasSP ? 2 : 0);
    if (nunits == 0) return RAE_OK;
    auto pick = [](int n) { int r = (n + 15) / 16; return r <= 1 ? 1 : r <= 2 ? 2 : r <= 4 ? 4 : 8; };
    const int RJ = pick(h->d), RK = pick(h->K);
    p.njt = (h->d + 16 * RJ - 1) / (16 * RJ);
    p.nkt = (h->K + 16 * RK - 1) / (16 * RK);
    const int blocks = nunits * p.nsplit * p.njt * p.nkt;
#define RAE_GC(J, Kk) k_grad_dense<J, Kk><<<blocks, 256, 0, st>>>(p)
#define RAE_GCJ(J)                                      \
    do {                                                \
        if (RK == 1) RAE_GC(J, 1);                      \
        else if (RK == 2) RAE_GC(J, 2);                 \
        else if (RK == 4) RAE_GC(J, 4);                 \
        else RAE_GC(J, 8);                              \
    } while (0)
    if (RJ == 1) RAE_GCJ(1);
    else if (RJ == 2) RAE_GCJ(2);
    else if (RJ == 4) RAE_GCJ(4);
    else RAE_GCJ(8);
#undef RAE_GCJ
#undef RAE_GC
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

}  // namespace rae
