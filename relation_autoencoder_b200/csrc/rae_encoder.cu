// Kernel 1: encoder.  One warp per example: gather-sum of the example's W rows from the binary CSR batch, + Wb,
// fused K-way softmax, entropy (training) or argmax (labelling).
//   reference: sparse.dot(x_feats, W) + Wb ; T.nnet.softmax ; T.argmax   learning/models/encoders/RelationClassifier.py:35-36,45-47
//              entropy = alpha * -sum(log(q) * q)                       learning/OieModel.py:81
// HBM-bound: nnz*K*4 bytes of W rows per batch; every lane keeps UNROLL x KT independent row loads in flight.
#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

template <int KT>
__global__ void __launch_bounds__(256) k_encoder_forward(const int32_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ indices,
                                                         const float* __restrict__ W, const float* __restrict__ Wb,
                                                         int B, int K, float alpha, float* __restrict__ q,
                                                         float* __restrict__ logq, float* __restrict__ sc_ent,
                                                         int sc_stride, int64_t* __restrict__ labels) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int beg = indptr[b], end = indptr[b + 1];
    float z[KT];
#pragma unroll
    for (int t = 0; t < KT; ++t) {
        int k = lane + 32 * t;
        z[t] = (k < K) ? Wb[k] : 0.f;
    }
    constexpr int UNROLL = (KT <= 4) ? 8 : 2;
    for (int base = beg; base < end; base += 32) {
        const int mine = (base + lane < end) ? indices[base + lane] : 0;
        const int cnt = min(32, end - base);
        for (int t0 = 0; t0 < cnt; t0 += UNROLL) {
            float v[UNROLL][KT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int f = __shfl_sync(kFull, mine, (t0 + u) & 31);
                const bool ok = (t0 + u) < cnt;
                const float* row = W + (size_t)f * K;
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    int k = lane + 32 * t;
                    v[u][t] = (ok && k < K) ? ld_nc(row + k) : 0.f;
                }
            }
            // fixed summation order (feature order of the CSR row) -> run-to-run reproducible
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int t = 0; t < KT; ++t) z[t] += v[u][t];
        }
    }
    // softmax over K
    float m = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < KT; ++t) {
        int k = lane + 32 * t;
        if (k < K && z[t] > m) { m = z[t]; arg = k; }   // ascending k per lane: first max wins
    }
    if (labels != nullptr) {
        float bm = m;
        int ba = arg;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(kFull, bm, o);
            int oa = __shfl_xor_sync(kFull, ba, o);
            if (om > bm || (om == bm && oa < ba)) { bm = om; ba = oa; }
        }
        if (lane == 0) labels[b] = (int64_t)ba;
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < KT; ++t) {
        int k = lane + 32 * t;
        if (k < K) s += expf(z[t] - m);
    }
    s = warp_sum(s);
    const float lse = m + logf(s);
    float ent = 0.f;
#pragma unroll
    for (int t = 0; t < KT; ++t) {
        int k = lane + 32 * t;
        if (k < K) {
            float lq = z[t] - lse;
            float p = expf(lq);
            q[(size_t)b * K + k] = p;
            if (logq != nullptr) logq[(size_t)b * K + k] = lq;
            ent -= p * lq;
        }
    }
    if (sc_ent != nullptr) {
        ent = warp_sum(ent);
        if (lane == 0) sc_ent[(size_t)b * sc_stride] = alpha * ent;
    }
}

// K % 4 == 0: every lane owns QT float4 (16-byte row loads, one instruction per W row for K <= 128) and UNROLL rows
// are in flight together.
template <int QT>
__global__ void __launch_bounds__(256) k_encoder_forward_v4(const int32_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices,
                                                            const float* __restrict__ W, const float* __restrict__ Wb,
                                                            int B, int K, float alpha, float* __restrict__ q,
                                                            float* __restrict__ logq, float* __restrict__ sc_ent,
                                                            int sc_stride, int64_t* __restrict__ labels) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int beg = indptr[b], end = indptr[b + 1];
    const int nq = K >> 2;
    float4 z[QT];
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const int qi = lane + 32 * t;
        z[t] = (qi < nq) ? __ldg(reinterpret_cast<const float4*>(Wb) + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int UNROLL = (QT == 1) ? 16 : (QT == 2 ? 8 : 4);
    for (int base = beg; base < end; base += 32) {
        const int mine = (base + lane < end) ? indices[base + lane] : 0;
        const int cnt = min(32, end - base);
        for (int t0 = 0; t0 < cnt; t0 += UNROLL) {
            float4 v[UNROLL][QT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int f = __shfl_sync(kFull, mine, (t0 + u) & 31);
                const bool ok = (t0 + u) < cnt;
                const float4* row = reinterpret_cast<const float4*>(W + (size_t)f * K);
#pragma unroll
                for (int t = 0; t < QT; ++t) {
                    const int qi = lane + 32 * t;
                    v[u][t] = (ok && qi < nq) ? __ldg(row + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // fixed summation order (feature order of the CSR row) -> run-to-run reproducible
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int t = 0; t < QT; ++t) {
                    z[t].x += v[u][t].x; z[t].y += v[u][t].y; z[t].z += v[u][t].z; z[t].w += v[u][t].w;
                }
        }
    }
    float m = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const int qi = lane + 32 * t;
        if (qi < nq) {
            const float e[4] = {z[t].x, z[t].y, z[t].z, z[t].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (e[c] > m) { m = e[c]; arg = 4 * qi + c; }     // ascending k per lane: first max wins
        }
    }
    if (labels != nullptr) {
        float bm = m;
        int ba = arg;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(kFull, bm, o);
            const int oa = __shfl_xor_sync(kFull, ba, o);
            if (om > bm || (om == bm && oa < ba)) { bm = om; ba = oa; }
        }
        if (lane == 0) labels[b] = (int64_t)ba;
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const int qi = lane + 32 * t;
        if (qi < nq) s += (expf(z[t].x - m) + expf(z[t].y - m)) + (expf(z[t].z - m) + expf(z[t].w - m));
    }
    s = warp_sum(s);
    const float lse = m + logf(s);
    float ent = 0.f;
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const int qi = lane + 32 * t;
        if (qi < nq) {
            const float4 lq = make_float4(z[t].x - lse, z[t].y - lse, z[t].z - lse, z[t].w - lse);
            const float4 pr = make_float4(expf(lq.x), expf(lq.y), expf(lq.z), expf(lq.w));
            reinterpret_cast<float4*>(q + (size_t)b * K)[qi] = pr;
            if (logq != nullptr) reinterpret_cast<float4*>(logq + (size_t)b * K)[qi] = lq;
            ent -= (pr.x * lq.x + pr.y * lq.y) + (pr.z * lq.z + pr.w * lq.w);
        }
    }
    if (sc_ent != nullptr) {
        ent = warp_sum(ent);
        if (lane == 0) sc_ent[(size_t)b * sc_stride] = alpha * ent;
    }
}

int launch_encoder_forward(rae_engine* h, const int32_t* indptr, const int32_t* indices, int B, float* q, float* logq,
                           float* sc_ent, int64_t* labels, cudaStream_t st) {
    const int K = h->K;
    const int threads = 256;
    const int blocks = (B * 32 + threads - 1) / threads;
    const float* W = h->P[RAE_P_W];
    const float* Wb = h->P[RAE_P_WB];
    const float alpha = (float)h->cfg.alpha;
    const int kt = (K + 31) / 32;
    // 16-byte path: rows of W, q and log q must be 16-byte aligned (K % 4 == 0 and aligned base pointers)
    const bool vec = (K & 3) == 0 && K <= 512 && (((uintptr_t)W | (uintptr_t)Wb | (uintptr_t)q | (uintptr_t)logq) & 15) == 0;
    if (vec) {
        const int qt = (K / 4 + 31) / 32;
#define RAE_ENC4(QT)                                                                                                 \
    launch_pdl(k_encoder_forward_v4<QT>, dim3(blocks), dim3(threads), 0, st, indptr, indices, W, Wb, B, K, alpha, q, logq, sc_ent, SC_N, \
                                                          labels)
        if (qt <= 1) RAE_ENC4(1);
        else if (qt <= 2) RAE_ENC4(2);
        else RAE_ENC4(4);
#undef RAE_ENC4
        h->launches++;
        RAE_CUDA(h, cudaGetLastError());
        return RAE_OK;
    }
#define RAE_ENC(KT)                                                                                              \
    k_encoder_forward<KT><<<blocks, threads, 0, st>>>(indptr, indices, W, Wb, B, K, alpha, q, logq, sc_ent, SC_N, \
                                                       labels)
    if (kt <= 1) RAE_ENC(1);
    else if (kt <= 2) RAE_ENC(2);
    else if (kt <= 4) RAE_ENC(4);
    else if (kt <= 8) RAE_ENC(8);
    else if (kt <= 16) RAE_ENC(16);
    else if (kt <= 32) RAE_ENC(32);
    else return fail(h, RAE_EINVAL, "K=%d > 1024 is not supported by the encoder kernel", K);
#undef RAE_ENC
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

}  // namespace rae
