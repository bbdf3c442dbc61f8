// Kernel 3: deterministic segmented scatter + sparse-row AdaGrad, dense-parameter update, cost reduction.
//
//   reference: T.grad(cost, params) + AdaGrad.update / SGD.update   learning/Optimizers.py:18-33,38-52
//              (Theano accumulates duplicate rows through AdvancedIncSubtensor1 BEFORE the optimiser squares the
//               gradient, so duplicates must be summed first: sort-by-row, segment-reduce in sorted order, then ONE
//               read-modify-write per unique row.  No atomics anywhere -> bitwise reproducible.)
//              regulariser  learning/OieModel.py:54-62, learning/OieInduction.py:131-135
//
// Rows with zero gradient are fixed points of Optimizers.py:29-32 (acc' = acc, p' = p - lr*0/(sqrt(acc)+1e-6) = p), so
// visiting only the touched rows is exactly the reference's dense sweep when lambda1 = lambda2 = 0.
#include <algorithm>
#include <cub/cub.cuh>

#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

namespace {

__global__ void k_entity_keys(const int32_t* __restrict__ a1, const int32_t* __restrict__ a2,
                              const int32_t* __restrict__ neg1, const int32_t* __restrict__ neg2, int64_t neg_ld, int B,
                              int S, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    // occurrence id o = slot*B + b ; slot 0: args1, 1: args2, 2+s: neg1[s], 2+S+s: neg2[s]
    const int n = (2 + 2 * S) * B;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        const int slot = o / B, b = o - slot * B;
        int row;
        if (slot == 0) row = a1[b];
        else if (slot == 1) row = a2[b];
        else if (slot < 2 + S) row = neg1[(size_t)(slot - 2) * neg_ld + b];
        else row = neg2[(size_t)(slot - 2 - S) * neg_ld + b];
        keys[o] = (uint32_t)row;
        vals[o] = (uint32_t)o;
    }
}

__global__ void k_feature_keys(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int B,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    // one warp per example: key = feature id, value = example index within the batch
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int base = indptr[0];
    for (int p = indptr[b] + lane; p < indptr[b + 1]; p += 32) {
        keys[p - base] = (uint32_t)indices[p];
        vals[p - base] = (uint32_t)b;
    }
}

__global__ void k_flag_heads(const uint32_t* __restrict__ keys_s, int n, int32_t* __restrict__ flags) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flags[i] = (i == 0 || keys_s[i] != keys_s[i - 1]) ? 1 : 0;
}

__global__ void k_scatter_heads(const int32_t* __restrict__ flags, const int32_t* __restrict__ pos, int n,
                                int32_t* __restrict__ seg_start, int32_t* __restrict__ n_seg) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (flags[i]) seg_start[pos[i]] = i;
        if (i == n - 1) {
            const int ns = pos[i] + flags[i];
            seg_start[ns] = n;
            *n_seg = ns;
        }
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) { seg_start[0] = 0; *n_seg = 0; }
}

// ---- entity rows: grad(row) = sum over the row's occurrences (sorted order) of coef * direction ----
struct EntArgs {
    const uint32_t* keys_s; const uint32_t* vals_s; const int32_t* seg_start; const int32_t* n_seg;
    const float* ev; const float* sc; const float* gn1; const float* gn2;
    float* A; float* Ab; float* accA; float* accAb;
    float* gA_dense; float* gAb_dense;
    int B, S, d, dp;
    float lr;
    int adagrad, emit, apply;
};

template <int DT>
__global__ void __launch_bounds__(256) k_entity_update(EntArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int nseg = *p.n_seg;
    for (int seg = gw; seg < nseg; seg += nw) {
        const int beg = p.seg_start[seg], end = p.seg_start[seg + 1];
        const uint32_t row = p.keys_s[beg];
        float g[DT];
#pragma unroll
        for (int t = 0; t < DT; ++t) g[t] = 0.f;
        float gb = 0.f;
        for (int base = beg; base < end; base += 32) {
            const uint32_t mine = (base + lane < end) ? p.vals_s[base + lane] : 0u;
            const int cnt = min(32, end - base);
            // per-lane decode of one occurrence, then broadcast
            const int slot_m = (int)(mine / (uint32_t)p.B);
            const int b_m = (int)(mine - (uint32_t)slot_m * (uint32_t)p.B);
            int vs_m; float coef_m, bias_m;
            if (slot_m == 0) { vs_m = E_GA1; coef_m = 1.f; bias_m = p.sc[(size_t)b_m * SC_N + SC_GU1]; }
            else if (slot_m == 1) { vs_m = E_GA2; coef_m = 1.f; bias_m = p.sc[(size_t)b_m * SC_N + SC_GU2]; }
            else if (slot_m < 2 + p.S) { vs_m = E_V1; coef_m = p.gn1[(size_t)(slot_m - 2) * p.B + b_m]; bias_m = coef_m; }
            else { vs_m = E_V2; coef_m = p.gn2[(size_t)(slot_m - 2 - p.S) * p.B + b_m]; bias_m = coef_m; }
            if (base + lane >= end) { coef_m = 0.f; bias_m = 0.f; }
            const int off_m = (b_m * E_NV + vs_m) * p.dp;
            constexpr int UN = 4;
            for (int t0 = 0; t0 < cnt; t0 += UN) {
                float x[UN][DT], cf[UN], bs[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int src = (t0 + u) & 31;
                    const int off = __shfl_sync(kFull, off_m, src);
                    cf[u] = __shfl_sync(kFull, coef_m, src);
                    bs[u] = __shfl_sync(kFull, bias_m, src);
                    const bool ok = (t0 + u) < cnt;
                    if (!ok) { cf[u] = 0.f; bs[u] = 0.f; }
#pragma unroll
                    for (int t = 0; t < DT; ++t) {
                        const int j = lane + 32 * t;
                        x[u][t] = (ok && j < p.d) ? p.ev[(size_t)off + j] : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    gb += bs[u];
#pragma unroll
                    for (int t = 0; t < DT; ++t) g[t] = fmaf(cf[u], x[u][t], g[t]);
                }
            }
        }
        if (p.emit) {
#pragma unroll
            for (int t = 0; t < DT; ++t) {
                const int j = lane + 32 * t;
                if (j < p.d) p.gA_dense[(size_t)row * p.d + j] = g[t];
            }
            if (lane == 0) p.gAb_dense[row] = gb;
        }
        if (p.apply) {
#pragma unroll
            for (int t = 0; t < DT; ++t) {
                const int j = lane + 32 * t;
                if (j < p.d) {
                    const size_t idx = (size_t)row * p.d + j;
                    float w = p.A[idx];
                    if (p.adagrad) {
                        float a = p.accA[idx];
                        adagrad_apply(w, a, g[t], p.lr);
                        p.accA[idx] = a;
                    } else {
                        w -= p.lr * g[t];
                    }
                    p.A[idx] = w;
                }
            }
            if (lane == 0) {
                float w = p.Ab[row];
                if (p.adagrad) {
                    float a = p.accAb[row];
                    adagrad_apply(w, a, gb, p.lr);
                    p.accAb[row] = a;
                } else {
                    w -= p.lr * gb;
                }
                p.Ab[row] = w;
            }
        }
    }
}

// ---- feature rows: grad W[f,:] = sum over examples containing f (sorted order) of dz[b,:] ----
struct WArgs {
    const uint32_t* keys_s; const uint32_t* vals_s; const int32_t* seg_start; const int32_t* n_seg;
    const float* dz;
    float* W; float* accW; float* gW_dense;
    int K;
    float lr;
    int adagrad, emit, apply;
};

template <int KT>
__global__ void __launch_bounds__(256) k_w_update(WArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int nseg = *p.n_seg;
    for (int seg = gw; seg < nseg; seg += nw) {
        const int beg = p.seg_start[seg], end = p.seg_start[seg + 1];
        const uint32_t row = p.keys_s[beg];
        float g[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) g[t] = 0.f;
        for (int base = beg; base < end; base += 32) {
            const uint32_t mine = (base + lane < end) ? p.vals_s[base + lane] : 0u;
            const int cnt = min(32, end - base);
            constexpr int UN = (KT <= 4) ? 8 : 2;
            for (int t0 = 0; t0 < cnt; t0 += UN) {
                float x[UN][KT];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const uint32_t b = __shfl_sync(kFull, mine, (t0 + u) & 31);
                    const bool ok = (t0 + u) < cnt;
#pragma unroll
                    for (int t = 0; t < KT; ++t) {
                        const int k = lane + 32 * t;
                        x[u][t] = (ok && k < p.K) ? p.dz[(size_t)b * p.K + k] : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u)
#pragma unroll
                    for (int t = 0; t < KT; ++t) g[t] += x[u][t];
            }
        }
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            const int k = lane + 32 * t;
            if (k < p.K) {
                const size_t idx = (size_t)row * p.K + k;
                if (p.emit) p.gW_dense[idx] = g[t];
                if (p.apply) {
                    float w = p.W[idx];
                    if (p.adagrad) {
                        float a = p.accW[idx];
                        adagrad_apply(w, a, g[t], p.lr);
                        p.accW[idx] = a;
                    } else {
                        w -= p.lr * g[t];
                    }
                    p.W[idx] = w;
                }
            }
        }
    }
}

// ---- dense parameters ----
// sum the batch-split partials of dC/dC1/dC2 and the per-CTA partials of dWb into the flat dense gradient buffer
__global__ void k_dense_finalize(const float* __restrict__ part, int nsplit, size_t n_units_elems,
                                 const float* __restrict__ dzsum_part, int n_dz_part, int K, float* __restrict__ out,
                                 size_t off_wb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_units_elems) {
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * n_units_elems + i];
        out[i] = s;
    } else if (i < n_units_elems + (size_t)K) {
        const int k = (int)(i - n_units_elems);
        float s = 0.f;
        for (int c = 0; c < n_dz_part; ++c) s += dzsum_part[(size_t)c * K + k];
        out[off_wb + k] = s;
    }
}

// elementwise optimiser over a dense tensor: g = grad (+ adj*(l1*sign(p) + 2*l2*p)); AdaGrad or SGD.
// Also used for W when the regulariser makes its gradient dense (OieModel.py:54-56).
__global__ void k_dense_apply(float* __restrict__ p, float* __restrict__ acc, float* __restrict__ grad, size_t n,
                              float lr, float reg_l1, float reg_l2, int adagrad, int write_back_grad) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float w = p[i];
        float g = grad[i];
        if (reg_l1 != 0.f || reg_l2 != 0.f) {
            const float sgn = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : 0.f);
            g += reg_l1 * sgn + 2.f * reg_l2 * w;
            if (write_back_grad) grad[i] = g;
        }
        if (adagrad) {
            float a = acc[i];
            adagrad_apply(w, a, g, lr);
            acc[i] = a;
        } else {
            w -= lr * g;
        }
        p[i] = w;
    }
}

// deterministic partial sums of |p| and p^2 (regulariser value, OieModel.py:54-56,60-62)
__global__ void k_reg_norms(const float* __restrict__ p, size_t n, double* __restrict__ part /* [grid][2] */) {
    __shared__ double s1[256], s2[256];
    double a = 0.0, b = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = (double)p[i];
        a += fabs(v);
        b += v * v;
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s1[0]; part[2 * blockIdx.x + 1] = s2[0]; }
}

// cost = -(sum of score partials)/Z + adj*(l1*L1 + l2*L2)       (OieModel.py:90, OieInduction.py:134-135)
__global__ void k_cost(const double* __restrict__ loss_part, int n_loss, const double* __restrict__ reg_part, int n_reg,
                       double invZ, double adj_l1, double adj_l2, double* __restrict__ cost) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n_loss; ++i) s += loss_part[i];
        double l1 = 0.0, l2 = 0.0;
        for (int i = 0; i < n_reg; ++i) { l1 += reg_part[2 * i]; l2 += reg_part[2 * i + 1]; }
        *cost = -s * invZ + adj_l1 * l1 + adj_l2 * l2;
    }
}

}  // namespace

size_t segwork_temp_bytes(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 32, (cudaStream_t)0);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, (cudaStream_t)0);
    return (a > b ? a : b) + 256;
}

int build_entity_keys(rae_engine* h, const int32_t* a1, const int32_t* a2, const int32_t* neg1, const int32_t* neg2,
                      int64_t neg_ld, cudaStream_t st) {
    const int n = (2 + 2 * h->S) * h->B;
    const int blocks = min((n + 255) / 256, 4 * h->num_sms);
    k_entity_keys<<<blocks, 256, 0, st>>>(a1, a2, neg1, neg2, neg_ld, h->B, h->S, h->ent.keys, h->ent.vals);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int build_feature_keys(rae_engine* h, const int32_t* indptr, const int32_t* indices, cudaStream_t st) {
    const int blocks = (h->B * 32 + 255) / 256;
    k_feature_keys<<<blocks, 256, 0, st>>>(indptr, indices, h->B, h->feat.keys, h->feat.vals);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int sort_and_segment(rae_engine* h, SegWork& w, int64_t n, cudaStream_t st) {
    if (n > w.capacity) return fail(h, RAE_EINVAL, "internal: segment workspace too small (%lld > %lld)", (long long)n, (long long)w.capacity);
    if (n > 0) {
        size_t bytes = h->cub_bytes;
        // LSD radix sort is stable: equal rows keep ascending occurrence order == np.argsort(kind='stable')
        RAE_CUDA(h, cub::DeviceRadixSort::SortPairs(h->cub_tmp, bytes, w.keys, w.keys_s, w.vals, w.vals_s, (int)n, 0,
                                                    w.key_bits, st));
        h->launches += (w.key_bits + 7) / 8 + 2;
        const int blocks = min((int)((n + 255) / 256), 4 * h->num_sms);
        k_flag_heads<<<blocks, 256, 0, st>>>(w.keys_s, (int)n, w.flags);
        bytes = h->cub_bytes;
        RAE_CUDA(h, cub::DeviceScan::ExclusiveSum(h->cub_tmp, bytes, w.flags, w.pos, (int)n, st));
        k_scatter_heads<<<blocks, 256, 0, st>>>(w.flags, w.pos, (int)n, w.seg_start, w.n_seg);
        h->launches += 4;
    } else {
        k_scatter_heads<<<1, 32, 0, st>>>(w.flags, w.pos, 0, w.seg_start, w.n_seg);
        h->launches++;
    }
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_entity_update(rae_engine* h, const uint32_t* keys_s, const uint32_t* vals_s, const int32_t* seg_start,
                         const int32_t* n_seg, int64_t n_occ, bool emit_dense, bool apply, cudaStream_t st) {
    EntArgs p{};
    p.keys_s = keys_s; p.vals_s = vals_s; p.seg_start = seg_start; p.n_seg = n_seg;
    p.ev = h->ev; p.sc = h->sc; p.gn1 = h->gn1; p.gn2 = h->gn2;
    p.A = h->P[RAE_P_A]; p.Ab = h->P[RAE_P_AB]; p.accA = h->ACC[RAE_P_A]; p.accAb = h->ACC[RAE_P_AB];
    p.gA_dense = h->gA_dense; p.gAb_dense = h->gAb_dense;
    p.B = h->B; p.S = h->S; p.d = h->d; p.dp = h->dp;
    p.lr = (float)h->cfg.lr; p.adagrad = h->adagrad; p.emit = emit_dense; p.apply = apply;
    const int64_t warps = n_occ < 1 ? 1 : n_occ;
    const int blocks = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->num_sms * 8);
    const int dt = (h->d + 31) / 32;
    if (dt <= 1) k_entity_update<1><<<blocks, 256, 0, st>>>(p);
    else if (dt <= 2) k_entity_update<2><<<blocks, 256, 0, st>>>(p);
    else if (dt <= 4) k_entity_update<4><<<blocks, 256, 0, st>>>(p);
    else k_entity_update<8><<<blocks, 256, 0, st>>>(p);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_w_update(rae_engine* h, const uint32_t* keys_s, const uint32_t* vals_s, const int32_t* seg_start,
                    const int32_t* n_seg, int64_t nnz, bool emit_dense, bool apply, cudaStream_t st) {
    WArgs p{};
    p.keys_s = keys_s; p.vals_s = vals_s; p.seg_start = seg_start; p.n_seg = n_seg;
    p.dz = h->dz; p.W = h->P[RAE_P_W]; p.accW = h->ACC[RAE_P_W]; p.gW_dense = h->gW_dense;
    p.K = h->K; p.lr = (float)h->cfg.lr; p.adagrad = h->adagrad; p.emit = emit_dense; p.apply = apply;
    const int64_t warps = nnz < 1 ? 1 : nnz;
    const int blocks = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->num_sms * 8);
    const int kt = (h->K + 31) / 32;
    if (kt <= 1) k_w_update<1><<<blocks, 256, 0, st>>>(p);
    else if (kt <= 2) k_w_update<2><<<blocks, 256, 0, st>>>(p);
    else if (kt <= 4) k_w_update<4><<<blocks, 256, 0, st>>>(p);
    else if (kt <= 8) k_w_update<8><<<blocks, 256, 0, st>>>(p);
    else if (kt <= 16) k_w_update<16><<<blocks, 256, 0, st>>>(p);
    else k_w_update<32><<<blocks, 256, 0, st>>>(p);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_dense_finalize(rae_engine* h, cudaStream_t st) {
    const size_t n_units = (size_t)h->off_gWb;   // elements of [C | C1 | C2]
    const size_t total = n_units + (size_t)h->K;
    const int blocks = (int)((total + 255) / 256);
    k_dense_finalize<<<blocks, 256, 0, st>>>(h->gC_part, h->gC_nsplit, n_units, h->dzsum_part, h->dz_part_used,
                                             h->K, h->dense_grad, (size_t)h->off_gWb);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

static int apply_one(rae_engine* h, int pid, float* grad, size_t n, bool regularised, bool write_back, cudaStream_t st) {
    if (n == 0) return RAE_OK;
    const float r1 = regularised ? (float)(h->cfg.adj * h->cfg.l1) : 0.f;
    const float r2 = regularised ? (float)(h->cfg.adj * h->cfg.l2) : 0.f;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->num_sms * 8);
    k_dense_apply<<<blocks, 256, 0, st>>>(h->P[pid], h->ACC[pid], grad, n, (float)h->cfg.lr, r1, r2, h->adagrad, write_back ? 1 : 0);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_dense_apply(rae_engine* h, cudaStream_t st) {
    const bool reg = (h->cfg.l1 != 0.0 || h->cfg.l2 != 0.0);
    const bool regdec = reg && h->cfg.ext_reg;
    const size_t dd = (size_t)h->d * h->d * h->K, dk = (size_t)h->d * h->K;
    int rc;
    if (h->hasM && (rc = apply_one(h, RAE_P_C, h->dense_grad + h->off_gC, dd, regdec, h->debug_dense, st))) return rc;
    if (h->hasSP) {
        if ((rc = apply_one(h, RAE_P_C1, h->dense_grad + h->off_gC1, dk, regdec, h->debug_dense, st))) return rc;
        if ((rc = apply_one(h, RAE_P_C2, h->dense_grad + h->off_gC2, dk, regdec, h->debug_dense, st))) return rc;
    }
    if ((rc = apply_one(h, RAE_P_WB, h->dense_grad + h->off_gWb, (size_t)h->K, false, false, st))) return rc;
    if (h->dense_w) {
        // regulariser makes dW dense: gW_dense holds the data gradient (zero rows elsewhere)
        if ((rc = apply_one(h, RAE_P_W, h->gW_dense, (size_t)h->cfg.F * h->K, reg, h->debug_dense, st))) return rc;
    }
    return RAE_OK;
}

int launch_cost(rae_engine* h, cudaStream_t st) {
    const bool reg = (h->cfg.l1 != 0.0 || h->cfg.l2 != 0.0);
    int n_reg = 0;
    if (reg) {
        // fixed grid per tensor -> fixed summation order
        auto norms = [&](int pid, size_t n) -> int {
            if (n == 0) return RAE_OK;
            const int blocks = 64;
            if (n_reg + blocks > h->n_reg_part) return fail(h, RAE_EINVAL, "internal: reg_part too small");
            k_reg_norms<<<blocks, 256, 0, st>>>(h->P[pid], n, h->reg_part + 2 * (size_t)n_reg);
            n_reg += blocks;
            h->launches++;
            return RAE_OK;
        };
        int rc;
        if ((rc = norms(RAE_P_W, (size_t)h->cfg.F * h->K))) return rc;
        if (h->cfg.ext_reg) {
            if (h->hasM && (rc = norms(RAE_P_C, (size_t)h->d * h->d * h->K))) return rc;
            if (h->hasSP) {
                if ((rc = norms(RAE_P_C1, (size_t)h->d * h->K))) return rc;
                if ((rc = norms(RAE_P_C2, (size_t)h->d * h->K))) return rc;
            }
        }
    }
    const int n_loss = (h->B + 7) / 8;
    k_cost<<<1, 32, 0, st>>>(h->loss_part, n_loss, h->reg_part, n_reg, 1.0 / h->Z, h->cfg.adj * h->cfg.l1,
                             h->cfg.adj * h->cfg.l2, h->cost_dev);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_zero(rae_engine* h, void* p, size_t bytes, cudaStream_t st) {
    RAE_CUDA(h, cudaMemsetAsync(p, 0, bytes, st));
    return RAE_OK;
}

}  // namespace rae
