// Kernel 3: deterministic segmented scatter + sparse-row AdaGrad, dense-parameter update, cost reduction.
//
//   reference: T.grad(cost, params) + AdaGrad.update / SGD.update   learning/Optimizers.py:18-33,38-52
//              (Theano accumulates duplicate rows through AdvancedIncSubtensor1 BEFORE the optimiser squares the
//               gradient, so duplicates must be summed first: sort-by-row, segment-reduce in sorted order, then ONE
//               read-modify-write per unique row.  No atomics in the data path -> bitwise reproducible.)
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

__global__ void k_row_keys(const int32_t* __restrict__ rows, int n, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        keys[i] = (uint32_t)rows[i];
        vals[i] = (uint32_t)i;
    }
}

// out[i,:] = table[rows[i],:]; one warp per row, 16-byte vectors when the row pitch allows it
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ table, int width, const int32_t* __restrict__ rows,
                                                     int n, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = gw; i < n; i += nw) {
        const float* src = table + (size_t)rows[i] * width;
        float* dst = out + (size_t)i * width;
        if ((width & 3) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int q = lane; q < (width >> 2); q += 32) d4[q] = s4[q];
        } else {
            for (int k = lane; k < width; k += 32) dst[k] = src[k];
        }
    }
}

// ---- segment boundaries (introspection / parity tests only: the update kernels work from the sorted keys) ----
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

// integer count of distinct keys (statistics only; integer atomics do not affect any result)
__global__ void k_count_heads(const uint32_t* __restrict__ keys_s, int n, int32_t* __restrict__ out) {
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        c += (i == 0 || keys_s[i] != keys_s[i - 1]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---- segmented reduction of sorted occurrences, balanced over warps -------------------------------------------------
// Level 1: one warp per chunk of 32 sorted positions.  Runs (equal keys) that begin and end inside the chunk are
// complete segments: reduced in position order and applied at once (fused optimiser RMW).  Runs that continue from the
// previous chunk or into the next one write a partial row to scratch slot 2c (run starts at the chunk's first position)
// or 2c+1.  Level 2: the warp of the chunk in which a multi-chunk segment STARTS walks the following chunks while their
// first key continues the row, adds the partials in chunk order and applies the update.  Hot rows (Zipf) are thereby
// spread over many warps; the summation order is a fixed function of the sorted layout -> bitwise reproducible.
struct EntArgs {
    const uint32_t* keys_s; const uint32_t* vals_s;
    const float* ev; const float* sc; const float* gn1; const float* gn2;
    float* A; float* Ab; float* accA; float* accAb;
    float* gA_dense; float* gAb_dense;
    float* part;     // [2*nchunks][PE], PE = dp + 4, bias partial at [dp]
    int B, S, d, dp, n;
    float lr;
    int adagrad, emit, apply;
};

template <int DT>
__device__ __forceinline__ void entity_finish(const EntArgs& p, uint32_t row, const float (&g)[DT], float gb, int lane) {
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

template <int DT>
__global__ void __launch_bounds__(256) k_entity_chunks(EntArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = p.n;
    const int nchunks = (n + 31) >> 5;
    const int PE = p.dp + 4;
    for (int c = gw; c < nchunks; c += nw) {
        const int p0 = c << 5;
        const int cnt = min(32, n - p0);
        const bool live = lane < cnt;
        const uint32_t key = live ? p.keys_s[p0 + lane] : 0u;
        const uint32_t mine = live ? p.vals_s[p0 + lane] : 0u;
        // decode this lane's occurrence: o = slot*B + b
        const int slot_m = (int)(mine / (uint32_t)p.B);
        const int b_m = (int)(mine - (uint32_t)slot_m * (uint32_t)p.B);
        int vs_m; float coef_m, bias_m;
        if (slot_m == 0) { vs_m = E_GA1; coef_m = 1.f; bias_m = p.sc[(size_t)b_m * SC_N + SC_GU1]; }
        else if (slot_m == 1) { vs_m = E_GA2; coef_m = 1.f; bias_m = p.sc[(size_t)b_m * SC_N + SC_GU2]; }
        else if (slot_m < 2 + p.S) { vs_m = E_V1; coef_m = p.gn1[(size_t)(slot_m - 2) * p.B + b_m]; bias_m = coef_m; }
        else { vs_m = E_V2; coef_m = p.gn2[(size_t)(slot_m - 2 - p.S) * p.B + b_m]; bias_m = coef_m; }
        if (!live) { coef_m = 0.f; bias_m = 0.f; }
        const int off_m = (b_m * E_NV + vs_m) * p.dp;
        const uint32_t up = __shfl_up_sync(kFull, key, 1);
        const unsigned heads = __ballot_sync(kFull, live && (lane == 0 || key != up));
        const uint32_t key0 = __shfl_sync(kFull, key, 0), keyl = __shfl_sync(kFull, key, cnt - 1);
        const bool cont_prev = p0 > 0 && p.keys_s[p0 - 1] == key0;
        const bool cont_next = p0 + cnt < n && p.keys_s[p0 + cnt] == keyl;
        unsigned m = heads;
        while (m) {
            const int r0 = __ffs(m) - 1;
            m &= m - 1;
            const int r1 = m ? (__ffs(m) - 1) : cnt;
            const uint32_t row = __shfl_sync(kFull, key, r0);
            float g[DT];
#pragma unroll
            for (int t = 0; t < DT; ++t) g[t] = 0.f;
            float gb = 0.f;
            constexpr int UN = 4;
            for (int t0 = r0; t0 < r1; t0 += UN) {
                float x[UN][DT], cf[UN], bs[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int src = (t0 + u) & 31;
                    const int off = __shfl_sync(kFull, off_m, src);
                    cf[u] = __shfl_sync(kFull, coef_m, src);
                    bs[u] = __shfl_sync(kFull, bias_m, src);
                    const bool ok = (t0 + u) < r1;
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
            const bool partial = (r0 == 0 && cont_prev) || (r1 == cnt && cont_next);
            if (!partial) {
                entity_finish<DT>(p, row, g, gb, lane);
            } else {
                float* o = p.part + (size_t)(2 * c + (r0 == 0 ? 0 : 1)) * PE;
#pragma unroll
                for (int t = 0; t < DT; ++t) {
                    const int j = lane + 32 * t;
                    if (j < p.d) o[j] = g[t];
                }
                if (lane == 0) o[p.dp] = gb;
            }
        }
    }
}

// does a multi-chunk segment start in chunk c?  returns its first partial slot (or -1) and its row
__device__ __forceinline__ int long_segment_start(const uint32_t* __restrict__ keys_s, int n, int c, uint32_t* row) {
    const int p0 = c << 5;
    const int cnt = min(32, n - p0);
    if (p0 + cnt >= n) return -1;
    const uint32_t keyl = keys_s[p0 + cnt - 1];
    if (keys_s[p0 + cnt] != keyl) return -1;          // last run ends with the chunk
    *row = keyl;
    if (keys_s[p0] == keyl) {                          // run covers the chunk from its first position
        if (p0 > 0 && keys_s[p0 - 1] == keyl) return -1;   // ... and continues an earlier chunk: not the start
        return 2 * c;
    }
    return 2 * c + 1;
}

template <int DT>
__global__ void __launch_bounds__(256) k_entity_long(EntArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = p.n;
    const int nchunks = (n + 31) >> 5;
    const int PE = p.dp + 4;
    for (int c0 = gw; c0 < nchunks; c0 += nw) {
        uint32_t row = 0;
        const int slot0 = long_segment_start(p.keys_s, n, c0, &row);
        if (slot0 < 0) continue;
        float g[DT];
        const float* s0 = p.part + (size_t)slot0 * PE;
#pragma unroll
        for (int t = 0; t < DT; ++t) {
            const int j = lane + 32 * t;
            g[t] = (j < p.d) ? s0[j] : 0.f;
        }
        float gb = s0[p.dp];
        constexpr int UN = 4;
        bool more = true;
        for (int c = c0 + 1; more; c += UN) {
            float x[UN][DT], bs[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                // chunk c+u belongs to the segment iff every chunk up to it starts with the row
                const bool ok = more && (c + u) < nchunks && p.keys_s[(size_t)(c + u) << 5] == row;
                more = ok;
                const float* s = p.part + (size_t)(2 * (ok ? c + u : c0)) * PE;
                bs[u] = ok ? s[p.dp] : 0.f;
#pragma unroll
                for (int t = 0; t < DT; ++t) {
                    const int j = lane + 32 * t;
                    x[u][t] = (ok && j < p.d) ? s[j] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                gb += bs[u];
#pragma unroll
                for (int t = 0; t < DT; ++t) g[t] += x[u][t];
            }
        }
        entity_finish<DT>(p, row, g, gb, lane);
    }
}

// ---- feature rows: grad W[f,:] = sum over the examples containing f (sorted order) of dz[b,:] ----
struct WArgs {
    const uint32_t* keys_s; const uint32_t* vals_s;
    const float* dz;
    float* W; float* accW; float* gW_dense;
    float* part;     // [2*nchunks][K]
    int K, n;
    float lr;
    int adagrad, emit, apply;
};

template <int KT>
__device__ __forceinline__ void w_finish(const WArgs& p, uint32_t row, const float (&g)[KT], int lane) {
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

template <int KT>
__global__ void __launch_bounds__(256) k_w_chunks(WArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = p.n;
    const int nchunks = (n + 31) >> 5;
    for (int c = gw; c < nchunks; c += nw) {
        const int p0 = c << 5;
        const int cnt = min(32, n - p0);
        const bool live = lane < cnt;
        const uint32_t key = live ? p.keys_s[p0 + lane] : 0u;
        const uint32_t mine = live ? p.vals_s[p0 + lane] : 0u;
        const uint32_t up = __shfl_up_sync(kFull, key, 1);
        const unsigned heads = __ballot_sync(kFull, live && (lane == 0 || key != up));
        const uint32_t key0 = __shfl_sync(kFull, key, 0), keyl = __shfl_sync(kFull, key, cnt - 1);
        const bool cont_prev = p0 > 0 && p.keys_s[p0 - 1] == key0;
        const bool cont_next = p0 + cnt < n && p.keys_s[p0 + cnt] == keyl;
        unsigned m = heads;
        while (m) {
            const int r0 = __ffs(m) - 1;
            m &= m - 1;
            const int r1 = m ? (__ffs(m) - 1) : cnt;
            const uint32_t row = __shfl_sync(kFull, key, r0);
            float g[KT];
#pragma unroll
            for (int t = 0; t < KT; ++t) g[t] = 0.f;
            constexpr int UN = (KT <= 4) ? 8 : 2;
            for (int t0 = r0; t0 < r1; t0 += UN) {
                float x[UN][KT];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const uint32_t b = __shfl_sync(kFull, mine, (t0 + u) & 31);
                    const bool ok = (t0 + u) < r1;
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
            const bool partial = (r0 == 0 && cont_prev) || (r1 == cnt && cont_next);
            if (!partial) {
                w_finish<KT>(p, row, g, lane);
            } else {
                float* o = p.part + (size_t)(2 * c + (r0 == 0 ? 0 : 1)) * p.K;
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    const int k = lane + 32 * t;
                    if (k < p.K) o[k] = g[t];
                }
            }
        }
    }
}

template <int KT>
__global__ void __launch_bounds__(256) k_w_long(WArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = p.n;
    const int nchunks = (n + 31) >> 5;
    for (int c0 = gw; c0 < nchunks; c0 += nw) {
        uint32_t row = 0;
        const int slot0 = long_segment_start(p.keys_s, n, c0, &row);
        if (slot0 < 0) continue;
        float g[KT];
        const float* s0 = p.part + (size_t)slot0 * p.K;
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            const int k = lane + 32 * t;
            g[t] = (k < p.K) ? s0[k] : 0.f;
        }
        constexpr int UN = (KT <= 4) ? 8 : 2;
        bool more = true;
        for (int c = c0 + 1; more; c += UN) {
            float x[UN][KT];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const bool ok = more && (c + u) < nchunks && p.keys_s[(size_t)(c + u) << 5] == row;
                more = ok;
                const float* s = p.part + (size_t)(2 * (ok ? c + u : c0)) * p.K;
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    const int k = lane + 32 * t;
                    x[u][t] = (ok && k < p.K) ? s[k] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int t = 0; t < KT; ++t) g[t] += x[u][t];
        }
        w_finish<KT>(p, row, g, lane);
    }
}

// ---- lane-per-position formulation (fast path) ------------------------------------------------------------------------
// One warp per chunk of 32 sorted positions, lane = position.  The row payload is streamed 8 floats at a time; a warp
// segmented inclusive scan (5 shuffle steps, predicates precomputed from the keys) gives every run's sum on its last
// lane, which applies the optimiser (complete run) or stores the partial (run crosses a chunk boundary).  All runs of a
// chunk are in flight together, loads are sector-aligned, and the reduction tree is a fixed function of the layout.
struct ScanCtx {
    unsigned same;      // bit i: lane - 2^i belongs to the same run
    bool live, is_last, partial;
    int slot;           // partial slot 2c or 2c+1 (valid on is_last && partial)
    uint32_t key;
};

__device__ __forceinline__ ScanCtx scan_ctx(const uint32_t* __restrict__ keys_s, int n, int c, int lane) {
    ScanCtx s;
    const int p0 = c << 5;
    const int cnt = min(32, n - p0);
    s.live = lane < cnt;
    s.key = s.live ? keys_s[p0 + lane] : 0xffffffffu;
    s.same = 0u;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint32_t up = __shfl_up_sync(kFull, s.key, 1 << i);
        if (s.live && lane >= (1 << i) && up == s.key) s.same |= 1u << i;
    }
    const uint32_t dn = __shfl_down_sync(kFull, s.key, 1);
    s.is_last = s.live && (lane == cnt - 1 || dn != s.key);
    const unsigned heads = __ballot_sync(kFull, s.live && !(s.same & 1u));
    const int rs = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));     // first lane of this lane's run
    const uint32_t key0 = __shfl_sync(kFull, s.key, 0), keyl = __shfl_sync(kFull, s.key, cnt - 1);
    const bool cont_prev = p0 > 0 && keys_s[p0 - 1] == key0;
    const bool cont_next = p0 + cnt < n && keys_s[p0 + cnt] == keyl;
    s.partial = (rs == 0 && cont_prev) || (lane == cnt - 1 && cont_next);
    s.slot = 2 * c + (rs == 0 ? 0 : 1);
    return s;
}

__device__ __forceinline__ float seg_scan(float v, unsigned same) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float t = __shfl_up_sync(kFull, v, 1 << i);
        if (same & (1u << i)) v += t;
    }
    return v;
}

__device__ __forceinline__ void opt_apply4(float4& w, float4& a, const float4& g, float lr, int adagrad) {
    if (adagrad) {
        adagrad_apply(w.x, a.x, g.x, lr); adagrad_apply(w.y, a.y, g.y, lr);
        adagrad_apply(w.z, a.z, g.z, lr); adagrad_apply(w.w, a.w, g.w, lr);
    } else {
        w.x -= lr * g.x; w.y -= lr * g.y; w.z -= lr * g.z; w.w -= lr * g.w;
    }
}

// requires K % 4 == 0
__global__ void __launch_bounds__(256) k_w_scan(WArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int nchunks = (p.n + 31) >> 5;
    const int nq = p.K >> 2;
    for (int c = gw; c < nchunks; c += nw) {
        const ScanCtx s = scan_ctx(p.keys_s, p.n, c, lane);
        const uint32_t b = s.live ? p.vals_s[(c << 5) + lane] : 0u;
        const float4* src = reinterpret_cast<const float4*>(p.dz + (size_t)b * p.K);
        const size_t rowoff = (size_t)s.key * p.K;
#pragma unroll 2
        for (int q0 = 0; q0 < nq; q0 += 2) {
            const bool two = q0 + 1 < nq;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (s.live) {
                v0 = src[q0];
                if (two) v1 = src[q0 + 1];
            }
            v0.x = seg_scan(v0.x, s.same); v0.y = seg_scan(v0.y, s.same); v0.z = seg_scan(v0.z, s.same); v0.w = seg_scan(v0.w, s.same);
            v1.x = seg_scan(v1.x, s.same); v1.y = seg_scan(v1.y, s.same); v1.z = seg_scan(v1.z, s.same); v1.w = seg_scan(v1.w, s.same);
            if (s.is_last) {
                if (s.partial) {
                    float4* o = reinterpret_cast<float4*>(p.part + (size_t)s.slot * p.K);
                    o[q0] = v0;
                    if (two) o[q0 + 1] = v1;
                } else {
                    if (p.emit) {
                        float4* o = reinterpret_cast<float4*>(p.gW_dense + rowoff);
                        o[q0] = v0;
                        if (two) o[q0 + 1] = v1;
                    }
                    if (p.apply) {
                        float4* wp = reinterpret_cast<float4*>(p.W + rowoff);
                        float4* ap = reinterpret_cast<float4*>(p.accW + rowoff);
                        float4 w0 = wp[q0], a0 = p.adagrad ? ap[q0] : v0;
                        float4 w1 = v1, a1 = v1;
                        if (two) { w1 = wp[q0 + 1]; if (p.adagrad) a1 = ap[q0 + 1]; }
                        opt_apply4(w0, a0, v0, p.lr, p.adagrad);
                        wp[q0] = w0;
                        if (p.adagrad) ap[q0] = a0;
                        if (two) {
                            opt_apply4(w1, a1, v1, p.lr, p.adagrad);
                            wp[q0 + 1] = w1;
                            if (p.adagrad) ap[q0 + 1] = a1;
                        }
                    }
                }
            }
        }
    }
}

// level 2, one CTA per chunk: if a multi-chunk segment starts here, lanes of warp 0 probe the first key of the next
// chunks 32 at a time to find its extent, the 8 warps sum disjoint strided subsets of its partial rows, and the CTA
// combines them in warp order.
template <int ROWLEN_MAX>
__device__ __forceinline__ int long_extent(const uint32_t* __restrict__ keys_s, int nchunks, int c0, uint32_t row, int lane) {
    int m = 0;
    for (int base = c0 + 1; base < nchunks; base += 32) {
        const int c = base + lane;
        const bool cont = c < nchunks && keys_s[(size_t)c << 5] == row;
        const unsigned bal = __ballot_sync(kFull, cont);
        if (bal == 0xffffffffu) { m += 32; continue; }
        m += __ffs(~bal) - 1;
        break;
    }
    return m;
}

__global__ void __launch_bounds__(256) k_w_long2(WArgs p) {
    extern __shared__ float red[];      // [8][K]
    __shared__ int sh_slot, sh_m;
    __shared__ uint32_t sh_row;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (p.n + 31) >> 5;
    for (int c0 = blockIdx.x; c0 < nchunks; c0 += gridDim.x) {
        if (warp == 0) {
            uint32_t row = 0;
            const int slot0 = long_segment_start(p.keys_s, p.n, c0, &row);
            int m = 0;
            if (slot0 >= 0) m = long_extent<0>(p.keys_s, nchunks, c0, row, lane);
            if (lane == 0) { sh_slot = slot0; sh_m = m; sh_row = row; }
        }
        __syncthreads();
        const int slot0 = sh_slot, m = sh_m;
        const uint32_t row = sh_row;
        if (slot0 >= 0) {
            // warp w sums partial rows of chunks c0+1+w, c0+1+w+8, ...
            for (int k0 = 0; k0 < p.K; k0 += 32) {
                const int k = k0 + lane;
                float acc = 0.f;
                constexpr int UN = 8;
                for (int i = warp; i < m; i += 8 * UN) {
                    float x[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int ii = i + 8 * u;
                        x[u] = (ii < m && k < p.K) ? p.part[(size_t)(2 * (c0 + 1 + ii)) * p.K + k] : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) acc += x[u];
                }
                if (k < p.K) red[warp * p.K + k] = acc;
            }
            __syncthreads();
            for (int k = threadIdx.x; k < p.K; k += 256) {
                float g = p.part[(size_t)slot0 * p.K + k];
#pragma unroll
                for (int w = 0; w < 8; ++w) g += red[w * p.K + k];
                const size_t idx = (size_t)row * p.K + k;
                if (p.emit) p.gW_dense[idx] = g;
                if (p.apply) {
                    float wv = p.W[idx];
                    if (p.adagrad) {
                        float a = p.accW[idx];
                        adagrad_apply(wv, a, g, p.lr);
                        p.accW[idx] = a;
                    } else {
                        wv -= p.lr * g;
                    }
                    p.W[idx] = wv;
                }
            }
        }
        __syncthreads();
    }
}

// entity rows, lane = occurrence: payload = coef * direction row (ev, stride dp, float4-aligned) and the bias scalar
__global__ void __launch_bounds__(256) k_entity_scan(EntArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int nchunks = (p.n + 31) >> 5;
    const int PE = p.dp + 4;
    const int nq = p.dp >> 2;
    const bool vec = (p.d & 3) == 0;
    for (int c = gw; c < nchunks; c += nw) {
        const ScanCtx s = scan_ctx(p.keys_s, p.n, c, lane);
        const uint32_t mine = s.live ? p.vals_s[(c << 5) + lane] : 0u;
        const int slot_m = (int)(mine / (uint32_t)p.B);
        const int b_m = (int)(mine - (uint32_t)slot_m * (uint32_t)p.B);
        int vs; float coef, bias;
        if (slot_m == 0) { vs = E_GA1; coef = 1.f; bias = p.sc[(size_t)b_m * SC_N + SC_GU1]; }
        else if (slot_m == 1) { vs = E_GA2; coef = 1.f; bias = p.sc[(size_t)b_m * SC_N + SC_GU2]; }
        else if (slot_m < 2 + p.S) { vs = E_V1; coef = p.gn1[(size_t)(slot_m - 2) * p.B + b_m]; bias = coef; }
        else { vs = E_V2; coef = p.gn2[(size_t)(slot_m - 2 - p.S) * p.B + b_m]; bias = coef; }
        if (!s.live) { coef = 0.f; bias = 0.f; }
        const float4* src = reinterpret_cast<const float4*>(p.ev + (size_t)(b_m * E_NV + vs) * p.dp);
        const float gb = seg_scan(bias, s.same);
        const size_t rowoff = (size_t)s.key * p.d;
        if (s.is_last) {
            if (s.partial) {
                p.part[(size_t)s.slot * PE + p.dp] = gb;
            } else {
                if (p.emit) p.gAb_dense[s.key] = gb;
                if (p.apply) {
                    float w = p.Ab[s.key];
                    if (p.adagrad) {
                        float a = p.accAb[s.key];
                        adagrad_apply(w, a, gb, p.lr);
                        p.accAb[s.key] = a;
                    } else {
                        w -= p.lr * gb;
                    }
                    p.Ab[s.key] = w;
                }
            }
        }
#pragma unroll 2
        for (int q0 = 0; q0 < nq; q0 += 2) {
            const bool two = q0 + 1 < nq;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (s.live) {
                v0 = src[q0];
                if (two) v1 = src[q0 + 1];
            }
            v0.x *= coef; v0.y *= coef; v0.z *= coef; v0.w *= coef;
            v1.x *= coef; v1.y *= coef; v1.z *= coef; v1.w *= coef;
            v0.x = seg_scan(v0.x, s.same); v0.y = seg_scan(v0.y, s.same); v0.z = seg_scan(v0.z, s.same); v0.w = seg_scan(v0.w, s.same);
            v1.x = seg_scan(v1.x, s.same); v1.y = seg_scan(v1.y, s.same); v1.z = seg_scan(v1.z, s.same); v1.w = seg_scan(v1.w, s.same);
            if (s.is_last) {
                if (s.partial) {
                    float4* o = reinterpret_cast<float4*>(p.part + (size_t)s.slot * PE);
                    o[q0] = v0;
                    if (two) o[q0 + 1] = v1;
                } else if (vec) {
                    if (p.emit) {
                        float4* o = reinterpret_cast<float4*>(p.gA_dense + rowoff);
                        o[q0] = v0;
                        if (two) o[q0 + 1] = v1;
                    }
                    if (p.apply) {
                        float4* wp = reinterpret_cast<float4*>(p.A + rowoff);
                        float4* ap = reinterpret_cast<float4*>(p.accA + rowoff);
                        float4 w0 = wp[q0], a0 = p.adagrad ? ap[q0] : v0;
                        float4 w1 = v1, a1 = v1;
                        if (two) { w1 = wp[q0 + 1]; if (p.adagrad) a1 = ap[q0 + 1]; }
                        opt_apply4(w0, a0, v0, p.lr, p.adagrad);
                        wp[q0] = w0;
                        if (p.adagrad) ap[q0] = a0;
                        if (two) {
                            opt_apply4(w1, a1, v1, p.lr, p.adagrad);
                            wp[q0 + 1] = w1;
                            if (p.adagrad) ap[q0 + 1] = a1;
                        }
                    }
                } else {
                    const float gv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int j = 4 * q0 + u;
                        if (j < p.d) {
                            const size_t idx = rowoff + j;
                            if (p.emit) p.gA_dense[idx] = gv[u];
                            if (p.apply) {
                                float w = p.A[idx];
                                if (p.adagrad) {
                                    float a = p.accA[idx];
                                    adagrad_apply(w, a, gv[u], p.lr);
                                    p.accA[idx] = a;
                                } else {
                                    w -= p.lr * gv[u];
                                }
                                p.A[idx] = w;
                            }
                        }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_entity_long2(EntArgs p) {
    extern __shared__ float red[];      // [8][PE]
    __shared__ int sh_slot, sh_m;
    __shared__ uint32_t sh_row;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (p.n + 31) >> 5;
    const int PE = p.dp + 4;
    for (int c0 = blockIdx.x; c0 < nchunks; c0 += gridDim.x) {
        if (warp == 0) {
            uint32_t row = 0;
            const int slot0 = long_segment_start(p.keys_s, p.n, c0, &row);
            int m = 0;
            if (slot0 >= 0) m = long_extent<0>(p.keys_s, nchunks, c0, row, lane);
            if (lane == 0) { sh_slot = slot0; sh_m = m; sh_row = row; }
        }
        __syncthreads();
        const int slot0 = sh_slot, m = sh_m;
        const uint32_t row = sh_row;
        if (slot0 >= 0) {
            for (int j0 = 0; j0 <= p.dp; j0 += 32) {       // column dp holds the bias partial
                const int j = j0 + lane;
                float acc = 0.f;
                constexpr int UN = 8;
                for (int i = warp; i < m; i += 8 * UN) {
                    float x[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int ii = i + 8 * u;
                        x[u] = (ii < m && j <= p.dp) ? p.part[(size_t)(2 * (c0 + 1 + ii)) * PE + j] : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) acc += x[u];
                }
                if (j <= p.dp) red[warp * PE + j] = acc;
            }
            __syncthreads();
            for (int j = threadIdx.x; j <= p.dp; j += 256) {
                if (j >= p.d && j != p.dp) continue;
                float g = p.part[(size_t)slot0 * PE + j];
#pragma unroll
                for (int w = 0; w < 8; ++w) g += red[w * PE + j];
                float* P = (j == p.dp) ? p.Ab + row : p.A + (size_t)row * p.d + j;
                float* AC = (j == p.dp) ? p.accAb + row : p.accA + (size_t)row * p.d + j;
                if (p.emit) {
                    if (j == p.dp) p.gAb_dense[row] = g; else p.gA_dense[(size_t)row * p.d + j] = g;
                }
                if (p.apply) {
                    float wv = *P;
                    if (p.adagrad) {
                        float a = *AC;
                        adagrad_apply(wv, a, g, p.lr);
                        *AC = a;
                    } else {
                        wv -= p.lr * g;
                    }
                    *P = wv;
                }
            }
        }
        __syncthreads();
    }
}

// ---- dense parameters ----
// sum the batch-split partials of dC/dC1/dC2 and the per-CTA partials of dWb into the flat dense gradient buffer.
// blocks [0, n_elem_blocks): 4 consecutive elements per thread (float4 when aligned), the splits summed in split order;
// blocks [n_elem_blocks, +K): one CTA per column k of dWb - thread t sums partial rows t, t+256, ... and a fixed-shape
// shared-memory tree combines the 256 values (deterministic).
__global__ void __launch_bounds__(256) k_dense_finalize(const float* __restrict__ part, int nsplit, size_t n_units_elems,
                                                        const float* __restrict__ dzsum_part, int n_dz_part, int K,
                                                        float* __restrict__ out, size_t off_wb, int n_elem_blocks) {
    if ((int)blockIdx.x < n_elem_blocks) {
        const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (i0 >= n_units_elems) return;
        if ((n_units_elems & 3) == 0) {
            float4 s = *reinterpret_cast<const float4*>(part + i0);
            for (int sp = 1; sp < nsplit; ++sp) {
                const float4 x = *reinterpret_cast<const float4*>(part + (size_t)sp * n_units_elems + i0);
                s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
            }
            *reinterpret_cast<float4*>(out + i0) = s;
        } else {
            for (size_t i = i0; i < min(i0 + 4, n_units_elems); ++i) {
                float s = 0.f;
                for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * n_units_elems + i];
                out[i] = s;
            }
        }
        return;
    }
    __shared__ float red[256];
    const int k = (int)blockIdx.x - n_elem_blocks;
    float s = 0.f;
    for (int c = threadIdx.x; c < n_dz_part; c += 256) s += dzsum_part[(size_t)c * K + k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[off_wb + k] = red[0];
}

// elementwise optimiser over up to 5 dense tensors in one launch (blockIdx.y = tensor):
// g = grad (+ adj*(l1*sign(p) + 2*l2*p)); AdaGrad (Optimizers.py:29-32) or SGD (:51).
// Also used for W when the regulariser makes its gradient dense (OieModel.py:54-56).
struct DenseJob { float* p; float* acc; float* grad; size_t n; float reg_l1, reg_l2; int write_back; };
struct DenseJobs { DenseJob j[5]; int count; float lr; int adagrad; };

__global__ void __launch_bounds__(256) k_dense_apply(DenseJobs jobs) {
    const DenseJob jb = jobs.j[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < jb.n; i += (size_t)gridDim.x * blockDim.x) {
        float w = jb.p[i];
        float g = jb.grad[i];
        if (jb.reg_l1 != 0.f || jb.reg_l2 != 0.f) {
            const float sgn = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : 0.f);
            g += jb.reg_l1 * sgn + 2.f * jb.reg_l2 * w;
            if (jb.write_back) jb.grad[i] = g;
        }
        if (jobs.adagrad) {
            float a = jb.acc[i];
            adagrad_apply(w, a, g, jobs.lr);
            jb.acc[i] = a;
        } else {
            w -= jobs.lr * g;
        }
        jb.p[i] = w;
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
// one CTA; thread t owns elements t, t+256, ... and the tree below has a fixed shape -> deterministic
__global__ void __launch_bounds__(256) k_cost(const double* __restrict__ loss_part, int n_loss,
                                              const double* __restrict__ reg_part, int n_reg, double invZ,
                                              double adj_l1, double adj_l2, double* __restrict__ cost) {
    __shared__ double s0[256], s1[256], s2[256];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n_loss; i += 256) a += loss_part[i];
    for (int i = threadIdx.x; i < n_reg; i += 256) { b += reg_part[2 * i]; c += reg_part[2 * i + 1]; }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b; s2[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s0[threadIdx.x] += s0[threadIdx.x + o];
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *cost = -s0[0] * invZ + adj_l1 * s1[0] + adj_l2 * s2[0];
}

int ensure_part(rae_engine* h, float** buf, size_t* cap, size_t need) {
    if (need <= *cap) return RAE_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = need + need / 4;
    cudaError_t e = cudaMalloc((void**)buf, want * sizeof(float));
    if (e != cudaSuccess) return fail(h, RAE_ENOMEM, "cudaMalloc(partials %zu floats) failed: %s", want, cudaGetErrorString(e));
    *cap = want;
    return RAE_OK;
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
    const int blocks = std::min((n + 255) / 256, 4 * h->num_sms);
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

int sort_pairs(rae_engine* h, SegWork& w, int64_t n, cudaStream_t st, void* tmp, size_t tmp_bytes) {
    if (n > w.capacity) return fail(h, RAE_EINVAL, "internal: sort workspace too small (%lld > %lld)", (long long)n, (long long)w.capacity);
    if (n <= 0) return RAE_OK;
    size_t bytes = tmp_bytes;
    // LSD radix sort is stable: equal rows keep ascending occurrence order == np.argsort(kind='stable')
    RAE_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp, bytes, w.keys, w.keys_s, w.vals, w.vals_s, (int)n, 0,
                                                w.key_bits, st));
    h->launches += (w.key_bits + 7) / 8 + 2;
    return RAE_OK;
}

int segment_heads(rae_engine* h, const uint32_t* keys_s, int64_t n, SegWork& w, cudaStream_t st) {
    if (n > w.capacity) return fail(h, RAE_EINVAL, "internal: segment workspace too small");
    if (n > 0) {
        const int blocks = std::min((int)((n + 255) / 256), 4 * h->num_sms);
        k_flag_heads<<<blocks, 256, 0, st>>>(keys_s, (int)n, w.flags);
        size_t bytes = h->cub_bytes;
        RAE_CUDA(h, cub::DeviceScan::ExclusiveSum(h->cub_tmp, bytes, w.flags, w.pos, (int)n, st));
        k_scatter_heads<<<blocks, 256, 0, st>>>(w.flags, w.pos, (int)n, w.seg_start, w.n_seg);
    } else {
        k_scatter_heads<<<1, 32, 0, st>>>(w.flags, w.pos, 0, w.seg_start, w.n_seg);
    }
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int count_unique(rae_engine* h, const uint32_t* keys_s, int64_t n, int32_t* out_dev, cudaStream_t st) {
    RAE_CUDA(h, cudaMemsetAsync(out_dev, 0, sizeof(int32_t), st));
    if (n > 0) {
        const int blocks = std::min((int)((n + 255) / 256), 4 * h->num_sms);
        k_count_heads<<<blocks, 256, 0, st>>>(keys_s, (int)n, out_dev);
        RAE_CUDA(h, cudaGetLastError());
    }
    return RAE_OK;
}

int launch_entity_update(rae_engine* h, const uint32_t* keys_s, const uint32_t* vals_s, int64_t n_occ, bool emit_dense,
                         bool apply, cudaStream_t st) {
    if (n_occ <= 0) return RAE_OK;
    EntArgs p{};
    p.keys_s = keys_s; p.vals_s = vals_s;
    p.ev = h->ev; p.sc = h->sc; p.gn1 = h->gn1; p.gn2 = h->gn2;
    p.A = h->P[RAE_P_A]; p.Ab = h->P[RAE_P_AB]; p.accA = h->ACC[RAE_P_A]; p.accAb = h->ACC[RAE_P_AB];
    p.gA_dense = h->gA_dense; p.gAb_dense = h->gAb_dense;
    p.B = h->B; p.S = h->S; p.d = h->d; p.dp = h->dp; p.n = (int)n_occ;
    p.lr = (float)h->cfg.lr; p.adagrad = h->adagrad; p.emit = emit_dense; p.apply = apply;
    const int64_t nchunks = (n_occ + 31) / 32;
    int rc = ensure_part(h, &h->ent_part, &h->ent_part_cap, (size_t)(2 * nchunks) * (h->dp + 4));
    if (rc) return rc;
    p.part = h->ent_part;
    const int blocks = (int)std::min<int64_t>((nchunks + 7) / 8, (int64_t)h->num_sms * 8);
    const int dt = (h->d + 31) / 32;
    (void)dt;
    {
        const int blocks2 = (int)std::min<int64_t>(nchunks, (int64_t)h->num_sms * 16);
        k_entity_scan<<<blocks, 256, 0, st>>>(p);
        k_entity_long2<<<blocks2, 256, sizeof(float) * 8 * (h->dp + 4), st>>>(p);
    }
    h->launches += 2;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_w_update(rae_engine* h, const uint32_t* keys_s, const uint32_t* vals_s, int64_t nnz, bool emit_dense,
                    bool apply, cudaStream_t st) {
    if (nnz <= 0) return RAE_OK;
    WArgs p{};
    p.keys_s = keys_s; p.vals_s = vals_s;
    p.dz = h->dz; p.W = h->P[RAE_P_W]; p.accW = h->ACC[RAE_P_W]; p.gW_dense = h->gW_dense;
    p.K = h->K; p.n = (int)nnz; p.lr = (float)h->cfg.lr; p.adagrad = h->adagrad; p.emit = emit_dense; p.apply = apply;
    const int64_t nchunks = (nnz + 31) / 32;
    int rc = ensure_part(h, &h->feat_part, &h->feat_part_cap, (size_t)(2 * nchunks) * h->K);
    if (rc) return rc;
    p.part = h->feat_part;
    const int blocks = (int)std::min<int64_t>((nchunks + 7) / 8, (int64_t)h->num_sms * 8);
    const int kt = (h->K + 31) / 32;
#define RAE_WU(KT)                                     \
    do {                                               \
        k_w_chunks<KT><<<blocks, 256, 0, st>>>(p);     \
        k_w_long<KT><<<blocks, 256, 0, st>>>(p);       \
    } while (0)
    if ((h->K & 3) == 0) {
        const int blocks2 = (int)std::min<int64_t>(nchunks, (int64_t)h->num_sms * 16);
        k_w_scan<<<blocks, 256, 0, st>>>(p);
        k_w_long2<<<blocks2, 256, sizeof(float) * 8 * h->K, st>>>(p);
    }
    else if (kt <= 1) RAE_WU(1);
    else if (kt <= 2) RAE_WU(2);
    else if (kt <= 4) RAE_WU(4);
    else if (kt <= 8) RAE_WU(8);
    else if (kt <= 16) RAE_WU(16);
    else RAE_WU(32);
#undef RAE_WU
    h->launches += 2;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int build_row_keys(rae_engine* h, const int32_t* rows, int64_t n, cudaStream_t st) {
    if (n > h->feat.capacity) return fail(h, RAE_EINVAL, "internal: row-key workspace too small");
    const int blocks = std::min((int)((n + 255) / 256), 4 * h->num_sms);
    k_row_keys<<<blocks, 256, 0, st>>>(rows, (int)n, h->feat.keys, h->feat.vals);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_gather_rows(rae_engine* h, const float* table, int64_t width, const int32_t* rows, int64_t n, float* out,
                       cudaStream_t st) {
    if (n <= 0) return RAE_OK;
    const int blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)h->num_sms * 8);
    k_gather_rows<<<blocks, 256, 0, st>>>(table, (int)width, rows, (int)n, out);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

// generic owner-side sparse-row optimiser step: table[row,:] updated with the sum of grads[val,:] over the row's
// sorted occurrences (the same kernels as the W update with dz := grads, K := width)
int launch_rows_apply(rae_engine* h, float* table, float* acc, int width, const uint32_t* keys_s, const uint32_t* vals_s,
                      const float* grads, int64_t n, cudaStream_t st) {
    if (n <= 0) return RAE_OK;
    WArgs p{};
    p.keys_s = keys_s; p.vals_s = vals_s;
    p.dz = grads; p.W = table; p.accW = acc; p.gW_dense = nullptr;
    p.K = width; p.n = (int)n; p.lr = (float)h->cfg.lr; p.adagrad = h->adagrad; p.emit = 0; p.apply = 1;
    const int64_t nchunks = (n + 31) / 32;
    int rc = ensure_part(h, &h->feat_part, &h->feat_part_cap, (size_t)(2 * nchunks) * width);
    if (rc) return rc;
    p.part = h->feat_part;
    const int blocks = (int)std::min<int64_t>((nchunks + 7) / 8, (int64_t)h->num_sms * 8);
    const int kt = (width + 31) / 32;
#define RAE_RA(KT)                                     \
    do {                                               \
        k_w_chunks<KT><<<blocks, 256, 0, st>>>(p);     \
        k_w_long<KT><<<blocks, 256, 0, st>>>(p);       \
    } while (0)
    if ((width & 3) == 0) {
        const int blocks2 = (int)std::min<int64_t>(nchunks, (int64_t)h->num_sms * 16);
        k_w_scan<<<blocks, 256, 0, st>>>(p);
        k_w_long2<<<blocks2, 256, sizeof(float) * 8 * width, st>>>(p);
    }
    else if (kt <= 1) RAE_RA(1);
    else if (kt <= 2) RAE_RA(2);
    else if (kt <= 4) RAE_RA(4);
    else if (kt <= 8) RAE_RA(8);
    else if (kt <= 16) RAE_RA(16);
    else RAE_RA(32);
#undef RAE_RA
    h->launches += 2;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_dense_finalize(rae_engine* h, cudaStream_t st) {
    const size_t n_units = (size_t)h->off_gWb;   // elements of [C | C1 | C2]
    const int n_elem_blocks = (int)((n_units + 1023) / 1024);
    k_dense_finalize<<<n_elem_blocks + h->K, 256, 0, st>>>(h->gC_part, h->gC_nsplit, n_units, h->dzsum_part, h->dz_part_used, h->K,
                                                          h->dense_grad, (size_t)h->off_gWb, n_elem_blocks);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_dense_apply(rae_engine* h, cudaStream_t st) {
    const bool reg = (h->cfg.l1 != 0.0 || h->cfg.l2 != 0.0);
    const bool regdec = reg && h->cfg.ext_reg;
    const size_t dd = (size_t)h->d * h->d * h->K, dk = (size_t)h->d * h->K;
    const float r1 = (float)(h->cfg.adj * h->cfg.l1), r2 = (float)(h->cfg.adj * h->cfg.l2);
    DenseJobs jobs{};
    jobs.lr = (float)h->cfg.lr;
    jobs.adagrad = h->adagrad;
    size_t nmax = 0;
    auto add = [&](int pid, float* grad, size_t n, bool regularised) {
        if (n == 0) return;
        DenseJob& j = jobs.j[jobs.count++];
        j.p = h->P[pid]; j.acc = h->ACC[pid]; j.grad = grad; j.n = n;
        j.reg_l1 = regularised ? r1 : 0.f;
        j.reg_l2 = regularised ? r2 : 0.f;
        j.write_back = h->debug_dense ? 1 : 0;
        nmax = std::max(nmax, n);
    };
    if (h->hasM) add(RAE_P_C, h->dense_grad + h->off_gC, dd, regdec);
    if (h->hasSP) {
        add(RAE_P_C1, h->dense_grad + h->off_gC1, dk, regdec);
        add(RAE_P_C2, h->dense_grad + h->off_gC2, dk, regdec);
    }
    add(RAE_P_WB, h->dense_grad + h->off_gWb, (size_t)h->K, false);
    // regulariser makes dW dense: gW_dense holds the data gradient (zero rows elsewhere)
    if (h->dense_w) add(RAE_P_W, h->gW_dense, (size_t)h->cfg.F * h->K, reg);
    const int bx = (int)std::min<size_t>((nmax + 255) / 256, (size_t)h->num_sms * 8);
    k_dense_apply<<<dim3(bx, jobs.count), 256, 0, st>>>(jobs);
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
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
    const int n_loss = h->n_loss_part;
    k_cost<<<1, 256, 0, st>>>(h->loss_part, n_loss, h->reg_part, n_reg, 1.0 / h->Z, h->cfg.adj * h->cfg.l1,
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
