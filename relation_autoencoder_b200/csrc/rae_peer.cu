// Peer-memory data path of the row-sharded multi-GPU step (one process per GPU, NVLink / NVSwitch, CUDA IPC).
//
// The reference is single-process (no counterpart).  What is exchanged follows from its optimiser: Theano sums duplicate
// rows before AdaGrad squares the gradient (AdvancedIncSubtensor1 + Optimizers.py:29-32), so the owner of a row must see
// the SUM of every rank's contribution before it applies the rule once.
//
//   k_fetch_rows : compact[i,:] = shard[id % world][id / world, :]      rows read straight from the owner's HBM over
//                                                                      NVLink (or locally), one warp per row, 16-byte lanes
//   k_pull_apply : owner-side: g = sum_{rank order} grad_rank[slot,:]   remote reads of the ranks' reduced gradient rows,
//                  then ONE optimiser read-modify-write per row        fixed order -> bitwise reproducible
//
// No atomics, no collective: the only synchronisation the sparse path needs is "all ranks have emitted" (given by the
// dense-gradient all-reduce that precedes the pull) and "all owners have applied" (the end-of-step barrier).
#include <string.h>

#include <algorithm>

#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

namespace {

struct PeerPtrs {
    const float* p[RAE_MAX_PEERS];
};

// remote gradient rows were written by another GPU during this step: read them past L1 (ld.global.cv)
__device__ __forceinline__ float4 ld_cv4(const float4* p) { return __ldcv(p); }
__device__ __forceinline__ float ld_cv(const float* p) { return __ldcv(p); }

template <bool VEC>
__global__ void __launch_bounds__(256) k_fetch_rows(PeerPtrs tabs, int world, int width, const int32_t* __restrict__ ids, int n,
                                                    float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    constexpr int UN = 4;                       // rows in flight per warp
    for (int i0 = gw * UN; i0 < n; i0 += nw * UN) {
        const float* src[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = min(i0 + u, n - 1);
            const int id = ids[i];
            const int owner = id % world;
            src[u] = tabs.p[owner] + (size_t)(id / world) * width;
        }
        if (VEC) {
            const int nq = width >> 2;
            for (int q0 = lane; q0 < nq; q0 += 32) {
                float4 v[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(src[u]) + q0);
#pragma unroll
                for (int u = 0; u < UN; ++u)
                    if (i0 + u < n) reinterpret_cast<float4*>(out + (size_t)(i0 + u) * width)[q0] = v[u];
            }
        } else {
            for (int k = lane; k < width; k += 32) {
                float v[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) v[u] = __ldg(src[u] + k);
#pragma unroll
                for (int u = 0; u < UN; ++u)
                    if (i0 + u < n) out[(size_t)(i0 + u) * width + k] = v[u];
            }
        }
    }
}

// width == 1 (bias table): one thread per row
__global__ void __launch_bounds__(256) k_fetch_scalars(PeerPtrs tabs, int world, const int32_t* __restrict__ ids, int n,
                                                       float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int id = ids[i];
        out[i] = __ldg(tabs.p[id % world] + id / world);
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_pull_apply(float* __restrict__ table, float* __restrict__ acc, int width,
                                                    const int32_t* __restrict__ rows_local, const int32_t* __restrict__ ent_off,
                                                    const int32_t* __restrict__ ent_src, const int32_t* __restrict__ ent_slot,
                                                    int n_rows, PeerPtrs grads, float lr, int adagrad) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = gw; i < n_rows; i += nw) {
        const int e0 = ent_off[i], e1 = ent_off[i + 1];
        const size_t rowoff = (size_t)rows_local[i] * width;
        // lanes 0..(e1-e0) hold the entries (at most `world` <= 16 per row)
        int my_src = 0, my_slot = 0;
        if (e0 + lane < e1) { my_src = ent_src[e0 + lane]; my_slot = ent_slot[e0 + lane]; }
        const int ne = e1 - e0;
        if (VEC) {
            const int nq = width >> 2;
            for (int q0 = lane; q0 < ((nq + 31) & ~31); q0 += 32) {
                const bool in = q0 < nq;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 w = g, a = g;
                if (in) {
                    w = reinterpret_cast<const float4*>(table + rowoff)[q0];
                    if (adagrad) a = reinterpret_cast<const float4*>(acc + rowoff)[q0];
                }
                for (int e = 0; e < ne; ++e) {                      // rank order: deterministic
                    const int s = __shfl_sync(kFull, my_src, e), slot = __shfl_sync(kFull, my_slot, e);
                    if (in) {
                        const float4 x = ld_cv4(reinterpret_cast<const float4*>(grads.p[s] + (size_t)slot * width) + q0);
                        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
                    }
                }
                if (in) {
                    if (adagrad) {
                        adagrad_apply(w.x, a.x, g.x, lr); adagrad_apply(w.y, a.y, g.y, lr);
                        adagrad_apply(w.z, a.z, g.z, lr); adagrad_apply(w.w, a.w, g.w, lr);
                        reinterpret_cast<float4*>(acc + rowoff)[q0] = a;
                    } else {
                        w.x -= lr * g.x; w.y -= lr * g.y; w.z -= lr * g.z; w.w -= lr * g.w;
                    }
                    reinterpret_cast<float4*>(table + rowoff)[q0] = w;
                }
            }
        } else {
            for (int k0 = 0; k0 < width; k0 += 32) {
                const int k = k0 + lane;
                const bool in = k < width;
                float g = 0.f, w = 0.f, a = 0.f;
                if (in) {
                    w = table[rowoff + k];
                    if (adagrad) a = acc[rowoff + k];
                }
                for (int e = 0; e < ne; ++e) {
                    const int s = __shfl_sync(kFull, my_src, e), slot = __shfl_sync(kFull, my_slot, e);
                    if (in) g += ld_cv(grads.p[s] + (size_t)slot * width + k);
                }
                if (in) {
                    if (adagrad) {
                        adagrad_apply(w, a, g, lr);
                        acc[rowoff + k] = a;
                    } else {
                        w -= lr * g;
                    }
                    table[rowoff + k] = w;
                }
            }
        }
    }
}

// width == 1: one thread per row
__global__ void __launch_bounds__(256) k_pull_apply_scalars(float* __restrict__ table, float* __restrict__ acc,
                                                            const int32_t* __restrict__ rows_local,
                                                            const int32_t* __restrict__ ent_off, const int32_t* __restrict__ ent_src,
                                                            const int32_t* __restrict__ ent_slot, int n_rows, PeerPtrs grads,
                                                            float lr, int adagrad) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
        float g = 0.f;
        for (int e = ent_off[i]; e < ent_off[i + 1]; ++e) g += ld_cv(grads.p[ent_src[e]] + ent_slot[e]);
        const int r = rows_local[i];
        float w = table[r];
        if (adagrad) {
            float a = acc[r];
            adagrad_apply(w, a, g, lr);
            acc[r] = a;
        } else {
            w -= lr * g;
        }
        table[r] = w;
    }
}

int fill_peers(rae_engine* h, const void* const* ptrs, int world, PeerPtrs* out, const char* what) {
    if (world < 1 || world > RAE_MAX_PEERS) return fail(h, RAE_EINVAL, "%s: world %d out of range [1, %d]", what, world, RAE_MAX_PEERS);
    memset(out, 0, sizeof(*out));
    for (int r = 0; r < world; ++r) {
        if (!ptrs[r]) return fail(h, RAE_EINVAL, "%s: null table pointer for rank %d", what, r);
        out->p[r] = static_cast<const float*>(ptrs[r]);
    }
    return RAE_OK;
}

}  // namespace

int launch_fetch_rows(rae_engine* h, const void* const* tables, int world, int64_t width, const int32_t* ids, int64_t n,
                      float* out, cudaStream_t st) {
    if (n <= 0) return RAE_OK;
    PeerPtrs pp;
    int rc = fill_peers(h, tables, world, &pp, "rae_fetch_rows");
    if (rc) return rc;
    if (width == 1) {
        const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->num_sms * 8);
        k_fetch_scalars<<<blocks, 256, 0, st>>>(pp, world, ids, (int)n, out);
    } else {
        const int64_t warps = (n + 3) / 4;
        const int blocks = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->num_sms * 8);
        if ((width & 3) == 0) k_fetch_rows<true><<<blocks, 256, 0, st>>>(pp, world, (int)width, ids, (int)n, out);
        else k_fetch_rows<false><<<blocks, 256, 0, st>>>(pp, world, (int)width, ids, (int)n, out);
    }
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

int launch_pull_apply(rae_engine* h, float* table, float* acc, int64_t width, const int32_t* rows_local, const int32_t* ent_off,
                      const int32_t* ent_src, const int32_t* ent_slot, int64_t n_rows, const void* const* grads, int world,
                      cudaStream_t st) {
    if (n_rows <= 0) return RAE_OK;
    PeerPtrs pp;
    int rc = fill_peers(h, grads, world, &pp, "rae_pull_apply");
    if (rc) return rc;
    const float lr = (float)h->cfg.lr;
    const int adagrad = h->adagrad ? 1 : 0;
    if (width == 1) {
        const int blocks = (int)std::min<int64_t>((n_rows + 255) / 256, (int64_t)h->num_sms * 8);
        k_pull_apply_scalars<<<blocks, 256, 0, st>>>(table, acc, rows_local, ent_off, ent_src, ent_slot, (int)n_rows, pp, lr, adagrad);
    } else {
        const int blocks = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)h->num_sms * 16);
        if ((width & 3) == 0)
            k_pull_apply<true><<<blocks, 256, 0, st>>>(table, acc, (int)width, rows_local, ent_off, ent_src, ent_slot, (int)n_rows, pp, lr, adagrad);
        else
            k_pull_apply<false><<<blocks, 256, 0, st>>>(table, acc, (int)width, rows_local, ent_off, ent_src, ent_slot, (int)n_rows, pp, lr, adagrad);
    }
    h->launches++;
    RAE_CUDA(h, cudaGetLastError());
    return RAE_OK;
}

}  // namespace rae

using namespace rae;

extern "C" {

int rae_peer_alloc(int64_t bytes, void** ptr_out, void* handle_out) {
    if (!ptr_out || bytes <= 0) return fail(nullptr, RAE_EINVAL, "rae_peer_alloc: bad argument");
    *ptr_out = nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return fail(nullptr, RAE_ENOMEM, "rae_peer_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    if ((e = cudaMemset(p, 0, (size_t)bytes)) != cudaSuccess) {
        cudaFree(p);
        return fail(nullptr, RAE_ECUDA, "rae_peer_alloc: cudaMemset failed: %s", cudaGetErrorString(e));
    }
    if (handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == RAE_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t hd;
        if ((e = cudaIpcGetMemHandle(&hd, p)) != cudaSuccess) {
            cudaFree(p);
            return fail(nullptr, RAE_ECUDA, "rae_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        }
        memcpy(handle_out, &hd, sizeof(hd));
    }
    *ptr_out = p;
    return RAE_OK;
}

int rae_peer_free(void* ptr) {
    if (!ptr) return RAE_OK;
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) return fail(nullptr, RAE_ECUDA, "rae_peer_free: %s", cudaGetErrorString(e));
    return RAE_OK;
}

int rae_peer_open(const void* handle, void** ptr_out) {
    if (!handle || !ptr_out) return fail(nullptr, RAE_EINVAL, "rae_peer_open: null argument");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, sizeof(hd));
    cudaError_t e = cudaIpcOpenMemHandle(ptr_out, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        *ptr_out = nullptr;
        return fail(nullptr, RAE_ECUDA, "rae_peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    }
    return RAE_OK;
}

int rae_peer_close(void* ptr) {
    if (!ptr) return RAE_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) return fail(nullptr, RAE_ECUDA, "rae_peer_close: %s", cudaGetErrorString(e));
    return RAE_OK;
}

int rae_fetch_rows(rae_engine* h, const void* const* tables, int32_t world, int64_t width, const int32_t* ids, int64_t n,
                   float* out, void* stream) {
    if (!h || !tables || width < 1 || n < 0 || (n > 0 && (!ids || !out))) return fail(h, RAE_EINVAL, "rae_fetch_rows: bad argument");
    return launch_fetch_rows(h, tables, world, width, ids, n, out, (cudaStream_t)stream);
}

int rae_pull_apply(rae_engine* h, float* table, float* acc, int64_t width, const int32_t* rows_local, const int32_t* ent_off,
                   const int32_t* ent_src, const int32_t* ent_slot, int64_t n_rows, const void* const* grads, int32_t world,
                   void* stream) {
    if (!h || !table || !grads || width < 1 || n_rows < 0) return fail(h, RAE_EINVAL, "rae_pull_apply: bad argument");
    if (h->adagrad && !acc) return fail(h, RAE_EINVAL, "rae_pull_apply: AdaGrad needs the accumulator table");
    if (n_rows > 0 && (!rows_local || !ent_off || !ent_src || !ent_slot)) return fail(h, RAE_EINVAL, "rae_pull_apply: null plan arrays");
    return launch_pull_apply(h, table, acc, width, rows_local, ent_off, ent_src, ent_slot, n_rows, grads, world, (cudaStream_t)stream);
}

}  // extern "C"
