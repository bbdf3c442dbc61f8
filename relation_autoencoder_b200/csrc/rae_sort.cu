// Stable LSD radix sort of the step's (row, occurrence) pairs: ONE cooperative kernel for all passes.
//
//   reference: the duplicate-row accumulation of AdvancedIncSubtensor1 behind T.grad, learning/Optimizers.py:27-32 - here a
//              deterministic segmented scatter: sort the occurrences by row (equal rows keep ascending occurrence order,
//              == np.argsort(kind='stable')), then reduce every row's segment in that order (rae_update.cu).
//
// The batches are small for a device-wide sort (49 k pairs at config 2, 172 k at the target shape): a library radix sort
// spends its time in launches and in the serial look-back chain of its few large tiles (3 x 12.7 us per pass + histogram +
// scan kernels measured).  Here the pairs are dealt over up to one CTA per SM (two thousand pairs each), each warp owns a
// contiguous run of its CTA's tile and keeps it in registers, and a pass is
//   sweep 1   digit histogram per warp (shared-memory counters owned by the warp)
//   exchange  CTA totals -> global table [CTA][256]; grid barrier; every CTA reads the table (coalesced, 148 rows) and
//             derives its own start offset per digit: digits below + same digit in CTAs before it + warps before it
//   sweep 2   the warp re-reads its pairs in order; __match_any_sync ranks equal digits inside the round by lane, the
//             running per-warp offsets rank them across rounds -> stable scatter into the other buffer; grid barrier.
// 8-bit digits, ceil(key_bits / 8) passes, ping-pong between the output arrays and the scratch so that the last pass lands
// in (keys_s, vals_s).  Integer atomics appear only on the warp-private shared-memory counters of sweep 1 (counts: order-free).
// The entity variant builds the pairs on the fly in pass 0 (key = entity id of occurrence o, value = o) instead of reading
// arrays written by a separate kernel.
#include <cooperative_groups.h>

#include <algorithm>

#include "rae_common.cuh"
#include "rae_internal.h"

namespace cg = cooperative_groups;

namespace rae {

namespace {

constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_RG = 8;             // rounds (of 32 pairs) a warp keeps in registers at a time
constexpr int SORT_TB = 22;            // table rows a thread has in flight in the exchange

struct SortArgs {
    // pass-0 source: arrays (mode 0) or the entity occurrence table (mode 1: slot 0 args1, 1 args2, 2+s neg1[s], 2+S+s neg2[s])
    const uint32_t* keys; const uint32_t* vals;
    const int32_t* a1; const int32_t* a2; const int32_t* neg1; const int32_t* neg2; long long neg_ld; int B, S;
    uint32_t* k_out; uint32_t* v_out;      // final sorted pairs
    uint32_t* k_tmp; uint32_t* v_tmp;      // scratch of the same size
    uint32_t* hist;                        // [gridDim.x][256]
    int n, passes;
    int L;                                 // pairs per warp run (a multiple of 32)
};

template <int MODE>
__device__ __forceinline__ void load_pair(const SortArgs& p, const uint32_t* ks, const uint32_t* vs, int pass, int i, uint32_t& key,
                                          uint32_t& val) {
    if (pass > 0 || MODE == 0) {
        key = ks[i];
        val = vs[i];
        return;
    }
    const int slot = i / p.B, b = i - slot * p.B;
    int row;
    if (slot == 0) row = p.a1[b];
    else if (slot == 1) row = p.a2[b];
    else if (slot < 2 + p.S) row = p.neg1[(size_t)(slot - 2) * p.neg_ld + b];
    else row = p.neg2[(size_t)(slot - 2 - p.S) * p.neg_ld + b];
    key = (uint32_t)row;
    val = (uint32_t)i;
}

template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS) k_radix_sort(SortArgs p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t wcnt[SORT_WARPS][256];      // sweep 1: per-warp digit counts; sweep 2: per-warp running offsets
    __shared__ uint32_t tot2[2][256], bef2[2][256];
    __shared__ uint32_t wsum[8];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int G = gridDim.x, me = blockIdx.x;
    // the warp runs of L pairs tile [0, n) in order: CTA me, warp w owns run me * SORT_WARPS + w
    const int L = p.L, R = L >> 5;
    const long long wbeg_ll = ((long long)me * SORT_WARPS + warp) * L;
    const int wbeg = (int)min(wbeg_ll, (long long)p.n), wend = (int)min(wbeg_ll + L, (long long)p.n);
    const unsigned lt = (1u << lane) - 1u;
    const bool resident = R <= SORT_RG;      // the warp's whole run stays in registers from sweep 1 to sweep 2

    for (int pass = 0; pass < p.passes; ++pass) {
        const int shift = 8 * pass;
        // the last pass writes the final arrays; earlier ones alternate backwards from there
        const bool to_out = ((p.passes - 1 - pass) & 1) == 0;
        // this pass's source = the previous pass's target; pass 0 reads the caller's arrays (mode 0) or builds the pairs (mode 1)
        const uint32_t* ks = pass == 0 ? p.keys : (to_out ? p.k_tmp : p.k_out);
        const uint32_t* vs = pass == 0 ? p.vals : (to_out ? p.v_tmp : p.v_out);
        uint32_t* kd = to_out ? p.k_out : p.k_tmp;
        uint32_t* vd = to_out ? p.v_out : p.v_tmp;
        // all loads of a group of rounds are issued before the first use: one memory latency per group
        auto load_group = [&](int r0, uint32_t (&key)[SORT_RG], uint32_t (&val)[SORT_RG]) {
#pragma unroll
            for (int u = 0; u < SORT_RG; ++u) {
                const int i = wbeg + ((r0 + u) << 5) + lane;
                key[u] = 0u; val[u] = 0u;
                if (r0 + u < R && i < wend) load_pair<MODE>(p, ks, vs, pass, i, key[u], val[u]);
            }
        };
        uint32_t rkey[SORT_RG], rval[SORT_RG];

        for (int i = t; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcnt[0][0])[i] = 0u;
        __syncthreads();
        // ---- sweep 1: per-warp histogram ----
        for (int r0 = 0; r0 < R; r0 += SORT_RG) {
            load_group(r0, rkey, rval);
#pragma unroll
            for (int u = 0; u < SORT_RG; ++u)
                if (r0 + u < R && wbeg + ((r0 + u) << 5) + lane < wend) atomicAdd(&wcnt[warp][(rkey[u] >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (t < 256) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) {
                const uint32_t c = wcnt[w][t];
                wcnt[w][t] = run;              // exclusive over the warps of the CTA
                run += c;
            }
            p.hist[(size_t)me * 256 + t] = run;
        }
        grid.sync();
        // ---- this CTA's start offset per digit: every thread sums half of the table's rows for one digit ----
        {
            const int part = t >> 8, dgt = t & 255;
            const int Gq = (G + 1) >> 1, c0 = part * Gq, c1 = min(G, c0 + Gq);
            uint32_t tot = 0, before = 0;
            for (int cb = c0; cb < c1; cb += SORT_TB) {
                uint32_t v[SORT_TB];
#pragma unroll
                for (int u = 0; u < SORT_TB; ++u) v[u] = cb + u < c1 ? p.hist[(size_t)(cb + u) * 256 + dgt] : 0u;
#pragma unroll
                for (int u = 0; u < SORT_TB; ++u) {
                    tot += v[u];
                    before += cb + u < me ? v[u] : 0u;
                }
            }
            tot2[part][dgt] = tot;
            bef2[part][dgt] = before;
        }
        __syncthreads();
        if (t < 256) {
            const uint32_t tot = tot2[0][t] + tot2[1][t];
            const uint32_t before = bef2[0][t] + bef2[1][t];
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) wsum[warp] = incl;
            // (only warps 0..7 take part: a named barrier over their 256 threads)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            uint32_t base = incl - tot;
            for (int w = 0; w < warp; ++w) base += wsum[w];
            const uint32_t start = base + before;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) wcnt[w][t] += start;
        }
        __syncthreads();
        // ---- sweep 2: stable scatter ----
        for (int r0 = 0; r0 < R; r0 += SORT_RG) {
            if (!resident) load_group(r0, rkey, rval);
#pragma unroll
            for (int u = 0; u < SORT_RG; ++u) {
                if (r0 + u >= R) continue;
                const bool valid = wbeg + ((r0 + u) << 5) + lane < wend;
                const uint32_t dg = (rkey[u] >> shift) & 255u;
                const unsigned mask = __match_any_sync(kFull, valid ? dg : (256u + (uint32_t)lane));
                const uint32_t base = valid ? wcnt[warp][dg] : 0u;
                const uint32_t rank = (uint32_t)__popc(mask & lt);
                __syncwarp();
                if (valid && rank == 0u) wcnt[warp][dg] = base + (uint32_t)__popc(mask);
                __syncwarp();
                if (valid) {
                    kd[base + rank] = rkey[u];
                    vd[base + rank] = rval[u];
                }
            }
        }
        if (pass + 1 < p.passes) grid.sync();
    }
}

}  // namespace

size_t radix_sort_temp_bytes(int64_t n, int num_sms) {
    return (size_t)n * 8 + (size_t)num_sms * 256 * sizeof(uint32_t) + 512;
}

// mode 0: (keys, vals) arrays; mode 1: entity occurrences.  tmp must hold radix_sort_temp_bytes(n).
static int radix_sort_launch(rae_engine* h, SortArgs& a, int mode, int64_t n, int key_bits, cudaStream_t st, void* tmp, size_t tmp_bytes) {
    if (n <= 0) return RAE_OK;
    if (tmp_bytes < radix_sort_temp_bytes(n, h->num_sms)) return fail(h, RAE_EINVAL, "internal: sort scratch too small");
    a.n = (int)n;
    a.passes = std::max(1, (key_bits + 7) / 8);
    uint8_t* t8 = static_cast<uint8_t*>(tmp);
    a.k_tmp = reinterpret_cast<uint32_t*>(t8);
    a.v_tmp = a.k_tmp + n;
    a.hist = reinterpret_cast<uint32_t*>(t8 + (((size_t)n * 8 + 255) & ~(size_t)255));
    // at least four rounds of 32 pairs per warp before another CTA is worth its row of the exchange table (172 k pairs:
    // 84 CTAs of 16 KB shared memory and half an SM's registers, so they fit beside the kernels of the other streams)
    const int G = (int)std::min<int64_t>(h->num_sms, std::max<int64_t>(1, (n + SORT_THREADS * 4 - 1) / (SORT_THREADS * 4)));
    const int64_t runs = (int64_t)G * SORT_WARPS;
    a.L = (int)(((n + runs - 1) / runs + 31) / 32 * 32);
    void* args[] = {&a};
    const void* fn = mode == 0 ? (const void*)k_radix_sort<0> : (const void*)k_radix_sort<1>;
    RAE_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(SORT_THREADS), args, 0, st));
    h->launches++;
    return RAE_OK;
}

int radix_sort_pairs(rae_engine* h, const uint32_t* keys, const uint32_t* vals, uint32_t* keys_s, uint32_t* vals_s, int64_t n,
                     int key_bits, cudaStream_t st, void* tmp, size_t tmp_bytes) {
    SortArgs a{};
    a.keys = keys; a.vals = vals; a.k_out = keys_s; a.v_out = vals_s;
    return radix_sort_launch(h, a, 0, n, key_bits, st, tmp, tmp_bytes);
}

int radix_sort_entities(rae_engine* h, const int32_t* a1, const int32_t* a2, const int32_t* neg1, const int32_t* neg2,
                        int64_t neg_ld, uint32_t* keys_s, uint32_t* vals_s, int key_bits, cudaStream_t st, void* tmp,
                        size_t tmp_bytes) {
    SortArgs a{};
    a.a1 = a1; a.a2 = a2; a.neg1 = neg1; a.neg2 = neg2; a.neg_ld = neg_ld; a.B = h->B; a.S = h->S;
    a.k_out = keys_s; a.v_out = vals_s;
    return radix_sort_launch(h, a, 1, (int64_t)(2 + 2 * h->S) * h->B, key_bits, st, tmp, tmp_bytes);
}

}  // namespace rae
