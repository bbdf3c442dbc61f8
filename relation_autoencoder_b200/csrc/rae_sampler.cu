// Device-side negative sampler, BIT-EXACT with the reference's host recipe
//   ids = cum.searchsorted(rng.uniform(0, cum[-1], S * n)) -> int32 -> reshape (S, n)
// (learning/NegativeExampleGenerator.py:24,32 over the freq^0.75 cumulative distribution of learning/OieData.py:57-59),
// where rng is the run's legacy numpy.random.RandomState: MT19937, double = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two
// consecutive 32-bit outputs, uniform(low, high) = low + (high - low) * double.
//
//   k_mt19937_words : ONE CTA advances the generator exactly as numpy does (the 624-word state is regenerated in the three
//                     dependency waves [0,227) [227,454) [454,624), each fully parallel) and writes tempered words;
//   k_search_cum    : every thread turns two words into its uniform and binary-searches the cumulative distribution
//                     (side = 'left': first index with cum[i] >= u), float64 compares -> identical integers.
// The generator state (key[624], pos) lives in a caller-owned device buffer and is handed back to the host generator after
// the call, so host and device draws can be interleaved in one stream of random numbers.
#include <algorithm>

#include "rae_common.cuh"
#include "rae_internal.h"

namespace rae {

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MT_A = 0x9908b0dfu, MT_UP = 0x80000000u, MT_LO = 0x7fffffffu;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & MT_UP) | (nxt & MT_LO);
    return far ^ (y >> 1) ^ ((y & 1u) ? MT_A : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// state[0..623] = key, state[624] = pos (numpy: next word is key[pos]; pos == 624 -> regenerate first)
__global__ void __launch_bounds__(256) k_mt19937_words(uint32_t* __restrict__ state, uint32_t* __restrict__ out, long long n_words) {
    __shared__ uint32_t mt[MT_N];
    const int t = threadIdx.x;
    for (int i = t; i < MT_N; i += 256) mt[i] = state[i];
    int pos = (int)state[MT_N];
    __syncthreads();
    long long done = 0;
    while (done < n_words) {
        if (pos >= MT_N) {
            // wave 1: kk in [0, 227) reads old mt[kk], mt[kk+1], mt[kk+397]
            uint32_t v = 0;
            if (t < 227) v = mt_twist(mt[t], mt[t + 1], mt[t + MT_M]);
            __syncthreads();
            if (t < 227) mt[t] = v;
            __syncthreads();
            // wave 2: kk in [227, 454) reads old mt[kk], mt[kk+1] and NEW mt[kk-227]
            if (t < 227) v = mt_twist(mt[227 + t], mt[228 + t], mt[t]);
            __syncthreads();
            if (t < 227) mt[227 + t] = v;
            __syncthreads();
            // wave 3: kk in [454, 624) reads old mt[kk], mt[kk+1] (kk = 623: NEW mt[0]) and NEW mt[kk-227]
            if (t < 170) {
                const int kk = 454 + t;
                v = mt_twist(mt[kk], kk == MT_N - 1 ? mt[0] : mt[kk + 1], mt[kk - 227]);
            }
            __syncthreads();
            if (t < 170) mt[454 + t] = v;
            __syncthreads();
            pos = 0;
        }
        const int take = (int)min((long long)(MT_N - pos), n_words - done);
        for (int i = t; i < take; i += 256) out[done + i] = mt_temper(mt[pos + i]);
        pos += take;
        done += take;
    }
    __syncthreads();
    for (int i = t; i < MT_N; i += 256) state[i] = mt[i];
    if (t == 0) state[MT_N] = (uint32_t)pos;
}

__global__ void __launch_bounds__(256) k_search_cum(const uint32_t* __restrict__ words, const double* __restrict__ cum, int n_cum,
                                                    double low, double range, int32_t* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t a = words[2 * i] >> 5, b = words[2 * i + 1] >> 6;
        const double x = __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
        const double u = __dadd_rn(low, __dmul_rn(range, x));          // no fused multiply-add: numpy rounds twice
        int lo = 0, hi = n_cum;                                         // first index with cum[idx] >= u
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum[mid] < u) lo = mid + 1; else hi = mid;
        }
        out[i] = lo;
    }
}

}  // namespace

}  // namespace rae

using namespace rae;

extern "C" int rae_sample_negatives(rae_engine* h, uint32_t* mt_state, const double* cum, int64_t n_cum, double cum_last,
                                    int32_t* out, int64_t n, uint32_t* scratch_words, void* stream) {
    // h may be NULL (the sampler needs no engine state): errors then land in the global message of rae_last_error(NULL)
    if (!mt_state || !cum || !out || !scratch_words || n < 0 || n_cum < 1 || n_cum >= ((int64_t)1 << 31))
        return fail(h, RAE_EINVAL, "rae_sample_negatives: bad argument");
    if (n == 0) return RAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    k_mt19937_words<<<1, 256, 0, st>>>(mt_state, scratch_words, (long long)(2 * n));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)(h ? h->num_sms : 148) * 16);
    k_search_cum<<<blocks, 256, 0, st>>>(scratch_words, cum, (int)n_cum, 0.0, cum_last - 0.0, out, (long long)n);
    if (h) h->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, RAE_ECUDA, "rae_sample_negatives: %s", cudaGetErrorString(e));
    return RAE_OK;
}
