// Sampling helpers shared by the generate-loop kernels (decode.cu: one launch per token; decode_persist.cu: inside the
// persistent per-event kernel): compaction of the non-zero probabilities, histogram top-k preselection, bitonic sort,
// top-p on the un-renormalised mass + top-k + draw (midi_model.py:152-165).  NT = threads of the calling CTA.
#pragma once
#include "common.cuh"

namespace smp {

constexpr int SMP_MAXV = 4096;

// ---------------------------------------------------------------------------------------------
// sampler
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ bool key_before(float pa, int ia, float pb, int ib) {   // sort order: prob desc, id asc
    return (pa > pb) || (pa == pb && ia < ib);
}

// Shared tail: s_p/s_i hold `n` candidate (prob, id) pairs (prob > 0), unsorted.  Sort, apply top-p on the
// un-renormalised mass and top-k, renormalise, draw with uniform u.  Returns the chosen id (all threads).
template <int NT>
__device__ int sample_tail(float* s_p, int* s_i, int n, float top_p, int top_k, float u, bool bf16_sem) {
    int n_sort = 32;
    while (n_sort < n) n_sort <<= 1;
    for (int i = n + threadIdx.x; i < n_sort; i += blockDim.x) { s_p[i] = -1.f; s_i[i] = 0x7fffffff; }
    __syncthreads();
    for (int k = 2; k <= n_sort; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_sort; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const float pa = s_p[i], pb = s_p[ixj];
                    const int ia = s_i[i], ib = s_i[ixj];
                    const bool a_first = key_before(pa, ia, pb, ib);
                    if (up ? !a_first : a_first) { s_p[i] = pb; s_p[ixj] = pa; s_i[i] = ib; s_i[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    }
    // Only ranks < kk can survive.  Sequential scan by one thread over <= kk entries is cheap for the
    // usual top_k (20); large k falls back to the same loop (still correct).
    const int kk = min(n, top_k);
    __shared__ int s_choice;
    if (threadIdx.x == 0) {
        float cum = 0.f, total = 0.f;
        const float pth = bf16_sem ? bf16_round(top_p) : top_p;
        int last = 0;
        for (int i = 0; i < kk; i++) {
            const float pi = s_p[i];
            cum += pi;
            const float cs = bf16_sem ? bf16_round(cum) : cum;
            const float before = bf16_sem ? bf16_round(cs - pi) : cs - pi;
            const float w = (before > pth) ? 0.f : pi;   // midi_model.py:155-156
            s_p[i] = w;   // weights overwrite the sorted probabilities in place
            total += w;
            if (w > 0.f) last = i;
        }
        int choice = 0;
        if (total > 0.f) {
            const float target = u * total;
            float run = 0.f;
            choice = last;
            for (int i = 0; i <= last; i++) {
                run += s_p[i];
                if (s_p[i] > 0.f && run > target) { choice = i; break; }
            }
        }
        s_choice = (n > 0) ? s_i[choice] : 0;
    }
    __syncthreads();
    return s_choice;
}

// compact the non-zero entries of s_p (indexed by id) into the front of (s_p, s_i); returns count.
template <int NT>
__device__ int compact_nonzero(float* s_p, int* s_i, int V, int* s_cnt) {
    const int per = (V + NT - 1) / NT;
    const int beg = threadIdx.x * per;
    const int end = min(V, beg + per);
    float loc_p[SMP_MAXV / NT];
    int c = 0;
    for (int i = beg; i < end; i++) {
        const float p = s_p[i];
        loc_p[i - beg] = p;
        if (p > 0.f) c++;
    }
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < NT; i++) { const int t = s_cnt[i]; s_cnt[i] = run; run += t; }
        s_cnt[NT] = run;
    }
    __syncthreads();
    int pos = s_cnt[threadIdx.x];
    const int total = s_cnt[NT];
    __syncthreads();   // everyone has read its chunk of s_p into registers before it is overwritten
    for (int i = beg; i < end; i++) {
        const float p = loc_p[i - beg];
        if (p > 0.f) { s_p[pos] = p; s_i[pos] = i; pos++; }
    }
    __syncthreads();
    return total;
}


// Top-k preselection: the sampling tail only looks at ranks < top_k, so instead of sorting all n candidates we keep the
// ones at or above the k-th largest value (found with two shared-memory histograms over the fp32 bit pattern: exponent,
// then the 8 mantissa bits below it -- probabilities are bf16-rounded so 8 mantissa bits separate all distinct values)
// and sort only those.  Returns the new candidate count (n itself when the selection would not shrink the set).
template <int NT>
__device__ int preselect_topk(float* s_p, int* s_i, int n, int top_k, int* s_hist /*[257]*/) {
    if (n <= 64 || top_k >= n || top_k > 64) return n;
    __shared__ int s_sel[2];       // {exponent bin, mantissa bin} of the k-th largest value
    __shared__ int s_count;
    for (int i = threadIdx.x; i < 257; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_hist[(__float_as_uint(s_p[i]) >> 23) & 0xFF], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int cum = 0, e = 255;
        for (; e > 0; e--) { cum += s_hist[e]; if (cum >= top_k) break; }
        s_sel[0] = e;
        s_hist[256] = cum - s_hist[e];      // candidates strictly above the boundary exponent
    }
    __syncthreads();
    const int e_star = s_sel[0];
    const int above = s_hist[256];
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned u = __float_as_uint(s_p[i]);
        if (((u >> 23) & 0xFF) == (unsigned)e_star) atomicAdd(&s_hist[(u >> 15) & 0xFF], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cum = above, m = 255;
        for (; m > 0; m--) { cum += s_hist[m]; if (cum >= top_k) break; }
        s_sel[1] = m;
        s_count = 0;
    }
    __syncthreads();
    const unsigned thr = ((unsigned)e_star << 23) | ((unsigned)s_sel[1] << 15);     // keep p with bit pattern >= thr
    // count first: if the selection is not small, keep everything
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine += (__float_as_uint(s_p[i]) >= thr) ? 1 : 0;
    atomicAdd(&s_count, mine);
    __syncthreads();
    const int m_sel = s_count;
    __syncthreads();
    if (m_sel > 256 || m_sel >= n) return n;
    // compact the selected entries to the front (their order is fixed later by the (prob desc, id asc) sort)
    float keep_p[SMP_MAXV / NT];
    int keep_i[SMP_MAXV / NT];
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (__float_as_uint(s_p[i]) >= thr) { keep_p[c] = s_p[i]; keep_i[c] = s_i[i]; c++; }
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int base = atomicAdd(&s_count, c);
    __syncthreads();       // all reads of s_p / s_i are done (values are in registers) before the overwrite
    for (int j = 0; j < c; j++) { s_p[base + j] = keep_p[j]; s_i[base + j] = keep_i[j]; }
    __syncthreads();
    return m_sel;
}


}   // namespace smp
