// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA into 128B-swizzled shared memory), persistent, warp-specialised:
//   warp 0     : TMA producer (one elected lane)
//   warp 1     : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5 : epilogue (TMEM -> registers -> fused epilogue -> global)
// D[M,N] = A . B^T with fp32 accumulation.  Each operand is either "K-major"
// (row-major [rows, K], what nn.Linear's forward needs: hf modeling_llama.py:177-184,
// 238-264) or "MN-major" (stored [K, rows]); the latter serves dgrad (B = W as stored)
// and wgrad (A = dY, B = X as stored) without any transposed copies.
// Epilogues: bf16 store; bf16 store + residual (two roundings, like `x + o_proj(..)`
// in hf :325/:331); fp32 split-K partials reduced by splitk_reduce_kernel.
#include <math.h>

#include "tc05.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARP0 = 2;
// threads = TMA warp + MMA warp + EG groups of four epilogue warps (a group covers the 128 TMEM lanes once)
__host__ __device__ constexpr int num_threads(int eg) { return 64 + 128 * eg; }

enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_PARTIAL_F32 = 2, EPI_ROPE = 3, EPI_SWIGLU = 4 };

struct GemmParams {
    bf16* C;
    const bf16* R;
    float* ws;
    int M, N, K;
    int ldc, ldr;
    int epilogue;
    int tma_store;             // EPI_STORE only: bf16 tile staged in shared memory and written with cp.async.bulk.tensor
    int splits, kb_per_split, num_kb;
    // tail split: the tiles of the last, partial wave (index >= tail_tile0) are cut into tail_splits K-slices whose fp32
    // partials go to tail_ws[tile - tail_tile0][slice][128][BLOCK_N] and are summed by tail_reduce_kernel
    int tail_tile0, tail_splits, tail_kb, total_items;
    float* tail_ws;
    int m_tiles, n_tiles;
    uint32_t mn_lbo, mn_sbo;   // MN-major descriptor byte offsets (overridable for bring-up)
    // EPI_ROPE: rotary embedding applied to columns [0, rope_cols) (the q and k thirds of a packed QKV row)
    const bf16* rope_cos;
    const bf16* rope_sin;
    int rope_S, rope_D, rope_cols;
    // EPI_SWIGLU: B = [gate | up] weight rows; tile n covers gate rows [128n, 128n+128) and up rows [I + 128n, ...)
    bf16* act;
    int swiglu_I, ld_act;
};

using namespace tc05;

// CG = CTAs per UMMA (1, or 2 = CTA pair: cta_group::2, M = 256 per MMA, each CTA stages its 128 rows of A and HALF of B)
template <int BLOCK_N, int CG = 1>
struct SmemLayout {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
    static constexpr int B_BYTES = (BLOCK_N / CG) * BLOCK_K * 2;   // 16/32 KB (half of it per CTA of a pair)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (CG == 2) ? 5 : ((BLOCK_N == 256) ? 4 : 6);
    static constexpr int OUT_BYTES = 2 * BLOCK_M * 64 * 2;  // two [128 rows x 128 B] staging boxes for the TMA-store epilogue
    static constexpr int BAR_BYTES = 1024;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024;   // +1024 for manual alignment
};

struct WorkItem {
    int m_blk, n_blk, kb0, kb1, split;
    int tail;          // -1: regular item; otherwise index of the tile inside the tail region
};
__device__ __forceinline__ WorkItem decode_item(const GemmParams& p, int item) {
    WorkItem w;
    const int reg_items = p.tail_tile0 * p.splits;
    int t;
    if (item < reg_items) {
        w.split = item % p.splits;
        t = item / p.splits;
        w.kb0 = w.split * p.kb_per_split;
        w.kb1 = min(w.kb0 + p.kb_per_split, p.num_kb);
        w.tail = -1;
    } else {
        const int idx = item - reg_items;
        w.tail = idx / p.tail_splits;
        w.split = idx % p.tail_splits;
        t = p.tail_tile0 + w.tail;
        w.kb0 = w.split * p.tail_kb;
        w.kb1 = min(w.kb0 + p.tail_kb, p.num_kb);
    }
    w.n_blk = t % p.n_tiles;
    w.m_blk = t / p.n_tiles;
    return w;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EG, int CG>
__global__ void __launch_bounds__(num_threads(EG), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
    using L = SmemLayout<BLOCK_N, CG>;
    B200_PDL_TRIGGER();
    // CTA pair: this CTA's rank in its 2-CTA cluster (0 = leader: issues the MMAs, owns the full / tmem-empty barriers)
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader_cta = rank == 0;
    const int first_item = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_stride = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int STAGES = L::STAGES;
    constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;   // double-buffered accumulator (256 or 512 columns)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* out_stage = smem + STAGES * L::STAGE_BYTES;         // 1024-aligned: STAGE_BYTES is a multiple of 16 KB
    uint8_t* bar_base = out_stage + L::OUT_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4 * EG * CG);      // pair: the epilogue warps of BOTH CTAs release the leader's accumulator
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS);
        else tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();             // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, TMEM, tensor-map prefetch) may have run under the previous kernel's tail (PDL); from
    // here on this kernel reads and writes global memory
    B200_PDL_WAIT();

    const int total_items = p.total_items;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = first_item; item < total_items; item += item_stride) {
                const WorkItem wi = decode_item(p, item);
                const int n_blk = wi.n_blk, m_blk = wi.m_blk * CG + (int)rank, kb0 = wi.kb0, kb1 = wi.kb1;
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * L::STAGE_BYTES;
                    uint8_t* sB = sA + L::A_BYTES;
                    if (CG == 1) {
                        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        if (!A_MN) {
                            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                        } else {
#pragma unroll
                            for (int c = 0; c < BLOCK_M / 64; c++)
                                tma_load_2d(sA + c * (BLOCK_K * 128), &tmA, &full_bar[stage], m_blk * BLOCK_M + c * 64, kb * BLOCK_K);
                        }
                        if (!B_MN) {
                            if (BLOCK_N == 256 && p.epilogue == EPI_SWIGLU) {   // gate half | up half (tensor map box = 128 rows)
                                tma_load_2d(sB, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * 128);
                                tma_load_2d(sB + 128 * BLOCK_K * 2, &tmB, &full_bar[stage], kb * BLOCK_K, p.swiglu_I + n_blk * 128);
                            } else {
                                tma_load_2d(sB, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < BLOCK_N / 64; c++)
                                tma_load_2d(sB + c * (BLOCK_K * 128), &tmB, &full_bar[stage], n_blk * BLOCK_N + c * 64, kb * BLOCK_K);
                        }
                    } else {
                        // pair: both CTAs' loads are counted on the LEADER's full barrier (it alone waits on it); the leader
                        // announces the bytes of both.  This CTA stages its own 128 rows of A and its half of B's N range.
                        const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
                        if (leader_cta) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                        if (!A_MN) {
                            tma_load_2d_pair(sA, &tmA, fb, kb * BLOCK_K, m_blk * BLOCK_M);
                        } else {
#pragma unroll
                            for (int c = 0; c < BLOCK_M / 64; c++)
                                tma_load_2d_pair(sA + c * (BLOCK_K * 128), &tmA, fb, m_blk * BLOCK_M + c * 64, kb * BLOCK_K);
                        }
                        constexpr int HALF_N = BLOCK_N / 2;
                        if (!B_MN) {
                            if (p.epilogue == EPI_SWIGLU)      // leader: gate rows, peer: the matching up rows
                                tma_load_2d_pair(sB, &tmB, fb, kb * BLOCK_K, (int)rank * p.swiglu_I + n_blk * 128);
                            else
                                tma_load_2d_pair(sB, &tmB, fb, kb * BLOCK_K, n_blk * BLOCK_N + (int)rank * HALF_N);
                        } else {
#pragma unroll
                            for (int c = 0; c < HALF_N / 64; c++)
                                tma_load_2d_pair(sB + c * (BLOCK_K * 128), &tmB, fb, n_blk * BLOCK_N + (int)rank * HALF_N + c * 64, kb * BLOCK_K);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && leader_cta) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M * CG, BLOCK_N, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = first_item; item < total_items; item += item_stride) {
                const WorkItem wi = decode_item(p, item);
                const int kb0 = wi.kb0, kb1 = wi.kb1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t sB = sA + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
                        const uint64_t adesc = A_MN ? make_smem_desc(sA + k * (UMMA_K * 128), p.mn_lbo, p.mn_sbo)
                                                    : make_smem_desc(sA + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t bdesc = B_MN ? make_smem_desc(sB + k * (UMMA_K * 128), p.mn_lbo, p.mn_sbo)
                                                    : make_smem_desc(sB + k * (UMMA_K * 2), 16, 1024);
                        if (CG == 2) umma_f16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else umma_f16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    if (CG == 2) umma_commit_pair(&empty_bar[stage]);   // frees the slot in BOTH CTAs
                    else umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) umma_commit_pair(&tfull_bar[acc]);   // accumulator halves ready for both CTAs' epilogues
                else umma_commit(&tfull_bar[acc]);          // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        // EG == 2: two groups of four epilogue warps split the tile's 64-column boxes between them (group g takes boxes
        // g, g + 2, ...; for SwiGLU feature half g), each group with its own staging box, named barrier and store leader:
        // twice the threads for the conversions / SFU work of the fused epilogues, and the two groups' TMA stores overlap.
        const int grp = (warp - EPI_WARP0) >> 2;
        int acc = 0, out_buf = (EG == 2) ? grp : 0;
        uint32_t acc_phase = 0;
        for (int item = first_item; item < total_items; item += item_stride) {
            const WorkItem wi = decode_item(p, item);
            const int split = wi.split, n_blk = wi.n_blk, m_blk = wi.m_blk * CG + (int)rank;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * BLOCK_M + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
            // ---- staged output: the thread's 64 consecutive columns (8 x 16 B) of row r_t go into a 128B-swizzled
            // [128 x 64] staging box, one TMA store per box (see the plain-store branch below for the why)
            const int r_t = quarter * 32 + lane;
            const bool leader = (warp == EPI_WARP0 + 4 * grp && lane == 0);
            const int bar_id = 1 + grp;
            auto box_row = [&]() -> uint8_t* {
                if (EG == 2) {
                    // one box per group: the previous store of this group must have finished reading it
                    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
                return out_stage + out_buf * (BLOCK_M * 128) + r_t * 128;
            };
            auto box_put = [&](uint8_t* dst, int chunk, const float* f8) {
                *reinterpret_cast<uint4*>(dst + ((chunk ^ (r_t & 7)) << 4)) = pack8(f8);
            };
            auto box_send = [&](const CUtensorMap* tm, int gcol) {
                fence_proxy_async_smem();
                // EG == 1: two boxes ping-pong; the previous box's store has finished reading the other buffer before anyone
                // writes it again
                if (EG == 1 && leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (leader) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(tm), "r"(smem_u32(out_stage + out_buf * (BLOCK_M * 128))), "r"(gcol), "r"(m_blk * BLOCK_M)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (EG == 1) out_buf ^= 1;
            };
            if (wi.tail >= 0) {
                // K-slice of a tail tile: fp32 partial, tile-local layout [128][BLOCK_N]
                float* dstp = p.tail_ws + ((size_t)(wi.tail * p.tail_splits + split) * BLOCK_M + r_t) * BLOCK_N;
#pragma unroll 1
                for (int c = grp; c < BLOCK_N / 32; c += EG) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        *reinterpret_cast<uint4*>(dstp + c * 32 + v * 4) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                }
            } else if (BLOCK_N == 256 && p.epilogue == EPI_SWIGLU && p.tma_store) {
                // gate|up projection with SwiGLU fused (hf modeling_llama.py:183), staged: per 64 features three boxes --
                // gate, up (both kept for the backward) and act = bf16(silu(gate)) * up.  TMEM is re-read for the act
                // pass instead of holding 128 values per thread in registers.
#pragma unroll 1
                for (int fb = grp; fb < 2; fb += EG) {
                    const int f0 = n_blk * 128 + fb * 64;
                    if (f0 >= p.swiglu_I) break;
#pragma unroll 1
                    for (int which = 0; which < 2; which++) {          // 0: gate columns, 1: up columns
                        uint8_t* dst = box_row();
#pragma unroll
                        for (int hc = 0; hc < 2; hc++) {
                            uint32_t r[32];
                            tmem_ld32(taddr + which * 128 + fb * 64 + hc * 32, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int v = 0; v < 4; v++) {
                                float f[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) f[q] = __uint_as_float(r[8 * v + q]);
                                box_put(dst, hc * 4 + v, f);
                            }
                        }
                        box_send(&tmC, which * p.swiglu_I + f0);
                    }
                    uint8_t* dst = box_row();
#pragma unroll
                    for (int hc = 0; hc < 2; hc++) {
                        uint32_t r1[32], r2[32];
                        tmem_ld32(taddr + fb * 64 + hc * 32, r1);
                        tmem_ld32(taddr + 128 + fb * 64 + hc * 32, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            float a[8];
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                const float g = bf16_round(__uint_as_float(r1[8 * v + q]));
                                const float u = bf16_round(__uint_as_float(r2[8 * v + q]));
                                a[q] = bf16_round(silu_f(g)) * u;
                            }
                            box_put(dst, hc * 4 + v, a);
                        }
                    }
                    box_send(&tmD, f0);
                }
            } else if (p.epilogue == EPI_ROPE && p.tma_store) {
                // QKV projection with RoPE fused (hf modeling_llama.py:262-268), staged, head_dim 64: one box = one head, the
                // (d, d+32) pairs are the two 32-column halves of the thread's 64 values.
                const int pos = row_ok ? row % p.rope_S : 0;
                const bf16* cp = p.rope_cos + (size_t)pos * 32;
                const bf16* sp = p.rope_sin + (size_t)pos * 32;
#pragma unroll 1
                for (int bx = grp; bx < BLOCK_N / 64; bx += EG) {
                    const int col0 = n_blk * BLOCK_N + bx * 64;
                    if (col0 >= p.N) break;
                    uint8_t* dst = box_row();
                    uint32_t r1[32], r2[32];
                    tmem_ld32(taddr + bx * 64, r1);
                    tmem_ld32(taddr + bx * 64 + 32, r2);
                    tmem_ld_wait();
                    const bool rot = col0 < p.rope_cols;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        float x1[8], x2[8], o1[8], o2[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            x1[q] = bf16_round(__uint_as_float(r1[8 * v + q]));
                            x2[q] = bf16_round(__uint_as_float(r2[8 * v + q]));
                        }
                        if (rot) {
                            float cc[8], ss[8];
                            unpack8(*reinterpret_cast<const uint4*>(cp + v * 8), cc);
                            unpack8(*reinterpret_cast<const uint4*>(sp + v * 8), ss);
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                o1[q] = bf16_round(x1[q] * cc[q]) + bf16_round(-x2[q] * ss[q]);
                                o2[q] = bf16_round(x2[q] * cc[q]) + bf16_round(x1[q] * ss[q]);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; q++) { o1[q] = x1[q]; o2[q] = x2[q]; }
                        }
                        box_put(dst, v, o1);
                        box_put(dst, 4 + v, o2);
                    }
                    box_send(&tmC, col0);
                }
            } else if (BLOCK_N == 256 && p.epilogue == EPI_SWIGLU) {
                // gate|up projection with SwiGLU fused (hf modeling_llama.py:183): accumulator columns [0,128) are gate
                // features, [128,256) the matching up features.  g, u = bf16(acc) are stored (backward needs them) and
                // act = bf16(bf16(silu(g)) * u) -- same rounding points as the stand-alone kernel.
#pragma unroll 1
                for (int c = grp; c < 4; c += EG) {
                    uint32_t r1[32], r2[32];
                    tmem_ld32(taddr + c * 32, r1);
                    tmem_ld32(taddr + 128 + c * 32, r2);
                    tmem_ld_wait();
                    const int f0 = n_blk * 128 + c * 32;
                    if (row_ok && f0 < p.swiglu_I) {
                        bf16* dg = p.C + (size_t)row * p.ldc + f0;
                        bf16* du = dg + p.swiglu_I;
                        bf16* da = p.act + (size_t)row * p.ld_act + f0;
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            float g[8], u[8], a[8];
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                g[q] = bf16_round(__uint_as_float(r1[8 * v + q]));
                                u[q] = bf16_round(__uint_as_float(r2[8 * v + q]));
                                a[q] = bf16_round(silu_f(g[q])) * u[q];
                            }
                            *reinterpret_cast<uint4*>(dg + v * 8) = pack8(g);
                            *reinterpret_cast<uint4*>(du + v * 8) = pack8(u);
                            *reinterpret_cast<uint4*>(da + v * 8) = pack8(a);
                        }
                    }
                }
            } else if (p.epilogue == EPI_ROPE) {
                // QKV projection with RoPE fused (hf modeling_llama.py:262-268): x = bf16(acc) (the Linear's rounding),
                // then o1 = bf16(bf16(x1*c) + bf16(-x2*s)), o2 = bf16(bf16(x2*c) + bf16(x1*s)) on (d, d + D/2) pairs.
                const int half = p.rope_D >> 1;
                const int cph = half >> 5;                      // 32-column chunks per half head (1 for d=64, 4 for d=256)
                const int pos = row_ok ? row % p.rope_S : 0;
#pragma unroll 1
                for (int c0 = grp * 2 * cph; c0 < BLOCK_N / 32; c0 += EG * 2 * cph) {
#pragma unroll 1
                    for (int j = 0; j < cph; j++) {
                        uint32_t r1[32], r2[32];
                        tmem_ld32(taddr + (c0 + j) * 32, r1);
                        tmem_ld32(taddr + (c0 + j + cph) * 32, r2);
                        tmem_ld_wait();
                        const int col1 = n_blk * BLOCK_N + (c0 + j) * 32;
                        const int col2 = col1 + half;
                        if (row_ok && col1 < p.N) {
                            bf16* d1 = p.C + (size_t)row * p.ldc + col1;
                            bf16* d2 = p.C + (size_t)row * p.ldc + col2;
                            const bool rot = col1 < p.rope_cols;
                            const bf16* cp = p.rope_cos + (size_t)pos * half + j * 32;
                            const bf16* sp = p.rope_sin + (size_t)pos * half + j * 32;
#pragma unroll
                            for (int v = 0; v < 4; v++) {
                                float x1[8], x2[8], o1[8], o2[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) {
                                    x1[q] = bf16_round(__uint_as_float(r1[8 * v + q]));
                                    x2[q] = bf16_round(__uint_as_float(r2[8 * v + q]));
                                }
                                if (rot) {
                                    float cc[8], ss[8];
                                    unpack8(*reinterpret_cast<const uint4*>(cp + v * 8), cc);
                                    unpack8(*reinterpret_cast<const uint4*>(sp + v * 8), ss);
#pragma unroll
                                    for (int q = 0; q < 8; q++) {
                                        o1[q] = bf16_round(x1[q] * cc[q]) + bf16_round(-x2[q] * ss[q]);
                                        o2[q] = bf16_round(x2[q] * cc[q]) + bf16_round(x1[q] * ss[q]);
                                    }
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; q++) { o1[q] = x1[q]; o2[q] = x2[q]; }
                                }
                                *reinterpret_cast<uint4*>(d1 + v * 8) = pack8(o1);
                                if (col2 < p.N) *reinterpret_cast<uint4*>(d2 + v * 8) = pack8(o2);
                            }
                        }
                    }
                }
            } else if (p.tma_store) {
                // plain bf16 store: row-per-thread 16-byte global stores touch 32 different rows per instruction and
                // made the epilogue longer than a K=1024 main loop (tensor pipe 64% busy); instead the tile goes
                // through two 128B-swizzled [128 x 64] staging boxes and leaves with one TMA store per box.
#pragma unroll 1
                for (int bx = grp; bx < BLOCK_N / 64; bx += EG) {
                    const int col0 = n_blk * BLOCK_N + bx * 64;
                    if (col0 >= p.N) break;                                   // uniform within the group
                    uint8_t* dst = box_row();
#pragma unroll
                    for (int hc = 0; hc < 2; hc++) {
                        uint32_t r[32];
                        tmem_ld32(taddr + bx * 64 + hc * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) f[j] = __uint_as_float(r[8 * v + j]);
                            box_put(dst, hc * 4 + v, f);
                        }
                    }
                    box_send(&tmC, col0);
                }
            } else
#pragma unroll 1
            for (int c = grp; c < BLOCK_N / 32; c += EG) {
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                tmem_ld_wait();
                const int col0 = n_blk * BLOCK_N + c * 32;
                if (row_ok && col0 < p.N) {
                    if (p.epilogue == EPI_PARTIAL_F32) {
                        float* dst = p.ws + ((size_t)split * p.M + row) * p.N + col0;
#pragma unroll
                        for (int v = 0; v < 8; v++) {
                            if (col0 + v * 4 < p.N) {
                                uint4 u = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                                *reinterpret_cast<uint4*>(dst + v * 4) = u;
                            }
                        }
                    } else {
                        bf16* dst = p.C + (size_t)row * p.ldc + col0;
                        const bf16* res = (p.epilogue == EPI_RESIDUAL) ? p.R + (size_t)row * p.ldr + col0 : nullptr;
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            if (col0 + v * 8 < p.N) {
                                float f[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) f[j] = __uint_as_float(r[8 * v + j]);
                                if (res != nullptr) {
                                    float rr[8];
                                    unpack8(*reinterpret_cast<const uint4*>(res + v * 8), rr);
#pragma unroll
                                    for (int j = 0; j < 8; j++) f[j] = bf16_round(f[j]) + rr[j];
                                }
                                *reinterpret_cast<uint4*>(dst + v * 8) = pack8(f);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 1 || leader_cta) mbar_arrive(&tempty_bar[acc]);
                else mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[acc]), 0));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_store && ((warp - EPI_WARP0) & 3) == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();             // nobody leaves while the peer's MMAs / commits may still touch this CTA
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, bf16* __restrict__ out, size_t n, int splits,
                                     int accumulate) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        float4 a = *reinterpret_cast<const float4*>(ws + i);
        for (int s = 1; s < splits; s++) {
            float4 b = *reinterpret_cast<const float4*>(ws + (size_t)s * n + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        bf162* o = reinterpret_cast<bf162*>(out + i);
        if (accumulate) {   // grad accumulation across micro-batches: bf16 += bf16(new), like autograd's AccumulateGrad
            float2 o0 = __bfloat1622float2(o[0]), o1 = __bfloat1622float2(o[1]);
            a.x = bf16_round(a.x) + o0.x; a.y = bf16_round(a.y) + o0.y;
            a.z = bf16_round(a.z) + o1.x; a.w = bf16_round(a.w) + o1.y;
        }
        o[0] = __floats2bfloat162_rn(a.x, a.y);
        o[1] = __floats2bfloat162_rn(a.z, a.w);
    }
}

// sums the K-slices of the tail tiles and writes bf16 into C (rows < M, columns < N, N a multiple of 8)
__global__ void tail_reduce_kernel(const float* __restrict__ ws, bf16* __restrict__ C, int M, int N, int ldc, int n_tiles,
                                   int tail_tile0, int n_tail, int splits, int block_n) {
    const int per_tile = BLOCK_M * block_n / 8;                 // 8-column groups per tile
    const long long total = (long long)n_tail * per_tile;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int tt = (int)(i / per_tile), e = (int)(i % per_tile);
        const int r = e / (block_n / 8), c8 = e % (block_n / 8);
        const int t = tail_tile0 + tt;
        const int row = (t / n_tiles) * BLOCK_M + r, col = (t % n_tiles) * block_n + c8 * 8;
        if (row >= M || col >= N) continue;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < splits; s++) {
            const float* src = ws + ((size_t)(tt * splits + s) * BLOCK_M + r) * block_n + c8 * 8;
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        *reinterpret_cast<uint4*>(C + (size_t)row * ldc + col) = pack8(acc);
    }
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EG, int CG>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmD, const GemmParams& p,
           cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N, CG>;
    auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, EG, CG>;
    static bool configured = false;
    if (!configured) {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL), "gemm smem attr");
        configured = true;
    }
    const int items = p.total_items;
    static int pdl = -1;
    if (pdl < 0) {
        // B200_PDL=1: programmatic dependent launch of the GEMMs.  Measured neutral on the train step (51.1/51.8 ms off,
        // 51.3/51.8 ms on, same box, profiles/r2_pdl_ab.txt): the GEMM prologue is ~2 us of a >=100 us kernel and the
        // producer kernels' tails are short, so it stays off by default.
        const char* e = getenv("B200_PDL");
        pdl = (e && e[0] == '1') ? 1 : 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(num_threads(EG));
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (CG == 1) {
        cfg.gridDim = dim3(items < b200_num_sms() ? items : b200_num_sms());
    } else {
        // CTA pairs: clusters of two CTAs (same TPC), one 256-row tile pair per cluster at a time
        const int pairs = b200_num_sms() / 2;
        cfg.gridDim = dim3(2 * (items < pairs ? items : pairs));
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = 2;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        n_attr++;
    }
    if (pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        n_attr++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmD, p), "gemm launch");
    B200_CHECK_LAUNCH("gemm_tcgen05");
    return B200_OK;
}

}   // namespace

// ---------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no -lcuda link dependency)
// ---------------------------------------------------------------------------
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of pitch `ld` elements, 128B swizzle, zero OOB fill.
}   // namespace

int tc05::make_tmap_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                       uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        b200_set_error("gemm: cuTensorMapEncodeTiled entry point unavailable");
        return B200_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        b200_set_error("gemm: cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu ld=%llu", (int)r, ptr,
                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

// 3-D variant {cols, rows-per-sequence, batch}: rows outside a sequence are zero-filled (attention tiles)
static int make_tmap_3d_any(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t cols, uint64_t rows,
                            uint64_t batch, uint64_t row_pitch, uint64_t batch_pitch, uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        b200_set_error("tmap3d: cuTensorMapEncodeTiled entry point unavailable");
        return B200_ERR_CUDA;
    }
    cuuint64_t gdim[3] = {cols, rows, batch};
    cuuint64_t gstride[2] = {row_pitch * esize, batch_pitch * esize};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, dt, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        b200_set_error("tmap3d: cuTensorMapEncodeTiled failed (%d) ptr=%p cols=%llu rows=%llu batch=%llu pitch=%llu/%llu",
                       (int)r, ptr, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch,
                       (unsigned long long)row_pitch, (unsigned long long)batch_pitch);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

// bf16 [batch][rows][cols] with element pitches; box = box_cols x box_rows x 1, 128B swizzle (box_cols * 2 == 128)
int tc05_make_tmap_3d(CUtensorMap* tm, const void* ptr, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t row_pitch,
                      uint64_t batch_pitch, uint32_t box_cols, uint32_t box_rows) {
    return make_tmap_3d_any(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, cols, rows, batch, row_pitch, batch_pitch, box_cols, box_rows);
}

// fp32 variant (box_cols * 4 == 128): target of the attention backward's dQ reduce-add
int tc05_make_tmap_3d_f32(CUtensorMap* tm, const void* ptr, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t row_pitch,
                          uint64_t batch_pitch, uint32_t box_cols, uint32_t box_rows) {
    return make_tmap_3d_any(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, cols, rows, batch, row_pitch, batch_pitch, box_cols, box_rows);
}

// ---------------------------------------------------------------------------
// C ABI (declared in include/midi_b200.h)
// ---------------------------------------------------------------------------
extern "C" size_t b200_gemm_workspace_bytes(int M, int N, int splits) {
    return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}

// Tail split.  tiles = full * sms + r: the last wave keeps only r of the sms CTAs busy for a whole tile time.  Cutting
// those r tiles into s K-slices makes the tail ceil(r*s/sms)/s tile times long (e.g. 512 tiles on 148 SMs: r = 68,
// s = 2 -> 3.5 instead of 4 waves).  Returns s (1 = leave the tail alone).
static int plan_tail(int tiles, int num_kb, int sms) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("B200_GEMM_TAIL_SPLIT");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    const int full = tiles / sms, r = tiles % sms;
    // measured (profiles/r1_gemm_tail_ab.txt): +3 % at K = 8192, neutral at K = 4096, a loss at K <= 3072 (the partial round
    // trip and the extra launch cost more than the half wave saves) and for 4-way splits -> long-K, 2-way only
    if (!enabled || r == 0 || full == 0 || full > 8 || num_kb < 96) return 1;
    const double cost = (double)((r * 2 + sms - 1) / sms) / 2 + 0.06;
    return cost < 0.75 ? 2 : 1;
}

// bytes of fp32 workspace b200_gemm_bf16 can use to split the tiles of the last partial wave along K (0: no tail split
// for this shape).  Optional: without the workspace the GEMM runs unsplit.
extern "C" size_t b200_gemm_tail_workspace_bytes(int M, int N, int K, int block_n) {
    if (block_n != 128 && block_n != 256) return 0;
    const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + block_n - 1) / block_n);
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int sms = b200_num_sms();
    const int s = plan_tail(tiles, num_kb, sms);
    return s > 1 ? (size_t)(tiles % sms) * s * BLOCK_M * block_n * sizeof(float) : 0;
}

// Tile / split-K planner.  Work items = tiles(block_n) x splits run persistently on `sms` CTAs, so the cost is
// (waves of items) x (k-blocks per item) in MMA clocks, plus pipeline fill, the exposed last epilogue and, for
// splits > 1, the fp32 partial round trip.  Picks the cheapest (block_n, splits): this is what removes the
// wave-quantisation loss of the small-output wgrad problems (e.g. 192 tiles on 148 SMs -> 3 splits, 3.9 waves).
static bool pair_enabled() {
    static int ok = -1;
    if (ok < 0) {
        const char* e = getenv("B200_GEMM_CTA_PAIR");
        ok = (e && e[0] == '0') ? 0 : 1;
    }
    return ok != 0;
}
static double plan_cost(int M, int N, int K, int block_n, int s, int sms) {
    // CTA pairs (256-wide tiles, more than one row tile): the schedulable unit is a 256 x 256 tile pair on sms / 2 clusters
    const bool pair = pair_enabled() && block_n == 256 && M > BLOCK_M;
    const int rows = pair ? 2 * BLOCK_M : BLOCK_M;
    const int workers = pair ? sms / 2 : sms;
    const double tiles = (double)((M + rows - 1) / rows) * ((N + block_n - 1) / block_n);
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per = (num_kb + s - 1) / s;
    const int s_eff = (num_kb + kb_per - 1) / kb_per;
    const double items = tiles * s_eff;
    const double waves = ceil(items / workers);
    // 4 UMMAs of 128 x block_n x 16 per k-block.  Measured on B200 (profiles/r1_*): 128-wide tiles top out near
    // 800 TFLOP/s (the 128x128x16 SS-mode UMMA re-reads 8 KB of smem operands per 64 clk = the 128 B/clk smem limit)
    // while 256-wide tiles reach 1.3-1.48 PFLOP/s, hence the 1.7x cost factor.
    const double kb_clk = 2.0 * block_n * (block_n == 128 ? 1.7 : 1.0);
    const double tile_ovh = 150.0 + (s_eff > 1 ? 1.0 : 0.5) * block_n * 6.0;   // accumulator hand-over + epilogue pressure
    double c = waves * (kb_per * kb_clk + tile_ovh) + 2500.0;   // + fill and exposed tail epilogue
    if (s_eff > 1) c += ((double)s_eff * M * N * 8.0 + (double)M * N * 2.0) / 5000.0 + 4000.0;   // bytes / (B/clk, mostly L2) + launch
    return c;
}

extern "C" int b200_gemm_plan(int M, int N, int K, int allow_split, int* block_n_out, int* splits_out) {
    const int sms = b200_num_sms();
    double best = 1e300;
    int bb = 128, bs = 1;
    for (int bn = 128; bn <= 256; bn += 128) {
        if (bn == 256 && N < 256) continue;
        const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
        const int smax = allow_split ? (num_kb / 8 > 16 ? 16 : (num_kb / 8 > 0 ? num_kb / 8 : 1)) : 1;
        for (int s = 1; s <= smax; s++) {
            const double c = plan_cost(M, N, K, bn, s, sms);
            if (c < best) { best = c; bb = bn; bs = s; }
        }
    }
    *block_n_out = bb;
    *splits_out = bs;
    return B200_OK;
}

// legacy helper (kept for ABI stability): split factor for a fixed block_n
extern "C" int b200_gemm_suggest_splits(int M, int N, int K, int block_n) {
    const int sms = b200_num_sms();
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int smax = num_kb / 8 > 16 ? 16 : (num_kb / 8 > 0 ? num_kb / 8 : 1);
    double best = 1e300;
    int bs = 1;
    for (int s = 1; s <= smax; s++) {
        const double c = plan_cost(M, N, K, block_n, s, sms);
        if (c < best) { best = c; bs = s; }
    }
    return bs;
}

static int gemm_impl(const void* A, const void* B, void* C, const void* R, int M, int N, int K, int lda, int ldb, int ldc,
                     int ldr, int a_mn_major, int b_mn_major, int accumulate, int block_n, int splits, void* workspace,
                     size_t workspace_bytes, const bf16* rope_cos, const bf16* rope_sin, int rope_S, int rope_D,
                     int rope_cols, cudaStream_t stream);

extern "C" int b200_gemm_bf16(const void* A, const void* B, void* C, const void* R, int M, int N, int K, int lda, int ldb,
                              int ldc, int ldr, int a_mn_major, int b_mn_major, int accumulate, int block_n, int splits,
                              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return gemm_impl(A, B, C, R, M, N, K, lda, ldb, ldc, ldr, a_mn_major, b_mn_major, accumulate, block_n, splits, workspace,
                     workspace_bytes, nullptr, nullptr, 0, 0, 0, stream);
}

// Fused QKV projection + RoPE: C[M,N] = rope(A . B^T) with both operands K-major; rows are positions r % S of their
// sequence, columns [0, rope_cols) are heads of width head_dim that get rotated, the rest (v) is stored as is.
extern "C" int b200_gemm_bf16_rope(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                                   const void* rope_cos, const void* rope_sin, int S, int head_dim, int rope_cols,
                                   cudaStream_t stream) {
    B200_CHECK_ARG(head_dim == 64 || head_dim == 128 || head_dim == 256, "gemm_rope: head_dim %d unsupported", head_dim);
    B200_CHECK_ARG(N % 256 == 0 && rope_cols % 256 == 0, "gemm_rope: N and rope_cols must be multiples of 256");
    B200_CHECK_ARG(S > 0 && rope_cos && rope_sin, "gemm_rope: missing tables");
    return gemm_impl(A, B, C, nullptr, M, N, K, lda, ldb, ldc, 0, 0, 0, 0, 256, 1, nullptr, 0, (const bf16*)rope_cos,
                     (const bf16*)rope_sin, S, head_dim, rope_cols, stream);
}

// Fused gate|up projection + SwiGLU: gu[M, 2I] = A . Wgu^T (stored, the backward pass needs g and u) and
// act[M, I] = bf16(bf16(silu(g)) * u) written by the same epilogue.  Wgu = [gate rows | up rows], K-major operands.
extern "C" int b200_gemm_bf16_swiglu(const void* A, const void* Wgu, void* gu, void* act, int M, int I, int K, int lda,
                                     int ldw, int ld_gu, int ld_act, cudaStream_t stream) {
    B200_CHECK_ARG(I % 128 == 0, "gemm_swiglu: intermediate size %d must be a multiple of 128", I);
    B200_CHECK_ARG(ld_act % 8 == 0 && (uintptr_t)act % 16 == 0, "gemm_swiglu: act must be 16-byte aligned");
    return gemm_impl(A, Wgu, gu, nullptr, M, 2 * I, K, lda, ldw, ld_gu, 0, 0, 0, 0, 256, 1, act, (size_t)I | ((size_t)ld_act << 32),
                     nullptr, nullptr, -1, 0, 0, stream);
}

static int gemm_impl(const void* A, const void* B, void* C, const void* R, int M, int N, int K, int lda, int ldb, int ldc,
                     int ldr, int a_mn_major, int b_mn_major, int accumulate, int block_n, int splits, void* workspace,
                     size_t workspace_bytes, const bf16* rope_cos, const bf16* rope_sin, int rope_S, int rope_D,
                     int rope_cols, cudaStream_t stream) {
    const bool swiglu = (rope_S == -1);      // internal marker set by b200_gemm_bf16_swiglu (workspace = act, bytes = I | ld<<32)
    B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    // N need not be a multiple of 8: B rows >= N are out of bounds for the tensor map (zero-filled), so the
    // epilogue may store whole 16-byte vectors up to roundup8(N) (zeros) as long as the row pitch covers them.
    const int N8 = (N + 7) / 8 * 8;
    B200_CHECK_ARG(ldc % 8 == 0 && ldc >= N8, "gemm: ldc (%d) must be a multiple of 8 and >= roundup8(N=%d)", ldc, N);
    B200_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 (16-byte TMA strides)");
    B200_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0),
                   "gemm: operands must be 16-byte aligned");
    B200_CHECK_ARG(block_n == 128 || block_n == 256, "gemm: block_n must be 128 or 256");
    B200_CHECK_ARG(R == nullptr || (ldr % 8 == 0 && (uintptr_t)R % 16 == 0 && N % 8 == 0),
                   "gemm: residual must be 16-byte aligned and N a multiple of 8");
    if (splits < 1) splits = 1;
    GemmParams p;
    p.C = (bf16*)C;
    p.R = (const bf16*)R;
    p.ws = (float*)workspace;
    p.M = M; p.N = N8; p.K = K;
    p.ldc = ldc; p.ldr = ldr;
    p.num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    if (splits > p.num_kb) splits = p.num_kb;
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;   // no empty splits
    p.splits = splits;
    // CTA pairs (cta_group::2, 256-row tile pairs) for the 256-wide tiles: halves the B-operand shared-memory traffic per
    // MMA and the B bytes each CTA pulls through TMA.  B200_GEMM_CTA_PAIR=0 keeps one CTA per tile.
    const int cg = (pair_enabled() && block_n == 256 && M > BLOCK_M) ? 2 : 1;
    p.m_tiles = (M + BLOCK_M * cg - 1) / (BLOCK_M * cg);
    p.n_tiles = (N + block_n - 1) / block_n;
    p.mn_lbo = BLOCK_K * 128;
    p.mn_sbo = 1024;
    if (const char* e = getenv("B200_GEMM_MN_SWAP")) {
        if (e[0] == '1') { p.mn_lbo = 1024; p.mn_sbo = BLOCK_K * 128; }
    }
    if (!swiglu && (splits > 1 || accumulate)) {
        B200_CHECK_ARG(R == nullptr, "gemm: split-K/accumulate cannot be combined with a residual epilogue");
        B200_CHECK_ARG(ldc == N && N % 8 == 0, "gemm: split-K/accumulate output must be contiguous (ldc == N, N %% 8 == 0)");
        const size_t need = (size_t)splits * M * N * sizeof(float);
        B200_CHECK_ARG(workspace != nullptr && workspace_bytes >= need, "gemm: workspace too small (%zu < %zu)",
                       workspace_bytes, need);
        p.epilogue = EPI_PARTIAL_F32;
    } else {
        p.epilogue = R ? EPI_RESIDUAL : EPI_STORE;
    }
    p.rope_cos = rope_cos; p.rope_sin = rope_sin; p.rope_S = rope_S; p.rope_D = rope_D; p.rope_cols = rope_cols;
    if (rope_cos) p.epilogue = EPI_ROPE;
    p.act = nullptr; p.swiglu_I = 0; p.ld_act = 0;
    if (swiglu) {
        p.epilogue = EPI_SWIGLU;
        p.act = (bf16*)workspace;
        p.swiglu_I = (int)(workspace_bytes & 0xffffffffu);
        p.ld_act = (int)(workspace_bytes >> 32);
        p.n_tiles = p.swiglu_I / 128;         // one tile = 128 gate + 128 up features
    }

    CUtensorMap tmA, tmB, tmC;
    int rc;
    static int tma_store_ok = -1;
    if (tma_store_ok < 0) {
        const char* e = getenv("B200_GEMM_TMA_STORE");       // 0 = per-thread global stores (bring-up / A-B timing)
        tma_store_ok = (e && e[0] == '0') ? 0 : 1;
    }
    // staged (TMA-store) epilogues: plain store always; SwiGLU when I is a multiple of 128 and the outputs are 16-byte
    // pitched; RoPE for head_dim 64 (one staging box = one head)
    p.tma_store = 0;
    if (tma_store_ok) {
        if (p.epilogue == EPI_STORE) p.tma_store = 1;
        else if (p.epilogue == EPI_SWIGLU && block_n == 256 && p.ld_act % 8 == 0 && ((uintptr_t)p.act % 16 == 0)) p.tma_store = 1;
        else if (p.epilogue == EPI_ROPE && rope_D == 64 && rope_cols % 64 == 0) p.tma_store = 1;
    }
    // tail split (plain bf16 store, no split-K, no accumulate) when the caller provided the optional workspace
    p.total_items = p.m_tiles * p.n_tiles * p.splits;
    p.tail_tile0 = p.m_tiles * p.n_tiles;
    p.tail_splits = 1; p.tail_kb = p.num_kb; p.tail_ws = nullptr;
    int n_tail = 0;
    if (cg == 1 && p.epilogue == EPI_STORE && p.splits == 1 && !accumulate && workspace != nullptr) {
        const int tiles = p.m_tiles * p.n_tiles, sms = b200_num_sms();
        const int ts = plan_tail(tiles, p.num_kb, sms);
        const size_t need = (size_t)(tiles % sms) * ts * BLOCK_M * block_n * sizeof(float);
        if (ts > 1 && workspace_bytes >= need && ((uintptr_t)workspace % 16 == 0)) {
            n_tail = tiles % sms;
            p.tail_tile0 = tiles - n_tail;
            p.tail_splits = ts;
            p.tail_kb = (p.num_kb + ts - 1) / ts;
            p.tail_splits = (p.num_kb + p.tail_kb - 1) / p.tail_kb;      // no empty slices
            p.tail_ws = (float*)workspace;
            p.total_items = p.tail_tile0 + n_tail * p.tail_splits;
        }
    }
    CUtensorMap tmD;
    if (p.tma_store) {
        if ((rc = tc05::make_tmap_2d(&tmC, C, N8, M, ldc, 64, BLOCK_M))) return rc;
        if (p.epilogue == EPI_SWIGLU) {
            if ((rc = tc05::make_tmap_2d(&tmD, p.act, p.swiglu_I, M, p.ld_act, 64, BLOCK_M))) return rc;
        } else tmD = tmC;
    }
    if (!a_mn_major) rc = tc05::make_tmap_2d(&tmA, A, K, M, lda, BLOCK_K, BLOCK_M);
    else             rc = tc05::make_tmap_2d(&tmA, A, M, K, lda, 64, BLOCK_K);
    if (rc) return rc;
    if (!b_mn_major) rc = tc05::make_tmap_2d(&tmB, B, K, N, ldb, BLOCK_K, swiglu ? 128 : block_n / cg);
    else             rc = tc05::make_tmap_2d(&tmB, B, N, K, ldb, 64, BLOCK_K);
    if (rc) return rc;
    if (!p.tma_store) { tmC = tmA; tmD = tmA; }      // unused by the kernel, but must be valid maps

    // epilogue warp groups: one (4 warps) for the plain epilogues, two (8 warps) for the fused RoPE / SwiGLU epilogues whose
    // conversions and SFU work would otherwise outlast a K = 1024 main loop (B200_GEMM_EG_PLAIN / B200_GEMM_EG_FUSED override)
    static int eg_plain = -1, eg_fused = -1;
    if (eg_plain < 0) {
        const char* e1 = getenv("B200_GEMM_EG_PLAIN");
        const char* e2 = getenv("B200_GEMM_EG_FUSED");
        eg_plain = (e1 && e1[0] == '2') ? 2 : 1;
        eg_fused = (e2 && e2[0] == '1') ? 1 : 2;
    }
    const int eg = (p.epilogue == EPI_ROPE || p.epilogue == EPI_SWIGLU) ? eg_fused : eg_plain;
#define B200_DISPATCH(BN, EGV, CGV)                                                                   \
    do {                                                                                    \
        if (!a_mn_major && !b_mn_major) rc = launch<BN, false, false, EGV, CGV>(tmA, tmB, tmC, tmD, p, stream);  \
        else if (!a_mn_major && b_mn_major) rc = launch<BN, false, true, EGV, CGV>(tmA, tmB, tmC, tmD, p, stream); \
        else if (a_mn_major && b_mn_major) rc = launch<BN, true, true, EGV, CGV>(tmA, tmB, tmC, tmD, p, stream);  \
        else rc = launch<BN, true, false, EGV, CGV>(tmA, tmB, tmC, tmD, p, stream);                              \
    } while (0)
    if (block_n == 256) {
        if (cg == 2) { if (eg == 2) B200_DISPATCH(256, 2, 2); else B200_DISPATCH(256, 1, 2); }
        else { if (eg == 2) B200_DISPATCH(256, 2, 1); else B200_DISPATCH(256, 1, 1); }
    } else { if (eg == 2) B200_DISPATCH(128, 2, 1); else B200_DISPATCH(128, 1, 1); }
#undef B200_DISPATCH
    if (rc) return rc;

    if (n_tail > 0) {
        const long long groups = (long long)n_tail * BLOCK_M * block_n / 8;
        int blocks = (int)((groups + 255) / 256);
        if (blocks > b200_num_sms() * 8) blocks = b200_num_sms() * 8;
        tail_reduce_kernel<<<blocks, 256, 0, stream>>>(p.tail_ws, p.C, M, N8, ldc, p.n_tiles, p.tail_tile0, n_tail, p.tail_splits, block_n);
        B200_CHECK_LAUNCH("gemm_tail_reduce");
    }
    if (p.epilogue == EPI_PARTIAL_F32) {
        const size_t n = (size_t)M * N;
        int blocks = (int)((n / 4 + 255) / 256);
        if (blocks > b200_num_sms() * 8) blocks = b200_num_sms() * 8;
        splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p.ws, p.C, n, splits, accumulate);
        B200_CHECK_LAUNCH("splitk_reduce");
    }
    return B200_OK;
}
