// Gather-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), the contraction behind every
// Conv1d / Linear of the ppvector hot path (reference: ppvector/models/utils.py:65-77 Conv1d.forward,
// reached from TDNNBlock utils.py:147, Res2NetBlock ecapa_tdnn.py:36-47, ASP pooling.py:107, fc).
//
//   out[r, n] = epilogue( sum_s  A_{map(s)}[r + row_off(s), a_col(s) : a_col(s)+64] . W[n, 64 s : 64 s + 64] )
//
// A "k-step" s is one 64-wide K slice: a conv tap is a row offset in the padded time layout, a channel
// concat is a different source tensor, `x_i + y_{i-1}` (Res2Net) is two sources sharing the same weights.
// Operands are split-bf16 planes (hi, lo); PPV_PREC_BF16X3 issues hi*hi + lo*hi + hi*lo into one fp32
// TMEM accumulator (fp32-grade), PPV_PREC_BF16 issues hi*hi only.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):
//   warp 0    TMA producer  : cp.async.bulk.tensor 3-D tiles (SWIZZLE_128B) into a STAGES-deep smem ring
//   warp 1    MMA issuer    : one lane issues tcgen05.mma (M=128, N=BN, K=16) ; tcgen05.commit frees the slot
//   warp 2    TMEM allocator
//   warps 4-11 epilogue     : (two warps per TMEM lane quarter, alternating 32-column chunks) tcgen05.ld 32 columns at a time -> bias / ReLU / BN affine / tanh ->
//                             split-bf16 (or fp32) stores incl. the reflect-halo rows
// TMEM holds two BN-column fp32 accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace ppv {

// BK = K elements per pipeline stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B).  The smaller
// stage keeps the same bytes per MMA but doubles the number of ring slots, i.e. more TMA bytes in flight for the same
// shared memory: the K=512 layers are bound by load latency x bytes-in-flight, not by the tensor pipe.
template <int BN, int NSPLIT, int BK>
struct GemmCfg {
    static constexpr int NA = (NSPLIT == 3) ? 2 : 1;  // A tiles per stage (hi[, lo])
    static constexpr int NB = NA;
    static constexpr int A_BYTES = GEMM_BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = NA * A_BYTES + NB * B_BYTES;
    static constexpr int STAGING_BYTES = EPI_STAGING_BYTES;  // TMA-store epilogue staging tile
    static constexpr int BAR_BYTES = 256;
    static constexpr int MAX_SMEM = 232448;  // 227 KB
    static constexpr int STAGES_RAW = (MAX_SMEM - 1024 - STAGING_BYTES - BAR_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: power of two >= 32
    static constexpr int MAX_STAGES = 8;      // barrier slots
    // pair mode (cta_group::2): a CTA's ring slot holds its A rows and HALF of the weight tile; the ring reuses the same bytes
    static constexpr int PAIR_STAGE_BYTES = NA * A_BYTES + NB * (B_BYTES / 2);
    static constexpr int PAIR_STAGES_RAW = (STAGES * STAGE_BYTES) / PAIR_STAGE_BYTES;
    static constexpr int PAIR_STAGES = PAIR_STAGES_RAW > MAX_STAGES ? MAX_STAGES : PAIR_STAGES_RAW;
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
    static_assert(BK == 64 || BK == 32, "BK");
};

template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    if (BK == 64) return make_sw128_kmajor_desc(smem_addr);
    // SWIZZLE_64B: rows of 32 bf16 (64 B), 8-row groups 512 B apart, layout type 4
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;
    return d;
}

// PAIR is a compile-time switch: a kernel that contains cta_group::2 instructions can only be launched in clusters of two.
template <int BN, int NSPLIT, int BK, bool PAIR = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams gp) {
    using Cfg = GemmCfg<BN, NSPLIT, BK>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment.
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t tiles_base = smem_base;
    const uint32_t staging = smem_base + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned (stage sizes are multiples of 1024)
    uint8_t* staging_gen = smem_gen + STAGES * Cfg::STAGE_BYTES;
    const uint32_t bar_base = staging + Cfg::STAGING_BYTES;
    // barrier layout (8 B each): full[8], empty[8], tmem_full[2], tmem_empty[2], tmem ptr, resident-weights barrier
    constexpr int MAXST = Cfg::MAX_STAGES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAXST + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAXST + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAXST + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAXST + 4);
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + 8 * (2 * MAXST + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Weight-stationary mode (gp.ws, narrow convs with one n-tile and a small K): the whole weight matrix is loaded ONCE per CTA
    // into the upper half of the ring's shared memory and the ring (4 slots) carries activation tiles only -- for the
    // 32-channel 3x3 convs of the 2-D models the per-tile reload of the weights was a third of the L2 -> SM traffic.
    // Pair mode (gp.pair, cta_group::2): the two CTAs of a cluster own the two 128-row halves of a 256 x BN tile.  Each loads its own
    // activation rows and HALF of the weight tile (BN / 2 rows); the leader's MMA thread issues 256 x BN x 16 MMAs that read both
    // CTAs' shared memory and write both CTAs' TMEM.  Shared-memory fill per CTA and k-step drops from A + B to A + B / 2 (for BN = 256:
    // 96 -> 64 KB), which is what bounds the K = 512 layers (TMA fill ~6300 B/clk chip-wide), and the ring holds more k-steps.
    constexpr bool pair = PAIR;
    uint32_t pair_rank = 0;
    if constexpr (PAIR) pair_rank = cluster_ctarank();
    const bool leader = pair_rank == 0;
    const int stage_bytes = pair ? Cfg::PAIR_STAGE_BYTES : Cfg::STAGE_BYTES;
    const int nst = gp.ws ? GEMM_WS_STAGES : (pair ? Cfg::PAIR_STAGES : STAGES);
    const uint32_t w_res = tiles_base + GEMM_WS_STAGES * Cfg::STAGE_BYTES;
    const uint32_t w_full = bar_base + 8u * (2 * MAXST + 5);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < GEMM_MAX_MAPS; ++i) prefetch_tmap(&gp.mapA[i]);
        prefetch_tmap(&gp.mapB);
        if (gp.epi.tma_store) prefetch_tmap(&gp.mapOut);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < MAXST; ++s) {
            mbar_init(full_bar(s), 1);   // pair: only the leader's full barriers are used (both CTAs' loads complete_tx there)
            mbar_init(empty_bar(s), 1);  // pair: released in both CTAs by the leader's multicast commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), pair ? 2 * EPI_WARP_ARRIVALS : EPI_WARP_ARRIVALS);  // one arrival per epilogue warp; pair: both CTAs' warps arrive at the leader
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) {
            tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before any complete_tx / multicast commit / remote arrive can reach them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    griddep_launch_dependents();  // PDL: the next kernel may begin its prologue
    griddep_wait();               // PDL: upstream activations are complete and visible

    // lin_splits > 0: "wgrad" mode -- the contraction runs over the COLUMNS of both operands (two transposed matrices
    // [M][R] and [N][R]), k-step s of split z reads columns (z * nk + s) * BK; there is no k-step table, and split z of
    // tile (m, n) accumulates into its own partial output block (rows shifted by z * lin_split_rows).
    const int mn_tiles = gp.m_tiles * gp.n_tiles;
    const int num_tiles = mn_tiles * (gp.lin_splits > 0 ? gp.lin_splits : 1);
    const int nk = gp.num_ksteps;
    // work list of this CTA: tile_first, tile_first + tile_step, ... < tile_end.  Default: tile = m * n_tiles + n over all CTAs.
    // pair: "pair tiles" pt = mp * n_tiles + n over the clusters; this CTA takes m = 2 * mp + rank (an odd last m-tile leaves one CTA of
    // the last pair on rows >= M: it still loads its weight half; its activation loads are zero-filled and its stores dropped).
    int tile_first = int(blockIdx.x), tile_step = int(gridDim.x);
    if constexpr (PAIR) {
        tile_first = int(cluster_id_x());
        tile_step = int(num_clusters_x());
    }
    const int tile_end = pair ? ((gp.m_tiles + 1) / 2) * gp.n_tiles : num_tiles;
    auto tile_m = [&](int t) { return pair ? 2 * (t / gp.n_tiles) + int(pair_rank) : (t % mn_tiles) / gp.n_tiles; };
    auto tile_n = [&](int t) { return pair ? t % gp.n_tiles : (t % mn_tiles) % gp.n_tiles; };

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        if (gp.ws && lane == 0) {  // resident weights: every k-slice, once
            mbar_arrive_expect_tx(w_full, nk * Cfg::NB * Cfg::B_BYTES);
            for (int s = 0; s < nk; ++s)
                for (int p = 0; p < Cfg::NB; ++p) tma_load_3d(w_res + (s * Cfg::NB + p) * Cfg::B_BYTES, &gp.mapB, w_full, s * BK, 0, p);
        }
        __syncwarp();
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            const int zsplit = pair ? 0 : tile / mn_tiles;
            const int m0 = tile_m(tile) * GEMM_BM;
            const int n0 = tile_n(tile) * BN;
            // L2 prefetch of the activation rows of this CTA's NEXT tile (one CTA per m-tile issues it): they come
            // from HBM, and a 2-4 slot ring alone cannot hide that latency.
            const int ntile = tile + tile_step;
            if (gp.l2_prefetch && ntile < tile_end && tile_n(ntile) == 0 && lane < nk) {
                const int nm0 = tile_m(ntile) * GEMM_BM;
                for (int s = lane; s < nk; s += 32) {
                    const KStep ks = gp.ksteps[s];
#pragma unroll
                    for (int p = 0; p < Cfg::NA; ++p) tma_prefetch_l2_3d(&gp.mapA[ks.map], ks.a_col, nm0 + ks.row_off, p);
                }
            }
            __syncwarp();
            for (int s = 0; s < nk; ++s) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (lane == 0) {
                    const uint32_t sa = tiles_base + stage * stage_bytes;
                    const uint32_t sb = sa + Cfg::NA * Cfg::A_BYTES;
                    const uint32_t fb = full_bar(stage);
                    if constexpr (PAIR) {
                        // both CTAs' boxes complete_tx on the LEADER's full barrier; the leader arms it with the bytes of both
                        if (leader) mbar_arrive_expect_tx(fb, 2 * Cfg::PAIR_STAGE_BYTES);
                        const uint32_t fbl = map_to_cta(fb, 0);
                        const KStep ks = gp.ksteps[s];
                        const CUtensorMap* ma = &gp.mapA[ks.map];
#pragma unroll
                        for (int p = 0; p < Cfg::NA; ++p) tma_load_3d_2sm(sa + p * Cfg::A_BYTES, ma, fbl, ks.a_col, m0 + ks.row_off, p);
#pragma unroll
                        for (int p = 0; p < Cfg::NB; ++p)
                            tma_load_3d_2sm(sb + p * (Cfg::B_BYTES / 2), &gp.mapBh, fbl, s * BK, n0 + int(pair_rank) * (BN / 2), p);
                    } else {
                    mbar_arrive_expect_tx(fb, gp.ws ? Cfg::NA * Cfg::A_BYTES : Cfg::STAGE_BYTES);
                    if (gp.ws) {
                        const KStep ks = gp.ksteps[s];
#pragma unroll
                        for (int p = 0; p < Cfg::NA; ++p) tma_load_3d(sa + p * Cfg::A_BYTES, &gp.mapA[ks.map], fb, ks.a_col, m0 + ks.row_off, p);
                    } else if (gp.lin_splits > 0) {
                        const int kcol = (zsplit * nk + s) * BK;
#pragma unroll
                        for (int p = 0; p < Cfg::NA; ++p) tma_load_3d(sa + p * Cfg::A_BYTES, &gp.mapA[0], fb, kcol, m0, p);
#pragma unroll
                        for (int p = 0; p < Cfg::NB; ++p)
                            tma_load_3d(sb + p * Cfg::B_BYTES, &gp.mapB, fb, kcol + gp.lin_b_col0, gp.lin_b_row0 + n0, p);
                    } else {
                        const KStep ks = gp.ksteps[s];
                        const CUtensorMap* ma = &gp.mapA[ks.map];
#pragma unroll
                        for (int p = 0; p < Cfg::NA; ++p)
                            tma_load_3d(sa + p * Cfg::A_BYTES, ma, fb, ks.a_col, m0 + ks.row_off, p);
#pragma unroll
                        for (int p = 0; p < Cfg::NB; ++p)
                            tma_load_3d(sb + p * Cfg::B_BYTES, &gp.mapB, fb, s * BK, n0, p);
                    }
                    }
                }
                __syncwarp();
                if (++stage == nst) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (PAIR && warp == 1) {
        // ===================== MMA issuer, pair mode: the leader CTA's thread drives both SMs =====================
        if constexpr (PAIR) if (leader) {
            constexpr uint32_t idesc2 = make_idesc_bf16(2 * GEMM_BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // BOTH CTAs' epilogues drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int s = 0; s < nk; ++s) {
                    mbar_wait(full_bar(stage), phase);  // both CTAs' operands of this k-step have landed
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = tiles_base + stage * Cfg::PAIR_STAGE_BYTES;
                        const uint32_t sb = sa + Cfg::NA * Cfg::A_BYTES;
                        const uint64_t a_hi = make_kmajor_desc<BK>(sa);
                        const uint64_t b_hi = make_kmajor_desc<BK>(sb);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc2, (s > 0 || k > 0) ? 1u : 0u);
                        if (NSPLIT == 3) {
                            const uint64_t a_lo = make_kmajor_desc<BK>(sa + Cfg::A_BYTES);
                            const uint64_t b_lo = make_kmajor_desc<BK>(sb + Cfg::B_BYTES / 2);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc2, 1u);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc2, 1u);
                        }
                        umma_commit_2sm(empty_bar(stage), uint16_t(3));                   // the slot is free in both CTAs
                        if (s == nk - 1) umma_commit_2sm(tfull_bar(acc), uint16_t(3));    // both halves of the accumulator are complete
                    }
                    __syncwarp();
                    if (++stage == nst) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        if (gp.ws) mbar_wait(w_full, 0);
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int s = 0; s < nk; ++s) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = tiles_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = gp.ws ? w_res + s * Cfg::NB * Cfg::B_BYTES : sa + Cfg::NA * Cfg::A_BYTES;
                    const uint64_t a_hi = make_kmajor_desc<BK>(sa);
                    const uint64_t b_hi = make_kmajor_desc<BK>(sb);
                    // K advance inside the 128-byte swizzle atom: +32 B (16 bf16) per UMMA_K => +2 in desc.lo
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (s > 0 || k > 0) ? 1u : 0u);
                    if (NSPLIT == 3) {
                        const uint64_t a_lo = make_kmajor_desc<BK>(sa + Cfg::A_BYTES);
                        const uint64_t b_lo = make_kmajor_desc<BK>(sb + Cfg::B_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                    }
                    umma_commit(empty_bar(stage));  // smem slot free when these MMAs retire
                    if (s == nk - 1) umma_commit(tfull_bar(acc));   // accumulator complete
                }
                __syncwarp();
                if (++stage == nst) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int etid = threadIdx.x - 128;  // 0..GEMM_EPI_THREADS-1
        const int ehalf = (warp - 4) >> 2;   // two warps share a TMEM lane quarter and split the column chunks
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            const int m0 = tile_m(tile) * GEMM_BM;
            const int n0 = tile_n(tile) * BN;
            epilogue_tile<BN>(gp.epi, &gp.mapOut, gp.M, gp.N, m0, n0, tmem_base + acc * BN, tfull_bar(acc), acc_phase, tempty_bar(acc), q,
                              lane, ehalf, etid, staging, staging_gen, pair ? 0 : int64_t(tile / mn_tiles) * gp.lin_split_rows, INT64_MIN, 0,
                              (PAIR && !leader) ? map_to_cta(tempty_bar(acc), 0) : 0u);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if ((etid & 127) == 0) bulk_wait_all();  // TMA stores issued by the two issuing threads have completed before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();  // no CTA retires while its peer may still read its shared memory / arrive on its barriers
    if (warp == 2) {
        if constexpr (PAIR) {
            tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
        } else {
            tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_planes_map_ex(CUtensorMap* m, const Planes& t, int box_cols, int box_rows, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(PPV_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(t.base) & 15) || (t.ld % 8) || (t.plane_stride % 8))
        return fail(PPV_EINVAL, "planes tensor not 16-byte aligned");
    cuuint64_t dims[3] = {cuuint64_t(t.ld), cuuint64_t(t.rows), 2};
    cuuint64_t strides[2] = {cuuint64_t(t.ld) * 2, cuuint64_t(t.plane_stride) * 2};
    cuuint32_t box[3] = {cuuint32_t(box_cols), cuuint32_t(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, t.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PPV_ECUDA, "cuTensorMapEncodeTiled failed, CUresult " + std::to_string(int(r)));
    return PPV_OK;
}

// 3-D map over split planes [2][rows][ld] bf16, box = {64 cols, box_rows, 1 plane}, SWIZZLE_128B.
int encode_planes_map(CUtensorMap* m, const Planes& t, int box_rows) { return encode_planes_map_ex(m, t, GEMM_BK, box_rows, 128); }

int gemm_build(GemmParams* gp, const GemmSource* srcs, int nsrc, const Planes& W, int M, int N, const Epilogue& epi,
               int BN, int BK) {
    PPV_REQUIRE(BK == 64 || BK == 32, "gemm_build: BK must be 64 or 32");
    PPV_REQUIRE(BN == 64 || BN == 128 || BN == 256, "gemm_build: BN must be 64/128/256");
    PPV_REQUIRE(epi.out_mode == OUT_F32 || N % 32 == 0, "gemm_build: planes output needs N % 32 == 0");
    PPV_REQUIRE(!epi.rowgrp_bias || N % 32 == 0, "gemm_build: row-group bias needs N % 32 == 0");
    memset(gp, 0, sizeof(*gp));
    // distinct A tensors -> maps
    const __nv_bfloat16* bases[GEMM_MAX_MAPS];
    int nmaps = 0;
    int ks = 0;
    for (int i = 0; i < nsrc; ++i) {
        const GemmSource& s = srcs[i];
        PPV_REQUIRE(s.ncols % BK == 0 && s.col0 % 8 == 0, "gemm_build: source K slice must be a multiple of the k-step");
        PPV_REQUIRE(s.col0 + s.ncols <= s.t.ld, "gemm_build: source K slice exceeds the row");
        int mi = -1;
        for (int j = 0; j < nmaps; ++j)
            if (bases[j] == s.t.base) mi = j;
        if (mi < 0) {
            PPV_REQUIRE(nmaps < GEMM_MAX_MAPS, "gemm_build: too many distinct A tensors");
            mi = nmaps++;
            bases[mi] = s.t.base;
            int rc = encode_planes_map_ex(&gp->mapA[mi], s.t, BK, GEMM_BM, BK * 2);
            if (rc) return rc;
        }
        for (int c = 0; c < s.ncols; c += BK) {
            PPV_REQUIRE(ks < GEMM_MAX_KSTEPS, "gemm_build: too many k-steps");
            gp->ksteps[ks].map = int16_t(mi);
            gp->ksteps[ks].row_off = int16_t(s.row_off);
            gp->ksteps[ks].a_col = s.col0 + c;
            ++ks;
        }
    }
    for (int j = nmaps; j < GEMM_MAX_MAPS; ++j) gp->mapA[j] = gp->mapA[0];
    PPV_REQUIRE(ks > 0, "gemm_build: empty K");
    PPV_REQUIRE(W.ld == ks * BK, "gemm_build: weight K does not match the k-steps");
    PPV_REQUIRE(W.rows >= N, "gemm_build: weight rows < N");
    int rc = encode_planes_map_ex(&gp->mapB, W, BK, BN, BK * 2);
    if (rc) return rc;
    gp->num_ksteps = ks;
    gp->bk = BK;
    gp->M = M;
    gp->N = N;
    gp->m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    gp->n_tiles = (N + BN - 1) / BN;
    gp->epi = epi;
    {
        // weight-stationary: one n-tile, all k-slices of W (both planes) fit next to a 4-slot ring, and enough tiles per CTA to pay
        const size_t stage_bytes = size_t(2) * GEMM_BM * BK * 2 + size_t(2) * BN * BK * 2;
        const size_t stages_full = std::min<size_t>(8, (232448 - 1024 - EPI_STAGING_BYTES - 256) / stage_bytes);
        const size_t w_bytes = size_t(ks) * 2 * BN * BK * 2;
        const char* wsenv = getenv("PPV_GEMM_WS");
        gp->ws = (gp->n_tiles == 1 && stages_full > GEMM_WS_STAGES && w_bytes <= (stages_full - GEMM_WS_STAGES) * stage_bytes && gp->m_tiles >= 296 &&
                  !(wsenv && wsenv[0] == '0'))
                     ? 1
                     : 0;
    }
    {
        // pair mode (cta_group::2, 256 x 256 tiles over two SMs): wide layers with many m-tiles that are not weight-stationary.
        // PPV_GEMM_PAIR=0 keeps them on single-CTA 128 x 256 tiles (A-B timing).
        const char* penv = getenv("PPV_GEMM_PAIR");
        gp->pair = (BN == 256 && !gp->ws && gp->m_tiles >= 64 && N % BN == 0 && !(penv && penv[0] == '0')) ? 1 : 0;
        if (gp->pair) {
            rc = encode_planes_map_ex(&gp->mapBh, W, BK, BN / 2, BK * 2);
            if (rc) return rc;
        }
    }
    {
        const char* ns = getenv("PPV_GEMM_NOSTORE");
        gp->epi.debug_nostore = (ns && ns[0] == '1') ? 1 : 0;
        const char* pf = getenv("PPV_GEMM_NO_L2PREFETCH");
        gp->l2_prefetch = (pf && pf[0] == '1') ? 0 : 1;
    }
    // Image-mode outputs on the SAME grid as the input (stride 1) can also go through the TMA-store staging tile: the rows that
    // are not interior positions are the zero border of the image and are stored as zeros (which is what they must hold).
    const bool img_same_grid = epi.img_Wp > 0 && epi.img_stride == 1 && epi.img_stride_w <= 1 && epi.out_Hp == epi.img_Hp && epi.out_Wp == epi.img_Wp;
    if (epi.out_mode == OUT_PLANES && !epi.halo && (N % 64) == 0 && (epi.img_Wp == 0 || img_same_grid)) {
        const char* nt = getenv("PPV_GEMM_NO_TMASTORE");
        if (!(nt && nt[0] == '1')) {
            Planes po;
            po.base = static_cast<__nv_bfloat16*>(epi.out);
            po.ld = int(epi.out_ld);
            po.plane_stride = epi.out_plane_stride;
            po.rows = epi.out_plane_stride / epi.out_ld;  // planes are allocated back to back
            if (po.rows * po.ld == po.plane_stride && (epi.out_col0 % 8) == 0) {
                rc = encode_planes_map_ex(&gp->mapOut, po, 32, GEMM_BM, 64);  // 32-column chunks, SWIZZLE_64B staging (gemm_epilogue.cuh)
                if (rc) return rc;
                gp->epi.tma_store = 1;
                if (img_same_grid) gp->epi.zero_invalid = 1;
            }
        }
    }
    if (epi.out_mode == OUT_PLANES) {
        PPV_REQUIRE((epi.out_ld % 16) == 0 && (epi.out_col0 % 16) == 0 && (epi.out_plane_stride % 16) == 0 &&
                        (reinterpret_cast<uintptr_t>(epi.out) & 31) == 0,
                    "gemm_build: planes output must be 32-byte aligned (256-bit stores)");
    } else {
        gp->epi.f32_vec_ok = ((epi.out_ld % 4) == 0 && (epi.out_col0 % 4) == 0 && (reinterpret_cast<uintptr_t>(epi.out) & 15) == 0) ? 1 : 0;
        if ((epi.out_ld % 8) == 0 && (epi.out_col0 % 8) == 0 && (reinterpret_cast<uintptr_t>(epi.out) & 31) == 0) gp->epi.f32_vec_ok = 2;
    }
    return PPV_OK;
}

// Weight-gradient GEMM: out[z][m, out_col0 + n] = sum over columns r of split z of  At[m, r] * Bt[b_row0 + n, r + b_col0]
// (At = transposed output gradient [M][R], Bt = transposed layer input [*][R]; b_col0 = conv tap offset in rows of the
// padded time layout).  Partial blocks are `split_rows` output rows apart; the caller sums them.
int gemm_build_wgrad(GemmParams* gp, const Planes& At, const Planes& Bt, int M, int N, int b_row0, int b_col0, int splits, float* out,
                     int64_t out_ld, int out_col0, int64_t split_rows, int BN) {
    PPV_REQUIRE(BN == 64 || BN == 128 || BN == 256, "gemm_build_wgrad: BN must be 64/128/256");
    PPV_REQUIRE(At.ld == Bt.ld && splits >= 1, "gemm_build_wgrad: operands must share the contraction length");
    memset(gp, 0, sizeof(*gp));
    int rc = encode_planes_map_ex(&gp->mapA[0], At, GEMM_BK, GEMM_BM, 128);
    if (rc) return rc;
    for (int j = 1; j < GEMM_MAX_MAPS; ++j) gp->mapA[j] = gp->mapA[0];
    rc = encode_planes_map_ex(&gp->mapB, Bt, GEMM_BK, BN, 128);
    if (rc) return rc;
    const int nk_total = (At.ld + GEMM_BK - 1) / GEMM_BK;
    gp->lin_splits = std::min(splits, nk_total);
    gp->num_ksteps = (nk_total + gp->lin_splits - 1) / gp->lin_splits;
    gp->lin_b_row0 = b_row0;
    gp->lin_b_col0 = b_col0;
    gp->lin_split_rows = split_rows;
    gp->bk = GEMM_BK;
    gp->M = M;
    gp->N = N;
    gp->m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    gp->n_tiles = (N + BN - 1) / BN;
    Epilogue ep;
    ep.out_mode = OUT_F32;
    ep.out = out;
    ep.out_ld = out_ld;
    ep.out_col0 = out_col0;
    gp->epi = ep;
    gp->epi.f32_vec_ok = ((out_ld % 4) == 0 && (out_col0 % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    if ((out_ld % 8) == 0 && (out_col0 % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) gp->epi.f32_vec_ok = 2;
    PPV_REQUIRE(N % 32 == 0, "gemm_build_wgrad: N % 32 == 0 required");
    return PPV_OK;
}

template <int BN, int NSPLIT, int BK>
static int launch_one(const GemmParams& gp, int num_sms, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, NSPLIT, BK>;
    PPV_ONCE_PER_DEVICE(PPV_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, NSPLIT, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES)));
    const int tiles = gp.m_tiles * gp.n_tiles * (gp.lin_splits > 0 ? gp.lin_splits : 1);
    if constexpr (BN == 256) if (gp.pair) {  // clusters of two CTAs; one cluster per pair tile at most
        PPV_ONCE_PER_DEVICE(PPV_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, NSPLIT, BK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES)));
        const int pair_tiles = ((gp.m_tiles + 1) / 2) * gp.n_tiles;
        const int clusters = std::min(pair_tiles, num_sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        PPV_PDL_OK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, NSPLIT, BK, true>, gp), "gemm_tcgen05_kernel (cta_group::2 pairs)");
        return PPV_OK;
    }
    const int grid = std::min(tiles, num_sms);
    PPV_PDL_OK(launch_pdl(gemm_tcgen05_kernel<BN, NSPLIT, BK>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, gp),
               "gemm_tcgen05_kernel");
    return PPV_OK;
}

template <int BN>
static int launch_bn(const GemmParams& gp, bool x3, int num_sms, cudaStream_t stream) {
    if (gp.bk == 32) return x3 ? launch_one<BN, 3, 32>(gp, num_sms, stream) : launch_one<BN, 1, 32>(gp, num_sms, stream);
    return x3 ? launch_one<BN, 3, 64>(gp, num_sms, stream) : launch_one<BN, 1, 64>(gp, num_sms, stream);
}

int gemm_launch(const GemmParams& gp, int BN, int precision, int num_sms, cudaStream_t stream) {
    const bool x3 = (precision == PPV_PREC_BF16X3);
    switch (BN) {
        case 64: return launch_bn<64>(gp, x3, num_sms, stream);
        case 128: return launch_bn<128>(gp, x3, num_sms, stream);
        case 256: return launch_bn<256>(gp, x3, num_sms, stream);
    }
    return fail(PPV_EINVAL, "gemm_launch: bad BN");
}

}  // namespace ppv
