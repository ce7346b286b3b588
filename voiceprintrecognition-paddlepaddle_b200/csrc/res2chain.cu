// The whole Res2Net chain of one SE-Res2Net block in ONE kernel: y_1 = f_1(x_1), y_j = f_j(x_j + y_{j-1}), j = 2..7, with
// f_j = BN(ReLU(conv_k3,d(.))) on 64-channel chunks (ppvector/models/ecapa_tdnn.py:36-47, TDNNBlock utils.py:147).
//
// The seven convs are strictly sequential, and as seven launches each one costs a full kernel latency (~24 us for 4 us of
// tensor work; profiles/launches_r1_final_summary.txt).  Here one CTA owns one UTTERANCE: its whole padded time axis
// (<= 384 rows x 64 channels, split-bf16) lives in shared memory in the SWIZZLE_128B layout the UMMA descriptors read, so
//   * conv taps are descriptor row offsets into that resident tile (as in res2conv.cu),
//   * the epilogue of conv j writes BN(ReLU(.)) to HBM once (the tdnn2 GEMM reads it) and writes x_{j+1} + y_j -- including
//     the reflect-padding halo rows -- straight back into the shared-memory tile as the A operand of conv j+1:
//     the intermediate sums never touch HBM and there is no inter-CTA dependency at all,
//   * the 48 KB of weights of conv j+1 stream into a second slot by TMA while conv j runs.
// Warp roles: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM, warps 4-19 epilogue (640 threads).  TMEM holds three
// 128 x 64 fp32 accumulators (the three 128-row tiles of the utterance).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace ppv {

constexpr int RC_PAD = 4;                       // rows loaded above the utterance (max dilation)
constexpr int RC_BOX_ROWS = 200;                // TMA box rows (<= 256); 200 * 128 B is a multiple of 1024
constexpr int RC_ROWS = 2 * RC_BOX_ROWS;        // 400 >= 4 + 384 + 4
constexpr int RC_A_PLANE = RC_ROWS * 128;       // 51200 B per plane
constexpr int RC_W_TILE = 64 * 128;             // [64 out ch x 64 k] bf16
constexpr int RC_MAX_TILES = 3;                 // 128-row output tiles per utterance
constexpr int RC_MAX_TP = RC_MAX_TILES * GEMM_BM;  // 384
constexpr int RC_S_PLANE = GEMM_BM * 128;       // staging tile: 128 rows x 64 channels bf16, SWIZZLE_128B
constexpr int RC_EPI_THREADS = 512;             // 16 epilogue warps: four per TMEM lane quarter, 16 channels each
constexpr int RC_THREADS = 128 + RC_EPI_THREADS;

template <int NSPLIT>
struct RCCfg {
    static constexpr int NP = (NSPLIT == 3) ? 2 : 1;
    static constexpr int A_BYTES = NP * RC_A_PLANE;
    static constexpr int W_SLOT = 3 * NP * RC_W_TILE;  // 3 taps x planes, single slot: the next conv's weights load under the epilogue
    static constexpr int S_BYTES = NP * RC_S_PLANE;    // one staging tile (hi, lo)
    static constexpr int SMEM_BYTES = 1024 + A_BYTES + W_SLOT + 2 * S_BYTES + 256;
};

#define RC_STAMP(role, conv, ev)                                                                   \
    do {                                                                                          \
        if (cp.trace && blockIdx.x == 0 && b == 0) cp.trace[((role) * 8 + (conv)) * 8 + (ev)] = clock64(); \
    } while (0)

// Per (conv, tile) a 32 KB staging tile S carries both directions of HBM traffic through TMA: the producer loads the slice
// of the NEXT chunk x_{j+2} into it, the epilogue reads its 16 channels from it and overwrites them with y_{j+1}, and warp 3
// stores the tile to HBM.  (Measured: with one row per thread every global load / store instruction touches 32 different
// cache lines, and the LSU wavefronts, not the tensor core, paced the kernel.)
template <int NSPLIT>
__global__ void __launch_bounds__(RC_THREADS, 1) res2chain_kernel(const __grid_constant__ Res2ChainParams cp) {
    using Cfg = RCCfg<NSPLIT>;
    constexpr int NP = Cfg::NP, BN = 64;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t a_base = smem_base;
    const uint32_t w_base = a_base + Cfg::A_BYTES;
    const uint32_t s_base = w_base + Cfg::W_SLOT;
    const uint32_t bar_base = s_base + 2 * Cfg::S_BYTES;
    const uint32_t x_full = bar_base, a_ready = bar_base + 8, a_free = bar_base + 16, x_free = bar_base + 24, w_full = bar_base + 32,
                   w_empty = bar_base + 40;
    auto s_full = [&](int i) { return bar_base + 48u + 8u * i; };
    auto s_empty = [&](int i) { return bar_base + 64u + 8u * i; };
    auto y_ready = [&](int i) { return bar_base + 80u + 8u * i; };
    auto tfull = [&](int t) { return bar_base + 96u + 8u * t; };
    auto tempty = [&](int t) { return bar_base + 120u + 8u * t; };
    const uint32_t tmem_slot = bar_base + 144u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&cp.mapX);
        prefetch_tmap(&cp.mapXt);
        prefetch_tmap(&cp.mapY);
        prefetch_tmap(&cp.mapYtail);
        for (int j = 0; j < cp.nconv; ++j) prefetch_tmap(&cp.mapW[j]);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(x_full, 1);
        mbar_init(a_ready, RC_EPI_THREADS);
        mbar_init(a_free, 1);
        mbar_init(x_free, 1);
        mbar_init(w_full, 1);
        mbar_init(w_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full(i), 1);
            mbar_init(s_empty(i), 1);
            mbar_init(y_ready(i), RC_EPI_THREADS);
        }
        for (int t = 0; t < RC_MAX_TILES; ++t) {
            mbar_init(tfull(t), 1);
            mbar_init(tempty(t), RC_EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);  // 3 accumulator tiles x (64 + 64) columns: [A_hi W_hi + A_lo W_hi | A_hi W_lo]
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    griddep_launch_dependents();
    const int nconv = cp.nconv, ntiles = cp.ntiles;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            griddep_wait();  // x comes from the previous kernel
            int g = 0, u = 0;
            uint32_t sn = 0;  // running staging-tile index (buffer = sn & 1)
            for (int b = blockIdx.x; b < cp.B; b += gridDim.x, ++u) {
                if (u > 0) mbar_wait(x_free, uint32_t(u - 1) & 1u);  // the last conv of the previous utterance has read the tile
                mbar_arrive_expect_tx(x_full, NP * 2 * RC_BOX_ROWS * 128);
                for (int pl = 0; pl < NP; ++pl)
                    for (int h = 0; h < 2; ++h)
                        tma_load_3d(a_base + pl * RC_A_PLANE + h * RC_BOX_ROWS * 128, &cp.mapX, x_full, cp.width /* chunk 1 */,
                                    b * cp.Tp - RC_PAD + h * RC_BOX_ROWS, pl);
                for (int j = 0; j < nconv; ++j, ++g) {
                    if (g > 0) mbar_wait(w_empty, uint32_t(g - 1) & 1u);
                    RC_STAMP(0, j, 0);
                    mbar_arrive_expect_tx(w_full, 3 * NP * RC_W_TILE);
                    for (int tap = 0; tap < 3; ++tap)
                        for (int pl = 0; pl < NP; ++pl) tma_load_3d(w_base + (tap * NP + pl) * RC_W_TILE, &cp.mapW[j], w_full, tap * 64, 0, pl);
                    for (int t = 0; t < ntiles; ++t, ++sn) {
                        const int i = sn & 1;
                        if (sn >= 2) mbar_wait(s_empty(i), ((sn >> 1) - 1) & 1u);  // the store of its previous use has read the tile
                        if (j + 1 < nconv) {  // slice of chunk j+2 that the epilogue adds to y_{j+1}
                            mbar_arrive_expect_tx(s_full(i), NP * RC_S_PLANE);
                            for (int pl = 0; pl < NP; ++pl)
                                tma_load_3d(s_base + i * Cfg::S_BYTES + pl * RC_S_PLANE, &cp.mapXt, s_full(i), (j + 2) * cp.width,
                                            b * cp.Tp + t * GEMM_BM, pl);
                        } else {
                            mbar_arrive(s_full(i));  // nothing to add after the last conv; the tile is used for the store only
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // split-bf16 product as TWO instructions per k-step (see conv3x3.cu): A_hi x [W_hi | W_lo] with N = 128 -- the hi and lo weight
        // tiles of a tap are adjacent in shared memory -- and A_lo x W_hi with N = 64; an N = 64 MMA is bound by its A-operand read,
        // so the N = 128 one costs the same.  The epilogue adds the two 64-column blocks.
        constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN), idesc2 = make_idesc_bf16(GEMM_BM, 2 * BN);
        int g = 0, u = 0;
        uint32_t aready_count = 0;
        for (int b = blockIdx.x; b < cp.B; b += gridDim.x, ++u) {
            for (int j = 0; j < nconv; ++j, ++g) {
                if (j == 0) mbar_wait(x_full, uint32_t(u) & 1u);
                mbar_wait(a_ready, aready_count & 1u);  // conv 0: reflect halo rows of the TMA-loaded tile are in place; else: operand rewritten
                ++aready_count;
                if (lane == 0) RC_STAMP(1, j, 0);  // operand ready
                mbar_wait(w_full, uint32_t(g) & 1u);
                if (lane == 0) RC_STAMP(1, j, 1);  // weights ready
                tc_fence_after();
                for (int t = 0; t < ntiles; ++t) {
                    if (g > 0) mbar_wait(tempty(t), uint32_t(g - 1) & 1u);  // epilogue drained this accumulator
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t d_tmem = tmem_base + t * 2 * BN;
                        uint32_t accumulate = 0;
#pragma unroll
                        for (int tap = 0; tap < 3; ++tap) {
                            const uint32_t roff = uint32_t(t * GEMM_BM + RC_PAD + (tap - 1) * cp.dil);
                            const uint64_t a_hi = make_sw128_kmajor_desc(a_base + roff * 128u);
                            const uint64_t b_hi = make_sw128_kmajor_desc(w_base + (tap * NP) * RC_W_TILE);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, NSPLIT == 3 ? idesc2 : idesc, accumulate);
                                accumulate = 1;
                            }
                            if (NSPLIT == 3) {
                                const uint64_t a_lo = make_sw128_kmajor_desc(a_base + RC_A_PLANE + roff * 128u);
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                            }
                        }
                        umma_commit(tfull(t));
                        RC_STAMP(1, j, 2 + t);  // tile t issued
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    umma_commit(w_empty);  // weight slot reusable
                    if (j == nconv - 1) umma_commit(x_free);
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ===================== store warp: staging tile -> HBM =====================
        if (lane == 0) {
            uint32_t sn = 0;
            for (int b = blockIdx.x; b < cp.B; b += gridDim.x) {
                for (int j = 0; j < nconv; ++j) {
                    for (int t = 0; t < ntiles; ++t, ++sn) {
                        const int i = sn & 1;
                        mbar_wait(y_ready(i), (sn >> 1) & 1u);  // all epilogue threads have written (and fenced) their part of the tile
                        const CUtensorMap* my = (t == ntiles - 1) ? &cp.mapYtail : &cp.mapY;  // never store beyond this utterance's rows
                        for (int pl = 0; pl < NP; ++pl)
                            tma_store_3d(my, s_base + i * Cfg::S_BYTES + pl * RC_S_PLANE, (j + 1) * cp.width, b * cp.Tp + t * GEMM_BM, pl);
                        bulk_commit_group();
                        bulk_wait_read0();
                        mbar_arrive(s_empty(i));
                    }
                }
            }
            bulk_wait_all();
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 16 warps, each thread one row x 16 channels =====================
        const int q = warp & 3, cq = (warp - 4) >> 2;
        const int c0 = cq * 16;
        griddep_wait();
        uint32_t g = 0, sn = 0;
        uint8_t* const a_gen = smem_gen;  // a_base == smem_base
        uint8_t* const s_gen = smem_gen + (s_base - smem_base);
        const int rloc = q * 32 + lane;  // row inside the tile == TMEM lane
        uint32_t ucount = 0;
        for (int b = blockIdx.x; b < cp.B; b += gridDim.x, ++ucount) {
            {
                // The producing GEMM stores valid frames only (TMA-store epilogue): build the reflect halo rows of the resident
                // tile here -- 2 P rows x 8 chunks x planes, one 16-byte copy per thread, in the swizzled layout.
                mbar_wait(x_full, ucount & 1u);
                const int e = threadIdx.x - 128;
                if (e < 2 * cp.P * 8 * NP) {
                    const int pl = e / (2 * cp.P * 8), r = (e / 8) % (2 * cp.P), c = e % 8;
                    const int k = (r % cp.P) + 1;
                    const int dst = (r < cp.P ? cp.P - k : cp.P + cp.T - 1 + k) + RC_PAD, src = (r < cp.P ? cp.P + k : cp.P + cp.T - 1 - k) + RC_PAD;
                    const uint4 val = *reinterpret_cast<const uint4*>(a_gen + pl * RC_A_PLANE + src * 128 + ((c ^ (src & 7)) << 4));
                    *reinterpret_cast<uint4*>(a_gen + pl * RC_A_PLANE + dst * 128 + ((c ^ (dst & 7)) << 4)) = val;
                }
                fence_proxy_async_smem();
                mbar_arrive(a_ready);
            }
            for (int j = 0; j < nconv; ++j, ++g) {
                const float* bias = cp.bias[j] + c0;
                const float* bsc = cp.bn_scale[j] + c0;
                const float* bsh = cp.bn_shift[j] + c0;
                const bool has_next = j + 1 < nconv;
                for (int t = 0; t < ntiles; ++t, ++sn) {
                    const int p = t * GEMM_BM + rloc;  // padded row inside the utterance
                    const int tt = p - cp.P;
                    const bool valid = tt >= 0 && tt < cp.T;
                    const int i = sn & 1;
                    uint8_t* const srow = s_gen + i * Cfg::S_BYTES + rloc * 128;
                    const int ch0 = ((cq * 2) ^ (rloc & 7)) << 4, ch1 = ((cq * 2 + 1) ^ (rloc & 7)) << 4;  // SWIZZLE_128B chunk offsets
                    mbar_wait(s_full(i), (sn >> 1) & 1u);
                    uint4 nh[2], nl[2];
                    if (has_next) {
                        nh[0] = *reinterpret_cast<const uint4*>(srow + ch0);
                        nh[1] = *reinterpret_cast<const uint4*>(srow + ch1);
                        if (NP == 2) {
                            nl[0] = *reinterpret_cast<const uint4*>(srow + RC_S_PLANE + ch0);
                            nl[1] = *reinterpret_cast<const uint4*>(srow + RC_S_PLANE + ch1);
                        } else {
                            nl[0] = nl[1] = make_uint4(0, 0, 0, 0);
                        }
                    }
                    mbar_wait(tfull(t), g & 1u);
                    if (threadIdx.x == 128) RC_STAMP(2, j, t);  // accumulator t complete
                    tc_fence_after();
                    uint32_t v[16];
                    __syncwarp();
                    tmem_ld16(tmem_base + t * 2 * BN + c0 + (uint32_t(q * 32) << 16), v);
                    if (NSPLIT == 3) {  // + the A_hi W_lo block
                        uint32_t v2[16];
                        tmem_ld16(tmem_base + t * 2 * BN + BN + c0 + (uint32_t(q * 32) << 16), v2);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(v2[k]));
                    }
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(tempty(t));  // values are in registers: the accumulator may be reused
                    float x[16];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // bias -> ReLU -> BatchNorm(eval) affine   (TDNNBlock, utils.py:147)
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + k);
                        const float4 s4 = __ldg(reinterpret_cast<const float4*>(bsc) + k);
                        const float4 h4 = __ldg(reinterpret_cast<const float4*>(bsh) + k);
                        x[4 * k + 0] = fmaf(fmaxf(__uint_as_float(v[4 * k + 0]) + b4.x, 0.f), s4.x, h4.x);
                        x[4 * k + 1] = fmaf(fmaxf(__uint_as_float(v[4 * k + 1]) + b4.y, 0.f), s4.y, h4.y);
                        x[4 * k + 2] = fmaf(fmaxf(__uint_as_float(v[4 * k + 2]) + b4.z, 0.f), s4.z, h4.z);
                        x[4 * k + 3] = fmaf(fmaxf(__uint_as_float(v[4 * k + 3]) + b4.w, 0.f), s4.w, h4.w);
                    }
                    {  // y_{j+1} -> staging tile (every row: rows outside the valid frames are halo / padding rows of the y buffer)
                        uint32_t h[8], l[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) split_pack_bf16x2(x[2 * k], x[2 * k + 1], h[k], l[k]);
                        *reinterpret_cast<uint4*>(srow + ch0) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(srow + ch1) = make_uint4(h[4], h[5], h[6], h[7]);
                        if (NP == 2) {
                            *reinterpret_cast<uint4*>(srow + RC_S_PLANE + ch0) = make_uint4(l[0], l[1], l[2], l[3]);
                            *reinterpret_cast<uint4*>(srow + RC_S_PLANE + ch1) = make_uint4(l[4], l[5], l[6], l[7]);
                        }
                        fence_proxy_async_smem();
                        mbar_arrive(y_ready(i));
                    }
                    // The resident tile is rewritten IN PLACE while later tiles of this conv are still being multiplied: the MMAs of
                    // tiles <= t are complete (tfull(t) was committed after them), and those of tile t+1 read this tile's last
                    // `dil` rows only.  So only the warps holding those rows (lane quarter 3) wait for tile t+1 -- which every
                    // warp is about to wait for anyway -- and nobody waits for the whole conv.
                    if (has_next && q == 3 && t + 1 < ntiles) mbar_wait(tfull(t + 1), g & 1u);
                    if (threadIdx.x == 128) RC_STAMP(2, j, 3 + (t > 0));
                    if (has_next && valid) {  // x_{j+2} + y_{j+1} -> operand of the next conv, in place, with its reflect halo rows
                        int rows[3] = {p + RC_PAD, -1, -1};
                        if (tt >= 1 && tt <= cp.P) rows[1] = cp.P - tt + RC_PAD;
                        const int uu = cp.T - 1 - tt;
                        if (uu >= 1 && uu <= cp.P) rows[2] = cp.P + cp.T - 1 + uu + RC_PAD;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t hw[4] = {nh[k].x, nh[k].y, nh[k].z, nh[k].w}, lw[4] = {nl[k].x, nl[k].y, nl[k].z, nl[k].w};
                            uint32_t oh[4], ol[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 hf = unpack_bf16x2(hw[e]), lf = unpack_bf16x2(lw[e]);
                                split_pack_bf16x2(x[8 * k + 2 * e] + (hf.x + lf.x), x[8 * k + 2 * e + 1] + (hf.y + lf.y), oh[e], ol[e]);
                            }
#pragma unroll
                            for (int r = 0; r < 3; ++r) {
                                if (rows[r] < 0) continue;
                                const int chunk = ((cq * 2 + k) ^ (rows[r] & 7)) << 4;
                                uint8_t* dst = a_gen + rows[r] * 128 + chunk;
                                *reinterpret_cast<uint4*>(dst) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                                if (NP == 2) *reinterpret_cast<uint4*>(dst + RC_A_PLANE) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                            }
                        }
                    }
                }
                if (threadIdx.x == 128) RC_STAMP(2, j, 5);  // all tiles written back
                if (has_next) {
                    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    mbar_arrive(a_ready);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int res2chain_build(Res2ChainParams* cp, const Planes& x, const Planes& y, const Planes* W, const float* const* bias, const float* const* bn_scale,
                    const float* const* bn_shift, int nconv, int B, int T, int P, int Tp, int dil) {
    PPV_REQUIRE(nconv >= 1 && nconv <= RES2CHAIN_MAX, "res2chain: 1..7 convs");
    PPV_REQUIRE(P == RC_PAD && Tp == T + 2 * P && Tp <= RC_MAX_TP, "res2chain: padded utterance must fit 384 rows with P == 4");
    PPV_REQUIRE(dil >= 1 && dil <= RC_PAD, "res2chain: dilation must be in [1,4]");
    PPV_REQUIRE(x.ld >= (nconv + 1) * 64 && y.ld >= (nconv + 1) * 64, "res2chain: buffers narrower than the chunks");
    memset(static_cast<void*>(cp), 0, sizeof(*cp));
    int rc = encode_planes_map_ex(&cp->mapX, x, 64, RC_BOX_ROWS, 128);
    if (rc) return rc;
    const int ntiles = (Tp + GEMM_BM - 1) / GEMM_BM;
    rc = encode_planes_map_ex(&cp->mapXt, x, 64, GEMM_BM, 128);
    if (rc) return rc;
    rc = encode_planes_map_ex(&cp->mapY, y, 64, GEMM_BM, 128);
    if (rc) return rc;
    rc = encode_planes_map_ex(&cp->mapYtail, y, 64, Tp - (ntiles - 1) * GEMM_BM, 128);
    if (rc) return rc;
    for (int j = 0; j < nconv; ++j) {
        PPV_REQUIRE(W[j].ld >= 192 && W[j].rows >= 64, "res2chain: weight layout mismatch");
        rc = encode_planes_map(&cp->mapW[j], W[j], 64);
        if (rc) return rc;
        cp->bias[j] = bias[j];
        cp->bn_scale[j] = bn_scale[j];
        cp->bn_shift[j] = bn_shift[j];
    }
    cp->x = x;
    cp->y = y;
    cp->nconv = nconv;
    cp->width = 64;
    cp->B = B;
    cp->T = T;
    cp->P = P;
    cp->Tp = Tp;
    cp->dil = dil;
    cp->ntiles = ntiles;
    if (getenv("PPV_RES2_TRACE")) {  // debug: leaked on purpose, read back by res2chain_trace_dump
        static unsigned long long* buf = nullptr;
        if (!buf) {
            cudaMalloc(reinterpret_cast<void**>(&buf), 3 * 8 * 8 * sizeof(unsigned long long));
            cudaMemset(buf, 0, 3 * 8 * 8 * sizeof(unsigned long long));
        }
        cp->trace = buf;
    }
    return PPV_OK;
}

template <int NSPLIT>
static int launch_rc(const Res2ChainParams& cp, int num_sms, cudaStream_t st) {
    using Cfg = RCCfg<NSPLIT>;
    PPV_ONCE_PER_DEVICE(PPV_CUDA_OK(cudaFuncSetAttribute(res2chain_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)));
    const int grid = std::min(cp.B, num_sms);
    PPV_PDL_OK(launch_pdl(res2chain_kernel<NSPLIT>, dim3(grid), dim3(RC_THREADS), Cfg::SMEM_BYTES, st, cp), "res2chain_kernel");
    return PPV_OK;
}

int res2chain_launch(const Res2ChainParams& cp, int precision, int num_sms, cudaStream_t st) {
    return precision == PPV_PREC_BF16X3 ? launch_rc<3>(cp, num_sms, st) : launch_rc<1>(cp, num_sms, st);
}
void res2chain_trace_dump(const Res2ChainParams& cp) {
    if (!cp.trace) return;
    unsigned long long h[3 * 8 * 8];
    cudaDeviceSynchronize();
    cudaMemcpy(h, cp.trace, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (unsigned long long v : h)
        if (v && v < t0) t0 = v;
    static const char* roles[3] = {"tma", "mma", "epi"};
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 7; ++j) {
            printf("res2chain trace %s conv %d:", roles[r], j);
            for (int e = 0; e < 8; ++e) printf(" %8lld", h[(r * 8 + j) * 8 + e] ? (long long)(h[(r * 8 + j) * 8 + e] - t0) : -1ll);
            printf("\n");
        }
}
bool res2chain_fits(int T, int P) { return P == RC_PAD && T + 2 * P <= RC_MAX_TP; }

}  // namespace ppv
