// Fused attentive-statistics pooling (K5): attention logits + softmax over time + weighted mean / std + asp_bn in
// ONE kernel; the [frames, 1536] logits never exist in HBM.
// Reference: ppvector/models/pooling.py:107-123 (attn = conv(tanh(tdnn(.))); masked softmax over T;
// mean = sum(a x); std = sqrt(clip(sum(a (x - mean)^2), eps))) and ppvector/models/ecapa_tdnn.py:271 (asp_bn).
//
// The GEMM is TRANSPOSED relative to the other layers: D[channel, frame] = W2[channel, :] . att[frame, :], i.e.
// A = conv weight slab (128 channels x K=128, K-major), B = attention-TDNN activations (128 frames x K, K-major:
// the activation matrix is already laid out that way).  A TMEM lane is then a CHANNEL and its columns are FRAMES, so
// each epilogue thread owns one channel and walks the frames of the utterance sequentially: the softmax over time and
// the weighted moments are per-thread running sums -- no cross-thread reduction, no atomics, deterministic.
//   * online softmax (running max, rescale once per 32-frame chunk)
//   * moments about the channel's global mean g (from asp_global): S0 = sum e, S1 = sum e (x-g), S2 = sum e (x-g)^2,
//     mean = g + S1/S0, var = S2/S0 - (S1/S0)^2   (shifted one-pass; the reference is two-pass)
//   * the conv bias is constant over time and cancels in the softmax: it is not even loaded.
// Work item = (utterance b, 128-channel slab); a CTA keeps the weight slab in smem for the item and streams the
// utterance in 64-frame tiles through a 2-deep TMA ring.  A ring stage carries BOTH the attention activations (the
// MMA's B operand) and the matching [64 frames x 128 channels] tile of x (un-swizzled, read by the epilogue with
// plain ld.shared), so the epilogue never waits on global memory; the accumulator is double-buffered in TMEM.
#include <stdio.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace ppv {

constexpr int AF_NT = 64;                         // frames per tile (MMA N)
constexpr int AF_A_TILE = 128 * 64 * 2;           // weight tile  [128 ch x 64 k]  bf16, SWIZZLE_128B
constexpr int AF_B_TILE = AF_NT * 64 * 2;         // att tile     [64 fr x 64 k]   bf16, SWIZZLE_128B
constexpr int AF_X_TILE = AF_NT * 128 * 2;        // x tile       [64 fr x 128 ch] bf16, no swizzle (one plane)
constexpr int AF_STAGES = 2;
constexpr float LOG2E = 1.4426950408889634f;

template <int NSPLIT>
__global__ void __launch_bounds__(384, 1) asp_fused_kernel(const __grid_constant__ AspFusedParams p) {
    constexpr int NP = (NSPLIT == 3) ? 2 : 1;  // planes loaded per MMA operand
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int ksteps = p.K / 64;  // 2 for K = 128
    const int a_bytes = ksteps * NP * AF_A_TILE;
    const int b_bytes = ksteps * NP * AF_B_TILE;
    const int stage_bytes = b_bytes + 2 * AF_X_TILE;  // x always needs hi and lo
    const uint32_t a_base = smem_base;
    const uint32_t s_base = smem_base + a_bytes;
    const uint32_t bar_base = s_base + AF_STAGES * stage_bytes;
    // barriers: a_full, a_empty, b_full[2], b_empty[2], tfull[2], tempty[2], tmem slot
    const uint32_t a_full = bar_base, a_empty = bar_base + 8;
    auto b_full = [&](int s) { return bar_base + 16u + 8u * s; };
    auto b_empty = [&](int s) { return bar_base + 32u + 8u * s; };
    auto tfull = [&](int a) { return bar_base + 48u + 8u * a; };
    auto tempty = [&](int a) { return bar_base + 64u + 8u * a; };
    const uint32_t tmem_slot = bar_base + 80u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapW);
        prefetch_tmap(&p.mapAtt);
        prefetch_tmap(&p.mapX);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < AF_STAGES; ++s) {
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1 + 256);  // MMA commit + the 256 epilogue threads that read the x tile
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull(a), 1);
            mbar_init(tempty(a), 256);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    griddep_launch_dependents();  // PDL
    griddep_wait();

    const int slabs = p.C / 128;
    const int items = slabs * p.B;
    const int ntiles = (p.T + AF_NT - 1) / AF_NT;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0, a_phase = 0;
        int cur_slab = -1;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int b = item / slabs, slab = item - b * slabs;  // utterance-major: att / x rows are shared by neighbouring items (L2)
            if (slab != cur_slab) {  // the weight slab stays in shared memory while consecutive items share it
                mbar_wait(a_empty, a_phase ^ 1u);
                if (lane == 0) {
                    mbar_arrive_expect_tx(a_full, a_bytes);
                    for (int ks = 0; ks < ksteps; ++ks)
                        for (int pl = 0; pl < NP; ++pl)
                            tma_load_3d(a_base + (ks * NP + pl) * AF_A_TILE, &p.mapW, a_full, ks * 64, slab * 128, pl);
                }
                __syncwarp();
                a_phase ^= 1u;
                cur_slab = slab;
            }
            {  // pull the NEXT item's x tiles (HBM) into L2 while this item is processed
                const int nitem = item + gridDim.x;
                if (nitem < items && lane < 2 * ntiles) {
                    const int nb = nitem / slabs, nslab = nitem - nb * slabs;
                    tma_prefetch_l2_3d(&p.mapX, nslab * 128, nb * p.Tp + p.P + (lane >> 1) * AF_NT, lane & 1);
                }
                __syncwarp();
            }
            const int row0 = b * p.Tp + p.P;
            for (int ft = 0; ft < ntiles; ++ft) {
                mbar_wait(b_empty(stage), phase ^ 1u);
                if (lane == 0) {
                    const uint32_t sb = s_base + stage * stage_bytes;
                    mbar_arrive_expect_tx(b_full(stage), stage_bytes);
                    for (int ks = 0; ks < ksteps; ++ks)
                        for (int pl = 0; pl < NP; ++pl)
                            tma_load_3d(sb + (ks * NP + pl) * AF_B_TILE, &p.mapAtt, b_full(stage), ks * 64, row0 + ft * AF_NT, pl);
                    for (int pl = 0; pl < 2; ++pl)
                        tma_load_3d(sb + b_bytes + pl * AF_X_TILE, &p.mapX, b_full(stage), slab * 128, row0 + ft * AF_NT, pl);
                }
                __syncwarp();
                if (++stage == AF_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16(128, AF_NT);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0, a_phase = 0;
        int cur_slab = -1;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int slab = item % slabs;
            if (slab != cur_slab) {
                mbar_wait(a_full, a_phase);
                a_phase ^= 1u;
                cur_slab = slab;
            }
            const int nitem = item + gridDim.x;
            const bool last_of_slab = (nitem >= items) || (nitem % slabs != slab);
            for (int ft = 0; ft < ntiles; ++ft) {
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                mbar_wait(b_full(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d_tmem = tmem_base + acc * AF_NT;
                    const uint32_t sb = s_base + stage * stage_bytes;
                    uint32_t accumulate = 0;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t a_hi = make_sw128_kmajor_desc(a_base + (ks * NP) * AF_A_TILE);
                        const uint64_t b_hi = make_sw128_kmajor_desc(sb + (ks * NP) * AF_B_TILE);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
                            accumulate = 1;
                        }
                        if (NSPLIT == 3) {
                            const uint64_t a_lo = make_sw128_kmajor_desc(a_base + (ks * NP + 1) * AF_A_TILE);
                            const uint64_t b_lo = make_sw128_kmajor_desc(sb + (ks * NP + 1) * AF_B_TILE);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                        }
                    }
                    umma_commit(b_empty(stage));  // one of the 129 arrivals: the att tiles are consumed
                    umma_commit(tfull(acc));
                    if (ft == ntiles - 1 && last_of_slab) umma_commit(a_empty);  // weight slab free once its last item's MMAs retire
                }
                __syncwarp();
                if (++stage == AF_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one thread = one channel =====================
        // 8 warps: warps 4-7 take the first 32 frames of every 64-frame tile, warps 8-11 the second 32; the two
        // partial (max, S0, S1, S2) of a channel are merged through shared memory at the end of the item.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int cl = q * 32 + lane;  // channel within the slab == TMEM lane
        float4* s_merge = reinterpret_cast<float4*>(smem_gen + (bar_base - smem_base) + 96);  // [128]
        int acc = 0, stage = 0;
        uint32_t acc_phase = 0, phase = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int b = item / slabs, slab = item - b * slabs;
            const int c = slab * 128 + cl;
            const int64_t goff = int64_t(b) * p.gstat.ld + c;
            const float g = __bfloat162float(p.gstat.hi()[goff]) + __bfloat162float(p.gstat.lo()[goff]);
            float m = -INFINITY, S0 = 0.f, S1 = 0.f, S2 = 0.f;
            const int Tb = p.nvalid ? max(1, min(p.T, p.nvalid[b])) : p.T;  // masked frames: softmax weight exactly 0
            for (int ft = 0; ft < ntiles; ++ft) {
                mbar_wait(b_full(stage), phase);  // acquire the TMA-written x tile directly (already complete by now)
                mbar_wait(tfull(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * AF_NT;
                const __nv_bfloat16* xs_hi = reinterpret_cast<const __nv_bfloat16*>(smem_gen + (s_base - smem_base) + stage * stage_bytes + b_bytes);
                const __nv_bfloat16* xs_lo = xs_hi + AF_X_TILE / 2;
                {
                    const int ch = half;
                    const int t0 = ft * AF_NT + ch * 32;
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld32(t_addr + ch * 32, v);
                    tmem_ld_wait();
                    const int nvalid = min(32, Tb - t0);  // warp-uniform
                    if (nvalid > 0) {
                        // chunk max as a 4-way tree (the serial fmax chain is latency-bound with 2 warps per scheduler)
                        float cm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) cm4[j & 3] = fmaxf(cm4[j & 3], __uint_as_float(v[j]));
                        const float cm = fmaxf(fmaxf(cm4[0], cm4[1]), fmaxf(cm4[2], cm4[3]));
                        if (cm > m) {
                            const float r = exp2f((m - cm) * LOG2E);  // exp(-inf) = 0 on the first chunk
                            S0 *= r;
                            S1 *= r;
                            S2 *= r;
                            m = cm;
                        }
                        // four independent accumulator sets (ILP), exp as one FFMA + ex2.approx
                        const float ms = m * LOG2E;
                        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < nvalid) {
                                const int o = (ch * 32 + j) * 128 + cl;
                                const float xv = __bfloat162float(xs_hi[o]) + __bfloat162float(xs_lo[o]);
                                const float e = exp2f(fmaf(__uint_as_float(v[j]), LOG2E, -ms));
                                const float d = xv - g;
                                const float ed = e * d;
                                a0[j & 3] += e;
                                a1[j & 3] += ed;
                                a2[j & 3] = fmaf(ed, d, a2[j & 3]);
                            }
                        }
                        S0 += (a0[0] + a0[1]) + (a0[2] + a0[3]);
                        S1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
                        S2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty(acc));
                mbar_arrive(b_empty(stage));  // done reading the x tile of this stage
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
                if (++stage == AF_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            // merge the two halves of the channel
            if (half == 1) s_merge[cl] = make_float4(m, S0, S1, S2);
            named_bar_sync(2, 256);
            if (half == 0) {
                const float4 o = s_merge[cl];
                const float mm = fmaxf(m, o.x);
                const float ra = exp2f((m - mm) * LOG2E), rb = (o.y > 0.f) ? exp2f((o.x - mm) * LOG2E) : 0.f;
                S0 = S0 * ra + o.y * rb;
                S1 = S1 * ra + o.z * rb;
                S2 = S2 * ra + o.w * rb;
            }
            named_bar_sync(2, 256);  // s_merge may be overwritten by the next item
            if (half == 1) continue;
            const float inv = 1.f / S0;
            const float dm = S1 * inv;
            const float mean = g + dm;
            const float sd = sqrtf(fmaxf(S2 * inv - dm * dm, p.eps));
            const int C = p.C;
            if (p.out_raw) {
                p.out_raw[int64_t(b) * 2 * C + c] = mean;
                p.out_raw[int64_t(b) * 2 * C + C + c] = sd;
            }
            __nv_bfloat16 h, l;
            split_bf16(fmaf(mean, p.bn_scale[c], p.bn_shift[c]), h, l);
            p.out.hi()[int64_t(b) * p.out.ld + c] = h;
            p.out.lo()[int64_t(b) * p.out.ld + c] = l;
            split_bf16(fmaf(sd, p.bn_scale[C + c], p.bn_shift[C + c]), h, l);
            p.out.hi()[int64_t(b) * p.out.ld + C + c] = h;
            p.out.lo()[int64_t(b) * p.out.ld + C + c] = l;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

int asp_fused_build(AspFusedParams* p, const Planes& W, const Planes& att, const Planes& x, const Planes& gstat, const float* bn_scale,
                    const float* bn_shift, const Planes& out, float* out_raw, int B, int T, int P, int Tp, int C, int K, float eps) {
    PPV_REQUIRE(C % 128 == 0 && K % 64 == 0 && K <= 128, "asp_fused: C % 128 == 0 and K in {64,128} required");
    memset(p, 0, sizeof(*p));
    PPV_REQUIRE(x.ld == C, "asp_fused: x row length must equal C");
    int rc = encode_planes_map(&p->mapW, W, 128);
    if (rc) return rc;
    rc = encode_planes_map(&p->mapAtt, att, AF_NT);
    if (rc) return rc;
    rc = encode_planes_map_ex(&p->mapX, x, 128, AF_NT, 0);
    if (rc) return rc;
    p->x = x;
    p->gstat = gstat;
    p->bn_scale = bn_scale;
    p->bn_shift = bn_shift;
    p->out = out;
    p->out_raw = out_raw;
    p->B = B;
    p->T = T;
    p->P = P;
    p->Tp = Tp;
    p->C = C;
    p->K = K;
    p->eps = eps;
    p->nvalid = nullptr;
    return PPV_OK;
}

int asp_fused_launch(const AspFusedParams& p, int precision, int num_sms, cudaStream_t st) {
    const bool x3 = precision == PPV_PREC_BF16X3;
    const int ksteps = p.K / 64, np = x3 ? 2 : 1;
    const int smem = 1024 + ksteps * np * AF_A_TILE + AF_STAGES * (ksteps * np * AF_B_TILE + 2 * AF_X_TILE) + 96 + 128 * 16 + 64;
    static bool attr3 = false, attr1 = false;
    bool& attr = x3 ? attr3 : attr1;
    if (!attr) {
        if (x3)
            PPV_CUDA_OK(cudaFuncSetAttribute(asp_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else
            PPV_CUDA_OK(cudaFuncSetAttribute(asp_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    const int items = (p.C / 128) * p.B;
    const int grid = std::min(items, num_sms);
    if (x3)
        PPV_PDL_OK(launch_pdl(asp_fused_kernel<3>, dim3(grid), dim3(384), size_t(smem), st, p), "asp_fused_kernel");
    else
        PPV_PDL_OK(launch_pdl(asp_fused_kernel<1>, dim3(grid), dim3(384), size_t(smem), st, p), "asp_fused_kernel");
    return PPV_OK;
}

}  // namespace ppv
