// Batched cosine scoring (K11).
// Reference: ppvector/predict.py:279-283 (contrast: dot / (|a||b|)), predict.py:173-187 (retrieval through
// sklearn cosine_similarity), ppvector/trainer.py:416-423 (eval: one 1-vs-all-enrol cosine row per trial in a
// Python loop).  Two forms:
//   all-pairs  [M,D] x [N,D] -> [M,N] : rows are L2-normalised into split-bf16 planes, then the same tcgen05
//              gather-GEMM as the model (K = D, fp32 out).  Tensor-core bound, 384 FLOP / pair.
//   pair-list  idx[P,2] -> [P]        : 8 lanes per pair, gather both rows (1536 B / pair), HBM/L2-bound.
#include "common.h"
#include "ptx.cuh"

namespace ppv {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// one warp per row: x / max(|x|, tiny) -> planes [rows_pad, Dp] (columns >= D zero)
__global__ void normalize_rows_kernel(const float* __restrict__ X, int rows, int D, Planes out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = X + int64_t(row) * D;
    float ss = 0.f;
    for (int i = lane; i < D; i += 32) ss = fmaf(x[i], x[i], ss);
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-30f);
    for (int i = lane; i < out.ld; i += 32) {
        __nv_bfloat16 h, l;
        split_bf16(i < D ? x[i] * inv : 0.f, h, l);
        out.hi()[int64_t(row) * out.ld + i] = h;
        out.lo()[int64_t(row) * out.ld + i] = l;
    }
}

size_t cosine_workspace_bytes(int M, int N, int D) {
    const size_t Dp = align_up(size_t(D), 64);
    const size_t a = align_up(size_t(M), 128) * Dp * 2 * sizeof(__nv_bfloat16);
    const size_t b = align_up(size_t(N), 128) * Dp * 2 * sizeof(__nv_bfloat16);
    return align_up(a, 256) + align_up(b, 256) + 256;
}

int cosine_matrix(const float* A, const float* Bm, int M, int N, int D, float* out, void* ws, size_t ws_bytes, int precision,
                  cudaStream_t st) {
    PPV_REQUIRE(A && Bm && out && ws, "cosine_matrix: null argument");
    PPV_REQUIRE(M > 0 && N > 0 && D > 0, "cosine_matrix: empty input");
    PPV_REQUIRE(ws_bytes >= cosine_workspace_bytes(M, N, D), "cosine_matrix: workspace too small");
    PPV_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "cosine_matrix: workspace must be 256-byte aligned");
    const int Dp = int(align_up(size_t(D), 64));
    Planes pa, pb;
    pa.rows = int64_t(align_up(size_t(M), 128));
    pa.ld = Dp;
    pa.plane_stride = pa.rows * Dp;
    pa.base = static_cast<__nv_bfloat16*>(ws);
    pb.rows = int64_t(align_up(size_t(N), 128));
    pb.ld = Dp;
    pb.plane_stride = pb.rows * Dp;
    pb.base = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(ws) + align_up(size_t(pa.plane_stride) * 4, 256));
    // rows beyond M / N must be finite (they only feed masked outputs, but keep them zero)
    PPV_CUDA_OK(cudaMemsetAsync(ws, 0, cosine_workspace_bytes(M, N, D), st));
    normalize_rows_kernel<<<(M + 7) / 8, 256, 0, st>>>(A, M, D, pa);
    PPV_LAUNCH_OK("normalize_rows_kernel(A)");
    normalize_rows_kernel<<<(N + 7) / 8, 256, 0, st>>>(Bm, N, D, pb);
    PPV_LAUNCH_OK("normalize_rows_kernel(B)");
    GemmSource src{pa, 0, Dp, 0};
    Epilogue ep;
    ep.out_mode = OUT_F32;
    ep.out = out;
    ep.out_ld = N;
    GemmParams gp;
    const int BN = 128;
    int rc = gemm_build(&gp, &src, 1, pb, M, N, ep, BN);
    if (rc) return rc;
    return gemm_launch(gp, BN, precision, device_sm_count(), st);
}

// pair list: 8 lanes per pair, float4 loads
__global__ void __launch_bounds__(256)
    cosine_pairlist_kernel(const float* __restrict__ E, const int32_t* __restrict__ idx, int64_t P, int n, int D, float* __restrict__ out) {
    const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t pair = gid >> 3;
    const int sub = int(gid & 7);
    float dot = 0.f, na = 0.f, nb = 0.f;
    const bool ok = pair < P;
    if (ok) {
        const int ia = idx[2 * pair], ib = idx[2 * pair + 1];
        if (ia >= 0 && ia < n && ib >= 0 && ib < n) {
            const float* a = E + int64_t(ia) * D;
            const float* b = E + int64_t(ib) * D;
            if ((D & 3) == 0) {
                const float4* a4 = reinterpret_cast<const float4*>(a);
                const float4* b4 = reinterpret_cast<const float4*>(b);
                for (int i = sub; i < D / 4; i += 8) {
                    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
                    dot += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
                    na += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                    nb += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
                }
            } else {
                for (int i = sub; i < D; i += 8) {
                    const float x = a[i], y = b[i];
                    dot += x * y;
                    na += x * x;
                    nb += y * y;
                }
            }
        } else {
            dot = NAN;
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        na += __shfl_xor_sync(0xffffffffu, na, o);
        nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if (ok && sub == 0) out[pair] = dot / (sqrtf(na) * sqrtf(nb));
}

int cosine_pairlist(const float* E, const int32_t* idx, int64_t P, int n, int D, float* out, cudaStream_t st) {
    if (P == 0) return PPV_OK;
    PPV_REQUIRE(E && idx && out, "cosine_pairlist: null argument");
    PPV_REQUIRE(P > 0 && n > 0 && D > 0, "cosine_pairlist: bad sizes");
    const int64_t threads = P * 8;
    cosine_pairlist_kernel<<<unsigned((threads + 255) / 256), 256, 0, st>>>(E, idx, P, n, D, out);
    PPV_LAUNCH_OK("cosine_pairlist_kernel");
    return PPV_OK;
}

}  // namespace ppv
