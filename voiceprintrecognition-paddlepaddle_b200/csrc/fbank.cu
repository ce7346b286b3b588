// Fused Kaldi Fbank front end (K1): framing -> DC removal -> pre-emphasis -> povey window -> zero-pad ->
// 512-point real FFT (one 256-point complex FFT) -> power -> sparse mel (501 non-zeros for 80 bins) -> log,
// then CMN over time (+ optional tail mask).
// Reference: ppvector/data_utils/featurizer.py:88-101 (KaldiFbank -> paddleaudio.compliance.kaldi.fbank,
// algorithm as torchaudio/compliance/kaldi.py:_get_window / fbank) and featurizer.py:43-59.
//
// fbank_logmel_kernel (round 2).  Persistent CTAs walk work items of 16 consecutive frames of one utterance.
//   * The item's waveform segment (15 * shift + win samples, 11 KB) arrives by ONE cp.async.bulk (TMA 1-D) on an mbarrier,
//     double-buffered: the copy of item i+2 is in flight while item i is in the butterflies.  HBM sees each sample ~1.09 times
//     (frames overlap 2.5x inside an item; only the item seams are re-read, from L2).
//   * 16 lanes own one frame (a warp = 2 frames): the 256-point complex FFT is two radix-16 passes held ENTIRELY in registers
//     (16 complex values per lane, radix-4 x radix-4 with constant twiddles) with one transpose through padded shared memory in
//     between -- no shuffle butterflies, no bit-reversal pass.  (Round 1 put one warp on one frame: 80 shuffles per lane
//     per frame plus a bank-conflicting bit-reversed round trip made it shuffle / LSU-issue bound at 4.6 % of the HBM roofline.)
//   * log-mel rows of the item are staged in shared memory and leave as 16-byte coalesced stores, together with the item's
//     column sums, so the separate mean pass over [B,T,F] is gone: fbank_finalize reads the partial sums.
#include <math.h>

#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace ppv {

constexpr int FB_NFFT = 512;          // padded window (round_to_power_of_two)
constexpr int FB_HALF = FB_NFFT / 2;  // complex FFT length
constexpr int FB_ITEM = 16;           // frames per work item
constexpr int FB_THREADS = 256;       // 8 warps x 2 frames
constexpr int FB_SCR = 17 * 16;       // float2 scratch per frame: [k1][n2] padded to 17 columns
constexpr int FB_SCR_PAIR = 2 * FB_SCR + 8;  // the two frames of a warp: the odd one starts 64 B (16 banks) past the even one's end
constexpr int FB_MAX_MELS = 128;
constexpr int FB_PART_ROWS = 32768;   // handle-owned partial-sum rows ([row][FB_MAX_MELS]): utterances x items per launch group
constexpr int FB_FIN_FRAMES = 32;     // frames per finalize block
constexpr int FB_TW512 = 144;         // e^{-2 pi i k / 512} is needed for k <= 143 only (bins k and 256 - k share a pair)

struct FbankTables {
    float* window = nullptr;    // [512] zero-padded
    float2* tw = nullptr;       // [16][16]  exp(-2 pi i k1 n2 / 256) at [k1*16 + n2]
    float2* tw512 = nullptr;    // [256]  exp(-2 pi i k / 512)
    float* mel_w = nullptr;     // [nnz]  (x 0.25: the kernel's spectrum is 2 X)
    int* mel_start = nullptr;   // [n_mels] first FFT bin
    int* mel_len = nullptr;     // [n_mels]
    int* mel_off = nullptr;     // [n_mels] offset into mel_w
    int nnz = 0;
};

struct Fbank {
    ppv_fbank_cfg cfg;
    int win = 0, shift = 0;
    FbankTables tb;
    float* part = nullptr;   // [FB_PART_ROWS, FB_MAX_MELS] per-item column sums
    float* part2 = nullptr;  // [FB_PART_ROWS / 64, FB_MAX_MELS] per-utterance sums (long utterances)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// a * exp(-2 pi i E / 16), E a compile-time constant: the trivial rotations cost nothing / two adds
template <int E>
__device__ __forceinline__ float2 rot16(float2 a) {
    constexpr int e = E & 15;
    constexpr float R = 0.70710678118654752f, C1 = 0.92387953251128674f, S1 = 0.38268343236508977f;
    if constexpr (e == 0) return a;
    else if constexpr (e == 4) return make_float2(a.y, -a.x);
    else if constexpr (e == 8) return make_float2(-a.x, -a.y);
    else if constexpr (e == 12) return make_float2(-a.y, a.x);
    else if constexpr (e == 2) return make_float2(R * (a.x + a.y), R * (a.y - a.x));
    else if constexpr (e == 6) return make_float2(R * (a.y - a.x), -R * (a.x + a.y));
    else if constexpr (e == 10) return make_float2(-R * (a.x + a.y), R * (a.x - a.y));
    else if constexpr (e == 14) return make_float2(R * (a.x - a.y), R * (a.x + a.y));
    else {
        // (x + iy)(c - is) = (xc + ys) + i(yc - xs), c = cos(2 pi e / 16), s = sin(2 pi e / 16)
        constexpr float c = (e == 1 || e == 15) ? C1 : (e == 3 || e == 13) ? S1 : (e == 5 || e == 11) ? -S1 : -C1;  // e in {1,3,5,7,9,11,13,15}
        constexpr float s = (e == 1 || e == 7) ? S1 : (e == 3 || e == 5) ? C1 : (e == 9 || e == 15) ? -S1 : -C1;
        return make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -(a.x * s)));
    }
}

// 4-point forward DFT in place: (x0,x1,x2,x3) -> (X0,X1,X2,X3)
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = make_float2(x0.x + x2.x, x0.y + x2.y), t1 = make_float2(x0.x - x2.x, x0.y - x2.y);
    const float2 t2 = make_float2(x1.x + x3.x, x1.y + x3.y), t3 = make_float2(x1.x - x3.x, x1.y - x3.y);
    x0 = make_float2(t0.x + t2.x, t0.y + t2.y);
    x2 = make_float2(t0.x - t2.x, t0.y - t2.y);
    x1 = make_float2(t1.x + t3.y, t1.y - t3.x);  // t1 - i t3
    x3 = make_float2(t1.x - t3.y, t1.y + t3.x);  // t1 + i t3
}

// 16-point forward DFT in registers (radix 4 x 4).  Input v[n] natural order; on return slot r holds X[fb_k_of_slot(r)].
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);  // slot 4c+b = sum_a x[4a+b] W4^{ac}
    v[5] = rot16<1>(v[5]);
    v[6] = rot16<2>(v[6]);
    v[7] = rot16<3>(v[7]);
    v[9] = rot16<2>(v[9]);
    v[10] = rot16<4>(v[10]);
    v[11] = rot16<6>(v[11]);
    v[13] = rot16<3>(v[13]);
    v[14] = rot16<6>(v[14]);
    v[15] = rot16<9>(v[15]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);  // slot 4c+d = X[c + 4d]
}
__host__ __device__ constexpr int fb_k_of_slot(int r) { return (r >> 2) + 4 * (r & 3); }

struct FbankSmem {  // byte offsets of the dynamic shared-memory carve-up (host and device agree through this struct)
    int tw, tw512, win, melw, mstart, mlen, moff, scr, out, seg, bar, total, seg_stride;
};
__host__ __device__ inline FbankSmem fbank_smem_layout(int nnz, int n_mels, int win, int shift) {
    FbankSmem L;
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += (bytes + 15) & ~15; return at; };
    L.tw = take(256 * 8);
    L.tw512 = take(FB_TW512 * 8);
    L.win = take(FB_NFFT * 4);
    L.melw = take(nnz * 4);
    L.mstart = take(n_mels * 4);
    L.mlen = take(n_mels * 4);
    L.moff = take(n_mels * 4);
    L.scr = take((FB_ITEM / 2) * FB_SCR_PAIR * 8);
    L.out = take(2 * FB_ITEM * n_mels * 4);
    L.seg_stride = (((FB_ITEM - 1) * shift + win) * 4 + 127) & ~127;
    L.seg = take(2 * L.seg_stride);
    L.bar = take(2 * 8);
    L.total = o;
    return L;
}

// raw log-mel out_raw [B, T, n_mels] + per-item column sums part[(b * nitem + item) * FB_MAX_MELS + m]
// WIN > 0: the window length is a compile-time constant (400 for 16 kHz / 25 ms): the bounds tests of the frame load disappear.
template <bool VEC, int WIN>
__global__ void __launch_bounds__(FB_THREADS, 3)
    fbank_logmel_kernel(const float* __restrict__ wav, int B, int L, int T, int win_rt, int shift, int n_mels, float preemph,
                        float log_floor, FbankTables tb, int use_tma, float* __restrict__ out_raw, float* __restrict__ part,
                        const int* __restrict__ valid_frames) {
    extern __shared__ __align__(128) uint8_t fb_smem[];
    const int win = WIN > 0 ? WIN : win_rt;
    const FbankSmem lay = fbank_smem_layout(tb.nnz, n_mels, win, shift);
    float2* s_tw = reinterpret_cast<float2*>(fb_smem + lay.tw);
    float2* s_tw512 = reinterpret_cast<float2*>(fb_smem + lay.tw512);
    float* s_win = reinterpret_cast<float*>(fb_smem + lay.win);
    float* s_melw = reinterpret_cast<float*>(fb_smem + lay.melw);
    int* s_mstart = reinterpret_cast<int*>(fb_smem + lay.mstart);
    int* s_mlen = reinterpret_cast<int*>(fb_smem + lay.mlen);
    int* s_moff = reinterpret_cast<int*>(fb_smem + lay.moff);
    float2* s_scr = reinterpret_cast<float2*>(fb_smem + lay.scr);
    float* s_out = reinterpret_cast<float*>(fb_smem + lay.out);
    const uint32_t bar0 = smem_u32(fb_smem + lay.bar);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = lane & 15;
    const int fl = warp * 2 + (lane >> 4);  // frame slot inside the item
    const int nitem = (T + FB_ITEM - 1) / FB_ITEM;
    const int total = B * nitem;

    for (int i = tid; i < 256; i += FB_THREADS) s_tw[i] = tb.tw[i];
    for (int i = tid; i < FB_TW512; i += FB_THREADS) s_tw512[i] = tb.tw512[i];
    for (int i = tid; i < FB_NFFT; i += FB_THREADS) s_win[i] = tb.window[i];
    for (int i = tid; i < tb.nnz; i += FB_THREADS) s_melw[i] = tb.mel_w[i];
    for (int i = tid; i < n_mels; i += FB_THREADS) {
        s_mstart[i] = tb.mel_start[i];
        s_mlen[i] = tb.mel_len[i];
        s_moff[i] = tb.mel_off[i];
    }
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        fence_mbar_init();
    }
    griddep_launch_dependents();
    griddep_wait();  // the tables above are constants; the waveform / output buffers may belong to the previous step
    __syncthreads();

    auto item_geom = [&](int it, int& b, int& f0, int& nf) {
        b = it / nitem;
        f0 = (it - b * nitem) * FB_ITEM;
        nf = min(FB_ITEM, T - f0);
    };
    auto issue = [&](int it, int buf) {  // one elected thread: TMA 1-D bulk copy of the item's waveform segment
        int b, f0, nf;
        item_geom(it, b, f0, nf);
        const uint32_t bytes = uint32_t((nf - 1) * shift + win) * 4u;
        mbar_arrive_expect_tx(bar0 + 8 * buf, bytes);
        bulk_load_1d(smem_u32(fb_smem + lay.seg + buf * lay.seg_stride), wav + int64_t(b) * L + int64_t(f0) * shift, bytes, bar0 + 8 * buf);
    };
    if (use_tma && tid == 0) {
        if (int(blockIdx.x) < total) issue(blockIdx.x, 0);
        if (int(blockIdx.x + gridDim.x) < total) issue(blockIdx.x + gridDim.x, 1);
    }

    const float inv_win = 1.f / float(win);
    // the two frames of a warp sit 16 banks apart (odd slot = even slot + FB_SCR + 8 float2): their 4-byte power stores / mel reads do not
    // collide; every pair owns its own 64 B of slack, no slot overlaps another
    float2* scr = s_scr + (fl >> 1) * FB_SCR_PAIR + (fl & 1) * (FB_SCR + 8);
    int n = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x, ++n) {
        const int buf = n & 1;
        int b, f0, nf;
        item_geom(it, b, f0, nf);
        const float* seg = reinterpret_cast<const float*>(fb_smem + lay.seg + buf * lay.seg_stride);
        if (use_tma) {
            mbar_wait(bar0 + 8 * buf, (n >> 1) & 1);
        } else {  // unaligned waveforms: plain loads (a [B, L] batch with L % 4 != 0)
            const int len = (nf - 1) * shift + win;
            const float* src = wav + int64_t(b) * L + int64_t(f0) * shift;
            float* dst = const_cast<float*>(seg);
            for (int i = tid; i < len; i += FB_THREADS) dst[i] = __ldg(src + i);
            __syncthreads();
        }
        const float* s = seg + min(fl, nf - 1) * shift;  // idle slots of a short last item recompute its last frame, unseen

        // ---- load: lane q holds z[16 n1 + q] = (y[32 n1 + 2q], y[32 n1 + 2q + 1]) for n1 = 0..15 ----
        float2 v[16];
        float prev[16];
        float acc = 0.f;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int j = 32 * n1 + 2 * q;
            float a0 = 0.f, a1 = 0.f, ap = 0.f;
            if (WIN > 0 && 32 * n1 + 32 <= WIN) {  // whole row inside the window: no test at all
                if constexpr (VEC) {
                    const float2 t = *reinterpret_cast<const float2*>(s + j);
                    a0 = t.x;
                    a1 = t.y;
                } else {
                    a0 = s[j];
                    a1 = s[j + 1];
                }
                ap = s[n1 == 0 ? max(j - 1, 0) : j - 1];
            } else if (WIN > 0 && 32 * n1 >= WIN) {
                // zero padding
            } else if (j + 1 < win) {
                if constexpr (VEC) {
                    const float2 t = *reinterpret_cast<const float2*>(s + j);
                    a0 = t.x;
                    a1 = t.y;
                } else {
                    a0 = s[j];
                    a1 = s[j + 1];
                }
                ap = s[j > 0 ? j - 1 : 0];
            } else if (j < win) {
                a0 = s[j];
                ap = s[j > 0 ? j - 1 : 0];
            }
            v[n1] = make_float2(a0, a1);
            prev[n1] = ap;
            acc += a0 + a1;
        }
        // DC offset over the frame's win samples (torchaudio kaldi.py:_get_window remove_dc_offset): 16-lane sum
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        const float cdc = (1.f - preemph) * (acc * inv_win);
        // pre-emphasis + povey window: ((s_j - mu) - p (s_{j-1} - mu)) w_j = (s_j - p s_{j-1} - (1 - p) mu) w_j; s_{-1} := s_0
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            if (WIN > 0 && 32 * n1 >= WIN) continue;  // v[n1] is already (0, 0)
            const float2 w = *reinterpret_cast<const float2*>(s_win + 32 * n1 + 2 * q);  // zero beyond win
            const float y0 = (fmaf(-preemph, prev[n1], v[n1].x) - cdc) * w.x;
            const float y1 = (fmaf(-preemph, v[n1].x, v[n1].y) - cdc) * w.y;
            v[n1] = make_float2(y0, y1);
        }
        // ---- pass 1: 16-point DFT over n1 (lane = n2), twiddle W256^{n2 k1}, transpose through shared memory ----
        fft16(v);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int k1 = fb_k_of_slot(r);
            scr[k1 * 17 + q] = k1 == 0 ? v[r] : cmul(v[r], s_tw[k1 * 16 + q]);
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) v[n2] = scr[q * 17 + n2];
        // ---- pass 2: 16-point DFT over n2 (lane = k1): slot r holds Z[q + 16 k2(r)] ----
        fft16(v);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 16; ++r) scr[q + 16 * fb_k_of_slot(r)] = v[r];
        __syncwarp();
        // ---- real-FFT untangling, bins k and 256-k from the same pair: with 2F = Zk + conj Z[256-k], 2G = -i (Zk - conj Z[256-k]),
        //      2 X[k] = 2F + W^k 2G and 2 X[256-k] = conj(2F - W^k 2G), W = e^{-2 pi i / 512}.  Lane q takes k = q + 16 i, i = 0..8
        //      (i = 8 is k = 128, lane 0 only). ----
        float pk[9], pn[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int k = q + 16 * i;  // <= 143
            const float2 zk = scr[k];
            const float2 zn = scr[(FB_HALF - k) & (FB_HALF - 1)];
            const float2 f = make_float2(zk.x + zn.x, zk.y - zn.y);
            const float2 g = make_float2(zk.y + zn.y, zn.x - zk.x);  // -i (zk - conj zn)
            const float2 wg = cmul(s_tw512[k], g);
            const float ar = f.x + wg.x, ai = f.y + wg.y, br = f.x - wg.x, bi = f.y - wg.y;
            pk[i] = fmaf(ar, ar, ai * ai);
            pn[i] = fmaf(br, br, bi * bi);
        }
        __syncwarp();
        float* ps = reinterpret_cast<float*>(scr);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int k = q + 16 * i;
            if (i < 8 || q == 0) {  // i = 8: only k = 128 is new (the pairs (128 + q, 128 - q) belong to lane 16 - q at i = 7)
                ps[k] = pk[i];
                if (k > 0 && k < FB_HALF / 2) ps[FB_HALF - k] = pn[i];  // k = 0: the partner is the Nyquist bin, unused by the mel banks
            }
        }
        __syncwarp();
        // ---- sparse mel + log into the staging tile ----
        float* so = s_out + buf * (FB_ITEM * n_mels);
        if (fl < nf) {
            for (int m = q; m < n_mels; m += 16) {
                const float* w = s_melw + s_moff[m];
                const float* p = ps + s_mstart[m];
                const int len = s_mlen[m];  // a multiple of 4 (rows are padded with zero weights)
                float e0 = 0.f, e1 = 0.f;
                for (int j = 0; j < len; j += 4) {
                    e0 = fmaf(w[j], p[j], e0);
                    e1 = fmaf(w[j + 1], p[j + 1], e1);
                    e0 = fmaf(w[j + 2], p[j + 2], e0);
                    e1 = fmaf(w[j + 3], p[j + 3], e1);
                }
                so[fl * n_mels + m] = logf(fmaxf(e0 + e1, log_floor));
            }
        }
        __syncthreads();  // staging tile complete; every warp is done with this item's waveform segment and scratch
        if (use_tma && tid == 0) {
            const int nxt = it + 2 * int(gridDim.x);
            if (nxt < total) issue(nxt, buf);
        }
        // ---- coalesced store of the item's rows + its column sums ----
        float* dst = out_raw + (int64_t(b) * T + f0) * n_mels;
        const int cnt = nf * n_mels;
        if (((reinterpret_cast<uintptr_t>(dst) | uintptr_t(cnt * 4)) & 15) == 0) {
            for (int i = tid; i < cnt / 4; i += FB_THREADS) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(so)[i];
        } else {
            for (int i = tid; i < cnt; i += FB_THREADS) dst[i] = so[i];
        }
        if (tid < n_mels) {
            // ragged batches: frames beyond the utterance's own count are padding and stay out of its time mean
            const int nsum = valid_frames ? max(0, min(nf, valid_frames[b] - f0)) : nf;
            float sum = 0.f;
            for (int f = 0; f < nsum; ++f) sum += so[f * n_mels + tid];
            part[int64_t(it) * FB_MAX_MELS + tid] = sum;
        }
    }
}

// long utterances: fold the per-item sums of each utterance into one row (nitem -> 1)
__global__ void __launch_bounds__(128) fbank_part_reduce_kernel(const float* __restrict__ part, int nitem, float* __restrict__ out) {
    griddep_launch_dependents();
    griddep_wait();
    const int b = blockIdx.x, f = threadIdx.x;
    float s = 0.f;
    for (int i = 0; i < nitem; ++i) s += part[(int64_t(b) * nitem + i) * FB_MAX_MELS + f];
    out[int64_t(b) * FB_MAX_MELS + f] = s;
}

// CMN + tail mask; writes fp32 [B,T,F] (out_f32, may alias raw) and/or split planes in the padded layout.
// mean[b][f] = (sum over the utterance's nsum partial rows) / T.
__global__ void __launch_bounds__(256)
    fbank_finalize_kernel(const float* __restrict__ raw, const float* __restrict__ part, int nsum, const float* __restrict__ lens_ratio,
                          const int* __restrict__ valid_frames, int B, int T, int F, float* out_f32, Planes out_pl, int P, int Tp) {
    __shared__ float s_mu[FB_MAX_MELS];
    griddep_launch_dependents();
    griddep_wait();
    const int b = blockIdx.y;
    if (threadIdx.x < FB_MAX_MELS) {
        float s = 0.f;
        if (int(threadIdx.x) < F)
            for (int i = 0; i < nsum; ++i) s += part[(int64_t(b) * nsum + i) * FB_MAX_MELS + threadIdx.x];
        s_mu[threadIdx.x] = s / float(valid_frames ? max(1, min(T, valid_frames[b])) : T);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int keep_n = T;
    if (lens_ratio) keep_n = int(lens_ratio[b] * float(T));  // featurizer.py:51: (ratio * T).astype(int32)
    if (valid_frames) keep_n = min(keep_n, valid_frames[b]);
    for (int t = blockIdx.x * FB_FIN_FRAMES + (threadIdx.x >> 5); t < min(T, (int(blockIdx.x) + 1) * FB_FIN_FRAMES); t += 8) {
        const int64_t frame = int64_t(b) * T + t;
        const bool keep = t < keep_n;
        const float* src = raw + frame * F;
        if (out_f32) {
            for (int f = lane; f < F; f += 32) out_f32[frame * F + f] = keep ? src[f] - s_mu[f] : 0.f;
        }
        if (out_pl.base) {
            const int64_t row = int64_t(b) * Tp + P + t;
            int64_t rows[3] = {row, -1, -1};
            if (t >= 1 && t <= P) rows[1] = row - 2 * t;
            const int u = T - 1 - t;
            if (u >= 1 && u <= P) rows[2] = row + 2 * u;
            for (int c = 2 * lane; c < out_pl.ld; c += 64) {
                const float a = (c < F && keep) ? src[c] - s_mu[c] : 0.f;
                const float bb = (c + 1 < F && keep) ? src[c + 1] - s_mu[c + 1] : 0.f;
                uint32_t hi, lo;
                split_pack_bf16x2(a, bb, hi, lo);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (rows[k] >= 0) {
                        *reinterpret_cast<uint32_t*>(out_pl.hi() + rows[k] * out_pl.ld + c) = hi;
                        *reinterpret_cast<uint32_t*>(out_pl.lo() + rows[k] * out_pl.ld + c) = lo;
                    }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
static int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
    PPV_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), v.size() * sizeof(T)));
    PPV_CUDA_OK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PPV_OK;
}

int fbank_create(const ppv_fbank_cfg* cfg, Fbank** out) {
    PPV_REQUIRE(cfg && out, "fbank_create: null argument");
    Fbank* h = new Fbank();
    h->cfg = *cfg;
    h->win = int(cfg->sample_rate * cfg->frame_length_ms * 0.001f);
    h->shift = int(cfg->sample_rate * cfg->frame_shift_ms * 0.001f);
    if (next_pow2(h->win) != FB_NFFT || h->shift <= 0 || h->win < 2) {
        delete h;
        return fail(PPV_EUNSUPPORTED, "fbank: only window sizes that pad to a 512-point FFT are implemented (16 kHz, 25 ms)");
    }
    if (cfg->n_mels < 4 || cfg->n_mels > FB_MAX_MELS) {
        delete h;
        return fail(PPV_EUNSUPPORTED, "fbank: n_mels must be in [4,128]");
    }
    const int win = h->win;
    // povey window: hann(periodic=False)^0.85
    std::vector<float> window(FB_NFFT, 0.f);  // zero beyond win: the kernel multiplies all 512 slots
    for (int i = 0; i < win; ++i) {
        const double hann = 0.5 - 0.5 * cos(2.0 * M_PI * i / (win - 1));
        window[i] = float(pow(hann, 0.85));
    }
    std::vector<float2> tw(256), tw512(256);
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n2 = 0; n2 < 16; ++n2)
            tw[k1 * 16 + n2] = make_float2(float(cos(-2.0 * M_PI * (k1 * n2) / 256.0)), float(sin(-2.0 * M_PI * (k1 * n2) / 256.0)));
    for (int k = 0; k < 256; ++k) tw512[k] = make_float2(float(cos(-2.0 * M_PI * k / 512.0)), float(sin(-2.0 * M_PI * k / 512.0)));
    // mel banks, evaluated in float32 like torchaudio kaldi.py:get_mel_banks (vtln_warp == 1)
    const int nb = cfg->n_mels;
    const float sr = float(cfg->sample_rate);
    const float nyq = 0.5f * sr;
    float high = cfg->high_freq;
    if (high <= 0.f) high += nyq;
    const float low = cfg->low_freq;
    if (!(low >= 0.f && low < nyq && high > 0.f && high <= nyq && low < high)) {
        delete h;
        return fail(PPV_EINVAL, "fbank: bad low_freq / high_freq");
    }
    const double mel_low = 1127.0 * log(1.0 + double(low) / 700.0);
    const double mel_high = 1127.0 * log(1.0 + double(high) / 700.0);
    const double delta = (mel_high - mel_low) / (nb + 1);
    const float bin_width = sr / float(FB_NFFT);
    std::vector<float> melw;
    std::vector<int> mstart(nb), mlen(nb), moff(nb);
    for (int m = 0; m < nb; ++m) {
        const float left = float(mel_low) + float(m) * float(delta);
        const float center = float(mel_low) + (float(m) + 1.0f) * float(delta);
        const float right = float(mel_low) + (float(m) + 2.0f) * float(delta);
        int first = -1, last = -1;
        std::vector<float> row(FB_HALF, 0.f);
        for (int k = 0; k < FB_HALF; ++k) {
            const float mel = 1127.0f * logf(1.0f + (bin_width * float(k)) / 700.0f);
            const float up = (mel - left) / (center - left);
            const float down = (right - mel) / (right - center);
            const float w = fmaxf(0.f, fminf(up, down));
            row[k] = w;
            if (w > 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        mstart[m] = first < 0 ? 0 : first;
        int len = first < 0 ? 0 : last - first + 1;
        len = (len + 3) & ~3;  // the kernel's mel loop is unrolled by 4: pad with zero weights over valid power bins
        if (mstart[m] + len > FB_HALF) mstart[m] = FB_HALF - len;
        mlen[m] = len;
        moff[m] = int(melw.size());
        for (int k = 0; k < len; ++k) melw.push_back(0.25f * row[mstart[m] + k]);  // the kernel's power spectrum is |2X|^2
    }
    if (melw.empty()) melw.push_back(0.f);
    h->tb.nnz = int(melw.size());
    int rc = upload(&h->tb.window, window);
    if (!rc) rc = upload(&h->tb.tw, tw);
    if (!rc) rc = upload(&h->tb.tw512, tw512);
    if (!rc) rc = upload(&h->tb.mel_w, melw);
    if (!rc) rc = upload(&h->tb.mel_start, mstart);
    if (!rc) rc = upload(&h->tb.mel_len, mlen);
    if (!rc) rc = upload(&h->tb.mel_off, moff);
    if (!rc && (cudaMalloc(reinterpret_cast<void**>(&h->part), size_t(FB_PART_ROWS) * FB_MAX_MELS * sizeof(float)) != cudaSuccess ||
                cudaMalloc(reinterpret_cast<void**>(&h->part2), size_t(FB_PART_ROWS / 64) * FB_MAX_MELS * sizeof(float)) != cudaSuccess))
        rc = fail(PPV_ECUDA, "fbank: cudaMalloc(partial sums) failed");
    if (rc) {
        delete h;
        return rc;
    }
    *out = h;
    return PPV_OK;
}

void fbank_destroy(Fbank* h) {
    if (!h) return;
    cudaFree(h->tb.window);
    cudaFree(h->tb.tw);
    cudaFree(h->tb.tw512);
    cudaFree(h->tb.mel_w);
    cudaFree(h->tb.mel_start);
    cudaFree(h->tb.mel_len);
    cudaFree(h->tb.mel_off);
    cudaFree(h->part);
    cudaFree(h->part2);
    delete h;
}

int fbank_num_frames(const Fbank* h, int L) {
    if (L < h->win) return 0;
    return 1 + (L - h->win) / h->shift;
}
int fbank_n_mels(const Fbank* h) { return h->cfg.n_mels; }

// raw: scratch [B,T,F] (may equal out_f32).  Exactly one or both of out_f32 / out_pl.
int fbank_run(Fbank* h, const float* wav, const float* lens_ratio, int B, int L, float* raw, float* out_f32,
              const Planes& out_pl, int P, int Tp, cudaStream_t st, const int* valid_frames) {
    PPV_REQUIRE(h && wav && raw, "fbank_run: null argument");
    PPV_REQUIRE(B > 0, "fbank_run: empty batch");
    const int T = fbank_num_frames(h, L);
    PPV_REQUIRE(T > 0, "fbank_run: waveform shorter than one frame");
    const int F = h->cfg.n_mels;
    const FbankSmem lay = fbank_smem_layout(h->tb.nnz, F, h->win, h->shift);
    PPV_REQUIRE(lay.total <= 200 * 1024, "fbank_run: shared memory budget exceeded (frame shift too large)");
    // three CTAs per SM need <= 76 800 B each (228 KB - 1 KB reserved per CTA): 76 784 B for 16 kHz / 25 ms / 10 ms / 80 bins
    const bool vec = (h->shift % 2) == 0;
    auto kern = (h->win == 400) ? (vec ? fbank_logmel_kernel<true, 400> : fbank_logmel_kernel<false, 400>)
                                : (vec ? fbank_logmel_kernel<true, 0> : fbank_logmel_kernel<false, 0>);
    PPV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total));  // per device: cheap, set every call
    // TMA 1-D bulk copies need 16-byte aligned sources and sizes: every item starts at b * L + 16 k * shift samples
    const int use_tma = (reinterpret_cast<uintptr_t>(wav) % 16 == 0) && (L % 4 == 0) && ((FB_ITEM * h->shift) % 4 == 0) &&
                        (h->shift % 4 == 0) && (h->win % 4 == 0);
    const int nitem = (T + FB_ITEM - 1) / FB_ITEM;
    PPV_REQUIRE(nitem <= FB_PART_ROWS, "fbank_run: utterance too long (more than 524288 frames)");
    const int group = std::max(1, std::min(B, FB_PART_ROWS / nitem));  // utterances per launch group (partial-sum rows)
    const int sms = device_sm_count();
    for (int b0 = 0; b0 < B; b0 += group) {
        const int nb = std::min(group, B - b0);
        const float* w = wav + int64_t(b0) * L;
        float* r = raw + int64_t(b0) * T * F;
        const int grid = std::min(nb * nitem, 3 * sms);
        PPV_PDL_OK(launch_pdl(kern, dim3(grid), dim3(FB_THREADS), size_t(lay.total), st, w, nb, L, T, h->win, h->shift, F, h->cfg.preemph,
                              h->cfg.log_floor, h->tb, use_tma, r, h->part, valid_frames ? valid_frames + b0 : (const int*)nullptr),
                   "fbank_logmel_kernel");
        const float* sums = h->part;
        int nsum = nitem;
        if (nitem > 64 && nb <= FB_PART_ROWS / 64) {  // long utterances: one row of sums per utterance instead of nitem per finalize block
            PPV_PDL_OK(launch_pdl(fbank_part_reduce_kernel, dim3(nb), dim3(FB_MAX_MELS), 0, st, (const float*)h->part, nitem, h->part2),
                       "fbank_part_reduce_kernel");
            sums = h->part2;
            nsum = 1;
        }
        Planes pl = out_pl;
        if (pl.base) {  // rows of this group start at b0 * Tp
            pl.base += int64_t(b0) * Tp * pl.ld;
        }
        PPV_PDL_OK(launch_pdl(fbank_finalize_kernel, dim3((T + FB_FIN_FRAMES - 1) / FB_FIN_FRAMES, nb), dim3(256), 0, st, (const float*)r, sums,
                              nsum, lens_ratio ? lens_ratio + b0 : (const float*)nullptr, valid_frames ? valid_frames + b0 : (const int*)nullptr, nb, T, F,
                              out_f32 ? out_f32 + int64_t(b0) * T * F : (float*)nullptr, pl, P, Tp),
                   "fbank_finalize_kernel");
    }
    return PPV_OK;
}

}  // namespace ppv
