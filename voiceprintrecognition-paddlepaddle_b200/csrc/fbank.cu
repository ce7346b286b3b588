// Fused Kaldi Fbank front end (K1): framing -> DC removal -> pre-emphasis -> povey window -> zero-pad ->
// 512-point real FFT (256-point complex FFT with warp-shuffle butterflies) -> power -> sparse mel (501
// non-zeros for 80 bins) -> log, then CMN over time (+ optional tail mask).
// Reference: ppvector/data_utils/featurizer.py:88-101 (KaldiFbank -> paddleaudio.compliance.kaldi.fbank,
// algorithm as torchaudio/compliance/kaldi.py:_get_window / fbank) and featurizer.py:43-59.
//
// One warp owns one frame.  A block stages the contiguous waveform segment of its 32 frames in shared memory
// once (frames overlap 2.5x), so HBM sees each sample once: 192 000 B in + 95 360 B out per 3 s utterance.
#include <math.h>

#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace ppv {

constexpr int FB_NFFT = 512;          // padded window (round_to_power_of_two)
constexpr int FB_HALF = FB_NFFT / 2;  // complex FFT length
constexpr int FB_FRAMES_PER_BLOCK = 32;
constexpr int FB_WARPS = 8;
constexpr int FB_ZPAD = FB_HALF + FB_HALF / 8;  // padded scratch: idx + (idx >> 3)
constexpr int FB_MAX_MELS = 128;
constexpr int FB_MEAN_ROWS = 4096;  // utterances per CMN chunk (handle-owned mean buffer)

struct FbankTables {
    float* window = nullptr;    // [win]
    float2* tw256 = nullptr;    // [128]  exp(-2 pi i j / 256)
    float2* tw512 = nullptr;    // [256]  exp(-2 pi i k / 512)
    float* mel_w = nullptr;     // [nnz]
    int* mel_start = nullptr;   // [n_mels] first FFT bin
    int* mel_len = nullptr;     // [n_mels]
    int* mel_off = nullptr;     // [n_mels] offset into mel_w
    int nnz = 0;
};

struct Fbank {
    ppv_fbank_cfg cfg;
    int win = 0, shift = 0;
    FbankTables tb;
    float* mean_buf = nullptr;  // [FB_MEAN_ROWS, FB_MAX_MELS]
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ int zidx(int p) { return p + (p >> 3); }

// raw log-mel: out_raw [B, T, n_mels]
__global__ void __launch_bounds__(FB_WARPS * 32)
    fbank_logmel_kernel(const float* __restrict__ wav, int L, int T, int win, int shift, int n_mels, float preemph,
                        float log_floor, FbankTables tb, float* __restrict__ out_raw) {
    extern __shared__ __align__(16) uint8_t fb_smem[];
    // carve
    float2* s_tw256 = reinterpret_cast<float2*>(fb_smem);                 // 128
    float2* s_tw512 = s_tw256 + 128;                                      // 256
    float2* s_z = s_tw512 + 256;                                          // FB_WARPS * FB_ZPAD
    float* s_p = reinterpret_cast<float*>(s_z + FB_WARPS * FB_ZPAD);      // FB_WARPS * 256
    float* s_win = s_p + FB_WARPS * FB_HALF;                              // win
    float* s_melw = s_win + win;                                          // nnz
    int* s_mstart = reinterpret_cast<int*>(s_melw + tb.nnz);              // n_mels
    int* s_mlen = s_mstart + n_mels;
    int* s_moff = s_mlen + n_mels;
    float* s_wav = reinterpret_cast<float*>(s_moff + n_mels);             // (FPB-1)*shift + win

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * FB_FRAMES_PER_BLOCK;
    const int nf = min(FB_FRAMES_PER_BLOCK, T - f0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < 128; i += blockDim.x) s_tw256[i] = tb.tw256[i];
    for (int i = tid; i < 256; i += blockDim.x) s_tw512[i] = tb.tw512[i];
    for (int i = tid; i < win; i += blockDim.x) s_win[i] = tb.window[i];
    for (int i = tid; i < tb.nnz; i += blockDim.x) s_melw[i] = tb.mel_w[i];
    for (int i = tid; i < n_mels; i += blockDim.x) {
        s_mstart[i] = tb.mel_start[i];
        s_mlen[i] = tb.mel_len[i];
        s_moff[i] = tb.mel_off[i];
    }
    griddep_launch_dependents();
    griddep_wait();  // tables above are constants; the waveform / output buffers may be shared with the previous step
    const int seg = (nf - 1) * shift + win;
    const float* src = wav + int64_t(b) * L + int64_t(f0) * shift;
    for (int i = tid; i < seg; i += blockDim.x) s_wav[i] = __ldg(src + i);
    __syncthreads();

    float2* zs = s_z + warp * FB_ZPAD;
    float* ps = s_p + warp * FB_HALF;
    const float inv_win = 1.f / float(win);

    for (int fl = warp; fl < nf; fl += FB_WARPS) {
        const float* s = s_wav + fl * shift;
        // DC offset (torchaudio kaldi.py:_get_window remove_dc_offset)
        float acc = 0.f;
        for (int j = lane; j < win; j += 32) acc += s[j];
        const float mu = warp_sum(acc) * inv_win;

        // pre-emphasis + window, packed as complex z[n] = y[2n] + i y[2n+1], n = r*32 + lane
        float2 z[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int j = 2 * (r * 32 + lane);
            float y0 = 0.f, y1 = 0.f;
            if (j < win) {
                const float c0 = s[j] - mu;
                const float cm = s[j > 0 ? j - 1 : 0] - mu;
                y0 = (c0 - preemph * cm) * s_win[j];
                if (j + 1 < win) {
                    const float c1 = s[j + 1] - mu;
                    y1 = (c1 - preemph * c0) * s_win[j + 1];
                }
            }
            z[r] = make_float2(y0, y1);
        }
        // 256-point complex FFT, decimation in frequency; index n = (r << 5) | lane.
        // stages on bits 7,6,5 live in registers
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // bit 7, half = 128, twiddle exponent n mod 128
            const float2 a = z[r], c = z[r + 4];
            z[r] = make_float2(a.x + c.x, a.y + c.y);
            z[r + 4] = cmul(make_float2(a.x - c.x, a.y - c.y), s_tw256[r * 32 + lane]);
        }
#pragma unroll
        for (int g = 0; g < 8; g += 4) {  // bit 6, half = 64, exponent 2 * (n mod 64)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float2 a = z[g + r], c = z[g + r + 2];
                z[g + r] = make_float2(a.x + c.x, a.y + c.y);
                z[g + r + 2] = cmul(make_float2(a.x - c.x, a.y - c.y), s_tw256[2 * (r * 32 + lane)]);
            }
        }
#pragma unroll
        for (int g = 0; g < 8; g += 2) {  // bit 5, half = 32, exponent 4 * lane
            const float2 a = z[g], c = z[g + 1];
            z[g] = make_float2(a.x + c.x, a.y + c.y);
            z[g + 1] = cmul(make_float2(a.x - c.x, a.y - c.y), s_tw256[4 * lane]);
        }
        // stages on lane bits 4..0: shuffle butterflies
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            const int mask = 16 >> st;
            const bool upper = (lane & mask) != 0;
            const float2 w = s_tw256[(lane & (mask - 1)) * (8 << st)];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float px = __shfl_xor_sync(0xffffffffu, z[r].x, mask);
                const float py = __shfl_xor_sync(0xffffffffu, z[r].y, mask);
                if (!upper) {
                    z[r] = make_float2(z[r].x + px, z[r].y + py);
                } else {
                    z[r] = cmul(make_float2(px - z[r].x, py - z[r].y), w);
                }
            }
        }
        // position n holds Z[bitrev8(n)]; park at the natural position, read back bit-reversed
#pragma unroll
        for (int r = 0; r < 8; ++r) zs[zidx(r * 32 + lane)] = z[r];
        __syncwarp();
        // real-FFT untangling: X[k] = (Zk + conj(Z[256-k]))/2 - (i/2) e^{-2 pi i k/512} (Zk - conj(Z[256-k]))
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = i * 32 + lane;
            const int kn = (FB_HALF - k) & (FB_HALF - 1);
            const float2 zk = zs[zidx(__brev(unsigned(k)) >> 24)];
            float2 zn = zs[zidx(__brev(unsigned(kn)) >> 24)];
            zn.y = -zn.y;
            const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
            const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y - zn.y));
            const float2 wd = cmul(s_tw512[k], d);  // multiply by -i: (x, y) -> (y, -x)
            const float xr = e.x + wd.y, xi = e.y - wd.x;
            ps[k] = xr * xr + xi * xi;
        }
        __syncwarp();
        // sparse mel + log
        float* dst = out_raw + (int64_t(b) * T + f0 + fl) * n_mels;
        for (int m = lane; m < n_mels; m += 32) {
            const float* w = s_melw + s_moff[m];
            const float* p = ps + s_mstart[m];
            float e = 0.f;
            for (int j = 0; j < s_mlen[m]; ++j) e = fmaf(w[j], p[j], e);
            dst[m] = logf(fmaxf(e, log_floor));
        }
        __syncwarp();
    }
}

// column means over time: mean[b, f]
__global__ void __launch_bounds__(256) fbank_mean_kernel(const float* __restrict__ raw, int T, int F, float* __restrict__ mean) {
    __shared__ float s_part[8][FB_MAX_MELS];
    griddep_launch_dependents();
    griddep_wait();
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[FB_MAX_MELS / 32] = {0.f, 0.f, 0.f, 0.f};
    const float* base = raw + int64_t(b) * T * F;
    for (int t = warp; t < T; t += 8) {
#pragma unroll
        for (int i = 0; i < FB_MAX_MELS / 32; ++i) {
            const int f = lane + 32 * i;
            if (f < F) acc[i] += base[int64_t(t) * F + f];
        }
    }
#pragma unroll
    for (int i = 0; i < FB_MAX_MELS / 32; ++i) s_part[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    if (threadIdx.x < F) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_part[w][threadIdx.x];
        mean[int64_t(b) * FB_MAX_MELS + threadIdx.x] = s / float(T);
    }
}

// CMN + tail mask; writes fp32 [B,T,F] (out_f32, may alias raw) and/or split planes in the padded layout.
__global__ void __launch_bounds__(256)
    fbank_finalize_kernel(const float* __restrict__ raw, const float* __restrict__ mean, const float* __restrict__ lens_ratio,
                          int B, int T, int F, float* out_f32, Planes out_pl, int P, int Tp) {
    griddep_launch_dependents();
    griddep_wait();
    const int64_t frame = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (frame >= int64_t(B) * T) return;
    const int b = int(frame / T);
    const int t = int(frame - int64_t(b) * T);
    bool keep = true;
    if (lens_ratio) keep = t < int(lens_ratio[b] * float(T));  // featurizer.py:51: (ratio * T).astype(int32)
    const float* src = raw + frame * F;
    const float* mu = mean + int64_t(b) * FB_MAX_MELS;
    if (out_f32) {
        for (int f = lane; f < F; f += 32) out_f32[frame * F + f] = keep ? src[f] - mu[f] : 0.f;
    }
    if (out_pl.base) {
        const int64_t row = int64_t(b) * Tp + P + t;
        int64_t rows[3] = {row, -1, -1};
        if (t >= 1 && t <= P) rows[1] = row - 2 * t;
        const int u = T - 1 - t;
        if (u >= 1 && u <= P) rows[2] = row + 2 * u;
        for (int c = 2 * lane; c < out_pl.ld; c += 64) {
            const float a = (c < F && keep) ? src[c] - mu[c] : 0.f;
            const float bb = (c + 1 < F && keep) ? src[c + 1] - mu[c + 1] : 0.f;
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(a, h0, l0);
            split_bf16(bb, h1, l1);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (rows[k] >= 0) {
                    *reinterpret_cast<uint32_t*>(out_pl.hi() + rows[k] * out_pl.ld + c) = pack_bf16x2(h0, h1);
                    *reinterpret_cast<uint32_t*>(out_pl.lo() + rows[k] * out_pl.ld + c) = pack_bf16x2(l0, l1);
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
    std::vector<float> window(win);
    for (int i = 0; i < win; ++i) {
        const double hann = 0.5 - 0.5 * cos(2.0 * M_PI * i / (win - 1));
        window[i] = float(pow(hann, 0.85));
    }
    std::vector<float2> tw256(128), tw512(256);
    for (int j = 0; j < 128; ++j) tw256[j] = make_float2(float(cos(-2.0 * M_PI * j / 256.0)), float(sin(-2.0 * M_PI * j / 256.0)));
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
        mlen[m] = first < 0 ? 0 : last - first + 1;
        moff[m] = int(melw.size());
        for (int k = 0; k < mlen[m]; ++k) melw.push_back(row[mstart[m] + k]);
    }
    if (melw.empty()) melw.push_back(0.f);
    h->tb.nnz = int(melw.size());
    int rc = upload(&h->tb.window, window);
    if (!rc) rc = upload(&h->tb.tw256, tw256);
    if (!rc) rc = upload(&h->tb.tw512, tw512);
    if (!rc) rc = upload(&h->tb.mel_w, melw);
    if (!rc) rc = upload(&h->tb.mel_start, mstart);
    if (!rc) rc = upload(&h->tb.mel_len, mlen);
    if (!rc) rc = upload(&h->tb.mel_off, moff);
    if (!rc && cudaMalloc(reinterpret_cast<void**>(&h->mean_buf), size_t(FB_MEAN_ROWS) * FB_MAX_MELS * sizeof(float)) != cudaSuccess)
        rc = fail(PPV_ECUDA, "fbank: cudaMalloc(mean_buf) failed");
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
    cudaFree(h->tb.tw256);
    cudaFree(h->tb.tw512);
    cudaFree(h->tb.mel_w);
    cudaFree(h->tb.mel_start);
    cudaFree(h->tb.mel_len);
    cudaFree(h->tb.mel_off);
    cudaFree(h->mean_buf);
    delete h;
}

int fbank_num_frames(const Fbank* h, int L) {
    if (L < h->win) return 0;
    return 1 + (L - h->win) / h->shift;
}
int fbank_n_mels(const Fbank* h) { return h->cfg.n_mels; }

static size_t fbank_smem_bytes(const Fbank* h) {
    size_t s = 0;
    s += 128 * sizeof(float2) + 256 * sizeof(float2);
    s += size_t(FB_WARPS) * FB_ZPAD * sizeof(float2);
    s += size_t(FB_WARPS) * FB_HALF * sizeof(float);
    s += size_t(h->win) * sizeof(float);
    s += size_t(h->tb.nnz) * sizeof(float);
    s += 3 * size_t(h->cfg.n_mels) * sizeof(int);
    s += (size_t(FB_FRAMES_PER_BLOCK - 1) * h->shift + h->win) * sizeof(float);
    return s;
}

// raw: scratch [B,T,F] (may equal out_f32).  Exactly one or both of out_f32 / out_pl.
int fbank_run(Fbank* h, const float* wav, const float* lens_ratio, int B, int L, float* raw, float* out_f32,
              const Planes& out_pl, int P, int Tp, cudaStream_t st) {
    PPV_REQUIRE(h && wav && raw, "fbank_run: null argument");
    PPV_REQUIRE(B > 0, "fbank_run: empty batch");
    const int T = fbank_num_frames(h, L);
    PPV_REQUIRE(T > 0, "fbank_run: waveform shorter than one frame");
    const int F = h->cfg.n_mels;
    static bool attr_set = false;
    const size_t smem = fbank_smem_bytes(h);
    if (!attr_set) {
        PPV_CUDA_OK(cudaFuncSetAttribute(fbank_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    PPV_REQUIRE(smem <= 100 * 1024, "fbank_run: shared memory budget exceeded");
    for (int b0 = 0; b0 < B; b0 += FB_MEAN_ROWS) {
        const int nb = std::min(FB_MEAN_ROWS, B - b0);
        const float* w = wav + int64_t(b0) * L;
        float* r = raw + int64_t(b0) * T * F;
        dim3 grid((T + FB_FRAMES_PER_BLOCK - 1) / FB_FRAMES_PER_BLOCK, nb);
        PPV_PDL_OK(launch_pdl(fbank_logmel_kernel, grid, dim3(FB_WARPS * 32), smem, st, w, L, T, h->win, h->shift, F, h->cfg.preemph,
                              h->cfg.log_floor, h->tb, r),
                   "fbank_logmel_kernel");
        PPV_PDL_OK(launch_pdl(fbank_mean_kernel, dim3(nb), dim3(256), 0, st, (const float*)r, T, F, h->mean_buf), "fbank_mean_kernel");
        Planes pl = out_pl;
        if (pl.base) {  // rows of this chunk start at b0 * Tp
            pl.base += int64_t(b0) * Tp * pl.ld;
        }
        const int64_t frames = int64_t(nb) * T;
        PPV_PDL_OK(launch_pdl(fbank_finalize_kernel, dim3(unsigned((frames + 7) / 8)), dim3(256), 0, st, (const float*)r,
                              (const float*)h->mean_buf, lens_ratio ? lens_ratio + b0 : (const float*)nullptr, nb, T, F,
                              out_f32 ? out_f32 + int64_t(b0) * T * F : (float*)nullptr, pl, P, Tp),
                   "fbank_finalize_kernel");
    }
    return PPV_OK;
}

}  // namespace ppv
