// Batched waveform preparation in front of the feature extractor: speed perturbation -> volume perturbation -> additive noise at a
// drawn SNR -> dB normalisation -> crop / zero-pad, one launch sequence per BATCH.
// Reference: ppvector/data_utils/reader.py:85-104 (resample, augment_audio :153-163, normalize(target_db), crop) -- per utterance, on
// the CPU, inside DataLoader workers, through the un-vendored yeaudio package; configs/augmentation.yml:1-33.  The semantics below
// are RECALLED yeaudio behaviour (restated in oracle/audio_prep.py), the random draws stay on the host:
//   speed   change_speed(r): new_len = int(len / r), samples = np.interp(linspace(0, len, new_len), arange(len), samples)
//   volume  samples *= 10^(gain_dB / 20)
//   noise   noise gain = min(rms_dB(signal) - rms_dB(noise) - snr_dB, 300) dB; samples += noise * gain (noise tiled over the utterance)
//   norm    gain = min(target_dB - rms_dB(samples), 300) dB over the WHOLE utterance (before the crop), rms_dB = 10 log10(mean x^2)
//   crop    [start, start + crop_len) of the result, zero-padded to the batch's output length
// Two passes over the samples: (1) per-utterance sums  S_xx, S_nn, S_xn  of the speed-changed signal and its noise segment, from
// which every gain follows in closed form (the mixture's energy is g^2 S_xx + 2 g g_n S_xn + g_n^2 S_nn); (2) the output.  HBM-bound:
// ~2 reads of the raw samples + 1 write of the crop.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace ppv {

namespace {

constexpr int AP_CHUNK = 8192;  // samples per block in pass 1

struct PrepItem {
    int raw_len, new_len, crop_start, crop_len, noise_off, noise_len, has_noise;
    float pos_step, vol_gain_db, snr_db;
};

__device__ __forceinline__ PrepItem load_item(const int32_t* ip, const float* fp, int b) {
    PrepItem it;
    const int32_t* i = ip + b * PPV_PREP_NI;
    const float* f = fp + b * PPV_PREP_NF;
    it.raw_len = i[0];
    it.new_len = i[1];
    it.crop_start = i[2];
    it.crop_len = i[3];
    it.noise_off = i[4];
    it.noise_len = i[5];
    it.has_noise = i[6];
    it.pos_step = f[0];
    it.vol_gain_db = f[1];
    it.snr_db = f[2];
    return it;
}

// sample j of the speed-changed signal: np.interp(j * pos_step, arange(raw_len), x) with np.interp's clamping beyond the last index
__device__ __forceinline__ float speed_sample(const float* __restrict__ x, const PrepItem& it, int j) {
    if (it.new_len == it.raw_len) return x[j];
    // linspace(0, raw_len, new_len)[j] in double (a float position loses the fraction beyond ~1e6 samples)
    const double pos = double(j) * (double(it.raw_len) / double(max(it.new_len - 1, 1)));
    int i0 = int(pos);
    if (i0 >= it.raw_len - 1) return x[it.raw_len - 1];
    const float fr = float(pos - double(i0));
    const float a = x[i0], c = x[i0 + 1];
    return a + fr * (c - a);
}

__global__ void __launch_bounds__(256) prep_stats_kernel(const float* __restrict__ wav, int64_t wav_ld, const int32_t* __restrict__ ip,
                                                         const float* __restrict__ fp, const float* __restrict__ noise, int nchunk,
                                                         double* __restrict__ partial) {
    __shared__ double red[3][8];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const PrepItem it = load_item(ip, fp, b);
    const float* x = wav + int64_t(b) * wav_ld;
    double sxx = 0.0, snn = 0.0, sxn = 0.0;
    const int j0 = chunk * AP_CHUNK, j1 = min(it.new_len, j0 + AP_CHUNK);
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const float v = speed_sample(x, it, j);
        sxx += double(v) * double(v);
        if (it.has_noise) {
            const float n = noise[it.noise_off + (j % it.noise_len)];
            snn += double(n) * double(n);
            sxn += double(v) * double(n);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
        snn += __shfl_xor_sync(0xffffffffu, snn, o);
        sxn += __shfl_xor_sync(0xffffffffu, sxn, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sxx;
        red[1][threadIdx.x >> 5] = snn;
        red[2][threadIdx.x >> 5] = sxn;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];  // fixed order: deterministic
        partial[(int64_t(b) * nchunk + chunk) * 3 + threadIdx.x] = s;
    }
}

// gains[b] = {signal gain, noise gain} including the dB normalisation
__global__ void prep_gains_kernel(const int32_t* __restrict__ ip, const float* __restrict__ fp, const double* __restrict__ partial, int nchunk,
                                  int B, float target_db, int normalize, float* __restrict__ gains) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const PrepItem it = load_item(ip, fp, b);
    double sxx = 0.0, snn = 0.0, sxn = 0.0;
    const int used = (it.new_len + AP_CHUNK - 1) / AP_CHUNK;
    for (int c = 0; c < used; ++c) {
        sxx += partial[(int64_t(b) * nchunk + c) * 3 + 0];
        snn += partial[(int64_t(b) * nchunk + c) * 3 + 1];
        sxn += partial[(int64_t(b) * nchunk + c) * 3 + 2];
    }
    const double n = double(max(it.new_len, 1));
    const double g = pow(10.0, double(it.vol_gain_db) / 20.0);
    double gn = 0.0;
    if (it.has_noise && snn > 0.0 && sxx > 0.0) {
        const double sig_db = 10.0 * log10(g * g * sxx / n), noise_db = 10.0 * log10(snn / n);
        gn = pow(10.0, fmin(sig_db - noise_db - double(it.snr_db), 300.0) / 20.0);
    }
    double gnorm = 1.0;
    if (normalize) {
        const double ms = (g * g * sxx + 2.0 * g * gn * sxn + gn * gn * snn) / n;
        if (ms > 0.0) gnorm = pow(10.0, fmin(double(target_db) - 10.0 * log10(ms), 300.0) / 20.0);
    }
    gains[2 * b] = float(g * gnorm);
    gains[2 * b + 1] = float(gn * gnorm);
}

__global__ void __launch_bounds__(256) prep_apply_kernel(const float* __restrict__ wav, int64_t wav_ld, const int32_t* __restrict__ ip,
                                                         const float* __restrict__ fp, const float* __restrict__ noise,
                                                         const float* __restrict__ gains, int Lout, float* __restrict__ out) {
    const int b = blockIdx.y;
    const PrepItem it = load_item(ip, fp, b);
    const float* x = wav + int64_t(b) * wav_ld;
    const float gs = gains[2 * b], gn = gains[2 * b + 1];
    float* dst = out + int64_t(b) * Lout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Lout; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < it.crop_len) {
            const int j = it.crop_start + i;
            v = speed_sample(x, it, j) * gs;
            if (it.has_noise) v = fmaf(noise[it.noise_off + (j % it.noise_len)], gn, v);
        }
        dst[i] = v;
    }
}

}  // namespace

size_t audio_prep_workspace_bytes(int B, int max_len) {
    if (B <= 0 || max_len <= 0) return 0;
    const size_t nchunk = size_t((max_len + AP_CHUNK - 1) / AP_CHUNK);
    return ((size_t(B) * nchunk * 3 * sizeof(double) + 255) / 256) * 256 + ((size_t(B) * 2 * sizeof(float) + 255) / 256) * 256;
}

// wav [B][wav_ld] raw (resampled) samples; iparams [B][PPV_PREP_NI]; fparams [B][PPV_PREP_NF]; noise: concatenated noise clips (may be null
// when no item has noise); out [B][Lout].  max_new_len >= every item's new_len (sizes the workspace).
int audio_prep(const float* wav, int64_t wav_ld, const int32_t* iparams, const float* fparams, const float* noise, int B, int max_new_len,
               float target_db, int normalize, int Lout, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    PPV_REQUIRE(wav && iparams && fparams && out && ws, "audio_prep: null argument");
    PPV_REQUIRE(B > 0 && max_new_len > 0 && Lout > 0, "audio_prep: empty batch");
    PPV_REQUIRE(ws_bytes >= audio_prep_workspace_bytes(B, max_new_len) && (reinterpret_cast<uintptr_t>(ws) & 255) == 0,
                "audio_prep: workspace too small / unaligned");
    const int nchunk = (max_new_len + AP_CHUNK - 1) / AP_CHUNK;
    double* partial = static_cast<double*>(ws);
    float* gains = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + ((size_t(B) * nchunk * 3 * sizeof(double) + 255) / 256) * 256);
    prep_stats_kernel<<<dim3(nchunk, B), 256, 0, st>>>(wav, wav_ld, iparams, fparams, noise, nchunk, partial);
    prep_gains_kernel<<<(B + 127) / 128, 128, 0, st>>>(iparams, fparams, partial, nchunk, B, target_db, normalize, gains);
    const int gx = std::max(1, std::min((Lout + 255) / 256, 64));
    prep_apply_kernel<<<dim3(gx, B), 256, 0, st>>>(wav, wav_ld, iparams, fparams, noise, gains, Lout, out);
    PPV_LAUNCH_OK("audio_prep kernels");
    return PPV_OK;
}

}  // namespace ppv
