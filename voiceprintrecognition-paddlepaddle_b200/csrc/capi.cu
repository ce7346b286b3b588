// extern "C" boundary of libppv_b200 (include/ppv_b200.h): argument checking, handle unwrapping, error text.
// No exceptions cross this boundary; everything returns a status code.
#include <string.h>

#include <new>

#include "common.h"

namespace ppv {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

int device_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;
    }
    return n;
}

static int check_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(PPV_ECUDA, "no CUDA device: libppv_b200 has no CPU fallback");
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) return fail(PPV_EUNSUPPORTED, "libppv_b200 is built for sm_100a only (found sm_" + std::to_string(major * 10 + minor) + ")");
    return PPV_OK;
}

}  // namespace ppv

using namespace ppv;

struct ppv_fbank {
    Fbank* impl;
};
struct ppv_spectral {
    Spectral* impl;
};

struct ppv_trainer {
    Trainer* impl;
};

struct ppv_model {
    int kind;
    EcapaModel* ecapa;
    ResNetSEModel* resnet;
    ERes2NetModel* eres;
    CamppModel* campp;
};

#define PPV_GUARD_BEGIN try {
#define PPV_GUARD_END                                                        \
    }                                                                        \
    catch (const std::bad_alloc&) { return fail(PPV_ECUDA, "host out of memory"); } \
    catch (const std::exception& e) { return fail(PPV_EINVAL, std::string("exception: ") + e.what()); } \
    catch (...) { return fail(PPV_EINVAL, "unknown exception"); }

extern "C" {

int ppv_version(void) { return 100; }

int ppv_last_error(char* buf, size_t n) {
    const std::string& e = g_last_error;
    if (buf && n > 0) {
        const size_t k = std::min(n - 1, e.size());
        memcpy(buf, e.data(), k);
        buf[k] = 0;
    }
    return int(e.size());
}

int ppv_device_sm_count(void) { return device_sm_count(); }

// ---------------------------------------------------------------- fbank
void ppv_fbank_default_cfg(ppv_fbank_cfg* c) {
    if (!c) return;
    c->sample_rate = 16000;
    c->n_mels = 80;
    c->frame_length_ms = 25.f;
    c->frame_shift_ms = 10.f;
    c->preemph = 0.97f;
    c->low_freq = 20.f;
    c->high_freq = 0.f;
    c->log_floor = 1.1920928955078125e-07f;
}

int ppv_fbank_create(const ppv_fbank_cfg* cfg, ppv_fbank_t** out) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(cfg && out, "ppv_fbank_create: null argument");
    int rc = check_device();
    if (rc) return rc;
    Fbank* impl = nullptr;
    rc = fbank_create(cfg, &impl);
    if (rc) return rc;
    *out = new ppv_fbank{impl};
    return PPV_OK;
    PPV_GUARD_END
}
int ppv_fbank_destroy(ppv_fbank_t* h) {
    if (!h) return PPV_OK;
    fbank_destroy(h->impl);
    delete h;
    return PPV_OK;
}
int ppv_fbank_num_frames(const ppv_fbank_t* h, int L) { return h ? fbank_num_frames(h->impl, L) : 0; }
int ppv_fbank_feature_dim(const ppv_fbank_t* h) { return h ? fbank_n_mels(h->impl) : 0; }

int ppv_fbank_forward(ppv_fbank_t* h, const float* wav, const float* lens_ratio, int B, int L, float* out, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && wav && out, "ppv_fbank_forward: null argument");
    PPV_REQUIRE(B > 0 && L > 0, "ppv_fbank_forward: empty input");
    // raw log-mel goes to `out`, then CMN is applied in place (each element is read and written by one thread)
    return fbank_run(h->impl, wav, lens_ratio, B, L, out, out, Planes(), 0, 0, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

int ppv_fbank_forward_ragged(ppv_fbank_t* h, const float* wav, const int32_t* valid_frames, int B, int L, float* out, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && wav && out && valid_frames, "ppv_fbank_forward_ragged: null argument");
    PPV_REQUIRE(B > 0 && L > 0, "ppv_fbank_forward_ragged: empty input");
    return fbank_run(h->impl, wav, nullptr, B, L, out, out, Planes(), 0, 0, static_cast<cudaStream_t>(stream), valid_frames);
    PPV_GUARD_END
}

size_t ppv_audio_prep_workspace_bytes(int B, int max_new_len) { return audio_prep_workspace_bytes(B, max_new_len); }
int ppv_audio_prep(const float* wav, int64_t wav_ld, const int32_t* iparams, const float* fparams, const float* noise, int B, int max_new_len,
                   float target_db, int normalize, int Lout, float* out, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    return audio_prep(wav, wav_ld, iparams, fparams, noise, B, max_new_len, target_db, normalize, Lout, out, ws, ws_bytes,
                      static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

// ---------------------------------------------------------------- model
void ppv_ecapa_default_cfg(ppv_ecapa_cfg* c) {
    if (!c) return;
    c->input_size = 80;
    c->embd_dim = 192;
    const int ch[5] = {512, 512, 512, 512, 1536}, ks[5] = {5, 3, 3, 3, 1}, dl[5] = {1, 2, 3, 4, 1};
    for (int i = 0; i < 5; ++i) {
        c->channels[i] = ch[i];
        c->kernel_sizes[i] = ks[i];
        c->dilations[i] = dl[i];
    }
    c->attention_channels = 128;
    c->res2net_scale = 8;
    c->se_channels = 128;
    c->precision = PPV_PREC_BF16X3;
    c->pooling = PPV_POOL_ASP;
    c->global_context = 1;
}

void ppv_eres2net_default_cfg(ppv_eres2net_cfg* c) {
    if (c) ppv_eres2net_default_cfg_impl(c);
}
void ppv_campplus_default_cfg(ppv_campplus_cfg* c) {
    if (c) ppv_campplus_default_cfg_impl(c);
}
void ppv_resnetse_default_cfg(ppv_resnetse_cfg* c) {
    if (c) ppv_resnetse_default_cfg_impl(c);
}

int ppv_model_create(int kind, const void* cfg, ppv_model_t** out) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(cfg && out, "ppv_model_create: null argument");
    if (kind != PPV_MODEL_ECAPA_TDNN && kind != PPV_MODEL_RESNET_SE && kind != PPV_MODEL_ERES2NET && kind != PPV_MODEL_CAMPPLUS)
        return fail(PPV_EUNSUPPORTED,
                    "ppv_model_create: implemented kinds are PPV_MODEL_ECAPA_TDNN, PPV_MODEL_RESNET_SE, PPV_MODEL_ERES2NET, PPV_MODEL_CAMPPLUS");
    int rc = check_device();
    if (rc) return rc;
    if (kind == PPV_MODEL_RESNET_SE) {
        ResNetSEModel* r = nullptr;
        rc = resnetse_create(static_cast<const ppv_resnetse_cfg*>(cfg), &r);
        if (rc) return rc;
        *out = new ppv_model{kind, nullptr, r, nullptr, nullptr};
        return PPV_OK;
    }
    if (kind == PPV_MODEL_ERES2NET) {
        ERes2NetModel* e = nullptr;
        rc = eres2net_create(static_cast<const ppv_eres2net_cfg*>(cfg), &e);
        if (rc) return rc;
        *out = new ppv_model{kind, nullptr, nullptr, e, nullptr};
        return PPV_OK;
    }
    if (kind == PPV_MODEL_CAMPPLUS) {
        CamppModel* c = nullptr;
        rc = campplus_create(static_cast<const ppv_campplus_cfg*>(cfg), &c);
        if (rc) return rc;
        *out = new ppv_model{kind, nullptr, nullptr, nullptr, c};
        return PPV_OK;
    }
    EcapaModel* m = nullptr;
    rc = ecapa_create(static_cast<const ppv_ecapa_cfg*>(cfg), &m);
    if (rc) return rc;
    *out = new ppv_model{kind, m, nullptr, nullptr, nullptr};
    return PPV_OK;
    PPV_GUARD_END
}
int ppv_model_destroy(ppv_model_t* h) {
    if (!h) return PPV_OK;
    ecapa_destroy(h->ecapa);
    resnetse_destroy(h->resnet);
    eres2net_destroy(h->eres);
    campplus_destroy(h->campp);
    delete h;
    return PPV_OK;
}
int ppv_model_load_weight(ppv_model_t* h, const char* name, const float* data, const int64_t* shape, int ndim) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_model_load_weight: null model");
    if (h->resnet) return resnetse_load_weight(h->resnet, name, data, shape, ndim);
    if (h->eres) return eres2net_load_weight(h->eres, name, data, shape, ndim);
    if (h->campp) return campplus_load_weight(h->campp, name, data, shape, ndim);
    return ecapa_load_weight(h->ecapa, name, data, shape, ndim);
    PPV_GUARD_END
}
int ppv_model_finalize(ppv_model_t* h) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_model_finalize: null model");
    if (h->resnet) return resnetse_finalize(h->resnet);
    if (h->eres) return eres2net_finalize(h->eres);
    if (h->campp) return campplus_finalize(h->campp);
    return ecapa_finalize(h->ecapa);
    PPV_GUARD_END
}
int ppv_model_set_precision(ppv_model_t* h, int precision) {
    PPV_REQUIRE(h, "ppv_model_set_precision: null model");
    if (h->resnet) return resnetse_set_precision(h->resnet, precision);
    if (h->eres) return eres2net_set_precision(h->eres, precision);
    if (h->campp) return campplus_set_precision(h->campp, precision);
    return ecapa_set_precision(h->ecapa, precision);
}
int ppv_model_embd_dim(const ppv_model_t* h) {
    if (!h) return 0;
    if (h->campp) return campplus_embd_dim(h->campp);
    return h->resnet ? resnetse_embd_dim(h->resnet) : h->eres ? eres2net_embd_dim(h->eres) : ecapa_embd_dim(h->ecapa);
}
size_t ppv_model_workspace_bytes(const ppv_model_t* h, int B, int T) {
    if (h && h->campp) return campplus_workspace_bytes(h->campp, B, T);
    return !h ? 0 : h->resnet ? resnetse_workspace_bytes(h->resnet, B, T) : h->eres ? eres2net_workspace_bytes(h->eres, B, T)
                                                                              : ecapa_workspace_bytes(h->ecapa, B, T);
}

int ppv_model_forward(ppv_model_t* h, const float* feat, int B, int T, float* emb, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && feat && emb, "ppv_model_forward: null argument");
    if (h->resnet) return resnetse_forward(h->resnet, feat, B, T, emb, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    if (h->eres) return eres2net_forward(h->eres, feat, B, T, emb, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    if (h->campp) return campplus_forward(h->campp, feat, B, T, emb, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    return ecapa_forward(h->ecapa, feat, nullptr, nullptr, nullptr, B, T, 0, emb, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_model_forward_lengths(ppv_model_t* h, const float* feat, const float* lengths, int B, int T, float* emb, void* ws, size_t ws_bytes,
                              void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && feat && emb, "ppv_model_forward_lengths: null argument");
    if (!h->ecapa) return fail(PPV_EUNSUPPORTED, "ppv_model_forward_lengths: only EcapaTdnn.forward takes lengths in the reference");
    return ecapa_forward(h->ecapa, feat, nullptr, nullptr, nullptr, B, T, 0, emb, ws, ws_bytes, static_cast<cudaStream_t>(stream), lengths);
    PPV_GUARD_END
}
int ppv_model_forward_wav(ppv_model_t* h, ppv_fbank_t* fb, const float* wav, const float* lens_ratio, int B, int L, float* emb,
                          void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && fb && wav && emb, "ppv_model_forward_wav: null argument");
    if (h->resnet || h->eres || h->campp) return fail(PPV_EUNSUPPORTED, "ppv_model_forward_wav: the fused waveform path exists for ECAPA-TDNN only; call ppv_fbank_forward + ppv_model_forward");
    const int T = fbank_num_frames(fb->impl, L);
    PPV_REQUIRE(T > 0, "ppv_model_forward_wav: waveform shorter than one frame");
    return ecapa_forward(h->ecapa, nullptr, fb->impl, wav, lens_ratio, B, T, L, emb, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_model_read_tap(ppv_model_t* h, const char* name, float* out, size_t out_elems, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_model_read_tap: null model");
    if (h->resnet) return resnetse_read_tap(h->resnet, name, out, out_elems, static_cast<cudaStream_t>(stream));
    if (h->eres) return eres2net_read_tap(h->eres, name, out, out_elems, static_cast<cudaStream_t>(stream));
    if (h->campp) return campplus_read_tap(h->campp, name, out, out_elems, static_cast<cudaStream_t>(stream));
    return ecapa_read_tap(h->ecapa, name, out, out_elems, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

size_t ppv_eer_workspace_bytes(int64_t n) { return eer_workspace_bytes(n); }
int ppv_eer_mindcf(const float* scores, const int32_t* labels, int64_t n, double p_target, double c_miss, double c_fa, double* out4, void* ws,
                   size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(labels, "ppv_eer_mindcf: null labels");
    return eer_mindcf(scores, labels, nullptr, nullptr, 0, n, p_target, c_miss, c_fa, out4, ws, ws_bytes, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_eer_mindcf_matrix(const float* scores, const int32_t* trial_labels, const int32_t* enroll_labels, int M, int N, double p_target,
                          double c_miss, double c_fa, double* out4, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(trial_labels && enroll_labels && M > 0 && N > 0, "ppv_eer_mindcf_matrix: bad argument");
    return eer_mindcf(scores, nullptr, trial_labels, enroll_labels, N, int64_t(M) * N, p_target, c_miss, c_fa, out4, ws, ws_bytes,
                      static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_row_argmax(const float* sim, int rows, int cols, int32_t* idx, float* best, void* stream) {
    PPV_GUARD_BEGIN
    return row_argmax(sim, rows, cols, idx, best, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

int ppv_model_profile(ppv_model_t* h, int enable) {
    PPV_REQUIRE(h && h->ecapa, "ppv_model_profile: ECAPA-TDNN model required");
    return ecapa_profile(h->ecapa, enable);
}
int ppv_model_profile_read(ppv_model_t* h, double* gemm_ms, double* other_ms, int64_t* gemm_launches, int64_t* other_launches) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && h->ecapa, "ppv_model_profile_read: ECAPA-TDNN model required");
    return ecapa_profile_read(h->ecapa, gemm_ms, other_ms, gemm_launches, other_launches);
    PPV_GUARD_END
}

// ---------------------------------------------------------------- STFT front ends, SpecAugment
void ppv_spectral_default_cfg(ppv_spectral_cfg* c, int method) {
    if (c) spectral_default_cfg(c, method);
}
int ppv_spectral_create(const ppv_spectral_cfg* cfg, ppv_spectral_t** out) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(cfg && out, "ppv_spectral_create: null argument");
    int rc = check_device();
    if (rc) return rc;
    Spectral* impl = nullptr;
    rc = spectral_create(cfg, &impl);
    if (rc) return rc;
    *out = new ppv_spectral{impl};
    return PPV_OK;
    PPV_GUARD_END
}
int ppv_spectral_destroy(ppv_spectral_t* h) {
    if (!h) return PPV_OK;
    spectral_destroy(h->impl);
    delete h;
    return PPV_OK;
}
int ppv_spectral_num_frames(const ppv_spectral_t* h, int L) { return h ? spectral_num_frames(h->impl, L) : 0; }
int ppv_spectral_feature_dim(const ppv_spectral_t* h) { return h ? spectral_feature_dim(h->impl) : 0; }
int ppv_spectral_forward(ppv_spectral_t* h, const float* wav, const float* lens_ratio, int B, int L, float* out, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h && wav && out, "ppv_spectral_forward: null argument");
    PPV_REQUIRE(B > 0 && L > 0, "ppv_spectral_forward: empty input");
    return spectral_run(h->impl, wav, lens_ratio, B, L, out, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_spec_augment(float* feat, const int32_t* params, int B, int T, int F, int n_freq_masks, int n_time_masks, int fill_mode, void* stream) {
    PPV_GUARD_BEGIN
    return spec_augment_run(feat, params, B, T, F, n_freq_masks, n_time_masks, fill_mode, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

// ---------------------------------------------------------------- training step
int ppv_trainer_create(const ppv_ecapa_cfg* cfg, int num_classes, ppv_trainer_t** out) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(cfg && out, "ppv_trainer_create: null argument");
    int rc = check_device();
    if (rc) return rc;
    Trainer* impl = nullptr;
    rc = trainer_create(cfg, num_classes, &impl);
    if (rc) return rc;
    *out = new ppv_trainer{impl};
    return PPV_OK;
    PPV_GUARD_END
}
int ppv_trainer_destroy(ppv_trainer_t* h) {
    if (!h) return PPV_OK;
    trainer_destroy(h->impl);
    delete h;
    return PPV_OK;
}
int64_t ppv_trainer_param_count(const ppv_trainer_t* h) { return h ? trainer_param_count(h->impl) : 0; }
int64_t ppv_trainer_stat_count(const ppv_trainer_t* h) { return h ? trainer_stat_count(h->impl) : 0; }
int ppv_trainer_lookup(const ppv_trainer_t* h, const char* name, int64_t* offset, int64_t* numel, int* is_stat) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_trainer_lookup: null handle");
    return trainer_lookup(h->impl, name, offset, numel, is_stat);
    PPV_GUARD_END
}
int ppv_set_pdl(int enabled) {
    const int prev = ppv::pdl_enabled();
    ppv::pdl_enabled() = enabled ? 1 : 0;
    return prev;
}
int ppv_trainer_set_precision(ppv_trainer_t* h, int precision) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_trainer_set_precision: null handle");
    return trainer_set_precision(h->impl, precision);
    PPV_GUARD_END
}
int ppv_trainer_bind(ppv_trainer_t* h, float* params, float* grads, float* stats) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_trainer_bind: null handle");
    return trainer_bind(h->impl, params, grads, stats);
    PPV_GUARD_END
}
size_t ppv_trainer_workspace_bytes(ppv_trainer_t* h, int B, int T) {
    try {
        return h ? trainer_workspace_bytes(h->impl, B, T) : 0;
    } catch (...) {
        return 0;
    }
}
int ppv_trainer_forward_backward(ppv_trainer_t* h, const float* feat, const int64_t* labels, int B, int T, float margin, float scale,
                                 int easy_margin, float label_smoothing, float* loss, float* logits, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_trainer_forward_backward: null handle");
    return trainer_forward_backward(h->impl, feat, labels, B, T, margin, scale, easy_margin, label_smoothing, loss, logits, ws, ws_bytes,
                                    static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_trainer_read_tap(ppv_trainer_t* h, const char* name, float* out, size_t out_elems, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(h, "ppv_trainer_read_tap: null handle");
    return trainer_read_tap(h->impl, name, out, out_elems, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int64_t step, float grad_scale, void* stream) {
    PPV_GUARD_BEGIN
    return adam_step(params, grads, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

// ---------------------------------------------------------------- cosine scoring
size_t ppv_cosine_workspace_bytes(int M, int N, int D) { return cosine_workspace_bytes(M, N, D); }
int ppv_cosine_matrix(const float* A, const float* Bm, int M, int N, int D, float* out, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    int rc = check_device();
    if (rc) return rc;
    return cosine_matrix(A, Bm, M, N, D, out, ws, ws_bytes, PPV_PREC_BF16X3, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_cosine_pairlist(const float* E, const int32_t* idx, int64_t P, int n, int D, float* out, void* stream) {
    PPV_GUARD_BEGIN
    return cosine_pairlist(E, idx, P, n, D, out, static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

// ---------------------------------------------------------------- AAM head
size_t ppv_aam_workspace_bytes(int B, int D, int S) { return aam_workspace_bytes(B, D, S); }
int ppv_aam_forward(const float* emb, const float* W, const int64_t* labels, int B, int D, int S, float margin, float scale,
                    int easy_margin, float label_smoothing, float* logits, float* loss, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    return aam_forward(emb, W, labels, B, D, S, margin, scale, easy_margin, label_smoothing, logits, loss, ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}
int ppv_aam_backward(const float* emb, const float* W, const int64_t* labels, const float* logits, int B, int D, int S, float margin,
                     float scale, int easy_margin, float label_smoothing, float* d_emb, float* d_W, void* ws, size_t ws_bytes,
                     void* stream) {
    PPV_GUARD_BEGIN
    return aam_backward(emb, W, labels, logits, B, D, S, margin, scale, easy_margin, label_smoothing, d_emb, d_W, ws, ws_bytes,
                        static_cast<cudaStream_t>(stream));
    PPV_GUARD_END
}

// ---------------------------------------------------------------- GEMM test hook
static inline size_t au(size_t x, size_t a) { return (x + a - 1) / a * a; }
size_t ppv_gemm_test_workspace_bytes(int M, int N, int K) {
    const size_t Kp = au(size_t(K), 64);
    return au(au(size_t(M), 128) * Kp * 4, 256) + au(au(size_t(N), 256) * Kp * 4, 256) + 256;
}
int ppv_gemm_test(const float* A, const float* W, const float* bias, const float* bn_scale, const float* bn_shift, int relu, int M,
                  int N, int K, int block_n, int block_k, int precision, float* out, void* ws, size_t ws_bytes, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(A && W && out && ws, "ppv_gemm_test: null argument");
    PPV_REQUIRE(ws_bytes >= ppv_gemm_test_workspace_bytes(M, N, K), "ppv_gemm_test: workspace too small");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kp = int(au(size_t(K), 64));
    Planes pa, pw;
    pa.rows = int64_t(au(size_t(M), 128));
    pa.ld = Kp;
    pa.plane_stride = pa.rows * Kp;
    pa.base = static_cast<__nv_bfloat16*>(ws);
    pw.rows = int64_t(au(size_t(N), 256));
    pw.ld = Kp;
    pw.plane_stride = pw.rows * Kp;
    pw.base = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(ws) + au(size_t(pa.plane_stride) * 4, 256));
    PPV_CUDA_OK(cudaMemsetAsync(ws, 0, ppv_gemm_test_workspace_bytes(M, N, K), st));
    rc = launch_f32_to_planes(A, M, K, pa, st);
    if (rc) return rc;
    rc = launch_f32_to_planes(W, N, K, pw, st);
    if (rc) return rc;
    GemmSource src{pa, 0, Kp, 0};
    Epilogue ep;
    ep.bias = bias;
    ep.relu = relu;
    ep.bn_scale = bn_scale;
    ep.bn_shift = bn_shift;
    ep.out_mode = OUT_F32;
    ep.out = out;
    ep.out_ld = N;
    GemmParams gp;
    rc = gemm_build(&gp, &src, 1, pw, M, N, ep, block_n, block_k);
    if (rc) return rc;
    return gemm_launch(gp, block_n, precision, device_sm_count(), st);
    PPV_GUARD_END
}

// Kernel-only timing of the gather-GEMM (tools/gemm_bench.py): operands are converted once, the kernel is launched
// `iters` times between two CUDA events on `stream`; *ms_per_launch receives the average.  out is fp32 [M,N] or, when
// planes_out != 0, a split-bf16 planes buffer inside ws (the layout every model layer writes).
int ppv_gemm_bench(int M, int N, int K, int block_n, int block_k, int precision, int planes_out, int iters, void* ws, size_t ws_bytes,
                   float* ms_per_launch, void* stream) {
    PPV_GUARD_BEGIN
    PPV_REQUIRE(ws && ms_per_launch && iters > 0, "ppv_gemm_bench: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t Kp = au(size_t(K), 64);
    const size_t a_bytes = au(au(size_t(M), 128) * Kp * 4, 256), w_bytes = au(au(size_t(N), 256) * Kp * 4, 256);
    const size_t o_bytes = au(au(size_t(M), 128) * size_t(N) * 4, 256);
    PPV_REQUIRE(ws_bytes >= a_bytes + w_bytes + o_bytes, "ppv_gemm_bench: workspace too small");
    Planes pa, pw, po;
    pa.rows = int64_t(au(size_t(M), 128)); pa.ld = int(Kp); pa.plane_stride = pa.rows * pa.ld; pa.base = static_cast<__nv_bfloat16*>(ws);
    pw.rows = int64_t(au(size_t(N), 256)); pw.ld = int(Kp); pw.plane_stride = pw.rows * pw.ld;
    pw.base = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(ws) + a_bytes);
    po.rows = pa.rows; po.ld = N; po.plane_stride = po.rows * po.ld;
    po.base = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(ws) + a_bytes + w_bytes);
    PPV_CUDA_OK(cudaMemsetAsync(ws, 0x11, a_bytes + w_bytes, st));  // small finite bf16 values
    GemmSource src{pa, 0, int(Kp), 0};
    Epilogue ep;
    ep.relu = 1;
    if (planes_out) {
        ep.out_mode = OUT_PLANES; ep.out = po.base; ep.out_ld = po.ld; ep.out_plane_stride = po.plane_stride;
    } else {
        ep.out_mode = OUT_F32; ep.out = po.base; ep.out_ld = N;
    }
    GemmParams gp;
    int rc = gemm_build(&gp, &src, 1, pw, M, N, ep, block_n, block_k);
    if (rc) return rc;
    const int sms = device_sm_count();
    for (int i = 0; i < 3; ++i) { rc = gemm_launch(gp, block_n, precision, sms, st); if (rc) return rc; }
    cudaEvent_t e0, e1;
    PPV_CUDA_OK(cudaEventCreate(&e0));
    PPV_CUDA_OK(cudaEventCreate(&e1));
    PPV_CUDA_OK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) { rc = gemm_launch(gp, block_n, precision, sms, st); if (rc) return rc; }
    PPV_CUDA_OK(cudaEventRecord(e1, st));
    PPV_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    PPV_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_launch = ms / float(iters);
    return PPV_OK;
    PPV_GUARD_END
}

}  // extern "C"
