// CAM++ (CAMPPlus) forward as a plan of tensor-core gather-GEMMs.
// Reference graph: ppvector/models/campplus.py:342-346 (CAMPPlus.forward), :278-289 (FCM head), :245-251 (BasicResBlock),
// :61-64 (TDNNLayer), :88-106 (CAMLayer + seg_pooling), :135-141 (CAMDenseTDNNLayer), :167-171 (dense concat),
// :183-186 (TransitLayer), :24-31 (statistics pooling), :196-204 (DenseLayer).  Eval mode, configs/cam++.yml defaults
// (growth 32, bn_size 4, init_channels 128, blocks 12/24/16 with dilations 1/2/2, 100-frame average segment pooling).
//
// Layouts:
//   * FCM head: zero-bordered NHWC images (H = frequency, W = time) as in resnet_se.cu; its convs stride the frequency
//     axis only, so a strided conv is computed on the input grid and stored on the (H/2, W) output grid;
//   * the head output [B, 32, F/8, T] is flattened to a time-major matrix that holds frame PAIRS: row (b, t') has the
//     320 channels of frame 2t' followed by those of frame 2t'+1.  The stride-2, kernel-5 TDNN conv is then five
//     K-sources (even|odd column windows at row offsets -1,-1,0,0,+1) and no output is computed and thrown away;
//   * D-TDNN part: padded time layout row = b*Tp + P + t with ZERO padding rows (Paddle's Conv1D pads with zeros).  One
//     buffer per dense block holds the growing concatenation; each layer's CAM output is stored into its 32-column window.
//
// Per dense layer (campplus.py:135-141):
//   relu(bn1(x))            one elementwise pass (the BN differs per layer, ReLU keeps it out of the weights)
//   linear1 + bn2 + relu    GEMM, BN folded into the weights, ReLU epilogue
//   context mask            one small kernel per utterance: segment sums -> mean + segment mean -> 128-64-32 MLP -> sigmoid
//   linear_local * mask     3-tap gather-GEMM whose epilogue multiplies by the mask of (utterance, segment)
#include <math.h>

#include <stdlib.h>

#include "common.h"
#include "model_common.h"
#include "ptx.cuh"

namespace ppv {

namespace {

constexpr int CP_P = 4;          // zero padding rows on each side of an utterance (>= max dilation)
constexpr int CP_SEG = 100;      // seg_pooling seg_len, campplus.py:95
constexpr int CP_MAX_SEG = 64;   // segments per utterance held in shared memory by the context kernel
constexpr int CP_MAX_LAYERS = 64;
constexpr int CP_NB = 3;
const int CP_LAYERS[CP_NB] = {12, 24, 16};
const int CP_DIL[CP_NB] = {1, 2, 2};

struct ResBlockW {
    GemmWeights conv1, conv2, sc;
    bool has_sc = false;
    int stride = 1, stage = 0;  // stage = index of the output geometry
};
struct DenseLayerW {
    float *bn1_scale = nullptr, *bn1_shift = nullptr;  // [Kp]
    GemmWeights linear1, local;
    float *w1t = nullptr, *b1 = nullptr, *w2t = nullptr, *b2 = nullptr;  // context MLP, transposed: [128][64], [64][32]
    int in_ch = 0, Kp = 0, block = 0, dil = 1;
};
struct TransitW {
    float *bn_scale = nullptr, *bn_shift = nullptr;
    GemmWeights linear;
    int C = 0;
};

struct CStep {
    enum Kind { STEM, GEMM, CONV3, PW, ADD_RELU, FLATTEN, BN_RELU, CONTEXT, STATS } kind;
    PwStep pw;  // PW: 1x1 conv with K <= 64 on the CUDA cores (pointwise.cu)
    GemmParams gp;
    Conv3x3Params c3;  // CONV3: the 32 -> 32 channel 3x3 convs of the FCM head (conv3x3.cu)
    int BN = 0;
    Planes a, b, d;
    const float *p0 = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr, *p4 = nullptr, *p5 = nullptr;
    float* fout = nullptr;
    int C = 0, img_rows = 0;
    int64_t rows = 0;
};

// ------------------------------------------------------------------------------------------------ kernels
// out[r, c] = relu(x[r, c] * scale[c] + shift[c]) for c < C (C % 8 == 0), every row
__global__ void __launch_bounds__(256)
    cp_bn_relu_kernel(Planes x, const float* __restrict__ scale, const float* __restrict__ shift, Planes out, int C, int64_t rows) {
    griddep_launch_dependents();
    griddep_wait();
    const int groups = C >> 3;
    const int64_t total = rows * groups;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / groups;
        const int c = int(i - r * groups) * 8;
        const uint4 h = *reinterpret_cast<const uint4*>(x.hi() + r * x.ld + c);
        const uint4 l = *reinterpret_cast<const uint4*>(x.lo() + r * x.ld + c);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + c)), b1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t oh[4], ol[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[k]));
            const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[k]));
            const float y0 = fmaxf(fmaf(hf.x + lf.x, sv[2 * k], bv[2 * k]), 0.f);
            const float y1 = fmaxf(fmaf(hf.y + lf.y, sv[2 * k + 1], bv[2 * k + 1]), 0.f);
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(y0, h0, l0);
            split_bf16(y1, h1, l1);
            oh[k] = pack_bf16x2(h0, h1);
            ol[k] = pack_bf16x2(l0, l1);
        }
        *reinterpret_cast<uint4*>(out.hi() + r * out.ld + c) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<uint4*>(out.lo() + r * out.ld + c) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
}

// Context mask of one utterance (campplus.py:88-93, :95-106): h [T, 128] ->
//   ctx[s] = mean_t h + mean_{t in segment s} h;  mask[s] = sigmoid(W2 relu(W1 ctx[s] + b1) + b2)   -> out [B * nseg, 32]
// Block = utterance, 256 threads: 16 channel groups of 8 x 16 frame lanes for the sums, then the tiny MLP in fp32.
__global__ void __launch_bounds__(256)
    cp_context_kernel(Planes h, int T, int P, int Tp, int nseg, const float* __restrict__ w1t, const float* __restrict__ b1,
                      const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ out) {
    __shared__ float s_part[16][128];
    __shared__ float s_seg[CP_MAX_SEG][128];  // segment sums, then ctx
    __shared__ float s_mean[128];
    __shared__ float s_hid[4][64];
    griddep_launch_dependents();
    griddep_wait();
    const int b = blockIdx.x;
    const int cg = threadIdx.x & 15, fl = threadIdx.x >> 4;
    const int64_t row0 = int64_t(b) * Tp + P;
    for (int s = 0; s < nseg; ++s) {
        const int t0 = s * CP_SEG, t1 = min(T, t0 + CP_SEG);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int t = t0 + fl; t < t1; t += 16) {
            const uint4 hv = *reinterpret_cast<const uint4*>(h.hi() + (row0 + t) * h.ld + cg * 8);
            const uint4 lv = *reinterpret_cast<const uint4*>(h.lo() + (row0 + t) * h.ld + cg * 8);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[k]));
                const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[k]));
                acc[2 * k] += hf.x + lf.x;
                acc[2 * k + 1] += hf.y + lf.y;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s_part[fl][cg * 8 + i] = acc[i];
        __syncthreads();
        if (threadIdx.x < 128) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) v += s_part[k][threadIdx.x];
            s_seg[s][threadIdx.x] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x < 128) {
        float tot = 0.f;
        for (int s = 0; s < nseg; ++s) tot += s_seg[s][threadIdx.x];
        s_mean[threadIdx.x] = tot / float(T);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nseg * 128; i += 256) {
        const int s = i >> 7, c = i & 127;
        const int cnt = min(T, (s + 1) * CP_SEG) - s * CP_SEG;
        s_seg[s][c] = s_mean[c] + s_seg[s][c] / float(cnt);
    }
    __syncthreads();
    // MLP: 4 segments at a time; thread (sl, j) computes hidden unit j of segment s0 + sl
    const int sl = threadIdx.x >> 6, j = threadIdx.x & 63;
    for (int s0 = 0; s0 < nseg; s0 += 4) {
        const int s = s0 + sl;
        if (s < nseg) {
            float a = __ldg(b1 + j);
#pragma unroll 8
            for (int c = 0; c < 128; ++c) a = fmaf(__ldg(w1t + c * 64 + j), s_seg[s][c], a);
            s_hid[sl][j] = fmaxf(a, 0.f);
        }
        __syncthreads();
        if (s < nseg && j < 32) {
            float a = __ldg(b2 + j);
#pragma unroll 8
            for (int k = 0; k < 64; ++k) a = fmaf(__ldg(w2t + k * 32 + j), s_hid[sl][k], a);
            out[(int64_t(b) * nseg + s) * 32 + j] = 1.f / (1.f + expf(-a));
        }
        __syncthreads();
    }
}

// [B, Hp, Wp, C] image -> frame-pair matrix: row b*Tp + P + (w >> 1), column (w & 1) * C*H + c*H + h
__global__ void __launch_bounds__(256) cp_flatten_pairs_kernel(Planes in, int B, int H, int W, int Hp, int Wp, int C, Planes out, int Tp, int P) {
    griddep_launch_dependents();
    griddep_wait();
    const int CH = C * H;
    const int64_t total = int64_t(B) * W * CH;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int col = int(i % CH);
        const int64_t bt = i / CH;
        const int c = col / H, hh = col % H;
        const int b = int(bt / W), w = int(bt % W);
        const int64_t src = ((int64_t(b) * Hp + hh + 1) * Wp + w + 1) * in.ld + c;
        const int64_t dst = (int64_t(b) * Tp + P + (w >> 1)) * out.ld + (w & 1) * CH + col;
        out.hi()[dst] = in.hi()[src];
        out.lo()[dst] = in.lo()[src];
    }
}

}  // namespace

struct CamppModel {
    ppv_campplus_cfg cfg;
    WeightMap raw;
    bool finalized = false;
    int precision = PPV_PREC_BF16X3;
    int num_sms = 148;
    void* arena = nullptr;
    float *stem_w = nullptr, *stem_b = nullptr;
    std::vector<ResBlockW> res;
    GemmWeights head_conv2, tdnn, dense;
    std::vector<DenseLayerW> layers;
    TransitW transit[CP_NB];
    int block_in[CP_NB], block_out[CP_NB];  // channels entering / leaving each dense block
    int head_ch = 0, final_ch = 0;
    // plan
    std::vector<CStep> steps;
    void* plan_ws = nullptr;
    int plan_B = 0, plan_T = 0, T2 = 0, Tp = 0, nseg = 0;
    ImageGeo geo[4];
    Planes stem_out, flat, tdnn_view, stats, final_x;
    Planes stage_out[4];
    Planes xblk[CP_NB], tr_out[CP_NB];
    float* emb_out = nullptr;
};

void ppv_campplus_default_cfg_impl(ppv_campplus_cfg* c) {
    c->input_size = 80;
    c->embd_dim = 192;
    c->growth_rate = 32;
    c->bn_size = 4;
    c->init_channels = 128;
    c->precision = PPV_PREC_BF16X3;
}

int campplus_create(const ppv_campplus_cfg* cfg, CamppModel** out) {
    PPV_REQUIRE(cfg && out, "campplus_create: null argument");
    if (cfg->growth_rate != 32 || cfg->bn_size != 4 || cfg->init_channels != 128)
        return fail(PPV_EUNSUPPORTED, "campplus: growth_rate 32, bn_size 4, init_channels 128 (configs/cam++.yml) are implemented");
    if (cfg->input_size % 8 || cfg->input_size < 8 || cfg->embd_dim % 32)
        return fail(PPV_EUNSUPPORTED, "campplus: input_size % 8 and embd_dim % 32 required");
    CamppModel* m = new CamppModel();
    m->cfg = *cfg;
    m->precision = cfg->precision;
    m->head_ch = 32 * (cfg->input_size / 8);
    m->num_sms = device_sm_count();
    *out = m;
    return PPV_OK;
}
void campplus_destroy(CamppModel* m) {
    if (!m) return;
    cudaFree(m->arena);
    delete m;
}
int campplus_embd_dim(const CamppModel* m) { return m->cfg.embd_dim; }
int campplus_set_precision(CamppModel* m, int precision) {
    PPV_REQUIRE(precision == PPV_PREC_BF16X3 || precision == PPV_PREC_BF16, "bad precision");
    m->precision = precision;
    return PPV_OK;
}
int campplus_load_weight(CamppModel* m, const char* name, const float* data, const int64_t* shape, int ndim) {
    PPV_REQUIRE(m, "campplus_load_weight: null model");
    if (m->finalized) return fail(PPV_ESTATE, "campplus_load_weight: model already finalized");
    return weight_map_load(&m->raw, name, data, shape, ndim);
}

// ------------------------------------------------------------------------------------------------ finalize
int campplus_finalize(CamppModel* m) {
    PPV_REQUIRE(m, "campplus_finalize: null model");
    if (m->finalized) return PPV_OK;
    ArenaBuilder ab;
    ab.wm = &m->raw;
    const ppv_campplus_cfg& cf = m->cfg;
    bool ok = true;
    const int G = cf.growth_rate, BC = cf.bn_size * cf.growth_rate;  // 32, 128

    // conv2d [N, Cin, k, k] + BN folded -> dense [N][k*k*Cin] (tap-major, taps in (dh, dw) order)
    auto conv2d_matrix = [&](GemmWeights* gw, const std::string& conv, const std::string& bn, int N, int Cin, int k) {
        const HostWeight* w = ab.get(conv + ".weight", {N, Cin, k, k});
        const HostWeight* b = ab.get(conv + ".bias", {N});
        std::vector<double> sc, sh;
        if (!w || !b || !ab.bn_affine(bn, N, &sc, &sh)) {
            ok = false;
            return;
        }
        const int taps = k * k, K = taps * Cin;
        std::vector<double> mtx(size_t(N) * K);
        std::vector<float> bias(std::max(N, 64), 0.f);
        for (int n = 0; n < N; ++n) {
            for (int t = 0; t < taps; ++t)
                for (int c = 0; c < Cin; ++c) mtx[size_t(n) * K + t * Cin + c] = double(w->v[(size_t(n) * Cin + c) * taps + t]) * sc[n];
            bias[n] = float(double(b->v[n]) * sc[n] + sh[n]);
        }
        ab.put_matrix(gw, mtx, N, K);
        ab.put_f32(&gw->bias, bias);
    };
    // conv1d [N, Cin, k] (+ optional BN folded) -> dense [N][k*Kp]; input channels beyond Cin (up to Kp) get zero weights
    auto conv1d_matrix = [&](GemmWeights* gw, const std::string& conv, const std::string& bn, int N, int Cin, int k, int Kp) {
        const HostWeight* w = ab.get(conv + ".weight", {N, Cin, k});
        const HostWeight* b = ab.get(conv + ".bias", {N});
        std::vector<double> sc(N, 1.0), sh(N, 0.0);
        if (!w || !b || (!bn.empty() && !ab.bn_affine(bn, N, &sc, &sh))) {
            ok = false;
            return;
        }
        const int K = k * Kp;
        std::vector<double> mtx(size_t(N) * K, 0.0);
        std::vector<float> bias(std::max(N, 64), 0.f);
        for (int n = 0; n < N; ++n) {
            for (int t = 0; t < k; ++t)
                for (int c = 0; c < Cin; ++c) mtx[size_t(n) * K + t * Kp + c] = double(w->v[(size_t(n) * Cin + c) * k + t]) * sc[n];
            bias[n] = float(double(b->v[n]) * sc[n] + sh[n]);
        }
        ab.put_matrix(gw, mtx, N, K);
        ab.put_f32(&gw->bias, bias);
    };
    auto bn_vectors = [&](const std::string& bn, int C, int Cp, float** scale, float** shift) {
        std::vector<double> sc, sh;
        if (!ab.bn_affine(bn, C, &sc, &sh)) {
            ok = false;
            return;
        }
        std::vector<float> s(Cp, 0.f), b(Cp, 0.f);
        for (int i = 0; i < C; ++i) {
            s[i] = float(sc[i]);
            b[i] = float(sh[i]);
        }
        ab.put_f32(scale, s);
        ab.put_f32(shift, b);
    };

    {  // head.conv1 + bn1 folded (1 -> 32 channels, CUDA-core stem)
        const HostWeight* w = ab.get("head.conv1.weight", {32, 1, 3, 3});
        const HostWeight* b = ab.get("head.conv1.bias", {32});
        std::vector<double> sc, sh;
        if (w && b && ab.bn_affine("head.bn1", 32, &sc, &sh)) {
            std::vector<float> w9(32 * 9), bb(32);
            for (int c = 0; c < 32; ++c) {
                for (int k = 0; k < 9; ++k) w9[c * 9 + k] = float(double(w->v[c * 9 + k]) * sc[c]);
                bb[c] = float(double(b->v[c]) * sc[c] + sh[c]);
            }
            ab.put_f32(&m->stem_w, w9);
            ab.put_f32(&m->stem_b, bb);
        } else {
            ok = false;
        }
    }
    m->res.clear();
    m->res.reserve(4);  // arena patches point into the elements
    for (int li = 1; li <= 2 && ok; ++li)
        for (int bi = 0; bi < 2 && ok; ++bi) {
            m->res.emplace_back();
            ResBlockW& rw = m->res.back();
            rw.stride = bi == 0 ? 2 : 1;
            rw.stage = li;
            rw.has_sc = bi == 0;
            const std::string p = "head.layer" + std::to_string(li) + "." + std::to_string(bi);
            conv2d_matrix(&rw.conv1, p + ".conv1", p + ".bn1", 32, 32, 3);
            conv2d_matrix(&rw.conv2, p + ".conv2", p + ".bn2", 32, 32, 3);
            if (rw.has_sc) conv2d_matrix(&rw.sc, p + ".shortcut.0", p + ".shortcut.1", 32, 32, 1);
        }
    if (ok) conv2d_matrix(&m->head_conv2, "head.conv2", "head.bn2", 32, 32, 3);
    if (ok) {
        // TDNN: out[t'] = sum_k w_k x[2t' + k - 2]; K-sources in tap order over the frame-pair matrix
        conv1d_matrix(&m->tdnn, "xvector.tdnn.linear", "xvector.tdnn.nonlinear.batchnorm", cf.init_channels, m->head_ch, 5, m->head_ch);
    }
    m->layers.clear();
    m->layers.reserve(CP_MAX_LAYERS);
    int channels = cf.init_channels;
    for (int bi = 0; bi < CP_NB && ok; ++bi) {
        m->block_in[bi] = channels;
        for (int li = 0; li < CP_LAYERS[bi] && ok; ++li) {
            m->layers.emplace_back();
            DenseLayerW& lw = m->layers.back();
            lw.block = bi;
            lw.dil = CP_DIL[bi];
            lw.in_ch = channels + li * G;
            lw.Kp = int(mc_align_up(size_t(lw.in_ch), 64));
            const std::string p = "xvector.block" + std::to_string(bi + 1) + ".tdnnd" + std::to_string(li + 1);
            bn_vectors(p + ".nonlinear1.batchnorm", lw.in_ch, lw.Kp, &lw.bn1_scale, &lw.bn1_shift);
            conv1d_matrix(&lw.linear1, p + ".linear1", p + ".nonlinear2.batchnorm", BC, lw.in_ch, 1, lw.Kp);
            conv1d_matrix(&lw.local, p + ".cam_layer.linear_local", "", G, BC, 3, BC);
            const HostWeight* w1 = ab.get(p + ".cam_layer.linear1.weight", {BC / 2, BC, 1});
            const HostWeight* b1 = ab.get(p + ".cam_layer.linear1.bias", {BC / 2});
            const HostWeight* w2 = ab.get(p + ".cam_layer.linear2.weight", {G, BC / 2, 1});
            const HostWeight* b2 = ab.get(p + ".cam_layer.linear2.bias", {G});
            if (!w1 || !b1 || !w2 || !b2) {
                ok = false;
                break;
            }
            std::vector<float> w1t(size_t(BC) * (BC / 2)), w2t(size_t(BC / 2) * G);
            for (int j = 0; j < BC / 2; ++j)
                for (int c = 0; c < BC; ++c) w1t[size_t(c) * (BC / 2) + j] = w1->v[size_t(j) * BC + c];
            for (int n = 0; n < G; ++n)
                for (int k = 0; k < BC / 2; ++k) w2t[size_t(k) * G + n] = w2->v[size_t(n) * (BC / 2) + k];
            ab.put_f32(&lw.w1t, w1t);
            ab.put_f32(&lw.b1, b1->v);
            ab.put_f32(&lw.w2t, w2t);
            ab.put_f32(&lw.b2, b2->v);
        }
        channels += CP_LAYERS[bi] * G;
        m->block_out[bi] = channels;
        if (!ok) break;
        TransitW& tw = m->transit[bi];
        tw.C = channels;
        const std::string p = "xvector.transit" + std::to_string(bi + 1);
        bn_vectors(p + ".nonlinear.batchnorm", channels, channels, &tw.bn_scale, &tw.bn_shift);
        // the last transit is followed directly by out_nonlinear (BN + ReLU): fold that BN into its weights
        conv1d_matrix(&tw.linear, p + ".linear", bi == CP_NB - 1 ? "xvector.out_nonlinear.batchnorm" : "", channels / 2, channels, 1, channels);
        channels /= 2;
    }
    m->final_ch = channels;
    if (ok) conv1d_matrix(&m->dense, "xvector.dense.linear", "xvector.dense.nonlinear.batchnorm", cf.embd_dim, 2 * channels, 1, 2 * channels);
    if (!ok) return fail(PPV_EINVAL, "campplus_finalize: " + (ab.err.empty() ? std::string("bad weights") : ab.err));
    int rc = ab.upload(&m->arena);
    if (rc) return rc;
    m->raw.clear();
    m->finalized = true;
    return PPV_OK;
}

// ------------------------------------------------------------------------------------------------ workspace / plan
namespace {

struct CpBuffers {
    Planes stem_out, flat, tmp, hbuf, stats;
    Planes c1[4], c2[4], sc[4], out[4];
    Planes head_out;
    Planes xblk[CP_NB], final_x;
    float *mask, *emb_out;
};

void cp_geometry(const CamppModel* m, int T, ImageGeo* geo) {
    int H = m->cfg.input_size;
    for (int l = 0; l < 4; ++l) {
        if (l > 0) H = (H - 1) / 2 + 1;
        geo[l].H = H;
        geo[l].W = T;
        geo[l].Hp = H + 2;
        geo[l].Wp = T + 2;
    }
}

void cp_carve(const CamppModel* m, WsCarver& cv, int B, int T, ImageGeo* geo, CpBuffers* cb) {
    cp_geometry(m, T, geo);
    const int T2 = (T - 1) / 2 + 1, Tp = T2 + 2 * CP_P, nseg = (T2 + CP_SEG - 1) / CP_SEG;
    const int64_t R = int64_t(B) * Tp;
    cb->stem_out = cv.planes(geo[0].rows(B), 32);
    for (size_t i = 0; i < m->res.size(); ++i) {
        const int64_t Ro = geo[m->res[i].stage].rows(B);
        cb->c1[i] = cv.planes(Ro, 32);
        cb->c2[i] = cv.planes(Ro, 32);
        if (m->res[i].has_sc) cb->sc[i] = cv.planes(Ro, 32);
        cb->out[i] = cv.planes(Ro, 32);
    }
    cb->head_out = cv.planes(geo[3].rows(B), 32);
    cb->flat = cv.planes(R, 2 * m->head_ch);
    int maxc = 0;
    for (int bi = 0; bi < CP_NB; ++bi) {
        cb->xblk[bi] = cv.planes(R, m->block_out[bi]);
        maxc = std::max(maxc, m->block_out[bi]);
    }
    cb->tmp = cv.planes(R, maxc);
    cb->hbuf = cv.planes(R, m->cfg.bn_size * m->cfg.growth_rate);
    cb->final_x = cv.planes(R, m->final_ch);
    cb->stats = cv.planes(B, 2 * m->final_ch);
    cb->mask = static_cast<float*>(cv.take(size_t(B) * nseg * m->cfg.growth_rate * sizeof(float)));
    cb->emb_out = static_cast<float*>(cv.take(mc_align_up(size_t(B), 128) * m->cfg.embd_dim * 4));
}

inline int cp_pick_bn(int N) { return (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 64; }

}  // namespace

size_t campplus_workspace_bytes(const CamppModel* m, int B, int T) {
    if (!m || !m->finalized || B <= 0 || T <= 0) return 0;
    WsCarver cv;
    ImageGeo geo[4];
    CpBuffers cb;
    cp_carve(m, cv, B, T, geo, &cb);
    return mc_align_up(cv.off, 256);
}

static int cp_build_plan(CamppModel* m, int B, int T, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = campplus_workspace_bytes(m, B, T);
    PPV_REQUIRE(ws && ws_bytes >= need, "campplus: workspace too small (see ppv_model_workspace_bytes)");
    PPV_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "campplus: workspace must be 256-byte aligned");
    PPV_REQUIRE(T >= 3, "campplus: too few frames (the statistics pooling needs at least two frames after the stride-2 TDNN)");
    const int T2 = (T - 1) / 2 + 1, Tp = T2 + 2 * CP_P, nseg = (T2 + CP_SEG - 1) / CP_SEG;
    PPV_REQUIRE(nseg <= CP_MAX_SEG, "campplus: utterance too long (more than 64 context segments of 100 frames)");
    PPV_REQUIRE(T + 3 < 32768, "campplus: utterance too long for 16-bit tap offsets");
    WsCarver cv;
    cv.base = static_cast<uint8_t*>(ws);
    CpBuffers cb;
    cp_carve(m, cv, B, T, m->geo, &cb);
    PPV_REQUIRE(m->geo[0].rows(B) < (int64_t(1) << 31), "campplus: batch too large for 32-bit row indices");
    PPV_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));  // zero borders, zero padding rows, zero padded columns
    m->steps.clear();
    const int M2 = int(int64_t(B) * Tp);

    auto img_epi = [&](const Planes& out, const ImageGeo& gin, const ImageGeo& gout, int stride_h) {
        Epilogue ep;
        ep.out_mode = OUT_PLANES;
        ep.out = out.base;
        ep.out_ld = out.ld;
        ep.out_plane_stride = out.plane_stride;
        ep.img_Hp = gin.Hp;
        ep.img_Wp = gin.Wp;
        ep.img_H = gin.H;
        ep.img_W = gin.W;
        ep.img_stride = stride_h;
        ep.img_stride_w = 1;
        ep.out_Hp = gout.Hp;
        ep.out_Wp = gout.Wp;
        return ep;
    };
    auto time_epi = [&](const Planes& out, int col0) {
        Epilogue ep;
        ep.out_mode = OUT_PLANES;
        ep.out = out.base;
        ep.out_ld = out.ld;
        ep.out_plane_stride = out.plane_stride;
        ep.out_col0 = col0;
        ep.Tp = Tp;
        ep.P = CP_P;
        ep.T = T2;
        ep.zero_invalid = 1;
        return ep;
    };
    auto add_gemm = [&](const GemmWeights& gw, const std::vector<GemmSource>& srcs, int M, Epilogue ep) -> int {
        ep.bias = gw.bias;
        if (pointwise_enabled() && ep.img_Wp > 0 && pointwise_supported(srcs.data(), int(srcs.size()), gw.N, ep)) {
            CStep sp;
            sp.kind = CStep::PW;
            for (size_t i = 0; i < srcs.size(); ++i) sp.pw.srcs[i] = srcs[i];
            sp.pw.nsrc = int(srcs.size());
            sp.pw.N = gw.N;
            sp.pw.M = M;
            sp.pw.W = gw.W;
            sp.pw.ep = ep;
            m->steps.push_back(sp);
            return PPV_OK;
        }
        CStep s;
        s.kind = CStep::GEMM;
        s.BN = cp_pick_bn(gw.N);
        int bk = 64;
        for (const GemmSource& g : srcs)
            if (g.ncols % 64) bk = 32;
        int rc = gemm_build(&s.gp, srcs.data(), int(srcs.size()), gw.W, M, gw.N, ep, s.BN, bk);
        if (rc) return rc;
        m->steps.push_back(s);
        return PPV_OK;
    };
    auto taps9 = [&](const Planes& p, const ImageGeo& g, std::vector<GemmSource>* v) {
        for (int dh = -1; dh <= 1; ++dh)
            for (int dw = -1; dw <= 1; ++dw) v->push_back(GemmSource{p, 0, 32, dh * g.Wp + dw});
    };
    const char* c3env = getenv("PPV_CONV3X3");  // 0 = run the 3x3 convs through the generic gather-GEMM (debugging / A-B timing)
    const bool use_c3 = !(c3env && c3env[0] == '0');
    // 3x3 conv, 32 -> 32 channels, on grid `g` of buffer `p`: the patch kernel (conv3x3.cu), else nine taps of the gather-GEMM
    auto add_conv3 = [&](const GemmWeights& gw, const Planes& p, const ImageGeo& g, Epilogue ep) -> int {
        if (use_c3 && gw.N == 32 && conv3x3_c32_supported(32, gw.N, g.H, g.W)) {
            ep.bias = gw.bias;
            CStep s;
            s.kind = CStep::CONV3;
            int rc = conv3x3_build(&s.c3, p, 0, gw.W, B, g.H, g.W, g.Hp, g.Wp, ep);
            if (rc) return rc;
            m->steps.push_back(s);
            return PPV_OK;
        }
        std::vector<GemmSource> t;
        taps9(p, g, &t);
        return add_gemm(gw, t, int(g.rows(B)), ep);
    };
    auto relu = [](Epilogue ep) {
        ep.relu = 1;
        return ep;
    };
    int rc;
    // ---- FCM head
    {
        CStep s;
        s.kind = CStep::STEM;
        m->steps.push_back(s);
    }
    Planes x = cb.stem_out;
    for (size_t i = 0; i < m->res.size(); ++i) {
        const ResBlockW& rw = m->res[i];
        const ImageGeo& gin = m->geo[rw.stride == 2 ? rw.stage - 1 : rw.stage];
        const ImageGeo& go = m->geo[rw.stage];
        rc = add_conv3(rw.conv1, x, gin, relu(img_epi(cb.c1[i], gin, go, rw.stride)));
        if (rc) return rc;
        rc = add_conv3(rw.conv2, cb.c1[i], go, img_epi(cb.c2[i], go, go, 1));
        if (rc) return rc;
        Planes resid = x;
        if (rw.has_sc) {
            rc = add_gemm(rw.sc, {GemmSource{x, 0, 32, 0}}, int(gin.rows(B)), img_epi(cb.sc[i], gin, go, rw.stride));
            if (rc) return rc;
            resid = cb.sc[i];
        }
        CStep s;
        s.kind = CStep::ADD_RELU;
        s.a = cb.c2[i];
        s.b = resid;
        s.d = cb.out[i];
        s.C = 32;
        s.img_rows = go.Hp * go.Wp;
        s.rows = go.rows(B);
        m->steps.push_back(s);
        x = cb.out[i];
        m->stage_out[rw.stage] = x;
    }
    {
        rc = add_conv3(m->head_conv2, x, m->geo[2], relu(img_epi(cb.head_out, m->geo[2], m->geo[3], 2)));
        if (rc) return rc;
        CStep s;
        s.kind = CStep::FLATTEN;
        s.a = cb.head_out;
        s.d = cb.flat;
        m->steps.push_back(s);
    }
    // ---- TDNN (k5, stride 2) over the frame-pair matrix -> first 128 columns of block 1's buffer
    {
        const int HC = m->head_ch;
        std::vector<GemmSource> srcs = {GemmSource{cb.flat, 0, HC, -1}, GemmSource{cb.flat, HC, HC, -1}, GemmSource{cb.flat, 0, HC, 0},
                                        GemmSource{cb.flat, HC, HC, 0}, GemmSource{cb.flat, 0, HC, 1}};
        rc = add_gemm(m->tdnn, srcs, M2, relu(time_epi(cb.xblk[0], 0)));
        if (rc) return rc;
    }
    // ---- dense blocks
    size_t li = 0;
    for (int bi = 0; bi < CP_NB; ++bi) {
        const Planes& xb = cb.xblk[bi];
        for (int l = 0; l < CP_LAYERS[bi]; ++l, ++li) {
            const DenseLayerW& lw = m->layers[li];
            CStep s;
            s.kind = CStep::BN_RELU;
            s.a = xb;
            s.d = cb.tmp;
            s.p0 = lw.bn1_scale;
            s.p1 = lw.bn1_shift;
            s.C = lw.Kp;
            s.rows = M2;
            m->steps.push_back(s);
            rc = add_gemm(lw.linear1, {GemmSource{cb.tmp, 0, lw.Kp, 0}}, M2, relu(time_epi(cb.hbuf, 0)));
            if (rc) return rc;
            CStep c;
            c.kind = CStep::CONTEXT;
            c.a = cb.hbuf;
            c.p0 = lw.w1t;
            c.p1 = lw.b1;
            c.p2 = lw.w2t;
            c.p3 = lw.b2;
            c.fout = cb.mask;
            m->steps.push_back(c);
            Epilogue ep = time_epi(xb, lw.in_ch);
            ep.seg_scale = cb.mask;
            ep.seg_len = CP_SEG;
            ep.nseg = nseg;
            const int BCc = cb.hbuf.ld;
            rc = add_gemm(lw.local, {GemmSource{cb.hbuf, 0, BCc, -lw.dil}, GemmSource{cb.hbuf, 0, BCc, 0}, GemmSource{cb.hbuf, 0, BCc, lw.dil}}, M2, ep);
            if (rc) return rc;
        }
        const TransitW& tw = m->transit[bi];
        CStep s;
        s.kind = CStep::BN_RELU;
        s.a = xb;
        s.d = cb.tmp;
        s.p0 = tw.bn_scale;
        s.p1 = tw.bn_shift;
        s.C = tw.C;
        s.rows = M2;
        m->steps.push_back(s);
        const bool last = bi == CP_NB - 1;
        Epilogue ep = time_epi(last ? cb.final_x : cb.xblk[bi + 1], 0);
        if (last) ep.relu = 1;  // out_nonlinear (BN folded into the transit weights) + ReLU
        rc = add_gemm(tw.linear, {GemmSource{cb.tmp, 0, tw.C, 0}}, M2, ep);
        if (rc) return rc;
        m->tr_out[bi] = last ? cb.final_x : cb.xblk[bi + 1];
    }
    {
        CStep s;
        s.kind = CStep::STATS;
        m->steps.push_back(s);
        Epilogue ep;
        ep.out_mode = OUT_F32;
        ep.out = cb.emb_out;
        ep.out_ld = m->cfg.embd_dim;
        rc = add_gemm(m->dense, {GemmSource{cb.stats, 0, 2 * m->final_ch, 0}}, B, ep);
        if (rc) return rc;
    }
    m->stem_out = cb.stem_out;
    m->flat = cb.flat;
    m->stats = cb.stats;
    m->final_x = cb.final_x;
    for (int bi = 0; bi < CP_NB; ++bi) m->xblk[bi] = cb.xblk[bi];
    m->emb_out = cb.emb_out;
    m->T2 = T2;
    m->Tp = Tp;
    m->nseg = nseg;
    m->plan_ws = ws;
    m->plan_B = B;
    m->plan_T = T;
    return PPV_OK;
}

// ------------------------------------------------------------------------------------------------ forward
int campplus_forward(CamppModel* m, const float* feat, int B, int T, float* emb, void* ws, size_t ws_bytes, cudaStream_t st) {
    PPV_REQUIRE(m && feat && emb, "campplus_forward: null argument");
    if (!m->finalized) return fail(PPV_ESTATE, "campplus_forward: call ppv_model_finalize first");
    PPV_REQUIRE(B > 0 && T > 0, "campplus_forward: empty batch");
    if (m->plan_ws != ws || m->plan_B != B || m->plan_T != T) {
        int rc = cp_build_plan(m, B, T, ws, ws_bytes, st);
        if (rc) {
            m->plan_ws = nullptr;
            return rc;
        }
    }
    int rc = PPV_OK;
    for (const CStep& s : m->steps) {
        switch (s.kind) {
            case CStep::STEM:
                rc = launch_stem_conv(feat, B, T, m->cfg.input_size, m->stem_w, m->stem_b, 32, m->stem_out, m->geo[0].Hp, m->geo[0].Wp, st);
                break;
            case CStep::GEMM: rc = gemm_launch(s.gp, s.BN, m->precision, m->num_sms, st); break;
            case CStep::CONV3: rc = conv3x3_launch(s.c3, m->precision, m->num_sms, st); break;
            case CStep::PW: rc = pointwise_launch(s.pw, m->num_sms, st); break;
            case CStep::ADD_RELU: rc = launch_se_scale_res(s.a, nullptr, s.b, 0, s.d, 0, s.C, s.img_rows, s.rows, m->num_sms, st, 1, 0.f); break;
            case CStep::FLATTEN: {
                const ImageGeo& g = m->geo[3];
                const int64_t total = int64_t(B) * g.W * 32 * g.H;
                const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(m->num_sms) * 16));
                PPV_PDL_OK(launch_pdl(cp_flatten_pairs_kernel, dim3(grid), dim3(256), 0, st, s.a, B, g.H, g.W, g.Hp, g.Wp, 32, s.d, m->Tp, CP_P),
                           "cp_flatten_pairs_kernel");
                break;
            }
            case CStep::BN_RELU: {
                const int64_t total = s.rows * (s.C / 8);
                const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(m->num_sms) * 16));
                PPV_PDL_OK(launch_pdl(cp_bn_relu_kernel, dim3(grid), dim3(256), 0, st, s.a, s.p0, s.p1, s.d, s.C, s.rows), "cp_bn_relu_kernel");
                break;
            }
            case CStep::CONTEXT:
                PPV_PDL_OK(launch_pdl(cp_context_kernel, dim3(B), dim3(256), 0, st, s.a, m->T2, CP_P, m->Tp, m->nseg, s.p0, s.p1, s.p2, s.p3, s.fout),
                           "cp_context_kernel");
                break;
            case CStep::STATS:
                rc = launch_colstats(m->final_x, 0, m->final_ch, B, m->T2, CP_P, m->Tp, 2, 0.f, nullptr, m->stats, st);
                break;
        }
        if (rc) return rc;
    }
    PPV_CUDA_OK(cudaMemcpyAsync(emb, m->emb_out, size_t(B) * m->cfg.embd_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return PPV_OK;
}

// taps: "head.layer1", "head.layer2" -> fp32 [B,H,W,32]; "tdnn" [B,T2,128]; "block1".."block3" [B,T2,C]; "transit1", "transit2"
// [B,T2,C/2]; "out_nonlinear" [B,T2,512] (transit3 + BN + ReLU); "stats" [B, 2*512]
int campplus_read_tap(CamppModel* m, const char* name, float* out, size_t out_elems, cudaStream_t st) {
    PPV_REQUIRE(m && name && out, "campplus_read_tap: null argument");
    if (!m->plan_ws) return fail(PPV_ESTATE, "campplus_read_tap: no forward has run");
    const std::string n(name);
    const int B = m->plan_B;
    if (n == "stats") {
        PPV_REQUIRE(out_elems >= size_t(B) * 2 * m->final_ch, "campplus_read_tap: output too small");
        return launch_planes_to_f32(m->stats, 0, 2 * m->final_ch, B, 1, 0, 1, out, st);
    }
    if (n == "head.layer1" || n == "head.layer2") {
        const int stage = n.back() - '0';
        const ImageGeo& g = m->geo[stage];
        PPV_REQUIRE(out_elems >= size_t(B) * g.H * g.W * 32, "campplus_read_tap: output too small");
        return launch_image_to_f32(m->stage_out[stage], B, g.H, g.W, g.Hp, g.Wp, 32, out, st);
    }
    Planes src;
    int C = 0;
    if (n == "tdnn") {
        src = m->xblk[0];
        C = m->cfg.init_channels;
    } else if (n.rfind("block", 0) == 0 && n.size() == 6 && n[5] >= '1' && n[5] <= '3') {
        src = m->xblk[n[5] - '1'];
        C = m->block_out[n[5] - '1'];
    } else if (n.rfind("transit", 0) == 0 && n.size() == 8 && n[7] >= '1' && n[7] <= '2') {
        src = m->tr_out[n[7] - '1'];
        C = m->block_out[n[7] - '1'] / 2;
    } else if (n == "out_nonlinear") {
        src = m->final_x;
        C = m->final_ch;
    } else {
        return fail(PPV_EINVAL, "campplus_read_tap: unknown tap " + n);
    }
    PPV_REQUIRE(out_elems >= size_t(B) * m->T2 * C, "campplus_read_tap: output too small");
    return launch_planes_to_f32(src, 0, C, B, m->T2, CP_P, m->Tp, out, st);
}

}  // namespace ppv
