// Host-side helpers shared by the model plans: named fp32 weights as loaded through the C ABI, an arena image that is
// built on the host (split-bf16 weight planes, fp32 vectors) and uploaded once, and the workspace carver.
#pragma once
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace ppv {

inline size_t mc_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct HostWeight {
    std::vector<float> v;
    std::vector<int64_t> shape;
};
using WeightMap = std::map<std::string, HostWeight>;

// Copies `data` (host or device fp32) into the map.
inline int weight_map_load(WeightMap* wm, const char* name, const float* data, const int64_t* shape, int ndim) {
    PPV_REQUIRE(wm && name && data && shape && ndim >= 1 && ndim <= 4, "load_weight: bad argument");
    HostWeight w;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        w.shape.push_back(shape[i]);
        n *= shape[i];
    }
    w.v.resize(size_t(n));
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, data);
    if (e == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
        PPV_CUDA_OK(cudaMemcpy(w.v.data(), data, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        cudaGetLastError();
        memcpy(w.v.data(), data, size_t(n) * sizeof(float));
    }
    (*wm)[name] = std::move(w);
    return PPV_OK;
}

struct GemmWeights {  // one conv / linear layer prepared for the gather-GEMM
    Planes W;         // [2][Npad][Ktot] split-bf16
    float* bias = nullptr;
    int N = 0, Ktot = 0;
};

struct ArenaBuilder {
    const WeightMap* wm = nullptr;
    std::vector<uint8_t> host;
    std::string err;
    struct Patch {
        void** dst;
        size_t off;
    };
    std::vector<Patch> patches;

    size_t reserve(size_t bytes) {
        const size_t off = mc_align_up(host.size(), 256);
        host.resize(off + bytes, 0);
        return off;
    }
    const HostWeight* get(const std::string& name, const std::vector<int64_t>& shape) {
        auto it = wm->find(name);
        if (it == wm->end()) {
            if (err.empty()) err = "missing weight " + name;
            return nullptr;
        }
        if (it->second.shape != shape) {
            if (err.empty()) err = "weight " + name + " has the wrong shape";
            return nullptr;
        }
        return &it->second;
    }
    template <typename T>
    void put_f32(T** dst, const std::vector<float>& v) {
        const size_t off = reserve(v.size() * sizeof(float));
        memcpy(host.data() + off, v.data(), v.size() * sizeof(float));
        patches.push_back({reinterpret_cast<void**>(dst), off});
    }
    // dense fp32 [N][K] (row-major) -> split planes [2][Npad][K]; rows N..Npad stay zero
    void put_matrix(GemmWeights* gw, const std::vector<double>& m, int N, int K, int n_align = 256) {
        const int Npad = int(mc_align_up(size_t(N), size_t(n_align)));
        const size_t plane = size_t(Npad) * K;
        const size_t off = reserve(2 * plane * sizeof(__nv_bfloat16));
        __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(host.data() + off);
        __nv_bfloat16* lo = hi + plane;
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        for (size_t i = 0; i < 2 * plane; ++i) hi[i] = z;
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const float x = float(m[size_t(n) * K + k]);
                const __nv_bfloat16 h = __float2bfloat16_rn(x);
                hi[size_t(n) * K + k] = h;
                lo[size_t(n) * K + k] = __float2bfloat16_rn(x - __bfloat162float(h));
            }
        gw->N = N;
        gw->Ktot = K;
        gw->W.rows = Npad;
        gw->W.ld = K;
        gw->W.plane_stride = int64_t(plane);
        patches.push_back({reinterpret_cast<void**>(&gw->W.base), off});
    }
    // BatchNorm(eval) as y = x * scale + shift (eps 1e-5), in double
    bool bn_affine(const std::string& prefix, int C, std::vector<double>* scale, std::vector<double>* shift) {
        const HostWeight *g = get(prefix + ".weight", {C}), *b = get(prefix + ".bias", {C}), *mu = get(prefix + "._mean", {C}),
                         *var = get(prefix + "._variance", {C});
        if (!g || !b || !mu || !var) return false;
        scale->resize(C);
        shift->resize(C);
        for (int i = 0; i < C; ++i) {
            const double s = double(g->v[i]) / sqrt(double(var->v[i]) + 1e-5);
            (*scale)[i] = s;
            (*shift)[i] = double(b->v[i]) - double(mu->v[i]) * s;
        }
        return true;
    }
    int upload(void** arena) {
        PPV_CUDA_OK(cudaMalloc(arena, host.size()));
        PPV_CUDA_OK(cudaMemcpy(*arena, host.data(), host.size(), cudaMemcpyHostToDevice));
        for (const Patch& p : patches) *p.dst = static_cast<uint8_t*>(*arena) + p.off;
        return PPV_OK;
    }
};

struct WsCarver {
    uint8_t* base = nullptr;
    size_t off = 0;
    void* take(size_t bytes) {
        off = mc_align_up(off, 256);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
    Planes planes(int64_t rows, int ld) {
        Planes p;
        p.rows = int64_t(mc_align_up(size_t(rows), 128));
        p.ld = ld;
        p.plane_stride = p.rows * ld;
        p.base = static_cast<__nv_bfloat16*>(take(size_t(2) * p.plane_stride * sizeof(__nv_bfloat16)));
        return p;
    }
};

}  // namespace ppv
