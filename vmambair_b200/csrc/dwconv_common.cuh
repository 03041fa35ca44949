// Depthwise 3x3 taps shared by the forward (oss_ops.cu) and backward (train_ops.cu) kernels.
#pragma once
#include "common.cuh"

namespace vmb {

template <typename in_t>
__device__ __forceinline__ float dw_at(const in_t* __restrict__ xc, const float* __restrict__ w9, int h, int w, int H, int W) {
    float acc = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int hh = h + dy;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int ww = w + dx;
            if (ww < 0 || ww >= W) continue;
            acc = fmaf(w9[(dy + 1) * 3 + dx + 1], to_f32<in_t>(xc[hh * W + ww]), acc);
        }
    }
    return acc;
}

// strip version: one thread = 8 consecutive pixels of one row (16 B vector loads of the 3 input rows + 2 halo scalars)
template <typename in_t>
__device__ __forceinline__ void dw_strip(const in_t* __restrict__ xc, const float* __restrict__ w9, int h, int w0, int H, int W,
                                         float bias, float* acc) {
    constexpr int V = Vec<in_t>::N;  // 8 (16-bit) or 4 (fp32): two vectors for fp32
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int hh = h + dy;
        if (hh < 0 || hh >= H) continue;
        const in_t* __restrict__ row = xc + (int64_t)hh * W;
        float v[10];
        v[0] = w0 > 0 ? to_f32<in_t>(row[w0 - 1]) : 0.f;
        v[9] = w0 + 8 < W ? to_f32<in_t>(row[w0 + 8]) : 0.f;
#pragma unroll
        for (int j = 0; j < 8 / V; ++j) load_vec<in_t>(row + w0 + j * V, v + 1 + j * V, V, true);
        const float k0 = w9[(dy + 1) * 3], k1 = w9[(dy + 1) * 3 + 1], k2 = w9[(dy + 1) * 3 + 2];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(k2, v[i + 2], fmaf(k1, v[i + 1], fmaf(k0, v[i], acc[i])));
    }
}


}  // namespace vmb
