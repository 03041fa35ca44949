// Pieces shared by the TMA-staged selective-scan kernels (scan_fwd_tma.cu, scan_bwd_tma.cu).
#pragma once
#include "scan_common.cuh"
#include "tma.cuh"

namespace vmb {

struct ScanTmaMaps {
    CUtensorMap u[4], d[4], b[4], c[4];  // plain operator: index 0 only (B/C group = z coordinate); grouped: one per group
};

// raw B or C rows (dense [16][CHUNK] box) -> fp32 pair layout float4 = (X[n][l], X[n+1][l], X[n][l+1], X[n+1][l+1])
template <typename in_t, int RB, int NT>
__device__ __forceinline__ void convert_bc_dense(float4* __restrict__ dst, const in_t* __restrict__ rawX, int tid, bool rev) {
    using Cfg = FwdCfg<RB>;
    constexpr int CHUNK = Cfg::CHUNK, V = Vec<in_t>::N, LG = CHUNK / V;
#pragma unroll
    for (int it = tid; it < 8 * LG; it += NT) {
        const int np = it / LG, l = (it % LG) * V;
        float f0[V], f1[V];
        load_vec_smem<in_t>(rawX + (2 * np) * CHUNK + l, f0);
        load_vec_smem<in_t>(rawX + (2 * np + 1) * CHUNK + l, f1);
        if (!rev) {
            float4* d = dst + np * Cfg::SLOTS + (l / T) * Cfg::SEGQ + (l % T) / 2;
#pragma unroll
            for (int i = 0; i < V / 2; ++i) d[i] = make_float4(f0[2 * i], f1[2 * i], f0[2 * i + 1], f1[2 * i + 1]);
        } else {  // raw index i <-> sequence position CHUNK-1-i
            const int s_hi = CHUNK - 1 - l;
#pragma unroll
            for (int j = 0; j < V / 2; ++j) {
                const int s = s_hi - 1 - 2 * j;
                dst[np * Cfg::SLOTS + (s / T) * Cfg::SEGQ + (s % T) / 2] = make_float4(f0[2 * j + 1], f1[2 * j + 1], f0[2 * j], f1[2 * j]);
            }
        }
    }
}

// softplus of two values: max(x,0) + log1p(e), e = exp(-|x|) in (0,1]  (== the reference's x<=20 ? log1p(exp(x)) : x to fp32
// rounding, selective_scan_fwd_kernel.cuh:117).  log1p(e): degree-6 series below 2^-3 (rel. error < 5e-7), ln2*lg2(1+e) above.
__device__ __forceinline__ float2 log1p_series2(float2 e) {
    float2 q = fma2(e, make_float2(-0.16666667f, -0.16666667f), make_float2(0.2f, 0.2f));
    q = fma2(q, e, make_float2(-0.25f, -0.25f));
    q = fma2(q, e, make_float2(0.33333334f, 0.33333334f));
    q = fma2(q, e, make_float2(-0.5f, -0.5f));
    q = fma2(q, e, make_float2(1.f, 1.f));
    return mul2(q, e);
}


// dt[t] <- softplus(dt[t] + bias) for the T positions of a lane (packed pairs; the lg2 MUFU is skipped when every exp(-|x|) of the
// warp is below 2^-3), or dt[t] + bias without softplus
__device__ __forceinline__ void softplus_block(float* dt, float bias, bool softplus) {
    if (softplus) {
        float e[T];
        float emax = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            dt[t] += bias;
            e[t] = ex2(-fabsf(dt[t]) * kLog2e);
            emax = fmaxf(emax, e[t]);
        }
        if (__all_sync(0xffffffffu, emax < 0.125f)) {
#pragma unroll
            for (int t = 0; t < T; t += 2) {
                const float2 lp = log1p_series2(make_float2(e[t], e[t + 1]));
                dt[t] = fmaxf(dt[t], 0.f) + lp.x;
                dt[t + 1] = fmaxf(dt[t + 1], 0.f) + lp.y;
            }
        } else {
#pragma unroll
            for (int t = 0; t < T; t += 2) {
                const float2 ee = make_float2(e[t], e[t + 1]);
                const float2 sm = log1p_series2(ee);
                const float2 w = add2(ee, make_float2(1.f, 1.f));
                const float2 bg = mul2(make_float2(lg2(w.x), lg2(w.y)), make_float2(kLn2, kLn2));
                dt[t] = fmaxf(dt[t], 0.f) + (e[t] < 0.125f ? sm.x : bg.x);
                dt[t + 1] = fmaxf(dt[t + 1], 0.f) + (e[t + 1] < 0.125f ? sm.y : bg.y);
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < T; ++t) dt[t] += bias;
    }
}

}  // namespace vmb
