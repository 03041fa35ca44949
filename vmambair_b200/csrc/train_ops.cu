// Backward stages of the fused OSS block (training path): the autograd of
//   LayerNorm (norm1 / norm2; reference SRGAN/VmambaIR/archs/MambaSISR6_arch.py:166-195),
//   out_norm + SiLU(z) gate + AdaptiveAvgPool2d sums (:433-441,493),
//   depthwise 3x3 + SiLU (:490-491) and depthwise 3x3 + exact-GELU gate (FeedForward :215-216),
//   the channel gate y*(1+c) / y+c in front of out_conv (:494-496)
// as hand-written kernels.  Activations are NCHW, pixel-contiguous (B, C, L) views addressed by (batch, channel) strides.
// Per-pixel reductions over channels run one thread per pixel (coalesced along L); per-channel reductions over pixels (the
// parameter gradients) run one CTA per (channel, image) with a block reduction and ONE fp32 atomicAdd per CTA.
// The 1x1-conv data gradients reuse vmb_pixlin with the transposed weight; the scan gradient is vmb_selective_scan_bwd.
#include "common.cuh"
#include "train_params.h"

namespace vmb {

__device__ __forceinline__ float sigmoid_t(float v) { return rcp_approx(1.f + ex2(-v * kLog2e)); }
__device__ __forceinline__ float silu_t(float v) { return v * sigmoid_t(v); }
__device__ __forceinline__ float dsilu_t(float v) {  // d/dv v*sigmoid(v)
    const float s = sigmoid_t(v);
    return s * fmaf(v, 1.f - s, 1.f);
}
__device__ __forceinline__ float erf_t(float x) {  // Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7 (same as the forward kernel)
    const float ax = fabsf(x);
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float r = 1.f - poly * t * ex2(-ax * ax * kLog2e);
    return copysignf(r, x);
}
__device__ __forceinline__ float gelu_t(float v) { return 0.5f * v * (1.f + erf_t(v * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_t(float v) {  // Phi(v) + v*phi(v)
    const float cdf = 0.5f * (1.f + erf_t(v * 0.70710678118654752f));
    return fmaf(v * 0.3989422804014327f, ex2(-0.5f * v * v * kLog2e), cdf);
}

__device__ __forceinline__ float block_sum_256(float v, float* sred) {  // 256 threads; result valid in thread 0
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sred[i];
    }
    __syncthreads();
    return t;
}

// ------------------------------------------------------------------------------------------ LayerNorm forward (materialised)
// y = (x - mu) * rstd * w + b  (mode 1, WithBias)   or   x * rstd * w  (mode 2, BiasFree; variance about the mean), two-pass.
template <typename in_t>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const LnFwdParams p) {
    pdl_trigger();
    pdl_wait();
    const int l = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
    if (l >= p.L) return;
    const in_t* __restrict__ x = reinterpret_cast<const in_t*>(p.x) + (int64_t)b * p.x_bs + l;
    float s = 0.f;
    for (int c = 0; c < p.C; ++c) s += to_f32<in_t>(x[(int64_t)c * p.x_cs]);
    const float mu = s / p.C;
    float v = 0.f;
    for (int c = 0; c < p.C; ++c) {
        const float d = to_f32<in_t>(x[(int64_t)c * p.x_cs]) - mu;
        v = fmaf(d, d, v);
    }
    const float rstd = rsqrtf(v / p.C + 1e-5f);
    if (p.stats) {
        p.stats[((int64_t)b * p.L + l) * 2] = mu;
        p.stats[((int64_t)b * p.L + l) * 2 + 1] = rstd;
    }
    if (!p.y) return;
    in_t* __restrict__ y = reinterpret_cast<in_t*>(p.y) + (int64_t)b * p.y_bs + l;
    for (int c = 0; c < p.C; ++c) {
        const float xv = to_f32<in_t>(x[(int64_t)c * p.x_cs]);
        const float o = p.mode == 1 ? fmaf((xv - mu) * rstd, p.w[c], p.b[c]) : xv * rstd * p.w[c];
        y[(int64_t)c * p.y_cs] = from_f32<in_t>(o);
    }
}

// ------------------------------------------------------------------------------------------ LayerNorm backward
// dx kernel: one thread per pixel.  mode 1: dx = rstd (gw - mean(gw) - xh mean(gw xh));  mode 2 (y = x rstd w):
// dx = rstd gw - rstd^3 (x - mu) mean(gw x).   gw = g*w, xh = (x - mu) rstd.   `add` (optional) is summed into dx
// (the residual branch's gradient).  Also writes (mu, rstd) per pixel for the parameter-gradient kernel.
template <typename in_t>
__global__ void __launch_bounds__(128) ln_bwd_dx_kernel(const LnBwdParams p) {
    pdl_trigger();
    pdl_wait();
    const int l = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
    if (l >= p.L) return;
    const in_t* __restrict__ x = reinterpret_cast<const in_t*>(p.x) + (int64_t)b * p.x_bs + l;
    const in_t* __restrict__ g = reinterpret_cast<const in_t*>(p.g) + (int64_t)b * p.g_bs + l;
    float s = 0.f;
    for (int c = 0; c < p.C; ++c) s += to_f32<in_t>(x[(int64_t)c * p.x_cs]);
    const float mu = s / p.C;
    float v = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < p.C; ++c) {
        const float xv = to_f32<in_t>(x[(int64_t)c * p.x_cs]);
        const float d = xv - mu;
        v = fmaf(d, d, v);
        const float gw = to_f32<in_t>(g[(int64_t)c * p.g_cs]) * p.w[c];
        s1 += gw;
        s2 = fmaf(gw, p.mode == 1 ? d : xv, s2);
    }
    const float rstd = rsqrtf(v / p.C + 1e-5f);
    p.stats[((int64_t)b * p.L + l) * 2] = mu;
    p.stats[((int64_t)b * p.L + l) * 2 + 1] = rstd;
    const float invC = 1.f / p.C;
    // mode 1: s2 = sum gw (x-mu) -> mean(gw xh) = s2 rstd / C ; term = xh * that = (x-mu) rstd^2 s2 / C
    const float k1 = s1 * invC, k2 = s2 * invC * rstd * rstd;
    const in_t* __restrict__ a = p.add ? reinterpret_cast<const in_t*>(p.add) + (int64_t)b * p.a_bs + l : nullptr;
    in_t* __restrict__ dx = reinterpret_cast<in_t*>(p.dx) + (int64_t)b * p.dx_bs + l;
    for (int c = 0; c < p.C; ++c) {
        const float xv = to_f32<in_t>(x[(int64_t)c * p.x_cs]);
        const float gw = to_f32<in_t>(g[(int64_t)c * p.g_cs]) * p.w[c];
        float o = p.mode == 1 ? rstd * (gw - k1 - (xv - mu) * k2) : rstd * (gw - (xv - mu) * k2);
        if (a) o += to_f32<in_t>(a[(int64_t)c * p.a_cs]);
        dx[(int64_t)c * p.dx_cs] = from_f32<in_t>(o);
    }
}

// parameter gradients: grid (C, B); dw[c] += sum_l g xh (mode 1) or g x rstd (mode 2); db[c] += sum_l g
template <typename in_t>
__global__ void __launch_bounds__(256) ln_bwd_dwdb_kernel(const LnBwdParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sred[8];
    const int c = blockIdx.x, b = blockIdx.y;
    const in_t* __restrict__ x = reinterpret_cast<const in_t*>(p.x) + (int64_t)b * p.x_bs + (int64_t)c * p.x_cs;
    const in_t* __restrict__ g = reinterpret_cast<const in_t*>(p.g) + (int64_t)b * p.g_bs + (int64_t)c * p.g_cs;
    const float2* __restrict__ st = reinterpret_cast<const float2*>(p.stats) + (int64_t)b * p.L;
    float aw = 0.f, ab = 0.f;
    for (int l = threadIdx.x; l < p.L; l += 256) {
        const float2 s = st[l];
        const float gv = to_f32<in_t>(g[l]), xv = to_f32<in_t>(x[l]);
        aw = fmaf(gv, (p.mode == 1 ? xv - s.x : xv) * s.y, aw);
        ab += gv;
    }
    const float tw = block_sum_256(aw, sred), tb = block_sum_256(ab, sred);
    if (threadIdx.x == 0) {
        atomicAdd(p.dw + c, tw);
        if (p.db) atomicAdd(p.db + c, tb);
    }
}

// ------------------------------------------------------------------------------------------ out_norm + gate backward
// forward (merge kernels): m = merged scan output (fp32), n = (m - mu) rstd w + b, y2 = n * silu(z), pooled[b,c] = sum_l y2.
// given dy2 and dpooled:  t = dy2 + dpooled[b,c];  dn = t silu(z);  dz = t n silu'(z);  dm = LayerNorm backward of dn.
template <typename in_t>
__global__ void __launch_bounds__(128) merge_bwd_dx_kernel(const MergeBwdParams p) {
    pdl_trigger();
    pdl_wait();
    const int l = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
    if (l >= p.L) return;
    const int C = p.C;
    const float* __restrict__ m = p.merged + (int64_t)b * C * p.L + l;
    const float2 st = reinterpret_cast<const float2*>(p.stats)[(int64_t)b * p.L + l];
    const float invC = 1.f / C;
    const float mu = st.x * invC;
    const float rstd = rsqrtf(fmaxf(st.y * invC - mu * mu, 0.f) + 1e-5f);
    const in_t* __restrict__ z = reinterpret_cast<const in_t*>(p.z) + (int64_t)b * p.z_bs + l;
    const in_t* __restrict__ dy2 = reinterpret_cast<const in_t*>(p.dy2) + (int64_t)b * C * p.L + l;
    const float* __restrict__ dpool = p.dpooled ? p.dpooled + (int64_t)b * C : nullptr;
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < C; ++c) {
        const float t = to_f32<in_t>(dy2[(int64_t)c * p.L]) + (dpool ? dpool[c] : 0.f);
        const float gw = t * silu_t(to_f32<in_t>(z[(int64_t)c * p.z_cs])) * p.w[c];
        s1 += gw;
        s2 = fmaf(gw, m[(int64_t)c * p.L] - mu, s2);
    }
    const float k1 = s1 * invC, k2 = s2 * invC * rstd * rstd;
    in_t* __restrict__ dm = reinterpret_cast<in_t*>(p.dm) + (int64_t)b * C * p.L + l;
    in_t* __restrict__ dz = reinterpret_cast<in_t*>(p.dz) + (int64_t)b * p.dz_bs + l;
    for (int c = 0; c < C; ++c) {
        const float t = to_f32<in_t>(dy2[(int64_t)c * p.L]) + (dpool ? dpool[c] : 0.f);
        const float zv = to_f32<in_t>(z[(int64_t)c * p.z_cs]);
        const float d = m[(int64_t)c * p.L] - mu;
        const float gw = t * silu_t(zv) * p.w[c];
        dm[(int64_t)c * p.L] = from_f32<in_t>(rstd * (gw - k1 - d * k2));
        const float n = fmaf(d * rstd, p.w[c], p.b[c]);
        dz[(int64_t)c * p.dz_cs] = from_f32<in_t>(t * n * dsilu_t(zv));
    }
}

template <typename in_t>
__global__ void __launch_bounds__(256) merge_bwd_dwdb_kernel(const MergeBwdParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sred[8];
    const int c = blockIdx.x, b = blockIdx.y, C = p.C;
    const float* __restrict__ m = p.merged + ((int64_t)b * C + c) * p.L;
    const float2* __restrict__ st = reinterpret_cast<const float2*>(p.stats) + (int64_t)b * p.L;
    const in_t* __restrict__ z = reinterpret_cast<const in_t*>(p.z) + (int64_t)b * p.z_bs + (int64_t)c * p.z_cs;
    const in_t* __restrict__ dy2 = reinterpret_cast<const in_t*>(p.dy2) + ((int64_t)b * C + c) * p.L;
    const float dp = p.dpooled ? p.dpooled[(int64_t)b * C + c] : 0.f;
    const float invC = 1.f / C;
    float aw = 0.f, ab = 0.f;
    for (int l = threadIdx.x; l < p.L; l += 256) {
        const float2 s = st[l];
        const float mu = s.x * invC;
        const float rstd = rsqrtf(fmaxf(s.y * invC - mu * mu, 0.f) + 1e-5f);
        const float dn = (to_f32<in_t>(dy2[l]) + dp) * silu_t(to_f32<in_t>(z[l]));
        aw = fmaf(dn, (m[l] - mu) * rstd, aw);
        ab += dn;
    }
    const float tw = block_sum_256(aw, sred), tb = block_sum_256(ab, sred);
    if (threadIdx.x == 0) {
        atomicAdd(p.dw + c, tw);
        atomicAdd(p.db + c, tb);
    }
}

// ------------------------------------------------------------------------------------------ depthwise 3x3 backward
template <typename in_t>
__device__ __forceinline__ float dw_at_t(const in_t* __restrict__ xc, const float* __restrict__ w9, int h, int w, int H, int W) {
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

// gradient w.r.t. the conv OUTPUT before the activation (the conv is recomputed, nothing but its input was saved):
//   mode 0: v = dw(x[c]) + b;  dv = g * silu'(v)
//   mode 1: v1 = dw(x[c]) + b[c], v2 = dw(x[c+Co]) + b[c+Co];  dv[c] = g * v2 * gelu'(v1),  dv[c+Co] = g * gelu(v1)
template <typename in_t>
__global__ void __launch_bounds__(256) dwconv_bwd_pre_kernel(const DwBwdParams p) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.y % p.Cout, b = blockIdx.y / p.Cout;
    const int L = p.H * p.W;
    const in_t* __restrict__ xb = reinterpret_cast<const in_t*>(p.x) + (int64_t)b * p.x_bs;
    const in_t* __restrict__ g = reinterpret_cast<const in_t*>(p.g) + (int64_t)b * p.g_bs + (int64_t)c * p.g_cs;
    in_t* __restrict__ dv = reinterpret_cast<in_t*>(p.dv) + (int64_t)b * p.dv_bs;
    float w0[9], w1[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        w0[i] = p.w[c * 9 + i];
        w1[i] = p.mode ? p.w[(c + p.Cout) * 9 + i] : 0.f;
    }
    const float b0 = p.bias ? p.bias[c] : 0.f, b1 = (p.bias && p.mode) ? p.bias[c + p.Cout] : 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < L; i += gridDim.x * 256) {
        const int h = i / p.W, w = i % p.W;
        const float v = dw_at_t<in_t>(xb + (int64_t)c * p.x_cs, w0, h, w, p.H, p.W) + b0;
        const float gv = to_f32<in_t>(g[i]);
        if (p.mode == 0) {
            dv[(int64_t)c * p.dv_cs + i] = from_f32<in_t>(gv * dsilu_t(v));
        } else {
            const float v2 = dw_at_t<in_t>(xb + (int64_t)(c + p.Cout) * p.x_cs, w1, h, w, p.H, p.W) + b1;
            dv[(int64_t)c * p.dv_cs + i] = from_f32<in_t>(gv * v2 * dgelu_t(v));
            dv[(int64_t)(c + p.Cout) * p.dv_cs + i] = from_f32<in_t>(gv * gelu_t(v));
        }
    }
}

// weight / bias gradient of the depthwise conv: grid (channels, B);  dw[c][j] += sum_pix dv[pix] * x[pix + tap_j];  db[c] += sum dv
template <typename in_t>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const DwBwdParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sred[8];
    const int c = blockIdx.x, b = blockIdx.y;
    const int L = p.H * p.W;
    const in_t* __restrict__ x = reinterpret_cast<const in_t*>(p.x) + (int64_t)b * p.x_bs + (int64_t)c * p.x_cs;
    const in_t* __restrict__ dv = reinterpret_cast<const in_t*>(p.dv) + (int64_t)b * p.dv_bs + (int64_t)c * p.dv_cs;
    float acc[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[j] = 0.f;
    for (int i = threadIdx.x; i < L; i += 256) {
        const int h = i / p.W, w = i % p.W;
        const float d = to_f32<in_t>(dv[i]);
        acc[9] += d;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int hh = h + dy;
            if (hh < 0 || hh >= p.H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int ww = w + dx;
                if (ww < 0 || ww >= p.W) continue;
                acc[(dy + 1) * 3 + dx + 1] = fmaf(d, to_f32<in_t>(x[hh * p.W + ww]), acc[(dy + 1) * 3 + dx + 1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const float t = block_sum_256(acc[j], sred);
        if (threadIdx.x == 0) {
            if (j < 9) atomicAdd(p.dwgt + c * 9 + j, t);
            else if (p.dbias) atomicAdd(p.dbias + c, t);
        }
    }
}

// ------------------------------------------------------------------------------------------ channel gate backward
// forward (prologue of out_conv): yg = y2 * (1 + c[b,k]) (mode 1) or y2 + c[b,k] (mode 2).  grid (C, B).
template <typename in_t>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const GateBwdParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sred[8];
    const int c = blockIdx.x, b = blockIdx.y;
    const in_t* __restrict__ dyg = reinterpret_cast<const in_t*>(p.dyg) + ((int64_t)b * p.C + c) * p.L;
    const in_t* __restrict__ y2 = reinterpret_cast<const in_t*>(p.y2) + ((int64_t)b * p.C + c) * p.L;
    in_t* __restrict__ dy2 = reinterpret_cast<in_t*>(p.dy2) + ((int64_t)b * p.C + c) * p.L;
    const float gate = p.gate[(int64_t)b * p.C + c];
    float acc = 0.f;
    for (int l = threadIdx.x; l < p.L; l += 256) {
        const float d = to_f32<in_t>(dyg[l]);
        if (p.mode == 1) {
            acc = fmaf(d, to_f32<in_t>(y2[l]), acc);
            dy2[l] = from_f32<in_t>(d * (1.f + gate));
        } else {
            acc += d;
            dy2[l] = dyg[l];
        }
    }
    const float t = block_sum_256(acc, sred);
    if (threadIdx.x == 0) p.dgate[(int64_t)b * p.C + c] = t;
}

// ------------------------------------------------------------------------------------------ fused Adam(W) + clip + EMA
// One pass over the flat fp32 parameter / gradient / moment buffers (reference: optimizer_g.step() + model_ema() +
// clip_grad_norm_, SRGAN/VmambaIR/models/MambaSISR_model.py:141-147, Deraining/basicsr/models/image_restoration_model.py:165-173,
// Deraining/basicsr/models/base_model.py:54-62) instead of ~1 500 per-tensor launches.  `state` (device, 4 floats):
// [0] step count (incremented by the trailing 1-thread kernel: graph replays advance it), [1] sum of squared gradients.
__global__ void __launch_bounds__(256) sqsum_kernel(const float* __restrict__ g, long n, float* __restrict__ state) {
    __shared__ float sred[8];
    float acc = 0.f;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) acc = fmaf(g[i], g[i], acc);
    const float t = block_sum_256(acc, sred);
    if (threadIdx.x == 0) atomicAdd(state + 1, t);
}

__global__ void __launch_bounds__(256) fused_adam_kernel(const AdamParams p) {
    const float step = p.state[0] + 1.f;
    const float bc1 = 1.f - powf(p.beta1, step), bc2 = 1.f - powf(p.beta2, step);
    float gs = p.grad_scale;
    if (p.max_norm > 0.f) {  // clip_grad_norm_: scale by max_norm / (total_norm + 1e-6), clamped to 1
        const float total = sqrtf(p.state[1]) * p.grad_scale;
        gs *= fminf(p.max_norm / (total + 1e-6f), 1.f);
    }
    const float step_size = p.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < p.n; i += (long)gridDim.x * 256) {
        float w = p.param[i];
        float g = p.grad[i] * gs;
        if (p.weight_decay != 0.f) {
            if (p.decoupled) w -= p.lr * p.weight_decay * w;  // AdamW
            else g = fmaf(p.weight_decay, w, g);                // Adam (L2)
        }
        const float m = fmaf(p.beta1, p.m[i], (1.f - p.beta1) * g);
        const float v = fmaf(p.beta2, p.v[i], (1.f - p.beta2) * g * g);
        p.m[i] = m;
        p.v[i] = v;
        w -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + p.eps);
        p.param[i] = w;
        if (p.ema) p.ema[i] = fmaf(p.ema_decay, p.ema[i], (1.f - p.ema_decay) * w);
        if (p.zero_grad) p.grad[i] = 0.f;
    }
}

__global__ void adam_step_inc_kernel(float* state) {
    state[0] += 1.f;
    state[1] = 0.f;
}

int fused_adam_launch(const AdamParams& p, cudaStream_t stream) {
    const int blocks = (int)((p.n + 255) / 256 < 148 * 8 ? (p.n + 255) / 256 : 148 * 8);
    if (p.max_norm > 0.f) sqsum_kernel<<<blocks, 256, 0, stream>>>(p.grad, p.n, p.state);
    fused_adam_kernel<<<blocks, 256, 0, stream>>>(p);
    adam_step_inc_kernel<<<1, 1, 0, stream>>>(p.state);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

// ---------------------------------------------------------------------------------------------------- launchers
#define VMB_DISPATCH(dtype, KERN, grid, block, ...)                                                    \
    switch (dtype) {                                                                                   \
        case VMB_F32: VMB_CUDA(launch_pdl(KERN<float>, grid, block, 0, stream, __VA_ARGS__)); break;    \
        case VMB_BF16: VMB_CUDA(launch_pdl(KERN<__nv_bfloat16>, grid, block, 0, stream, __VA_ARGS__)); break; \
        case VMB_F16: VMB_CUDA(launch_pdl(KERN<__half>, grid, block, 0, stream, __VA_ARGS__)); break;   \
        default: set_error("unsupported dtype %d", dtype); return VMB_ERR_INVALID;                     \
    }

int ln_fwd_launch(const LnFwdParams& p, int dtype, cudaStream_t stream) {
    VMB_CHECK(p.B <= 65535, "layernorm_fwd: batch > 65535");
    VMB_DISPATCH(dtype, ln_fwd_kernel, dim3((p.L + 127) / 128, p.B), dim3(128), p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

int ln_bwd_launch(const LnBwdParams& p, int dtype, cudaStream_t stream) {
    VMB_CHECK(p.B <= 65535 && p.C <= 65535, "layernorm_bwd: batch / channels > 65535");
    VMB_DISPATCH(dtype, ln_bwd_dx_kernel, dim3((p.L + 127) / 128, p.B), dim3(128), p);
    if (p.dw) VMB_DISPATCH(dtype, ln_bwd_dwdb_kernel, dim3(p.C, p.B), dim3(256), p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

int merge_bwd_launch(const MergeBwdParams& p, int dtype, cudaStream_t stream) {
    VMB_CHECK(p.B <= 65535 && p.C <= 65535, "merge_bwd: batch / channels > 65535");
    VMB_DISPATCH(dtype, merge_bwd_dx_kernel, dim3((p.L + 127) / 128, p.B), dim3(128), p);
    VMB_DISPATCH(dtype, merge_bwd_dwdb_kernel, dim3(p.C, p.B), dim3(256), p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

int dwconv_bwd_launch(const DwBwdParams& p, int dtype, cudaStream_t stream) {
    const int L = p.H * p.W;
    const int chans = p.mode ? 2 * p.Cout : p.Cout;
    VMB_CHECK((long)p.B * chans <= 65535, "dwconv_bwd: batch * channels > 65535");
    VMB_DISPATCH(dtype, dwconv_bwd_pre_kernel, dim3((L + 255) / 256, p.B * p.Cout), dim3(256), p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

int dwconv_wgrad_launch(const DwBwdParams& p, int dtype, cudaStream_t stream) {
    const int chans = p.mode ? 2 * p.Cout : p.Cout;
    VMB_CHECK(chans <= 65535 && p.B <= 65535, "dwconv_wgrad: grid too large");
    VMB_DISPATCH(dtype, dwconv_wgrad_kernel, dim3(chans, p.B), dim3(256), p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

int gate_bwd_launch(const GateBwdParams& p, int dtype, cudaStream_t stream) {
    VMB_CHECK(p.B <= 65535 && p.C <= 65535, "gate_bwd: grid too large");
    VMB_DISPATCH(dtype, gate_bwd_kernel, dim3(p.C, p.B), dim3(256), p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

}  // namespace vmb

// ---------------------------------------------------------------------------------------------------- C ABI
using namespace vmb;
static inline bool tdt_ok(int d) { return d == VMB_F32 || d == VMB_BF16 || d == VMB_F16; }

extern "C" int vmb_layernorm_fwd(const vmb_ln_fwd_args* a, void* stream) {
    VMB_CHECK(a && a->x && a->w && (a->y || a->stats), "layernorm_fwd: null pointer");
    VMB_CHECK(tdt_ok(a->dtype) && (a->mode == 1 || a->mode == 2) && (a->mode == 2 || a->b), "layernorm_fwd: bad dtype / mode / bias");
    VMB_CHECK(a->batch > 0 && a->C > 0 && a->L > 0, "layernorm_fwd: bad sizes");
    LnFwdParams p{a->x, a->w, a->b, a->y, a->stats, a->batch, a->C, a->L, a->mode, a->x_bs, a->x_cs, a->y_bs, a->y_cs};
    return ln_fwd_launch(p, a->dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int vmb_layernorm_bwd(const vmb_ln_bwd_args* a, void* stream) {
    VMB_CHECK(a && a->x && a->g && a->w && a->dx && a->stats, "layernorm_bwd: null pointer");
    VMB_CHECK(tdt_ok(a->dtype) && (a->mode == 1 || a->mode == 2), "layernorm_bwd: bad dtype / mode");
    VMB_CHECK(a->batch > 0 && a->C > 0 && a->L > 0, "layernorm_bwd: bad sizes");
    LnBwdParams p{a->x, a->g, a->add, a->w, a->dx, a->dw, a->db, a->stats, a->batch, a->C, a->L, a->mode,
                  a->x_bs, a->x_cs, a->g_bs, a->g_cs, a->a_bs, a->a_cs, a->dx_bs, a->dx_cs};
    return ln_bwd_launch(p, a->dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int vmb_merge_norm_gate_bwd(const vmb_merge_bwd_args* a, void* stream) {
    VMB_CHECK(a && a->merged && a->stats && a->z && a->dy2 && a->w && a->b && a->dm && a->dz && a->dw && a->db, "merge_bwd: null pointer");
    VMB_CHECK(tdt_ok(a->dtype) && a->batch > 0 && a->C > 0 && a->L > 0, "merge_bwd: bad arguments");
    MergeBwdParams p{a->merged, a->stats, a->z, a->dy2, a->dpooled, a->w, a->b, a->dm, a->dz, a->dw, a->db,
                     a->batch, a->C, a->L, a->z_bs, a->z_cs, a->dz_bs, a->dz_cs};
    return merge_bwd_launch(p, a->dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int vmb_dwconv3x3_bwd(const vmb_dwconv_bwd_args* a, void* stream) {
    VMB_CHECK(a && a->x && a->w && a->g && a->dv, "dwconv_bwd: null pointer");
    VMB_CHECK(tdt_ok(a->dtype) && (a->mode == 0 || a->mode == 1) && a->batch > 0 && a->c_out > 0 && a->H > 0 && a->W > 0, "dwconv_bwd: bad arguments");
    DwBwdParams p{a->x, a->w, a->bias, a->g, a->dv, a->dw, a->dbias, a->batch, a->c_out, a->H, a->W, a->mode,
                  a->x_bs, a->x_cs, a->g_bs, a->g_cs, a->dv_bs, a->dv_cs};
    int rc = dwconv_bwd_launch(p, a->dtype, static_cast<cudaStream_t>(stream));
    if (rc != VMB_OK || !a->dw) return rc;
    return dwconv_wgrad_launch(p, a->dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int vmb_fused_adam(const vmb_adam_args* a, void* stream) {
    VMB_CHECK(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq && a->state, "fused_adam: null pointer");
    VMB_CHECK(a->n > 0 && a->lr >= 0.f && a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f && a->eps > 0.f,
              "fused_adam: bad hyper-parameters");
    AdamParams p{a->param, a->grad, a->exp_avg, a->exp_avg_sq, a->ema, a->state, a->n, a->lr, a->beta1, a->beta2, a->eps,
                 a->weight_decay, a->decoupled_weight_decay, a->grad_scale, a->max_grad_norm, a->ema_decay, a->zero_grad};
    return fused_adam_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int vmb_channel_gate_bwd(const vmb_gate_bwd_args* a, void* stream) {
    VMB_CHECK(a && a->dyg && a->y2 && a->gate && a->dy2 && a->dgate, "channel_gate_bwd: null pointer");
    VMB_CHECK(tdt_ok(a->dtype) && (a->mode == 1 || a->mode == 2) && a->batch > 0 && a->C > 0 && a->L > 0, "channel_gate_bwd: bad arguments");
    GateBwdParams p{a->dyg, a->y2, a->gate, a->dy2, a->dgate, a->batch, a->C, a->L, a->mode};
    return gate_bwd_launch(p, a->dtype, static_cast<cudaStream_t>(stream));
}
