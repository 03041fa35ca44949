// Selective-scan forward for sm_100a -- replaces selective_scan_fwd_kernel
// (reference: Mamba/kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-172).
//
// Design (B200-first, not a port of the CUB block-scan kernel):
//   * one CTA = RB rows (channels d of one (batch, group)) x SEGS sequence segments of T=16
//     positions; lane -> (row = lane % RB, segment = warp*(32/RB) + lane/RB).
//   * B/C for the group are staged once per chunk in shared memory as fp32 and read by
//     BROADCAST (all rows of a warp read the same (n,l) word), so smem bandwidth is 1/RB of a
//     row-per-warp design and L2 traffic for B/C is 1/RB of the reference's row-per-CTA design.
//   * u/delta/out go straight between HBM and registers in full 32 B sectors (16 bf16 per lane).
//   * two passes per state pair, both states packed in one fma.rn.f32x2:
//       pass 1: a = ex2(A*log2e*dt) (MUFU), local end state with zero start;
//       segment scan: shuffles inside the warp, a short sequential fold over warp totals in smem;
//       pass 2: h = a*h + b from the true start state, y += h*C.
//     The decay of a whole segment is ex2(A * sum(dt)) -- one MUFU, no product chain.
//   * checkpoints of h every CKPT positions (fp32) feed the backward kernel.
#include "common.cuh"
#include "scan_params.h"

namespace vmb {

constexpr int T = kScanT;  // positions per lane per chunk

template <int RB, int W>
struct FwdCfg {
    static constexpr int SEGW = 32 / RB;
    static constexpr int SEGS = W * SEGW;
    static constexpr int CHUNK = SEGS * T;
    static constexpr int SEGQ = T / 2 + 1;  // float4 slots per segment (one pad slot: conflict-free broadcast)
    static constexpr int SLOTS = SEGS * SEGQ;
    static size_t smem_bytes(int npad) {
        return sizeof(float4) * (2 * 8 * SLOTS + 2 * W * RB) + sizeof(float) * (3 * RB * npad);
    }
};

// Stage B or C rows [n0, n0+16) x [c0, c0+CHUNK) of one (batch, group) into smem as
// float4 = (X[n][l], X[n+1][l], X[n][l+1], X[n+1][l+1]) at [n/2][seg*SEGQ + (l%T)/2].
template <typename in_t, typename Cfg, int NTHREADS>
__device__ __forceinline__ void stage_bc(float4* __restrict__ dst, const in_t* __restrict__ src,
                                         int64_t n_stride, int n0, int N, int c0, int L, bool vec_ok) {
    constexpr int V = Vec<in_t>::N;
    constexpr int LG = Cfg::CHUNK / V;  // l-groups per row
    for (int it = threadIdx.x; it < 8 * LG; it += NTHREADS) {
        const int np = it / LG, lg = it % LG;
        const int l = lg * V;  // offset inside chunk
        const int n = n0 + 2 * np;
        float f0[V], f1[V];
        const int valid = L - (c0 + l);
        if (n < N) load_vec<in_t>(src + (int64_t)n * n_stride + c0 + l, f0, valid, vec_ok);
        else {
#pragma unroll
            for (int i = 0; i < V; ++i) f0[i] = 0.f;
        }
        if (n + 1 < N) load_vec<in_t>(src + (int64_t)(n + 1) * n_stride + c0 + l, f1, valid, vec_ok);
        else {
#pragma unroll
            for (int i = 0; i < V; ++i) f1[i] = 0.f;
        }
        const int seg = l / T, tq = (l % T) / 2;
        float4* d = dst + np * Cfg::SLOTS + seg * Cfg::SEGQ + tq;
#pragma unroll
        for (int i = 0; i < V / 2; ++i) d[i] = make_float4(f0[2 * i], f1[2 * i], f0[2 * i + 1], f1[2 * i + 1]);
    }
}

template <typename in_t, int RB, int W>
__global__ void __launch_bounds__(W * 32, (W <= 8 ? 2 : 1))
scan_fwd_kernel(const ScanFwdParams p) {
    using Cfg = FwdCfg<RB, W>;
    constexpr int SEGW = Cfg::SEGW, CHUNK = Cfg::CHUNK, SEGQ = Cfg::SEGQ, SLOTS = Cfg::SLOTS;
    constexpr int NTHREADS = W * 32;
    constexpr int V = Vec<in_t>::N;

    extern __shared__ float4 smem_f4[];
    float4* sB = smem_f4;
    float4* sC = sB + 8 * SLOTS;
    float4* sAgg = sC + 8 * SLOTS;                                // [2][W][RB]
    float* sCarry = reinterpret_cast<float*>(sAgg + 2 * W * RB);  // [2][RB][npad]
    float* sA = sCarry + 2 * RB * p.npad;                         // [RB][npad]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane % RB, sl = lane / RB, seg = warp * SEGW + sl;
    const int b = blockIdx.y;
    const int d0 = blockIdx.x * RB, d = d0 + r;
    const int g = d0 / p.rows_per_group;
    const int N = p.N, L = p.L, npad = p.npad;

    const in_t* __restrict__ Bg = reinterpret_cast<const in_t*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
    const in_t* __restrict__ Cg = reinterpret_cast<const in_t*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
    const in_t* __restrict__ urow = reinterpret_cast<const in_t*>(p.u) + (int64_t)b * p.u_bs + (int64_t)d * p.u_ds;
    const in_t* __restrict__ drow = reinterpret_cast<const in_t*>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)d * p.dl_ds;
    in_t* __restrict__ orow = reinterpret_cast<in_t*>(p.out) + (int64_t)b * p.o_bs + (int64_t)d * p.o_ds;

    for (int i = tid; i < RB * npad; i += NTHREADS) {
        const int rr = i / npad, n = i % npad;
        sA[i] = n < N ? p.A[(int64_t)(d0 + rr) * N + n] * kLog2e : 0.f;
        sCarry[i] = 0.f;
    }
    const float Dval = p.D ? p.D[d] : 0.f;
    const float bias = p.bias ? p.bias[d] : 0.f;
    const int ntiles = npad / 16;

    int pc = 0;
    for (int c0 = 0; c0 < L; c0 += CHUNK, pc ^= 1) {
        const int l0 = c0 + seg * T;
        const int valid = min(max(L - l0, 0), T);
        float uv[T], dt[T];
#pragma unroll
        for (int v = 0; v < T / V; ++v) {
            load_vec<in_t>(urow + l0 + v * V, uv + v * V, valid - v * V, p.vec_ok);
            load_vec<in_t>(drow + l0 + v * V, dt + v * V, valid - v * V, p.vec_ok);
        }
        float sigma = 0.f;
        float dtu[T];
        float2 y2[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float x = dt[t] + bias;
            if (p.softplus) x = softplus_f(x);
            x = t < valid ? x : 0.f;  // identity element beyond the end: a=1, b=0
            dt[t] = x;
            dtu[t] = x * uv[t];
            sigma += x;
            y2[t] = make_float2(0.f, 0.f);
        }

        for (int nt = 0; nt < ntiles; ++nt) {
            __syncthreads();  // previous tile (and sAgg / carry buffers) fully consumed
            stage_bc<in_t, Cfg, NTHREADS>(sB, Bg, p.B_ns, nt * 16, N, c0, L, p.vec_ok);
            stage_bc<in_t, Cfg, NTHREADS>(sC, Cg, p.C_ns, nt * 16, N, c0, L, p.vec_ok);
            __syncthreads();

#pragma unroll 1
            for (int np = 0; np < 8; ++np) {
                const int n0 = nt * 16 + 2 * np;
                const float2 A2 = *reinterpret_cast<const float2*>(&sA[r * npad + n0]);
                const float4* __restrict__ bq = sB + np * SLOTS + seg * SEGQ;
                const float4* __restrict__ cq = sC + np * SLOTS + seg * SEGQ;
                // ---- pass 1: decay factors + local end state ----
                float2 a2[T];
                float2 hend = make_float2(0.f, 0.f);
#pragma unroll
                for (int t = 0; t < T; t += 2) {
                    const float4 Bq = bq[t / 2];
                    float2 e0 = mul2(A2, make_float2(dt[t], dt[t]));
                    float2 e1 = mul2(A2, make_float2(dt[t + 1], dt[t + 1]));
                    a2[t] = make_float2(ex2(e0.x), ex2(e0.y));
                    a2[t + 1] = make_float2(ex2(e1.x), ex2(e1.y));
                    hend = fma2(a2[t], hend, mul2(make_float2(dtu[t], dtu[t]), make_float2(Bq.x, Bq.y)));
                    hend = fma2(a2[t + 1], hend, mul2(make_float2(dtu[t + 1], dtu[t + 1]), make_float2(Bq.z, Bq.w)));
                }
                float2 P2 = mul2(A2, make_float2(sigma, sigma));
                P2 = make_float2(ex2(P2.x), ex2(P2.y));
                // ---- inclusive scan over the segments held by this warp ----
#pragma unroll
                for (int o = RB; o < 32; o <<= 1) {
                    const float2 Pp = shfl_up2(P2, o), Hp = shfl_up2(hend, o);
                    if (lane >= o) {
                        hend = fma2(P2, Hp, hend);
                        P2 = mul2(P2, Pp);
                    }
                }
                const int buf = np & 1;
                if (sl == SEGW - 1) sAgg[(buf * W + warp) * RB + r] = make_float4(P2.x, P2.y, hend.x, hend.y);
                float2 Pe = shfl_up2(P2, RB % 32), He = shfl_up2(hend, RB % 32);
                if (lane < RB) {
                    Pe = make_float2(1.f, 1.f);
                    He = make_float2(0.f, 0.f);
                }
                __syncthreads();
                // ---- fold the totals of the preceding warps onto the chunk-start state ----
                float2 st = *reinterpret_cast<const float2*>(&sCarry[(pc * RB + r) * npad + n0]);
                for (int w2 = 0; w2 < warp; ++w2) {
                    const float4 q = sAgg[(buf * W + w2) * RB + r];
                    st = fma2(make_float2(q.x, q.y), st, make_float2(q.z, q.w));
                }
                float2 h = fma2(Pe, st, He);
                if (warp == W - 1 && sl == SEGW - 1)
                    *reinterpret_cast<float2*>(&sCarry[((pc ^ 1) * RB + r) * npad + n0]) = fma2(P2, st, hend);
                // ---- pass 2: true states, output contraction ----
#pragma unroll
                for (int t = 0; t < T; t += 2) {
                    const float4 Bq = bq[t / 2];
                    const float4 Cq = cq[t / 2];
                    h = fma2(a2[t], h, mul2(make_float2(dtu[t], dtu[t]), make_float2(Bq.x, Bq.y)));
                    y2[t] = fma2(h, make_float2(Cq.x, Cq.y), y2[t]);
                    h = fma2(a2[t + 1], h, mul2(make_float2(dtu[t + 1], dtu[t + 1]), make_float2(Bq.z, Bq.w)));
                    y2[t + 1] = fma2(h, make_float2(Cq.z, Cq.w), y2[t + 1]);
                }
                if (p.ckpt != nullptr && l0 < L && ((l0 + T) % kScanCkpt) == 0) {
                    float* ck = p.ckpt + (((int64_t)b * p.dim + d) * p.n_ckpt + ((l0 + T) / kScanCkpt - 1)) * N + n0;
                    if (n0 < N) ck[0] = h.x;
                    if (n0 + 1 < N) ck[1] = h.y;
                }
            }
        }
        float yo[T];
#pragma unroll
        for (int t = 0; t < T; ++t) yo[t] = fmaf(Dval, uv[t], y2[t].x + y2[t].y);
#pragma unroll
        for (int v = 0; v < T / V; ++v) store_vec<in_t>(orow + l0 + v * V, yo + v * V, valid - v * V, p.vec_ok);
    }
}

template <typename in_t, int RB, int W>
static int launch_cfg(const ScanFwdParams& p, cudaStream_t stream) {
    using Cfg = FwdCfg<RB, W>;
    auto kern = scan_fwd_kernel<in_t, RB, W>;
    const size_t smem = Cfg::smem_bytes(p.npad);
    VMB_CHECK(smem <= 227 * 1024, "selective_scan_fwd: dstate=%d needs %zu B of shared memory", p.N, smem);
    VMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(p.dim / RB, p.batch);
    kern<<<grid, W * 32, smem, stream>>>(p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

// Pick rows-per-CTA: must divide the rows of a group; prefer long chunks (fewer barriers per
// position) unless the sequence is short or the grid would leave SMs idle.
static int pick_rb(const ScanFwdParams& p) {
    const int rpg = p.rows_per_group;
    int best = 0;
    double best_score = -1.0;
    const int cands[6] = {8, 16, 32, 4, 2, 1};
    for (int i = 0; i < 6; ++i) {
        const int rb = cands[i];
        if (rpg % rb) continue;
        const int w = rb >= 8 ? 8 : rb;  // threads = 32*w
        const int chunk = (32 / rb) * w * T;
        const long ctas = (long)p.batch * (p.dim / rb);
        const long slots = 148L * (rb >= 8 ? 2 : 4);
        const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
        // wasted lanes when the chunk overshoots a short sequence
        const int nchunks = (p.L + chunk - 1) / chunk;
        const double fill = (double)p.L / ((double)nchunks * chunk);
        double score = eff * fill * (rb >= 8 ? 1.0 : 0.5) * (rb == 8 ? 1.05 : 1.0);
        if (score > best_score) { best_score = score; best = rb; }
    }
    return best;
}

template <typename in_t>
static int launch_t(const ScanFwdParams& p, cudaStream_t stream) {
    switch (pick_rb(p)) {
        case 32: return launch_cfg<in_t, 32, 8>(p, stream);
        case 16: return launch_cfg<in_t, 16, 8>(p, stream);
        case 8: return launch_cfg<in_t, 8, 8>(p, stream);
        case 4: return launch_cfg<in_t, 4, 4>(p, stream);
        case 2: return launch_cfg<in_t, 2, 2>(p, stream);
        default: return launch_cfg<in_t, 1, 1>(p, stream);
    }
}

int scan_fwd_launch(const ScanFwdParams& p, int dtype, cudaStream_t stream) {
    switch (dtype) {
        case VMB_F32: return launch_t<float>(p, stream);
        case VMB_BF16: return launch_t<__nv_bfloat16>(p, stream);
        case VMB_F16: return launch_t<__half>(p, stream);
    }
    set_error("selective_scan_fwd: unsupported dtype %d", dtype);
    return VMB_ERR_INVALID;
}

}  // namespace vmb
