// Selective-scan forward for sm_100a -- replaces selective_scan_fwd_kernel
// (reference: Mamba/kernels/selective_scan/csrc/selective_scan/cus/selective_scan_fwd_kernel.cuh:61-172).
//
// Design (B200-first, not a port of the CUB block-scan kernel):
//   * one WARP (= one CTA of 32 threads) owns RB rows (channels d of one (batch, group)) and walks
//     the sequence in warp-chunks of (32/RB) segments x T=16 positions;
//     lane -> (row = lane % RB, segment = lane / RB).  No block-level barrier anywhere.
//   * B/C for the group are staged once per chunk in shared memory as fp32 and read by
//     BROADCAST (all rows of a warp read the same (n,l) word), so smem bandwidth is 1/RB of a
//     row-per-warp design and L2 traffic for B/C is 1/RB of the reference's row-per-CTA design.
//   * u/delta/out go straight between HBM and registers in full 32 B sectors (16 bf16 per lane).
//   * two passes per state pair, both states packed in one fma.rn.f32x2:
//       pass 1: a = ex2(A*log2e*dt) (MUFU), local end state with zero start;
//       segment scan: shuffles inside the warp; the chunk-to-chunk carry lives in the warp's smem;
//       pass 2: h = a*h + b from the true start state, y += h*C.
//     The decay of a whole segment is ex2(A * sum(dt)) -- one MUFU, no product chain.
//   * checkpoints of h every CKPT positions (fp32) feed the backward kernel.
#include <stdlib.h>

#include "scan_common.cuh"

namespace vmb {

template <typename in_t, int RB>
__global__ void __launch_bounds__(32, 16) scan_fwd_kernel(const ScanFwdParams p) {
    using Cfg = FwdCfg<RB>;
    using R = RawCfg<in_t, RB>;
    constexpr int SEGW = Cfg::SEGW, CHUNK = Cfg::CHUNK, SEGQ = Cfg::SEGQ, SLOTS = Cfg::SLOTS;
    constexpr int V = Vec<in_t>::N;

    extern __shared__ float4 smem_f4[];
    float4* sB = smem_f4;                                         // [8][SLOTS]
    float4* sC = sB + 8 * SLOTS;                                  // [8][SLOTS]
    in_t* raw = reinterpret_cast<in_t*>(sC + 8 * SLOTS);          // RawCfg rows
    float* sCarry = reinterpret_cast<float*>(reinterpret_cast<char*>(raw) + R::bytes);  // [RB][npad]
    float* sA = sCarry + RB * p.npad;                             // [RB][npad]  A * log2(e)

    const int lane = threadIdx.x;
    const int r = lane % RB, sl = lane / RB;
    const int blocks_per_batch = p.dim / RB;
    const int b = blockIdx.x / blocks_per_batch;
    const int d0 = (blockIdx.x % blocks_per_batch) * RB, d = d0 + r;
    const int g = d0 / p.rows_per_group;
    const int N = p.N, L = p.L, npad = p.npad;
    const bool async_ok = p.vec_ok;  // 16 B-aligned rows: asynchronous prefetch path

    const in_t* __restrict__ Bg = reinterpret_cast<const in_t*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
    const in_t* __restrict__ Cg = reinterpret_cast<const in_t*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
    const in_t* __restrict__ ublk = reinterpret_cast<const in_t*>(p.u) + (int64_t)b * p.u_bs + (int64_t)d0 * p.u_ds;
    const in_t* __restrict__ dblk = reinterpret_cast<const in_t*>(p.delta) + (int64_t)b * p.dl_bs + (int64_t)d0 * p.dl_ds;
    const in_t* __restrict__ urow = ublk + (int64_t)r * p.u_ds;
    const in_t* __restrict__ drow = dblk + (int64_t)r * p.dl_ds;
    in_t* __restrict__ orow = reinterpret_cast<in_t*>(p.out) + (int64_t)b * p.o_bs + (int64_t)d * p.o_ds;

    const in_t* io_base[2] = {ublk, dblk};
    const int64_t io_ds[2] = {p.u_ds, p.dl_ds};
    if (async_ok) prefetch_chunk<in_t, RB, 2>(raw, io_base, io_ds, Bg, Cg, p.B_ns, p.C_ns, N, 0, L, lane);
    for (int i = lane; i < RB * npad; i += 32) {
        const int rr = i / npad, n = i % npad;
        sA[i] = n < N ? p.A[(int64_t)(d0 + rr) * N + n] * kLog2e : 0.f;
        sCarry[i] = 0.f;
    }
    const float Dval = p.D ? p.D[d] : 0.f;
    const float bias = p.bias ? p.bias[d] : 0.f;
    const int ntiles = npad / 16;

    for (int c0 = 0; c0 < L; c0 += CHUNK) {
        const int l0 = c0 + sl * T;
        const int valid = min(max(L - l0, 0), T);
        float uv[T], dt[T];
        if (async_ok) {
            cp_async_wait_all();
            __syncwarp();
            convert_bc<in_t, RB, 2>(sB, raw + (2 * RB) * R::PITCH, lane);
            convert_bc<in_t, RB, 2>(sC, raw + (2 * RB + 16) * R::PITCH, lane);
#pragma unroll
            for (int v = 0; v < T / V; ++v) {
                load_vec_smem<in_t>(raw + r * R::PITCH + sl * T + v * V, uv + v * V);
                load_vec_smem<in_t>(raw + (RB + r) * R::PITCH + sl * T + v * V, dt + v * V);
            }
            __syncwarp();  // raw buffer fully consumed -> refill it with the next chunk while we compute
            if (c0 + CHUNK < L)
                prefetch_chunk<in_t, RB, 2>(raw, io_base, io_ds, Bg, Cg, p.B_ns, p.C_ns, N, c0 + CHUNK, L, lane);
        } else {
#pragma unroll
            for (int v = 0; v < T / V; ++v) {
                load_vec<in_t>(urow + l0 + v * V, uv + v * V, valid - v * V, false);
                load_vec<in_t>(drow + l0 + v * V, dt + v * V, valid - v * V, false);
            }
        }
        float sigma = 0.f;
        float dtu[T];
        float y[T];  // scalar accumulators (a packed pair per position would cost 16 more registers)
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float x = dt[t] + bias;
            if (p.softplus) x = softplus_f(x);
            x = t < valid ? x : 0.f;  // identity element beyond the end: a=1, b=0
            dt[t] = x;
            dtu[t] = x * uv[t];
            sigma += x;
            y[t] = Dval * uv[t];
        }

        for (int nt = 0; nt < ntiles; ++nt) {
            if (!async_ok || nt > 0) {  // synchronous staging: unaligned tensors, or state tiles beyond the first
                __syncwarp();
                stage_bc<in_t, Cfg>(sB, Bg, p.B_ns, nt * 16, N, c0, L, p.vec_ok, lane);
                stage_bc<in_t, Cfg>(sC, Cg, p.C_ns, nt * 16, N, c0, L, p.vec_ok, lane);
            }
            __syncwarp();

#pragma unroll 1
            for (int np = 0; np < 8; ++np) {
                const int n0 = nt * 16 + 2 * np;
                const float2 A2 = *reinterpret_cast<const float2*>(&sA[r * npad + n0]);
                const float4* __restrict__ bq = sB + np * SLOTS + sl * SEGQ;
                const float4* __restrict__ cq = sC + np * SLOTS + sl * SEGQ;
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
                float2 Pe = shfl_up2(P2, RB % 32), He = shfl_up2(hend, RB % 32);
                if (lane < RB) {
                    Pe = make_float2(1.f, 1.f);
                    He = make_float2(0.f, 0.f);
                }
                // ---- chunk-start state (written by the last segment's lanes one chunk ago) ----
                float2* carry = reinterpret_cast<float2*>(&sCarry[r * npad + n0]);
                const float2 st = *carry;
                float2 h = fma2(Pe, st, He);
                __syncwarp();
                if (sl == SEGW - 1) *carry = fma2(P2, st, hend);
                // ---- pass 2: true states, output contraction ----
#pragma unroll
                for (int t = 0; t < T; t += 2) {
                    const float4 Bq = bq[t / 2];
                    const float4 Cq = cq[t / 2];
                    h = fma2(a2[t], h, mul2(make_float2(dtu[t], dtu[t]), make_float2(Bq.x, Bq.y)));
                    y[t] = fmaf(h.y, Cq.y, fmaf(h.x, Cq.x, y[t]));
                    h = fma2(a2[t + 1], h, mul2(make_float2(dtu[t + 1], dtu[t + 1]), make_float2(Bq.z, Bq.w)));
                    y[t + 1] = fmaf(h.y, Cq.w, fmaf(h.x, Cq.z, y[t + 1]));
                }
                if (p.ckpt != nullptr && l0 < L && ((l0 + T) % kScanCkpt) == 0) {
                    float* ck = p.ckpt + (((int64_t)b * p.dim + d) * p.n_ckpt + ((l0 + T) / kScanCkpt - 1)) * N + n0;
                    if (n0 < N) ck[0] = h.x;
                    if (n0 + 1 < N) ck[1] = h.y;
                }
            }
        }
#pragma unroll
        for (int v = 0; v < T / V; ++v) store_vec<in_t>(orow + l0 + v * V, y + v * V, valid - v * V, p.vec_ok);
    }
}

template <typename in_t, int RB>
static int launch_cfg(const ScanFwdParams& p, cudaStream_t stream) {
    using Cfg = FwdCfg<RB>;
    auto kern = scan_fwd_kernel<in_t, RB>;
    const size_t smem = Cfg::smem_bytes(p.npad) + RawCfg<in_t, RB>::bytes;
    VMB_CHECK(smem <= 227 * 1024, "selective_scan_fwd: dstate=%d needs %zu B of shared memory", p.N, smem);
    if (smem > 48 * 1024) VMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long blocks = (long)p.batch * (p.dim / RB);
    kern<<<(unsigned)blocks, 32, smem, stream>>>(p);
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

// Rows per warp: must divide the rows of a group.  Larger RB = less smem/L2 traffic for B/C and
// fewer shuffle steps; smaller RB = more warps (parallelism) and longer warp-chunks.  Aim for
// >= ~12 warps per SM, fall back to the smallest admissible RB for small problems.
static int pick_rb(const ScanFwdParams& p) {
    const int rpg = p.rows_per_group;
    if (const char* e = getenv("VMB_SCAN_RB")) {  // tuning/debug override
        const int rb = atoi(e);
        if (rb > 0 && rb <= 32 && (rb & (rb - 1)) == 0 && rpg % rb == 0) return rb;
    }
    const long rows = (long)p.batch * p.dim;
    // measured on B200 (C=96, L=4096): RB=8 wins once there are >= ~6 warps per SM, RB=4 below that;
    // RB<4 only when the group has fewer rows (their warp-chunks need > 50 KB of smem per warp).
    int rb = 1;
    while (rb < 8 && rpg % (rb * 2) == 0) rb *= 2;
    if (rb == 8 && rows / 8 < 148L * 6 && rpg % 4 == 0) rb = 4;
    if (rb == 8 && rpg % 16 == 0 && rows / 16 >= 148L * 64) rb = 16;
    // short sequences: do not use a warp-chunk much longer than L
    while (rb < 32 && rpg % (rb * 2) == 0 && (32 / rb) * T >= 2 * p.L) rb *= 2;
    return rb;
}

template <typename in_t>
static int launch_t(const ScanFwdParams& p, cudaStream_t stream) {
    switch (pick_rb(p)) {
        case 32: return launch_cfg<in_t, 32>(p, stream);
        case 16: return launch_cfg<in_t, 16>(p, stream);
        case 8: return launch_cfg<in_t, 8>(p, stream);
        case 4: return launch_cfg<in_t, 4>(p, stream);
        case 2: return launch_cfg<in_t, 2>(p, stream);
        default: return launch_cfg<in_t, 1>(p, stream);
    }
}

int scan_fwd_launch(const ScanFwdParams& p, int dtype, cudaStream_t stream) {
    VMB_CHECK((long)p.batch * p.dim < (1L << 31), "selective_scan_fwd: batch*dim too large");
    switch (dtype) {
        case VMB_F32: return launch_t<float>(p, stream);
        case VMB_BF16: return launch_t<__nv_bfloat16>(p, stream);
        case VMB_F16: return launch_t<__half>(p, stream);
    }
    set_error("selective_scan_fwd: unsupported dtype %d", dtype);
    return VMB_ERR_INVALID;
}

}  // namespace vmb
