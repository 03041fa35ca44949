// Backward of the channel-direction OSS (training path): one CTA per image recomputes the forward of
// vmb_channel_branch (cforward_corev1, SRGAN/VmambaIR/archs/MambaSISR6_arch.py:438-483; Mamber32/33 and RealSR variants) in shared
// memory, then walks it backwards: channel_norm -> conv_cout -> un-flip / merge -> bidirectional selective scan over L = C
// (states kept per position: the whole (2 dc, 16, C) state history fits one CTA) -> dtc_proj + softplus -> xc_proj -> conv_cin -> pool.
// Replaces ~90 launches of torch autograd over (B, C)-sized tensors per block.  Parameter gradients are accumulated (fp32 atomics,
// the caller zero-fills); dpooled is fully written.
#include "common.cuh"
#include "oss_params.h"
#include "train_params.h"

namespace vmb {

constexpr int CHB_THREADS = 512;

__device__ __forceinline__ float chb_block_sum(float v, float* sRed, int lane, int warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sRed[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < CHB_THREADS / 32; ++i) t += sRed[i];
    return t;
}
__device__ __forceinline__ float chb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(CHB_THREADS) channel_branch_bwd_kernel(const ChannelBwdParams q) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    const ChannelParams& p = q.fwd;
    const int C = p.C, dc = p.dc, Rc = p.Rc, N = p.N, RN = Rc + 2 * N;
    const int CP = C | 1, rows = 2 * dc;
    float* sSeq = sm;                     // [dc][C]      xc (natural order)
    float* sDbl = sSeq + dc * C;          // [2][RN][CP]  xc_dbl per direction (direction order); later d(xc_dbl)
    float* sDt = sDbl + 2 * RN * CP;      // [rows][C]    softplus'ed dt; later d(seq) in its first dc rows
    float* sY = sDt + rows * C;           // [rows][C]    scan outputs; later du
    float* sDy = sY + rows * C;           // [rows][C]    d(scan output); later d(pre-softplus dt)
    float* sOut = sDy + rows * C;         // [C]          conv_cout output; later d(out)
    float* sMean = sOut + C;              // [C]          pooled means
    float* sRed = sMean + C;              // [64]
    float* sXp = sRed + 64;               // [2][RN][dc]
    float* sDw = sXp + 2 * RN * dc;       // [2][dc][Rc]
    float* sDb = sDw + 2 * dc * Rc;       // [2][dc]
    float* sCio = sDb + 2 * dc;           // cin_w[dc] cin_b[dc] cout_w[dc] cout_b[1]
    float* sH = sCio + 3 * dc + 1;        // [rows][16][CP] state history h_l; overwritten by dh dt u during the reverse scan
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = CHB_THREADS / 32;
    const bool has_cin = p.cin_w != nullptr, has_cout = p.cout_w != nullptr;

    for (int i = tid; i < 2 * RN * dc; i += CHB_THREADS) sXp[i] = p.xc_proj[i];
    for (int i = tid; i < 2 * dc * Rc; i += CHB_THREADS) sDw[i] = p.dtc_w[i];
    if (tid < 2 * dc) sDb[tid] = p.dtc_b[tid];
    if (tid < dc) {
        sCio[tid] = has_cin ? p.cin_w[tid] : 1.f;
        sCio[dc + tid] = has_cin ? p.cin_b[tid] : 0.f;
        sCio[2 * dc + tid] = has_cout ? p.cout_w[tid] : 1.f;
    }
    if (tid == 0) sCio[3 * dc] = has_cout ? p.cout_b[0] : 0.f;
    for (int l = tid; l < C; l += CHB_THREADS) sMean[l] = p.pooled[(int64_t)b * C + l] * p.inv_count;
    const int srow = tid >> 4, n = tid & 15;
    const bool sact = srow < rows && n < N;
    const float Araw = sact ? -__expf(p.Ac_logs[srow * N + n]) : 0.f;  // A = -exp(A_log)
    const float A2 = Araw * kLog2e;
    __syncthreads();
    // ================================ forward recompute (same phases as channel_branch_v2_kernel) ================================
    for (int j = warp; j < dc; j += NW)
        for (int l = lane; l < C; l += 32) sSeq[j * C + l] = fmaf(sMean[l], sCio[j], sCio[dc + j]);
    __syncthreads();
    for (int kc = warp; kc < 2 * RN; kc += NW) {
        const int k = kc >= RN;
        const float* wrow = sXp + kc * dc;
        for (int l = lane; l < C; l += 32) {
            const int ls = k ? C - 1 - l : l;
            float a = 0.f;
            for (int j = 0; j < dc; ++j) a = fmaf(wrow[j], sSeq[j * C + ls], a);
            sDbl[kc * CP + l] = a;
        }
    }
    __syncthreads();
    for (int kj = warp; kj < rows; kj += NW) {
        const int k = kj / dc;
        for (int l = lane; l < C; l += 32) {
            float a = sDb[kj];
            for (int r = 0; r < Rc; ++r) a = fmaf(sDw[kj * Rc + r], sDbl[(k * RN + r) * CP + l], a);
            sDt[kj * C + l] = softplus_f(a);
        }
    }
    __syncthreads();
    if (srow < rows) {  // sequential scan, thread = (row, state): keep every h_l
        const int k = srow / dc, j = srow - k * dc;
        const float* __restrict__ dtr = sDt + srow * C;
        const float* __restrict__ ur = sSeq + j * C;
        const float* __restrict__ Br = sDbl + (k * RN + Rc + (sact ? n : 0)) * CP;
        float* __restrict__ hh = sH + (srow * 16 + n) * CP;
        float h = 0.f;
        for (int l = 0; l < C; ++l) {
            const float dt = dtr[l];
            const float u = ur[k ? C - 1 - l : l];
            h = fmaf(ex2(dt * A2), h, sact ? dt * u * Br[l] : 0.f);
            hh[l] = h;
        }
    }
    __syncthreads();
    for (int row = warp; row < rows; row += NW) {  // y[row][l] = D u + sum_n h C
        const int k = row / dc, j = row - k * dc;
        const float Dv = p.Dsc[row];
        for (int l = lane; l < C; l += 32) {
            float y = 0.f;
            for (int s = 0; s < N; ++s) y = fmaf(sH[(row * 16 + s) * CP + l], sDbl[(k * RN + Rc + N + s) * CP + l], y);
            sY[row * C + l] = fmaf(Dv, sSeq[j * C + (k ? C - 1 - l : l)], y);
        }
    }
    __syncthreads();
    float part = 0.f;
    for (int l = tid; l < C; l += CHB_THREADS) {
        float acc = sCio[3 * dc];
        for (int j = 0; j < dc; ++j) acc = fmaf(sY[j * C + l] + sY[(dc + j) * C + (C - 1 - l)], sCio[2 * dc + j], acc);
        sOut[l] = acc;
        part += acc;
    }
    const float mu = chb_block_sum(part, sRed, lane, warp) / C;
    float vp = 0.f;
    for (int l = tid; l < C; l += CHB_THREADS) {
        const float d = sOut[l] - mu;
        vp += d * d;
    }
    const float rstd = rsqrtf(chb_block_sum(vp, sRed, lane, warp) / C + 1e-5f);
    // ================================ backward ================================
    // channel_norm: c = xh w + b
    const float* __restrict__ dcv = q.dc_out + (int64_t)b * C;
    float s1 = 0.f, s2 = 0.f;
    for (int l = tid; l < C; l += CHB_THREADS) {
        const float xh = (sOut[l] - mu) * rstd, g = dcv[l];
        atomicAdd(q.d_cn_w + l, g * xh);
        atomicAdd(q.d_cn_b + l, g);
        const float gw = g * p.cn_w[l];
        s1 += gw;
        s2 = fmaf(gw, xh, s2);
    }
    const float m1 = chb_block_sum(s1, sRed, lane, warp) / C, m2 = chb_block_sum(s2, sRed, lane, warp) / C;
    __syncthreads();
    float sb = 0.f;
    for (int l = tid; l < C; l += CHB_THREADS) {
        const float xh = (sOut[l] - mu) * rstd;
        const float d = rstd * (dcv[l] * p.cn_w[l] - m1 - xh * m2);
        sOut[l] = d;  // d(out)
        sb += d;
    }
    const float dbo = chb_block_sum(sb, sRed, lane, warp);
    if (tid == 0 && has_cout) atomicAdd(q.d_cout_b, dbo);
    __syncthreads();
    // conv_cout + merge: dy[0][j][l] = dout[l] wout[j], dy[1][j][l] = dout[C-1-l] wout[j]; d wout[j] = sum_l dout[l] ym[j][l]
    for (int j = warp; j < dc; j += NW) {
        float acc = 0.f;
        for (int l = lane; l < C; l += 32) {
            acc = fmaf(sOut[l], sY[j * C + l] + sY[(dc + j) * C + (C - 1 - l)], acc);
            sDy[j * C + l] = sOut[l] * sCio[2 * dc + j];
            sDy[(dc + j) * C + l] = sOut[C - 1 - l] * sCio[2 * dc + j];
        }
        acc = chb_warp_sum(acc);
        if (lane == 0 && has_cout) atomicAdd(q.d_cout_w + j, acc);
    }
    __syncthreads();
    // dD[row] = sum_l dy u
    for (int row = warp; row < rows; row += NW) {
        const int k = row / dc, j = row - k * dc;
        float acc = 0.f;
        for (int l = lane; l < C; l += 32) acc = fmaf(sDy[row * C + l], sSeq[j * C + (k ? C - 1 - l : l)], acc);
        acc = chb_warp_sum(acc);
        if (lane == 0) atomicAdd(q.d_Dsc + row, acc);
    }
    // reverse scan, thread = (row, state).  dh_l = C_l dy_l + a_{l+1} dh_{l+1};  g_l = a_l h_{l-1}
    //   dC[k][n][l] += dy h_l (per row, summed over the rows of the direction afterwards, staged in place of h)   -- see below
    if (srow < rows) {
        const int k = srow / dc, j = srow - k * dc;
        const float* __restrict__ dtr = sDt + srow * C;
        const float* __restrict__ ur = sSeq + j * C;
        const float* __restrict__ Br = sDbl + (k * RN + Rc + (sact ? n : 0)) * CP;
        const float* __restrict__ Cr = sDbl + (k * RN + Rc + N + (sact ? n : 0)) * CP;
        float* __restrict__ hh = sH + (srow * 16 + n) * CP;
        float* __restrict__ dyr = sDy + srow * C;   // read dy, then overwritten with d(pre-softplus dt) by lane n == 0
        float* __restrict__ dur = sY + srow * C;    // du
        const float Dv = p.Dsc[srow];
        float dh = 0.f, a_next = 0.f, dA = 0.f;
        for (int l = C - 1; l >= 0; --l) {
            const float dt = dtr[l], u = ur[k ? C - 1 - l : l], dy = dyr[l];
            const float a = ex2(dt * A2);
            const float hl = hh[l], hprev = l > 0 ? hh[l - 1] : 0.f;
            dh = sact ? fmaf(a_next, dh, Cr[l] * dy) : 0.f;
            a_next = a;
            const float g = a * hprev;
            float t1 = sact ? dh * Br[l] : 0.f;          // -> s1 = sum_n dh B
            float t2 = sact ? Araw * dh * g : 0.f;       // -> s2 = sum_n A dh a h_{l-1}
            dA = fmaf(dh * dt, g, dA);
            hh[l] = sact ? dy * hl : 0.f;                 // dC contribution of this row (h_l no longer needed: l-1 reads h[l-1], h[l-2])
            // the dB contribution dh dt u goes to the slot of h_{l+1}'s ... no free slot: keep it in a second array below
            const float vB = sact ? dh * dt * u : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                t1 += __shfl_xor_sync(0xffffffffu, t1, o, 16);
                t2 += __shfl_xor_sync(0xffffffffu, t2, o, 16);
            }
            __syncwarp();
            if (n == 0) {
                const float ddt = fmaf(u, t1, t2);
                dyr[l] = ddt * (1.f - ex2(-dt * kLog2e));  // d softplus = sigmoid(pre) = 1 - exp(-softplus(pre))
                dur[l] = fmaf(dt, t1, Dv * dy);
            }
            q.scratch_dB[(((int64_t)b * rows + srow) * 16 + n) * C + l] = vB;  // per-row dB term (global scratch, L2-resident)
        }
        if (sact) atomicAdd(q.d_Ac_logs + srow * N + n, dA * Araw);  // dA/dA_log = A
    }
    __syncthreads();
    // d(xc_dbl): dt_lr rows <- W_dt^T d(pre);  B rows <- sum_j dB[kj][n][l];  C rows <- sum_j dC[kj][n][l]  (direction order)
    // first the parameter gradients that still need the forward's dt_lr rows
    for (int t = warp; t < rows * Rc; t += NW) {
        const int kj = t / Rc, r = t - kj * Rc, k = kj / dc;
        float acc = 0.f;
        for (int l = lane; l < C; l += 32) acc = fmaf(sDy[kj * C + l], sDbl[(k * RN + r) * CP + l], acc);
        acc = chb_warp_sum(acc);
        if (lane == 0) atomicAdd(q.d_dtc_w + kj * Rc + r, acc);
    }
    for (int kj = warp; kj < rows; kj += NW) {
        float acc = 0.f;
        for (int l = lane; l < C; l += 32) acc += sDy[kj * C + l];
        acc = chb_warp_sum(acc);
        if (lane == 0) atomicAdd(q.d_dtc_b + kj, acc);
    }
    __syncthreads();
    for (int kc = warp; kc < 2 * RN; kc += NW) {
        const int k = kc >= RN, c = kc - k * RN;
        for (int l = lane; l < C; l += 32) {
            float v = 0.f;
            if (c < Rc) {
                for (int j = 0; j < dc; ++j) v = fmaf(sDw[(k * dc + j) * Rc + c], sDy[(k * dc + j) * C + l], v);
            } else if (c < Rc + N) {
                for (int j = 0; j < dc; ++j) v += q.scratch_dB[(((int64_t)b * rows + k * dc + j) * 16 + (c - Rc)) * C + l];
            } else {
                for (int j = 0; j < dc; ++j) v += sH[((k * dc + j) * 16 + (c - Rc - N)) * CP + l];
            }
            sDbl[kc * CP + l] = v;
        }
    }
    __syncthreads();
    // xc_proj: dW[k][c][j] = sum_l ddbl[k][c][l] xs[k][j][l];  dxs[k][j][l] = du[kj][l] + sum_c W[k][c][j] ddbl[k][c][l]
    for (int t = warp; t < 2 * RN * dc; t += NW) {
        const int kc = t / dc, j = t - kc * dc, k = kc >= RN;
        float acc = 0.f;
        for (int l = lane; l < C; l += 32) acc = fmaf(sDbl[kc * CP + l], sSeq[j * C + (k ? C - 1 - l : l)], acc);
        acc = chb_warp_sum(acc);
        if (lane == 0) atomicAdd(q.d_xc_proj + kc * dc + j, acc);
    }
    for (int kj = warp; kj < rows; kj += NW) {
        const int k = kj / dc, j = kj - k * dc;
        for (int l = lane; l < C; l += 32) {
            float v = sY[kj * C + l];
            for (int c = 0; c < RN; ++c) v = fmaf(sXp[(k * RN + c) * dc + j], sDbl[(k * RN + c) * CP + l], v);
            sY[kj * C + l] = v;  // dxs in direction order
        }
    }
    __syncthreads();
    // un-flip, conv_cin, pool
    for (int j = warp; j < dc; j += NW) {
        float aw = 0.f, ab = 0.f;
        for (int l = lane; l < C; l += 32) {
            const float d = sY[j * C + l] + sY[(dc + j) * C + (C - 1 - l)];
            sDt[j * C + l] = d;  // d(seq)
            aw = fmaf(d, sMean[l], aw);
            ab += d;
        }
        aw = chb_warp_sum(aw);
        ab = chb_warp_sum(ab);
        if (lane == 0 && has_cin) {
            atomicAdd(q.d_cin_w + j, aw);
            atomicAdd(q.d_cin_b + j, ab);
        }
    }
    __syncthreads();
    for (int l = tid; l < C; l += CHB_THREADS) {
        float v = 0.f;
        for (int j = 0; j < dc; ++j) v = fmaf(sDt[j * C + l], sCio[j], v);
        q.d_pooled[(int64_t)b * C + l] = v * p.inv_count;
    }
}

size_t channel_bwd_smem(int C, int dc, int Rc, int N) {
    const int RN = Rc + 2 * N, CP = C | 1, rows = 2 * dc;
    return sizeof(float) * ((size_t)dc * C + 2 * RN * CP + 3 * (size_t)rows * C + 2 * C + 64 + 2 * RN * dc + 2 * dc * Rc + 2 * dc +
                            3 * dc + 1 + (size_t)rows * 16 * CP);
}

int channel_bwd_launch(const ChannelBwdParams& q, cudaStream_t stream) {
    const ChannelParams& p = q.fwd;
    VMB_CHECK(p.N <= 16 && 2 * p.dc * 16 <= CHB_THREADS, "channel_branch_bwd: dstate <= 16 and 2*dc*16 <= %d", CHB_THREADS);
    const size_t smem = channel_bwd_smem(p.C, p.dc, p.Rc, p.N);
    VMB_CHECK(smem <= 227 * 1024, "channel_branch_bwd: C=%d needs %zu B of shared memory", p.C, smem);
    if (smem > 48 * 1024)
        VMB_CUDA(cudaFuncSetAttribute(channel_branch_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VMB_CUDA(launch_pdl(channel_branch_bwd_kernel, dim3(p.B), dim3(CHB_THREADS), smem, stream, q));
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

}  // namespace vmb

using namespace vmb;

extern "C" int64_t vmb_channel_branch_bwd_smem_bytes(int C, int dc, int Rc, int N) { return (int64_t)channel_bwd_smem(C, dc, Rc, N); }
extern "C" int64_t vmb_channel_branch_bwd_scratch_bytes(int batch, int dc, int C) { return 4 * (int64_t)batch * 2 * dc * 16 * C; }

extern "C" int vmb_channel_branch_bwd(const vmb_channel_bwd_args* a, void* stream) {
    VMB_CHECK(a && a->fwd.pooled && a->fwd.xc_proj && a->fwd.dtc_w && a->fwd.dtc_b && a->fwd.Ac_logs && a->fwd.Dsc && a->fwd.cn_w && a->fwd.cn_b,
              "channel_branch_bwd: null forward pointer");
    VMB_CHECK(a->dc_out && a->d_pooled && a->d_xc_proj && a->d_dtc_w && a->d_dtc_b && a->d_Ac_logs && a->d_Dsc && a->d_cn_w && a->d_cn_b && a->scratch,
              "channel_branch_bwd: null gradient pointer");
    VMB_CHECK((a->fwd.cin_w == nullptr) == (a->d_cin_w == nullptr) && (a->fwd.cout_w == nullptr) == (a->d_cout_w == nullptr),
              "channel_branch_bwd: conv_cin / conv_cout gradients must match the forward parameters");
    const vmb_channel_args& f = a->fwd;
    ChannelBwdParams q{};
    q.fwd = ChannelParams{f.pooled, f.inv_count, f.cin_w, f.cin_b, f.xc_proj, f.dtc_w, f.dtc_b, f.Ac_logs, f.Dsc, f.cout_w, f.cout_b,
                          f.cn_w, f.cn_b, nullptr, f.batch, f.C, f.dc, f.Rc, f.N};
    q.dc_out = a->dc_out; q.d_pooled = a->d_pooled; q.d_cin_w = a->d_cin_w; q.d_cin_b = a->d_cin_b; q.d_xc_proj = a->d_xc_proj;
    q.d_dtc_w = a->d_dtc_w; q.d_dtc_b = a->d_dtc_b; q.d_Ac_logs = a->d_Ac_logs; q.d_Dsc = a->d_Dsc; q.d_cout_w = a->d_cout_w;
    q.d_cout_b = a->d_cout_b; q.d_cn_w = a->d_cn_w; q.d_cn_b = a->d_cn_b; q.scratch_dB = a->scratch;
    return channel_bwd_launch(q, static_cast<cudaStream_t>(stream));
}
