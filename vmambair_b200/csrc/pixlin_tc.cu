// pixlin on the 5th-generation tensor cores: tcgen05.mma with the accumulator in TMEM.
//   out[b, m, p] = epi( sum_k W[m,k] * pro(x)[b,k,p] )      (same contract as pixlin.cu)
// D (128 output channels x 128 pixels, fp32) lives in TMEM; A = weight tile (K-major) and B = activation tile
// (MN-major: pixels are the contiguous dimension of NCHW) are read by the MMA straight from shared memory through
// UMMA shared-memory descriptors (SWIZZLE_NONE "interleave" canonical layout: 8 x 16 B core matrices).  One elected
// thread issues K/16 tcgen05.mma per output tile and commits to an mbarrier; the four warps then pull their 32 TMEM
// lanes with tcgen05.ld (one lane = one output channel = one contiguous pixel row) and apply the epilogue.
// Used for 16-bit I/O when the rows are 16 B aligned and K <= 384 (everything the OSS block needs); the mma.sync
// kernel in pixlin.cu remains the fallback for ragged shapes, the FFMA kernel for fp32.
#include <stdlib.h>

#include "common.cuh"
#include "oss_params.h"

namespace vmb {

constexpr int TC_PT = 128;       // pixels per CTA (UMMA N)
constexpr int TC_MT = 128;       // output channels per MMA tile (UMMA M)
constexpr int TC_THREADS = 128;  // 4 warps: warp w owns TMEM lanes [32w, 32w+32)
constexpr int TC_KMAX = 288;  // 6*128*kpad B of staging (X + double-buffered W) must fit 227 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// UMMA shared-memory descriptor, SWIZZLE_NONE, version 1 (Blackwell): start address, leading / stride byte offsets in
// 16 B units.  K-major operand: ((8,m),2):((1,SBO),LBO); MN-major operand: ((1,n),(8,k)):((X,SBO),(1,LBO)).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version_ = 1
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor, kind::f16: D=f32, A/B = bf16 (1) or f16 (0), A K-major, B MN-major, N>>3, M>>4
__device__ __forceinline__ uint32_t umma_idesc(int ab_fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
template <typename T> struct AbFmt;
template <> struct AbFmt<__nv_bfloat16> { static constexpr int v = 1; };
template <> struct AbFmt<__half> { static constexpr int v = 0; };
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float silu_tc(float v) { return v * rcp_approx(1.f + ex2(-v * kLog2e)); }
template <typename T> __device__ __forceinline__ float2 unpack2_tc(uint32_t v);
template <> __device__ __forceinline__ float2 unpack2_tc<__nv_bfloat16>(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2_tc<__half>(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc, int nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(nbytes) : "memory");
}

// element (k, n) of the activation tile, canonical MN-major no-swizzle layout: core (k/8, n/8) = 128 B, row k%8, col n%8
__device__ __forceinline__ int b_off(int k, int n) { return (((k >> 3) * (TC_PT / 8) + (n >> 3)) << 6) + ((k & 7) << 3) + (n & 7); }
// element (m, k) of the weight tile, canonical K-major no-swizzle layout: core (k/8, m/8) = 128 B, row m%8, col k%8
__device__ __forceinline__ int a_off(int m, int k) { return (((k >> 3) * (TC_MT / 8) + (m >> 3)) << 6) + ((m & 7) << 3) + (k & 7); }

template <typename in_t>
__global__ void __launch_bounds__(TC_THREADS) pixlin_tc_kernel(const PixlinParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int kpad = (p.K + 15) / 16 * 16;
    in_t* sB = reinterpret_cast<in_t*>(smem_raw);                 // [kpad/8][PT/8] cores
    in_t* sA = sB + kpad * TC_PT;                                 // [2][kpad/8][MT/8] cores
    float* sStat = reinterpret_cast<float*>(sA + 2 * kpad * TC_MT);  // [2][PT]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sStat + 2 * TC_PT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.z, p0 = blockIdx.x * TC_PT;
    const int mtiles = (p.M + TC_MT - 1) / TC_MT;
    const in_t* __restrict__ xb = reinterpret_cast<const in_t*>(p.x) + (int64_t)b * p.x_bs;
    const in_t* __restrict__ w = reinterpret_cast<const in_t*>(p.w);

    auto stage_a = [&](int mt, int buf) {  // weight tile rows [mt*128, +128): 16 B = 8 k of one row = one core-matrix row
        in_t* dst = sA + buf * kpad * TC_MT;
        for (int it = tid; it < TC_MT * (kpad / 8); it += TC_THREADS) {
            const int m = it % TC_MT, kg = it / TC_MT;
            const int mg = mt * TC_MT + m;
            const bool ok = mg < p.M;
            cp16(dst + a_off(m, kg * 8), ok ? (const void*)(w + (int64_t)mg * p.w_ld + kg * 8) : (const void*)w, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- one-time setup: TMEM allocation (warp 0), mbarrier ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TC_PT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- activation tile: 16 B = 8 pixels of one k row = one core-matrix row ----
    for (int it = tid; it < kpad * (TC_PT / 8); it += TC_THREADS) {
        const int ng = it % (TC_PT / 8), k = it / (TC_PT / 8);
        const bool ok = k < p.K && p0 + ng * 8 < p.P;
        cp16(sB + b_off(k, ng * 8), ok ? (const void*)(xb + (int64_t)k * p.x_cs + p0 + ng * 8) : (const void*)xb, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (blockIdx.y < mtiles) stage_a(blockIdx.y, 0);
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // activations landed (the weight tile may still be in flight)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // ---- prologue on the resident tile: one pixel per thread ----
    if (p.ln_mode) {
        const int n = tid;
        float s = 0.f;
        for (int k = 0; k < p.K; ++k) s += to_f32<in_t>(sB[b_off(k, n)]);
        const float mu = s / p.K;
        float v = 0.f;
        for (int k = 0; k < p.K; ++k) {
            const float d = to_f32<in_t>(sB[b_off(k, n)]) - mu;
            v = fmaf(d, d, v);
        }
        const float rstd = rsqrtf(v / p.K + 1e-5f);
        const bool wb = p.ln_mode == 1;
        for (int k = 0; k < p.K; ++k) {
            const float xv = to_f32<in_t>(sB[b_off(k, n)]);
            sB[b_off(k, n)] = from_f32<in_t>(wb ? fmaf((xv - mu) * rstd, p.ln_w[k], p.ln_b[k]) : xv * rstd * p.ln_w[k]);
        }
    }
    if (p.gate_mode) {
        const float* __restrict__ g = p.gate + (int64_t)b * p.g_bs;
        const int n = tid;
        for (int k = 0; k < p.K; ++k) {
            const float xv = to_f32<in_t>(sB[b_off(k, n)]);
            sB[b_off(k, n)] = from_f32<in_t>(p.gate_mode == 1 ? fmaf(xv, g[k], xv) : xv + g[k]);
        }
    }

    const uint32_t idesc = umma_idesc(AbFmt<in_t>::v, TC_MT, TC_PT);
    const in_t* __restrict__ res = p.residual ? reinterpret_cast<const in_t*>(p.residual) + (int64_t)b * p.r_bs : nullptr;
    in_t* __restrict__ ob = reinterpret_cast<in_t*>(p.out) + (int64_t)b * p.o_bs;
    uint32_t phase = 0;
    int buf = 0;
    for (int mt = blockIdx.y; mt < mtiles; mt += gridDim.y, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // tile + prologue complete; previous epilogue has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mt + (int)gridDim.y < mtiles) stage_a(mt + gridDim.y, buf ^ 1);
        if (tid == 0) {
            const uint32_t a0 = smem_u32(sA + buf * kpad * TC_MT), b0 = smem_u32(sB);
            for (int kk = 0; kk < kpad / 16; ++kk) {
                // K-major A: SBO = 128 B between 8-row groups, LBO = (MT/8)*128 B between the two 8-wide k groups
                const uint64_t ad = umma_desc(a0 + kk * 2 * (TC_MT / 8) * 128, (TC_MT / 8) * 128, 128);
                // MN-major B: SBO = 128 B between 8-pixel groups, LBO = (PT/8)*128 B between 8-row k groups
                const uint64_t bd = umma_desc(b0 + kk * 2 * (TC_PT / 8) * 128, (TC_PT / 8) * 128, 128);
                tc_mma(tmem_base, ad, bd, idesc, kk > 0 ? 1u : 0u);
            }
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: thread = TMEM lane = output channel; 32 pixels per tcgen05.ld ----
        const int mg = mt * TC_MT + tid;
        const float bs = (p.bias && mg < p.M) ? p.bias[mg] : 0.f;
        const bool act = mg >= p.act_from && mg < p.act_to;
#pragma unroll 1
        for (int c = 0; c < TC_PT; c += 32) {
            float v[32];
            tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
            if (mg >= p.M || p0 + c >= p.P) continue;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += bs;
            if (act) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = silu_tc(v[i]);
            }
            if (res) {
                const uint4* rp = reinterpret_cast<const uint4*>(res + (int64_t)mg * p.r_cs + p0 + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 r4 = __ldg(rp + j);
                    const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = unpack2_tc<in_t>(rw[q]);
                        v[j * 8 + 2 * q] += f.x;
                        v[j * 8 + 2 * q + 1] += f.y;
                    }
                }
            }
            uint4* op = reinterpret_cast<uint4*>(ob + (int64_t)mg * p.o_cs + p0 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                op[j] = make_uint4(pack2<in_t>(v[j * 8], v[j * 8 + 1]), pack2<in_t>(v[j * 8 + 2], v[j * 8 + 3]),
                                   pack2<in_t>(v[j * 8 + 4], v[j * 8 + 5]), pack2<in_t>(v[j * 8 + 6], v[j * 8 + 7]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_PT) : "memory");
}

bool pixlin_tc_applicable(const PixlinParams& p, int dtype, int out_dtype) {
    int mode = 0;  // 0: off (default: the mma.sync kernel is still faster), 1: when the grid fills the GPU, 2: whenever legal
    if (const char* e = getenv("VMB_PIXLIN_TC")) mode = atoi(e);
    if (mode == 0) return false;
    const int kpad = (p.K + 15) / 16 * 16;
    const bool legal = (dtype == VMB_BF16 || dtype == VMB_F16) && out_dtype == dtype && p.vec_ok && p.w_vec && kpad <= TC_KMAX &&
                       p.P % 32 == 0;
    return legal && (mode == 2 || (long)((p.P + TC_PT - 1) / TC_PT) * p.B >= 64);
}

int pixlin_tc_launch(const PixlinParams& p, int dtype, cudaStream_t stream) {
    const int kpad = (p.K + 15) / 16 * 16;
    const size_t smem = (size_t)2 * kpad * TC_PT + (size_t)2 * 2 * kpad * TC_MT + sizeof(float) * 2 * TC_PT + 64;
    const int ptiles = (p.P + TC_PT - 1) / TC_PT, mtiles = (p.M + TC_MT - 1) / TC_MT;
    int msplit = 1;
    while (msplit < mtiles && (long)ptiles * p.B * msplit < 148L * 3 / 2) ++msplit;
    dim3 grid(ptiles, msplit, p.B);
    if (dtype == VMB_BF16) {
        auto k = pixlin_tc_kernel<__nv_bfloat16>;
        VMB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TC_THREADS, smem, stream>>>(p);
    } else {
        auto k = pixlin_tc_kernel<__half>;
        VMB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TC_THREADS, smem, stream>>>(p);
    }
    VMB_CUDA(cudaGetLastError());
    return VMB_OK;
}

}  // namespace vmb
