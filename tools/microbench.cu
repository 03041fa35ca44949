// Pipe-throughput micro-benchmarks for the scan kernel's design decisions (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu && tools/microbench.bin
// Each test runs ONE CTA per SM with W warps; every warp executes ITERS x UNROLL instructions of the op under test on independent
// register chains and CTA 0 reports SM cycles (clock64) -> cycles per warp-instruction per SMSP (4 schedulers per SM).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

enum { T_FFMA, T_FFMA2, T_FMUL2B, T_MUFU, T_MIX_SCAN, T_SHFL, T_LDS128, T_MIX_POLY, T_FADD2, T_IMAD, T_LOP, T_MIX2, T_FMUL2S, T_FFMA2S, T_LDS_RB2, T_LDS_RB4, T_PASS1, T_COUNT };
const char* kNames[] = {"ffma (3 regs)", "ffma2", "fmul2 (pair x pair)", "mufu.ex2", "mix: 2 mufu + 3 ffma2", "shfl.up", "lds.128 broadcast",
                        "mix: 2 mufu + 7 ffma2", "fadd2", "imad", "lop3", "mix: 2 mufu + 3 ffma2 + 2 ffma + 1 lds128",
                        "fmul2 (pair x broadcast scalar)", "ffma2 (pair, bcast scalar, pair)", "lds.128 16 addr x2 lanes, 144 B stride",
                        "lds.128 8 addr x4 lanes, 144 B stride", "pass1-like: fmul2s, 2 mufu, fmul2s, ffma2 (loop-varying)"};
const int kInstrPerIter[] = {16, 16, 16, 16, 5 * 4, 16, 16, 9 * 4, 16, 16, 16, 8 * 4, 16, 16, 16, 16, 5 * 8};

template <int TEST>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cyc, float seed) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(seed, seed * 0.5f, seed * 0.25f, 1.f);
    __syncthreads();
    float2 a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(seed + i, seed - i);
        b[i] = make_float2(0.999f + 1e-4f * i, 1.001f - 1e-4f * i);
    }
    float2 c = make_float2(seed * 1e-3f, -seed * 1e-3f);
    int ia = threadIdx.x, ib = 12345;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (TEST == T_FFMA) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i].x = fmaf(a[i].x, b[i].x, c.x);
                a[i].y = fmaf(a[i].y, b[i].y, c.y);
            }
        } else if (TEST == T_FFMA2) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b[i], c);
        } else if (TEST == T_FMUL2B) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __fmul2_rn(a[i], b[i]);
        } else if (TEST == T_FADD2) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __fadd2_rn(a[i], b[i]);
        } else if (TEST == T_MUFU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i].x = ex2f(a[i].x);
                a[i].y = ex2f(a[i].y);
            }
        } else if (TEST == T_MIX_SCAN) {  // per state pair and position: e = A*dt ; a = ex2(e) x2 ; b = dtu*B ; h = a*h + b
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 e = __fmul2_rn(b[i], c);
                float2 aa = make_float2(ex2f(e.x), ex2f(e.y));
                float2 bb = __fmul2_rn(b[i + 4], c);
                a[i] = __ffma2_rn(aa, a[i], bb);
            }
        } else if (TEST == T_MIX_POLY) {  // 2 mufu + 7 packed FMA-pipe ops (a polynomial exp2 next to the MUFU one)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 e = __fmul2_rn(b[i], c);
                float2 aa = make_float2(ex2f(e.x), ex2f(e.y));
                float2 p = __ffma2_rn(e, c, b[i + 4]);
                p = __ffma2_rn(p, e, c);
                p = __ffma2_rn(p, e, b[i]);
                p = __ffma2_rn(p, e, c);
                p = __ffma2_rn(p, aa, c);
                a[i] = __ffma2_rn(aa, a[i], p);
            }
        } else if (TEST == T_MIX2) {  // pass-2-like mix with the y accumulation and a broadcast LDS
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 q = sm[(it + i) & 63];
                float2 e = __fmul2_rn(b[i], make_float2(q.x, q.y));
                float2 aa = make_float2(ex2f(e.x), ex2f(e.y));
                float2 bb = __fmul2_rn(b[i + 4], make_float2(q.z, q.w));
                a[i] = __ffma2_rn(aa, a[i], bb);
                a[i + 4].x = fmaf(a[i].x, q.x, a[i + 4].x);
                a[i + 4].y = fmaf(a[i].y, q.y, a[i + 4].y);
            }
        } else if (TEST == T_SHFL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i].x = __shfl_up_sync(0xffffffffu, a[i].x, 1);
                a[i].y = __shfl_up_sync(0xffffffffu, a[i].y, 2);
            }
        } else if (TEST == T_LDS128) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 q = sm[(it + i) & 63];
                a[i & 7].x += q.x;  // keep the load live (1 FADD per LDS; FADD rate >> LDS rate)
            }
        } else if (TEST == T_FMUL2S) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __fmul2_rn(a[i], make_float2(b[i].x, b[i].x));
        } else if (TEST == T_FFMA2S) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], make_float2(b[i].x, b[i].x), c);
        } else if (TEST == T_LDS_RB2 || TEST == T_LDS_RB4) {
            const int grp = (threadIdx.x & 31) / (TEST == T_LDS_RB2 ? 2 : 4);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 q = sm[grp * 9 + ((it + i) & 7)];
                a[i & 7].x += q.x;
            }
        } else if (TEST == T_PASS1) {  // pass-1 shape with loop-varying dt (c.x changes every iteration): nothing hoistable
            c.x += 1e-6f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float dti = c.x + b[i].y;
                float2 e = __fmul2_rn(b[i], make_float2(dti, dti));
                float2 aa = make_float2(ex2f(e.x), ex2f(e.y));
                float2 bb = __fmul2_rn(b[(i + 1) & 7], make_float2(c.x, c.x));
                a[0] = __ffma2_rn(aa, a[0], bb);
            }
        } else if (TEST == T_IMAD) {
#pragma unroll
            for (int i = 0; i < 16; ++i) ia = ia * ib + i;
        } else if (TEST == T_LOP) {
#pragma unroll
            for (int i = 0; i < 16; ++i) ia = (ia ^ ib) & (ia + i);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f || ia == 42) out[threadIdx.x] = s;  // keep everything live
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int TEST>
void run(int warps, float* out, long long* cyc, int sms) {
    bench<TEST><<<sms, warps * 32>>>(out, cyc, 0.5f);
    CK(cudaDeviceSynchronize());
    bench<TEST><<<sms, warps * 32>>>(out, cyc, 0.5f);
    CK(cudaDeviceSynchronize());
    long long h[1024];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += h[i];
    mean /= sms;
    const double warp_instr_per_smsp = (double)ITERS * kInstrPerIter[TEST] * warps / 4.0;
    printf("%-44s warps/SM=%2d  cycles=%9.0f  cyc/warp-instr/SMSP=%6.3f\n", kNames[TEST], warps, mean, mean / warp_instr_per_smsp);
}

template <int TEST>
void sweep(float* out, long long* cyc, int sms) {
    for (int w : {4, 8, 16, 32}) run<TEST>(w, out, cyc, sms);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, 4096));
    CK(cudaMalloc(&cyc, 8 * 1024));
    const int sms = p.multiProcessorCount;
    sweep<T_FFMA>(out, cyc, sms);
    sweep<T_FFMA2>(out, cyc, sms);
    sweep<T_FMUL2B>(out, cyc, sms);
    sweep<T_FADD2>(out, cyc, sms);
    sweep<T_MUFU>(out, cyc, sms);
    sweep<T_MIX_SCAN>(out, cyc, sms);
    sweep<T_MIX_POLY>(out, cyc, sms);
    sweep<T_MIX2>(out, cyc, sms);
    sweep<T_SHFL>(out, cyc, sms);
    sweep<T_LDS128>(out, cyc, sms);
    sweep<T_FMUL2S>(out, cyc, sms);
    sweep<T_FFMA2S>(out, cyc, sms);
    sweep<T_LDS_RB2>(out, cyc, sms);
    sweep<T_LDS_RB4>(out, cyc, sms);
    sweep<T_PASS1>(out, cyc, sms);
    sweep<T_IMAD>(out, cyc, sms);
    sweep<T_LOP>(out, cyc, sms);
    return 0;
}
