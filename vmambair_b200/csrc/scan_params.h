// Internal parameter blocks of the selective-scan kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vmb {

constexpr int kScanT = 16;      // sequence positions per lane per chunk
constexpr int kScanCkpt = 64;   // checkpoint interval (positions); see vmb_scan_ckpt_interval()

// Per-group sources (direction-aware fused path): group g of the scan reads its rows from its own tensors and may
// walk them backwards (rev: sequence position l <-> memory index L-1-l for u, delta, B, C and out).
struct ScanGroupDesc {
    const void *u, *delta, *Bm, *Cm;
    void* out;
    int rev;
};

struct ScanFwdParams {
    const void *u, *delta, *Bm, *Cm;
    const float *A, *D, *bias;
    void* out;
    float* ckpt;
    int batch, dim, L, N, G, npad, rows_per_group, n_ckpt;
    int64_t u_bs, u_ds, dl_bs, dl_ds, o_bs, o_ds;
    int64_t B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    int softplus;
    bool vec_ok;
    int ndesc;              // 0: flat tensors (B0 operator); else = G <= 4 group descriptors (strides shared, per-group bases)
    ScanGroupDesc grp[4];
    int num_sms;
};

struct ScanBwdParams {
    const void *u, *delta, *Bm, *Cm, *dout;
    const float *A, *D, *bias, *ckpt;
    void *du, *ddelta, *dB, *dC;  // I/O dtype
    float *dA, *dD, *dbias;      // fp32, accumulated into
    float* dBC;                  // fp32 scratch [b][g][npad/2][L][4], zeroed by the launcher
    int batch, dim, L, N, G, npad, rows_per_group, n_ckpt;
    int64_t u_bs, u_ds, dl_bs, dl_ds, do_bs, do_ds, du_bs, du_ds, dd_bs, dd_ds;
    int64_t B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    int softplus;
    bool vec_ok;
    int debug_nored;  // profiling only (VMB_BWD_NORED=1): skip the dB/dC global reductions
};

int scan_fwd_launch(const ScanFwdParams& p, int dtype, cudaStream_t stream);
bool scan_fwd_tma_pick(const ScanFwdParams& p, int& rb, int& ss);
int scan_fwd_tma_launch(const ScanFwdParams& p, int dtype, int rb, int ss, cudaStream_t stream);
int scan_bwd_launch(const ScanBwdParams& p, int dtype, cudaStream_t stream);
bool scan_bwd_tma_pick(const ScanBwdParams& p, int dtype, int& rb, int& ss);
int scan_bwd_tma_launch(const ScanBwdParams& p, int dtype, int rb, int ss, cudaStream_t stream);

}  // namespace vmb
