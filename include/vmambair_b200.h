/* vmambair_b200 -- C ABI of the B200-native Omni-Selective-Scan operator stack.
 *
 * Plain C: raw device pointers, explicit sizes/strides (in ELEMENTS), a dtype
 * enum and a cudaStream_t (passed as void*).  The caller owns every buffer;
 * the library allocates nothing persistent, never synchronises the host, and
 * is safe to call under CUDA-graph capture.  Every function returns 0 on
 * success; on failure a thread-local message is available from
 * vmb_last_error().
 *
 * Each entry point names the reference interface it replaces (paths relative
 * to the reference repo root).
 */
#ifndef VMAMBAIR_B200_H
#define VMAMBAIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMB_OK 0
#define VMB_ERR_INVALID 1 /* bad shape / dtype / stride / pointer */
#define VMB_ERR_CUDA 2    /* a CUDA runtime call or kernel launch failed */

typedef enum { VMB_F32 = 0, VMB_BF16 = 1, VMB_F16 = 2 } vmb_dtype;

const char* vmb_last_error(void);
/* library / build identification: "vmambair_b200 <ver> sm_100a" */
const char* vmb_version(void);

/* Checkpoint interval (in sequence positions) of the opaque `ckpt` buffer shared by
 * vmb_selective_scan_fwd and _bwd:  ckpt is fp32 (batch, dim, n_ckpt, dstate) with
 * n_ckpt = ceil(seqlen / interval).  Replaces the reference's `x`
 * (B, D, ceil(L/2048), 2N) -- Mamba/kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:217-220. */
int vmb_scan_ckpt_interval(void);

/* ---- B0 operator boundary ------------------------------------------------
 * selective_scan_cuda_core.fwd  (cus/selective_scan.cpp:157-239, kernel
 * cus/selective_scan_fwd_kernel.cuh:61-172).
 *   u, delta, out : (batch, dim, seqlen)   dtype `dt`, last-dim stride 1
 *   A             : (dim, dstate) fp32 contiguous
 *   Bm, Cm        : (batch, ngroups, dstate, seqlen) dtype `dt`, last-dim stride 1
 *   D, delta_bias : (dim) fp32 or NULL
 *   ckpt          : fp32 (batch, dim, n_ckpt, dstate) or NULL (inference: no checkpoints)
 * out = D*u + sum_n C_n * h_n,   h_n,l = exp(dt_l*A_n) h_n,l-1 + dt_l*u_l*B_n,l,
 * dt = softplus?(delta + delta_bias).  State and accumulation in fp32. */
typedef struct {
    const void* u; const void* delta; const float* A; const void* Bm; const void* Cm;
    const float* D; const float* delta_bias; void* out; float* ckpt;
    int batch, dim, seqlen, dstate, ngroups;
    int64_t u_bs, u_ds;         /* batch / dim strides of u      */
    int64_t delta_bs, delta_ds; /* ... of delta                   */
    int64_t out_bs, out_ds;     /* ... of out                     */
    int64_t B_bs, B_gs, B_ns;   /* batch / group / state strides  */
    int64_t C_bs, C_gs, C_ns;
    int delta_softplus;
    int dtype; /* vmb_dtype of u, delta, Bm, Cm, out */
} vmb_scan_fwd_args;
int vmb_selective_scan_fwd(const vmb_scan_fwd_args* a, void* stream);

/* selective_scan_cuda_core.bwd  (cus/selective_scan.cpp:241-349, kernel
 * cus/selective_scan_bwd_kernel.cuh:66-273).
 *   dout, du, ddelta : (batch, dim, seqlen) dtype `dt`
 *   dA (dim,dstate), dD (dim), ddelta_bias (dim): fp32, ACCUMULATED into (caller zero-fills,
 *       like the reference's torch::zeros_like, cpp:321-327)
 *   dB, dC : (batch, ngroups, dstate, seqlen) contiguous, dtype `dt`, fully written (the reference
 *       accumulates in fp32 and casts in torch, cpp:322-323,347; here the fp32 accumulation lives in
 *       `workspace` and the cast is done by the library)
 *   workspace : device scratch of vmb_scan_bwd_workspace_bytes(...) bytes, 16 B aligned; the library
 *       zero-fills it on `stream`
 *   ckpt  : as written by the forward (required when seqlen > interval). */
typedef struct {
    const void* u; const void* delta; const float* A; const void* Bm; const void* Cm;
    const float* D; const float* delta_bias; const void* dout; const float* ckpt;
    void* du; void* ddelta; float* dA; void* dB; void* dC; float* dD; float* ddelta_bias; void* workspace;
    int batch, dim, seqlen, dstate, ngroups;
    int64_t u_bs, u_ds, delta_bs, delta_ds, dout_bs, dout_ds;
    int64_t du_bs, du_ds, ddelta_bs, ddelta_ds;
    int64_t B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    int delta_softplus;
    int dtype;
} vmb_scan_bwd_args;
int vmb_selective_scan_bwd(const vmb_scan_bwd_args* a, void* stream);
int64_t vmb_scan_bwd_workspace_bytes(int batch, int ngroups, int dstate, int seqlen);

#ifdef __cplusplus
}
#endif
#endif /* VMAMBAIR_B200_H */
