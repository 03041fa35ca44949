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

/* Direction-aware variant for the fused OSS path: group g (<= 4 groups, dim = ngroups * rows_per_group) reads its
 * rows from its own tensors and, when rev[g] != 0, walks them backwards (sequence position l <-> memory index
 * L-1-l for u, delta, B, C and out), so the four scan orders of cross_scan_2d
 * (SRGAN/VmambaIR/archs/MambaSISR6_arch.py:401-404) need no flipped / gathered copies: directions 0,2 read the
 * natural (H,W) tensors, directions 1,3 the (W,H)-transposed ones.  Strides are shared by the groups; u/delta/out
 * rows are indexed inside the group.  Requires 16 B-aligned rows (seqlen % 8 == 0 for 16-bit I/O), dstate <= 16. */
typedef struct {
    const void* u[4]; const void* delta[4]; const void* Bm[4]; const void* Cm[4]; void* out[4]; int rev[4];
    const float* A; const float* D; const float* delta_bias;   /* (dim, dstate), (dim), (dim): indexed by g*rows+r */
    int batch, dim, seqlen, dstate, ngroups;
    int64_t u_bs, u_ds, delta_bs, delta_ds, out_bs, out_ds, B_bs, B_ns, C_bs, C_ns;
    int delta_softplus;
    int dtype;
} vmb_scan_grouped_args;
int vmb_selective_scan_fwd_grouped(const vmb_scan_grouped_args* a, void* stream);

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


/* ---- B1 module boundary: fused stages of one OSS block (inference) -------------------------------
 * These replace the un-fused PyTorch ops of SS2D_1 / MamberBlock / FeedForward
 * (SRGAN/VmambaIR/archs/MambaSISR6_arch.py:166-218, 395-515).  Activations are NCHW, pixel-contiguous. */

/* out[b,m,p] = epi( sum_k W[m,k] * pro(x)[b,k,p] ): 1x1 conv with fused LayerNorm / channel-gate prologue and
 * bias / SiLU-range / residual epilogue.  Replaces norm1+in_conv (:487-489), x_proj (:409-410), gate+out_conv+
 * residual (:494-498,:512), norm2+project_in (:213), project_out+residual (:217,:513). */
typedef struct {
    const void* x; const void* w; const float* bias; const void* residual; void* out;
    const float* ln_w; const float* ln_b;   /* LayerNorm over K per pixel (ln_mode 1: WithBias, 2: BiasFree) */
    const float* gate;                      /* (B,K) fp32: gate_mode 1: x*(1+g), 2: x+g */
    int ln_mode, gate_mode;
    int act_from, act_to;                   /* SiLU on output channels [act_from, act_to) */
    int batch, K, M, P;
    int64_t x_bs, x_cs, r_bs, r_cs, o_bs, o_cs, g_bs;
    int64_t w_ld;                           /* row stride of w (0: = K); rows padded to a multiple of 16 elements load vectorised */
    int dtype, out_dtype;                   /* x / w / residual dtype; out dtype (same, or VMB_F32) */
    int w_static;                           /* 1: w / bias / ln_w / ln_b were final before the preceding kernel of the stream was launched
                                             * (cached inference weights): the kernel may stage them while that kernel still runs
                                             * (programmatic dependent launch); 0: they may be its output -- fetched after it completes */
} vmb_pixlin_args;
int vmb_pixlin(const vmb_pixlin_args* a, void* stream);

/* depthwise 3x3 (pad 1) + bias, then mode 0: SiLU (SS2D_1.conv2d+act :490-491), mode 1: gelu(conv[c]) *
 * conv[c+C_out] (FeedForward.dwconv + gate :215-216), or mode 2: nothing (the transposed conv of the backward pass, called
 * with the spatially flipped taps).  w: (channels, 9) fp32. */
typedef struct {
    const void* x; const float* w; const float* bias; void* out;
    int batch, c_out, H, W, mode;
    int64_t x_bs, x_cs, o_bs, o_cs;
    int dtype;
} vmb_dwconv_args;
int vmb_dwconv3x3(const vmb_dwconv_args* a, void* stream);
/* mode 0 with a second output out_t (B, c_out, W*H) dense: every plane of `out` transposed -- the (W,H)-ordered copy of x that the
 * column-major scan directions (cross_scan_2d :402) and their x_proj GEMM read; replaces a vmb_transpose_hw pass behind the conv
 * (`out` must be dense). */
int vmb_dwconv3x3_t(const vmb_dwconv_args* a, void* out_t, void* stream);

/* four scan orders by index arithmetic (cross_scan_2d :401-404; CrossScan RealSR arch :325-343):
 * out[b][k][row][l] = src[k][b][row][pi_k(l)];  pi_0(l)=l, pi_1(w*H+h)=h*W+w, pi_2 = pi_0(L-1-l), pi_3 = pi_1(L-1-l). */
typedef struct {
    const void* src[4]; void* out;
    int batch, rows, H, W;
    int64_t src_bs, src_rs, out_bs;
    int dtype;
    int64_t out_ks; /* element stride between the four directions of `out` (0: rows * H * W, dense) */
} vmb_cross_scan_args;
int vmb_cross_scan(const vmb_cross_scan_args* a, void* stream);
/* 1..4 cross-scans of one geometry (same dtype, batch, H, W; own sources, row counts and destinations) in ONE launch: what the
 * training path needs per block -- x, delta and B|C gathered into the four scan orders (forward), du, ddelta, dB, dC scattered back
 * (backward; the inverse orders = H and W swapped).  Same values as nseg calls of vmb_cross_scan. */
int vmb_cross_scan_multi(const vmb_cross_scan_args* segs, int nseg, void* stream);

/* (B*C) planes of H x W -> W x H (the transposed copy of x that directions 1 and 3 scan). */
typedef struct { const void* x; void* out; int planes, H, W; int dtype; } vmb_transpose_args;
int vmb_transpose_hw(const vmb_transpose_args* a, void* stream);

/* PixelShuffle(2) of the SR tail (Upsampler, SRGAN/VmambaIR/archs/common.py:45-60; nn.PixelShuffle semantics:
 * out[b, c, 2h+i, 2w+j] = in[b, 4c + 2i + j, h, w]) on channels-last storage: x (B,H,W,4C) -> out (B,2H,2W,C), so the
 * 3x3 convs either side run on their native NHWC layout without layout-transform kernels.  A pure permutation:
 * results are bit-identical to torch.nn.functional.pixel_shuffle. */
typedef struct { const void* x; void* out; int batch, H, W, C; int dtype; } vmb_pixel_shuffle_args;
int vmb_pixel_shuffle2_nhwc(const vmb_pixel_shuffle_args* a, void* stream);
/* the same with bias (fp32, 4C values) added to the input channels first: the Upsampler convs (archs/common.py:52) run without
 * their bias and the permutation pass that follows them adds it -- one elementwise pass over the widest tensors of the net less. */
int vmb_pixel_shuffle2_nhwc_bias(const vmb_pixel_shuffle_args* a, const float* bias, void* stream);

/* Dense 3x3 convolution, stride 1, zero padding 1 -- the non-OSS convolutions of the U-Net:
 * OverlapPatchEmbed.proj (SRGAN/VmambaIR/archs/MambaSISR6_arch.py:520-528), Downsample.body = Conv2d + PixelUnshuffle(2)
 * (:533-541), Upsample.body = Conv2d + PixelShuffle(2) (:544-553), the SR tail's last conv + the nearest-neighbour up-sampled
 * input (:607,640; Upsampler in archs/common.py:45-60), Mamber32.output + inp_img (Deraining mamber32_arch.py forward).
 * What the reference runs as separate passes behind the conv is index arithmetic of the store:
 *   mode VMB_CONV_PLAIN        out[b, m, y, x]                                   (out may be a channel slice of a wider tensor:
 *                                                                                 the decoder's torch.cat([up, skip], 1) never runs)
 *        VMB_CONV_UNSHUFFLE2   out[b, 4m + 2(y%2) + x%2, y/2, x/2]               nn.PixelUnshuffle(2); H, W even
 *        VMB_CONV_SHUFFLE2     out[b, m/4, 2y + (m/2)%2, 2x + m%2]               nn.PixelShuffle(2); Cout % 4 == 0
 *        VMB_CONV_ADD_NEAREST  out[b, m, y, x] = conv + add[b, m, y/s, x/s]      F.interpolate(add, scale_factor=s, 'nearest'), s >= 1
 * x: NCHW view (x_bs, x_cs element strides; planes dense) or, in_nhwc = 1, dense NHWC (B,H,W,Cin) with Cin % 8 == 0.
 * out: NCHW planes dense with (o_bs, o_cs) strides; its spatial size is (H/2, W/2), (2H, 2W) or (H, W) by mode.
 * w: the Conv2d weight (Cout, Cin, 3, 3) re-packed by the caller as [tap = 3*ky + kx][Mpad][Kpad] in the I/O dtype, zero padded,
 *    Mpad = Cout rounded up to 64, Kpad = Cin rounded up to 16 (static parameters: the kernel may read them before the preceding
 *    kernel of the stream has finished).  bias: fp32 (Cout) or NULL.  16-bit I/O: tensor-core implicit GEMM, fp32 accumulation;
 *    fp32 I/O: FFMA. */
#define VMB_CONV_PLAIN 0
#define VMB_CONV_UNSHUFFLE2 1
#define VMB_CONV_SHUFFLE2 2
#define VMB_CONV_ADD_NEAREST 3
typedef struct {
    const void* x; const void* w; const float* bias; void* out; const void* add;
    int batch, Cin, Cout, H, W;
    int in_nhwc, mode, add_scale;
    int64_t x_bs, x_cs, o_bs, o_cs, add_bs, add_cs;
    int dtype;
} vmb_conv3x3_args;
int vmb_conv3x3(const vmb_conv3x3_args* a, void* stream);

/* inverse orders + 4-way fp32 sum (:427-430 / CrossMerge) + out_norm (:433) + gate y*SiLU(z) (:493) +
 * per-(b,c) sums of the result for AdaptiveAvgPool2d (:441; `pooled` fp32 (B,C), fully written).
 * workspace: vmb_merge_workspace_bytes() bytes of device scratch (fp32 merged values + per-pixel statistics + tile sums).
 * 16-bit I/O with H, W multiples of 16 and C <= 96 runs ONE kernel (all channels of a 16x16 pixel tile per CTA, no atomics on the
 * statistics); other shapes a channel-split sum kernel + a normalisation kernel. */
typedef struct {
    const void* ys; const void* z; const float* ln_w; const float* ln_b; void* y2; float* pooled;
    int batch, C, H, W;
    int64_t z_bs, z_cs;
    int dtype;
    void* workspace;
    int in_place_order; /* 0: ys[k] in scan order (k=2,3 reversed); 1: ys[0],ys[2] in natural (H,W) pixel order and
                           ys[1],ys[3] in transposed (W,H) pixel order (outputs of vmb_selective_scan_fwd_grouped) */
    int z_preact;       /* 0: z already holds SiLU(z) (inference: the in_conv epilogue applied it); 1: z is the pre-activation
                           and SiLU is applied here (training: the backward needs the pre-activation) */
    int save_ws;        /* 1: the workspace keeps the fp32 merged values (B,C,L) followed by the per-pixel (sum, sum of squares)
                           over C (B,L,2) for vmb_merge_norm_gate_bwd; 0: its content is unspecified afterwards */
} vmb_merge_args;
int vmb_merge_norm_gate(const vmb_merge_args* a, void* stream);
int64_t vmb_merge_workspace_bytes(int batch, int C, int H, int W);

/* the channel-direction OSS for one image per CTA (cforward_corev1 :438-483; Mamber32/33 and RealSR variants):
 * pooled means -> conv_cin -> xc_proj / dtc_proj -> bidirectional selective scan over L=C -> conv_cout ->
 * channel_norm.  All parameters fp32; c_out fp32 (B,C). cin_w / cout_w may be NULL (RealSR: dc=1). */
typedef struct {
    const float* pooled; float inv_count;
    const float* cin_w; const float* cin_b; const float* xc_proj; const float* dtc_w; const float* dtc_b;
    const float* Ac_logs; const float* Dsc; const float* cout_w; const float* cout_b; const float* cn_w; const float* cn_b;
    float* c_out;
    int batch, C, dc, Rc, N;
} vmb_channel_args;
int vmb_channel_branch(const vmb_channel_args* a, void* stream);

/* ---- training path: backward of the fused stages (autograd of the modules above) ----------------
 * The 1x1-conv data gradients are vmb_pixlin with the transposed weight, the scan gradient is vmb_selective_scan_bwd,
 * the direction gathers are vmb_cross_scan (pi_k^-1 = pi_k with H and W swapped).  Parameter gradients (dw, db, ...) are
 * fp32 and ACCUMULATED into (the caller zero-fills), like the reference's autograd accumulates .grad. */

/* y = LayerNorm_C(x) per pixel, materialised (the weight-gradient GEMMs of in_conv / project_in need it);
 * mode 1: WithBias, 2: BiasFree (MambaSISR6_arch.py:166-195).  y and/or stats (B, L, 2) = (mean, rstd) may be NULL. */
typedef struct {
    const void* x; const float* w; const float* b; void* y; float* stats;
    int batch, C, L, mode;
    int64_t x_bs, x_cs, y_bs, y_cs;
    int dtype;
} vmb_ln_fwd_args;
int vmb_layernorm_fwd(const vmb_ln_fwd_args* a, void* stream);

/* dx = LayerNorm backward of g (gradient w.r.t. the LayerNorm output) [+ add]; dw += sum g*xhat, db += sum g (dw/db may be NULL).
 * stats: fp32 scratch (B, L, 2), written by the data-gradient kernel.  Two-phase use (weight gradients on another stream):
 * dw = NULL -> dx (+ stats) only; dx = NULL -> dw / db only, from x, g and the stats of the earlier call. */
typedef struct {
    const void* x; const void* g; const void* add; const float* w; void* dx; float* dw; float* db; float* stats;
    int batch, C, L, mode;
    int64_t x_bs, x_cs, g_bs, g_cs, a_bs, a_cs, dx_bs, dx_cs;
    int dtype;
} vmb_ln_bwd_args;
int vmb_layernorm_bwd(const vmb_ln_bwd_args* a, void* stream);

/* backward of vmb_merge_norm_gate (z_preact = 1): merged / stats = the forward's workspace (fp32 merged scan output (B,C,L) and the
 * per-pixel (sum, sum of squares) over C); dy2 (B,C,L) dense, dpooled (B,C) fp32 or NULL.
 * -> dm (B,C,L) dense: gradient w.r.t. the merged scan output (gathered into the four scan orders by vmb_cross_scan),
 *    dz: gradient w.r.t. the pre-activation z, dw/db += out_norm parameter gradients.
 * Two-phase use: dw = db = NULL -> data gradients only; dm = dz = NULL -> parameter gradients only. */
typedef struct {
    const float* merged; const float* stats; const void* z; const void* dy2; const float* dpooled;
    const float* w; const float* b; void* dm; void* dz; float* dw; float* db;
    int batch, C, L;
    int64_t z_bs, z_cs, dz_bs, dz_cs;
    int dtype;
} vmb_merge_bwd_args;
int vmb_merge_norm_gate_bwd(const vmb_merge_bwd_args* a, void* stream);

/* backward of vmb_dwconv3x3 up to the conv output: dv = gradient w.r.t. the conv result before the activation
 * (mode 0: g * silu'(v); mode 1: dv[c] = g * v2 * gelu'(v1), dv[c+c_out] = g * gelu(v1)); the conv itself is recomputed from x.
 * dw (channels, 9) / dbias (channels) fp32 accumulated when dw != NULL.  The input gradient is vmb_dwconv3x3(dv, flipped taps, mode 2).
 * Two-phase use: dw = NULL -> dv only; g = NULL -> dv is an INPUT (an earlier call's) and only dw / dbias are accumulated. */
typedef struct {
    const void* x; const float* w; const float* bias; const void* g; void* dv; float* dw; float* dbias;
    int batch, c_out, H, W, mode;
    int64_t x_bs, x_cs, g_bs, g_cs, dv_bs, dv_cs;
    int dtype;
} vmb_dwconv_bwd_args;
int vmb_dwconv3x3_bwd(const vmb_dwconv_bwd_args* a, void* stream);

/* backward of the channel gate in front of out_conv (vmb_pixlin gate_mode 1: y*(1+c), 2: y+c): dyg, y2 (B,C,L) dense ->
 * dy2 (B,C,L), dgate (B,C) fp32 (fully written). */
typedef struct {
    const void* dyg; const void* y2; const float* gate; void* dy2; float* dgate;
    int batch, C, L, mode;
    int dtype;
} vmb_gate_bwd_args;
int vmb_channel_gate_bwd(const vmb_gate_bwd_args* a, void* stream);

/* backward of vmb_channel_branch (one CTA per image recomputes the forward in shared memory and walks it backwards).
 * fwd: the forward's arguments (c_out unused).  dc_out (B,C) fp32: gradient w.r.t. the forward's output.
 * -> d_pooled (B,C) fp32, fully written (gradient w.r.t. the pooled SUMS); every parameter gradient fp32, ACCUMULATED (zero-filled
 * by the caller), shaped like its parameter; d_cin_* / d_cout_* NULL exactly when the forward's are.
 * scratch: vmb_channel_branch_bwd_scratch_bytes(batch, dc, C) bytes.  Shared memory: vmb_channel_branch_bwd_smem_bytes(C, dc, Rc, N)
 * must be <= 227 KB (C <= 192 for dc = 4; wider levels keep the composed torch path). */
typedef struct {
    vmb_channel_args fwd;
    const float* dc_out; float* d_pooled;
    float* d_cin_w; float* d_cin_b; float* d_xc_proj; float* d_dtc_w; float* d_dtc_b; float* d_Ac_logs; float* d_Dsc;
    float* d_cout_w; float* d_cout_b; float* d_cn_w; float* d_cn_b;
    float* scratch;
} vmb_channel_bwd_args;
int vmb_channel_branch_bwd(const vmb_channel_bwd_args* a, void* stream);
int64_t vmb_channel_branch_bwd_smem_bytes(int C, int dc, int Rc, int N);
int64_t vmb_channel_branch_bwd_scratch_bytes(int batch, int dc, int C);

/* weight gradient of a 1x1 conv: out[m][k] += sum_{b,p} dy[b][m][p] * x[b][k][p]  (out fp32 (M,K), accumulated: the caller zero-fills;
 * per_batch != 0: out is (B,M,K) and batches are not summed -- the channel gate in front of out_conv scales them per image).
 * dy (B,M,L), x (B,K,L): bf16 / fp16 views with 16 B aligned, pixel-contiguous rows.  mma.sync tensor-core kernel, split over
 * pixels.  (fp32 activations: use the library GEMM.) */
typedef struct {
    const void* dy; const void* x; float* out;
    int batch, M, K, L;
    int64_t dy_bs, dy_cs, x_bs, x_cs;
    int per_batch;
    int dtype;
    float* dbias; /* optional (M) fp32, accumulated: dbias[m] += sum_{b,p} dy[b][m][p] (the conv's bias gradient) */
} vmb_wgrad_args;
int vmb_pixlin_wgrad(const vmb_wgrad_args* a, void* stream);

/* Per-step weight preparation of one OSS block in ONE launch: every job converts an fp32 parameter into the layout a kernel of
 * this library consumes.  type 0: (M,K) -> dst (M, ld) in `dtype`, columns >= K zero;  1: (M,K) -> dst (K, ld) = transpose, columns
 * >= M zero;  2: fold x_proj (src: (4, K+N2, M)) and dt_proj (src2: (4, M, K)) into big (dst: (4(M+N2), ld)) and its transpose
 * (dst2: (M, ld2)) -- big[k] = [W_dt,k W_x,k[:K] ; W_x,k[K:]], M = C, K = dt_rank, N2 = 2*d_state;  3: dst = -exp(src), fp32 (M,K);
 * 4: dst = the 9 taps of (M, 9) spatially flipped, fp32. */
#define VMB_PREP_MAX_JOBS 16
typedef struct {
    const float* src; const float* src2; void* dst; void* dst2;
    int type, M, K, N2, ld, ld2;
} vmb_prep_job;
typedef struct {
    vmb_prep_job jobs[VMB_PREP_MAX_JOBS];
    int njobs;
    int dtype;
} vmb_prep_args;
int vmb_prep_block_weights(const vmb_prep_args* a, void* stream);

/* out = add + x4[:,0] + x4[:,1] + x4[:,2] + x4[:,3]: x4 (batch, 4, per_dir) contiguous, add / out (batch, per_dir) contiguous
 * (the four un-permuted direction gradients of the scan input joined with the x_proj data gradient). */
int vmb_sum4_add(const void* x4, const void* add, void* out, int batch, long per_dir, int dtype, void* stream);

/* Fused Adam / AdamW step + gradient clipping + EMA over FLAT fp32 buffers of n elements (optimizer_g.step(), clip_grad_norm_,
 * model_ema(): SRGAN/VmambaIR/models/MambaSISR_model.py:141-147, Deraining/basicsr/models/image_restoration_model.py:165-173,
 * Deraining/basicsr/models/base_model.py:54-62).  grad is first multiplied by grad_scale (1/world after the all-reduce) and, when
 * max_grad_norm > 0, by min(1, max_grad_norm / (||grad|| + 1e-6)).  state: 4 floats on the device, zero-initialised by the
 * caller; state[0] is the step count, advanced by this call (so CUDA-graph replays keep the bias correction right).
 * ema (NULL: none): ema = ema_decay * ema + (1 - ema_decay) * param.  zero_grad != 0 clears the gradient buffer. */
typedef struct {
    float* param; float* grad; float* exp_avg; float* exp_avg_sq; float* ema; float* state;
    long n;
    float lr, beta1, beta2, eps, weight_decay;
    int decoupled_weight_decay;   /* 0: Adam (L2 added to the gradient), 1: AdamW */
    float grad_scale, max_grad_norm, ema_decay;
    int zero_grad;
} vmb_adam_args;
int vmb_fused_adam(const vmb_adam_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMAMBAIR_B200_H */
