/* CPU oracle for the selective-scan operator -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement (sequential recurrence, one (b,d) row at a time) of the
 * reference's selective_scan_ref
 * (Mamba/kernels/selective_scan/test_selective_scan.py:168-234) and of the
 * analytic backward evaluated by the reference bwd kernel
 * (csrc/selective_scan/cus/selective_scan_bwd_kernel.cuh:139-241).
 * Used by tests/ (checker), __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg.  Never linked into the product library.
 *
 * Layouts (all contiguous, fp32):  u,delta,out,dout: (B,D,L)   A: (D,N)
 *   Bm,Cm: (B,G,N,L)   Dv,bias: (D) or NULL.   Rows are independent -> OpenMP.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif
/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use every host core it can */
void scan_ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline float softplus_f(float x) { return x <= 20.f ? log1pf(expf(x)) : x; }
static inline double softplus_d(double x) { return x <= 20.0 ? log1p(exp(x)) : x; }

/* fp32 sequential forward (the reference oracle is fp32 throughout). */
void scan_ref_fwd_f32(const float *u, const float *delta, const float *A,
                      const float *Bm, const float *Cm, const float *Dv,
                      const float *bias, int softplus, float *out,
                      float *last_state /* (B,D,N) or NULL */,
                      int Bsz, int D, int L, int N, int G) {
    const int rep = D / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < Bsz; ++b)
        for (int d = 0; d < D; ++d) {
            const float *ur = u + ((size_t)b * D + d) * L;
            const float *dr = delta + ((size_t)b * D + d) * L;
            float *orow = out + ((size_t)b * D + d) * L;
            const int g = d / rep;
            const float *Bg = Bm + ((size_t)b * G + g) * N * L;
            const float *Cg = Cm + ((size_t)b * G + g) * N * L;
            const float *Ar = A + (size_t)d * N;
            float h[256];
            for (int n = 0; n < N; ++n) h[n] = 0.f;
            const float bi = bias ? bias[d] : 0.f;
            const float Dd = Dv ? Dv[d] : 0.f;
            for (int l = 0; l < L; ++l) {
                float dt = dr[l] + bi;
                if (softplus) dt = softplus_f(dt);
                const float dtu = dt * ur[l];
                float y = 0.f;
                for (int n = 0; n < N; ++n) {
                    h[n] = expf(dt * Ar[n]) * h[n] + dtu * Bg[(size_t)n * L + l];
                    y += h[n] * Cg[(size_t)n * L + l];
                }
                orow[l] = Dv ? y + ur[l] * Dd : y;
            }
            if (last_state)
                for (int n = 0; n < N; ++n) last_state[((size_t)b * D + d) * N + n] = h[n];
        }
}

/* fp64-internal forward ("ground truth" to show both fp32 results are equidistant). */
void scan_ref_fwd_f64(const float *u, const float *delta, const float *A,
                      const float *Bm, const float *Cm, const float *Dv,
                      const float *bias, int softplus, double *out,
                      int Bsz, int D, int L, int N, int G) {
    const int rep = D / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < Bsz; ++b)
        for (int d = 0; d < D; ++d) {
            const float *ur = u + ((size_t)b * D + d) * L;
            const float *dr = delta + ((size_t)b * D + d) * L;
            double *orow = out + ((size_t)b * D + d) * L;
            const int g = d / rep;
            const float *Bg = Bm + ((size_t)b * G + g) * N * L;
            const float *Cg = Cm + ((size_t)b * G + g) * N * L;
            const float *Ar = A + (size_t)d * N;
            double h[256];
            for (int n = 0; n < N; ++n) h[n] = 0.0;
            const double bi = bias ? bias[d] : 0.0;
            const double Dd = Dv ? Dv[d] : 0.0;
            for (int l = 0; l < L; ++l) {
                double dt = (double)dr[l] + bi;
                if (softplus) dt = softplus_d(dt);
                const double dtu = dt * ur[l];
                double y = 0.0;
                for (int n = 0; n < N; ++n) {
                    h[n] = exp(dt * Ar[n]) * h[n] + dtu * Bg[(size_t)n * L + l];
                    y += h[n] * Cg[(size_t)n * L + l];
                }
                orow[l] = y + ur[l] * Dd;
            }
        }
}

/* Analytic backward, fp64 internal, fp64 outputs.
 * du,ddelta: (B,D,L); dA: (D,N); dB,dC: (B,G,N,L); dD,dbias: (D) (may be NULL).
 * All outputs must be zero-initialised by the caller. */
void scan_ref_bwd_f64(const float *u, const float *delta, const float *A,
                      const float *Bm, const float *Cm, const float *Dv,
                      const float *bias, int softplus, const float *dout,
                      double *du, double *ddelta, double *dA, double *dB,
                      double *dC, double *dD, double *dbias,
                      int Bsz, int D, int L, int N, int G) {
    const int rep = D / G;
    /* parallel over (b, g): rows of one group write the same dB/dC slab. */
#pragma omp parallel
    {
        double *hs = (double *)malloc(sizeof(double) * (size_t)N * L);
        double *dts = (double *)malloc(sizeof(double) * (size_t)L);
        double *dA_loc = (double *)calloc((size_t)D * N, sizeof(double));
        double *dD_loc = (double *)calloc((size_t)D, sizeof(double));
        double *db_loc = (double *)calloc((size_t)D, sizeof(double));
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < Bsz; ++b)
            for (int g = 0; g < G; ++g) {
                const float *Bg = Bm + ((size_t)b * G + g) * N * L;
                const float *Cg = Cm + ((size_t)b * G + g) * N * L;
                double *dBg = dB + ((size_t)b * G + g) * N * L;
                double *dCg = dC + ((size_t)b * G + g) * N * L;
                for (int d = g * rep; d < (g + 1) * rep; ++d) {
                    const float *ur = u + ((size_t)b * D + d) * L;
                    const float *dr = delta + ((size_t)b * D + d) * L;
                    const float *gor = dout + ((size_t)b * D + d) * L;
                    double *dur = du + ((size_t)b * D + d) * L;
                    double *ddr = ddelta + ((size_t)b * D + d) * L;
                    const float *Ar = A + (size_t)d * N;
                    const double bi = bias ? bias[d] : 0.0;
                    const double Dd = Dv ? Dv[d] : 0.0;
                    double h[256], dh[256];
                    for (int n = 0; n < N; ++n) { h[n] = 0.0; dh[n] = 0.0; }
                    for (int l = 0; l < L; ++l) {
                        double dt = (double)dr[l] + bi;
                        if (softplus) dt = softplus_d(dt);
                        dts[l] = dt;
                        for (int n = 0; n < N; ++n) {
                            h[n] = exp(dt * Ar[n]) * h[n] + dt * ur[l] * Bg[(size_t)n * L + l];
                            hs[(size_t)n * L + l] = h[n];
                        }
                    }
                    for (int l = L - 1; l >= 0; --l) {
                        const double dt = dts[l], ul = ur[l], go = gor[l];
                        double du_acc = Dd * go, ddt_acc = 0.0;
                        for (int n = 0; n < N; ++n) {
                            if (l + 1 < L) dh[n] *= exp(dts[l + 1] * Ar[n]);
                            dh[n] += (double)Cg[(size_t)n * L + l] * go;
                            const double Bv = Bg[(size_t)n * L + l];
                            const double hl = hs[(size_t)n * L + l];
                            const double bl = dt * ul * Bv;
                            du_acc += dh[n] * Bv * dt;
                            ddt_acc += dh[n] * (ul * Bv + Ar[n] * (hl - bl));
                            dA_loc[(size_t)d * N + n] += dh[n] * dt * (hl - bl);
                            dBg[(size_t)n * L + l] += dh[n] * dt * ul;
                            dCg[(size_t)n * L + l] += go * hl;
                        }
                        dur[l] = du_acc;
                        if (softplus) {
                            const double x = (double)dr[l] + bi;
                            if (x <= 20.0) ddt_acc *= 1.0 / (1.0 + exp(-x));
                        }
                        ddr[l] = ddt_acc;
                        dD_loc[d] += go * ul;
                        db_loc[d] += ddt_acc;
                    }
                }
            }
#pragma omp critical
        {
            for (size_t i = 0; i < (size_t)D * N; ++i) dA[i] += dA_loc[i];
            if (dD) for (int d = 0; d < D; ++d) dD[d] += dD_loc[d];
            if (dbias) for (int d = 0; d < D; ++d) dbias[d] += db_loc[d];
        }
        free(hs); free(dts); free(dA_loc); free(dD_loc); free(db_loc);
    }
}
