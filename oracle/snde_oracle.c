/*
 * snde_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's rigid-warp + photometric + smoothness
 * loss path (Eisoc/Self-supervised-Normal-Depth-Estimation).  It is the checker
 * the CUDA path is compared against; nothing in the product imports, links or
 * executes this file (only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do).
 *
 * Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
 * section 4), so this oracle is pinned against outputs of the reference itself,
 * imported unmodified in the authoring container by tests/golden/make_golden.py;
 * the vectors it wrote are committed under tests/golden/ and checked by
 * tests/test_oracle_vs_golden.py.
 *
 * Every function follows the operation ORDER of the reference so that fp32
 * rounding matches it as closely as a re-implementation can:
 *   - "large" torch.matmul (inner dim 3 or 4 against h*w columns) == sequential
 *     FMA over k (probed: bit-exact on the reference's CPU path);
 *   - "small" 3x3/4x4 bmm == sequential multiply-then-add starting from 0;
 *   - elementwise chains are never contracted (build with -ffp-contract=off).
 *
 * Build twice: default (real=float) and -DSNDE_ORACLE_F64 (real=double, used to
 * measure the fp32 noise floor).  Layouts are the reference's logical ones:
 * images NHWC contiguous, depth (N,h,w), flow/coords (N,h,w,2).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef SNDE_ORACLE_F64
typedef double real;
#define R_FMA(a, b, c) fma((a), (b), (c))
#define R_FLOOR floor
#define R_FABS fabs
#define R_EXP exp
#define R_SIN sin
#define R_COS cos
#else
typedef float real;
#define R_FMA(a, b, c) fmaf((a), (b), (c))
#define R_FLOOR floorf
#define R_FABS fabsf
#define R_EXP expf
#define R_SIN sinf
#define R_COS cosf
#endif

#define API __attribute__((visibility("default")))

API int snde_oracle_real_bytes(void) { return (int)sizeof(real); }

/* ---------------------------------------------------------------------------
 * meshgrid 1-D tables.  reference: utils/utils_edited.py:307-317
 *   x_t = (linspace(-1,1,n) + 1) * 0.5 * (n-1)
 * torch.linspace (CPU): step=(end-start)/(n-1); first half start+step*i, second
 * half end-step*(n-1-i), each with a single rounding (FMA) -- probed bit-exact
 * for n in 2..700, 1024, 1242, 1248, 2048.
 * ------------------------------------------------------------------------- */
API void snde_oracle_grid_table(int n, real *out) {
    if (n == 1) { out[0] = (real)0; return; } /* linspace(-1,1,1) = [-1] -> (-1+1)*0.5*0 */
    const real start = (real)-1, end = (real)1;
    const real step = (end - start) / (real)(n - 1);
    const int half = n / 2;
    for (int i = 0; i < n; ++i) {
        real lin = (i < half) ? R_FMA(step, (real)i, start) : R_FMA(-step, (real)(n - i - 1), end);
        real t = lin + (real)1;
        t = t * (real)0.5;
        out[i] = t * (real)(n - 1);
    }
}

/* small dense matmul C[m,n] = A[m,k] B[k,n], "small bmm" order (see header) */
static void mm_small(int m, int k, int n, const real *A, const real *B, real *C) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            real acc = (real)0;
            for (int t = 0; t < k; ++t) acc = acc + A[i * k + t] * B[t * n + j];
            C[i * n + j] = acc;
        }
}

/* ---------------------------------------------------------------------------
 * euler2mat / pose_vec2mat.  reference: utils/utils_edited.py:171-214, 268-289
 * vec = [tx,ty,tz,rx,ry,rz];  R = Rx(rx) * Ry(ry) * Rz(rz);  T = [[R,t],[0,0,0,1]]
 * ------------------------------------------------------------------------- */
static void rot_xyz(real rx, real ry, real rz, real Rx[9], real Ry[9], real Rz[9]) {
    real cz = R_COS(rz), sz = R_SIN(rz), cy = R_COS(ry), sy = R_SIN(ry), cx = R_COS(rx), sx = R_SIN(rx);
    real z[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    real y[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    real x[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
    memcpy(Rz, z, sizeof z); memcpy(Ry, y, sizeof y); memcpy(Rx, x, sizeof x);
}

API void snde_oracle_pose_vec2mat(const real *vec, real *T /*16*/) {
    real Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
    rot_xyz(vec[3], vec[4], vec[5], Rx, Ry, Rz);
    mm_small(3, 3, 3, Rx, Ry, Rxy);
    mm_small(3, 3, 3, Rxy, Rz, R);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = vec[i];
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

/* torch.inverse (LAPACK getrf+getri).  reference call sites: utils_edited.py:231,345.
 * Restated as Gauss-Jordan with partial pivoting; not bit-identical to LAPACK
 * (nor is cuSOLVER, which the reference would use on a GPU). returns 0 / -1 singular */
API int snde_oracle_inverse(int n, const real *A, real *Ainv) {
    real a[16], inv[16];
    if (n > 4) return -2;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { a[i * n + j] = A[i * n + j]; inv[i * n + j] = (i == j) ? (real)1 : (real)0; }
    for (int c = 0; c < n; ++c) {
        int p = c; real best = R_FABS(a[c * n + c]);
        for (int r = c + 1; r < n; ++r) if (R_FABS(a[r * n + c]) > best) { best = R_FABS(a[r * n + c]); p = r; }
        if (best == (real)0) return -1;
        if (p != c) for (int j = 0; j < n; ++j) {
            real t = a[c * n + j]; a[c * n + j] = a[p * n + j]; a[p * n + j] = t;
            t = inv[c * n + j]; inv[c * n + j] = inv[p * n + j]; inv[p * n + j] = t;
        }
        real d = a[c * n + c];
        for (int j = 0; j < n; ++j) { a[c * n + j] = a[c * n + j] / d; inv[c * n + j] = inv[c * n + j] / d; }
        for (int r = 0; r < n; ++r) if (r != c) {
            real f = a[r * n + c];
            if (f != (real)0) for (int j = 0; j < n; ++j) {
                a[r * n + j] = a[r * n + j] - f * a[c * n + j];
                inv[r * n + j] = inv[r * n + j] - f * inv[c * n + j];
            }
        }
    }
    for (int i = 0; i < n * n; ++i) Ainv[i] = inv[i];
    return 0;
}

/* ---------------------------------------------------------------------------
 * Per-(source,batch) matrices of compute_rigid_flow.  reference:
 * utils/utils_edited.py:343-357 (pose_vec2mat, optional inverse, K4 @ pose) and
 * :231 (inverse(K)).  Output order is source-major n = src*B + b, as the caller
 * concatenates it (models/baseline.py:915-922).
 *   Pf[n] = (K4 @ T)[0:3,:]      (forward: target depth -> source pixels)
 *   Pb[n] = (K4 @ inv(T))[0:3,:] (backward: source depth -> target pixels)
 *   Kinv[b] = inv(K[b])
 * ------------------------------------------------------------------------- */
API int snde_oracle_pose_prep(int B, int S, const real *poses /*B,S,6*/, const real *K /*B,3,3*/,
                              real *Pf /*S*B,3,4*/, real *Pb /*S*B,3,4*/, real *Kinv /*B,3,3*/) {
    for (int b = 0; b < B; ++b) {
        if (snde_oracle_inverse(3, K + 9 * b, Kinv + 9 * b)) return -1;
        real K4[16] = {0};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K4[i * 4 + j] = K[9 * b + i * 3 + j];
        K4[15] = 1;
        for (int s = 0; s < S; ++s) {
            real T[16], Ti[16], P[16];
            snde_oracle_pose_vec2mat(poses + (b * S + s) * 6, T);
            if (snde_oracle_inverse(4, T, Ti)) return -1;
            int n = s * B + b;
            mm_small(4, 4, 4, K4, T, P);  memcpy(Pf + 12 * n, P, 12 * sizeof(real));
            mm_small(4, 4, 4, K4, Ti, P); memcpy(Pb + 12 * n, P, 12 * sizeof(real));
        }
    }
    return 0;
}

/* Backward of snde_oracle_pose_prep (what torch autograd does through
 * matmul / inverse / cat / stack / cos / sin in the reference).
 * gPf,gPb: (S*B,3,4) double; gKinv: (B,3,3) double (may be NULL)
 * -> gposes (B,S,6) double, gK (B,3,3) double (may be NULL). Accumulates (+=). */
static void mm_d(int m, int k, int n, const double *A, const double *B, double *C) {
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) {
        double acc = 0; for (int t = 0; t < k; ++t) acc += A[i * k + t] * B[t * n + j];
        C[i * n + j] = acc;
    }
}
static void tr_d(int m, int n, const double *A, double *At) {
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}
API int snde_oracle_pose_prep_bwd(int B, int S, const real *poses, const real *K,
                                  const double *gPf, const double *gPb, const double *gKinv,
                                  double *gposes, double *gK) {
    for (int b = 0; b < B; ++b) {
        real Kinv_r[9];
        if (snde_oracle_inverse(3, K + 9 * b, Kinv_r)) return -1;
        double Kd[9], Kinv[9];
        for (int i = 0; i < 9; ++i) { Kd[i] = K[9 * b + i]; Kinv[i] = Kinv_r[i]; }
        double gKb[9] = {0};
        if (gKinv) { /* Y = inv(K):  gK = -Y^T gY Y^T */
            double Yt[9], t1[9], t2[9];
            tr_d(3, 3, Kinv, Yt); mm_d(3, 3, 3, Yt, gKinv + 9 * b, t1); mm_d(3, 3, 3, t1, Yt, t2);
            for (int i = 0; i < 9; ++i) gKb[i] -= t2[i];
        }
        for (int s = 0; s < S; ++s) {
            int n = s * B + b;
            const real *vec = poses + (b * S + s) * 6;
            real Tr[16], Tir[16];
            snde_oracle_pose_vec2mat(vec, Tr);
            if (snde_oracle_inverse(4, Tr, Tir)) return -1;
            double T[16], Ti[16];
            for (int i = 0; i < 16; ++i) { T[i] = Tr[i]; Ti[i] = Tir[i]; }
            double gT[16] = {0};
            /* P = K4 @ X  (rows 0..2 carry gradient):  gX[k][c] += sum_r K[r][k] gP[r][c], k<3;  gK[r][k] += sum_c gP[r][c] X[k][c] */
            for (int dir = 0; dir < 2; ++dir) {
                const double *gP = (dir == 0 ? gPf : gPb) + 12 * n;
                const double *X = dir == 0 ? T : Ti;
                double gX[16] = {0};
                for (int k = 0; k < 3; ++k) for (int c = 0; c < 4; ++c) {
                    double acc = 0; for (int r = 0; r < 3; ++r) acc += Kd[r * 3 + k] * gP[r * 4 + c];
                    gX[k * 4 + c] = acc;
                }
                for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) {
                    double acc = 0; for (int c = 0; c < 4; ++c) acc += gP[r * 4 + c] * X[k * 4 + c];
                    gKb[r * 3 + k] += acc;
                }
                if (dir == 0) { for (int i = 0; i < 16; ++i) gT[i] += gX[i]; }
                else { /* X = inv(T): gT += -X^T gX X^T */
                    double Xt[16], t1[16], t2[16];
                    tr_d(4, 4, Ti, Xt); mm_d(4, 4, 4, Xt, gX, t1); mm_d(4, 4, 4, t1, Xt, t2);
                    for (int i = 0; i < 16; ++i) gT[i] -= t2[i];
                }
            }
            /* T -> (t, R);  R = Rx Ry Rz */
            double *gv = gposes + (b * S + s) * 6;
            for (int i = 0; i < 3; ++i) gv[i] += gT[i * 4 + 3];
            double gR[9];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) gR[i * 3 + j] = gT[i * 4 + j];
            real Rxr[9], Ryr[9], Rzr[9]; rot_xyz(vec[3], vec[4], vec[5], Rxr, Ryr, Rzr);
            double Rx[9], Ry[9], Rz[9];
            for (int i = 0; i < 9; ++i) { Rx[i] = Rxr[i]; Ry[i] = Ryr[i]; Rz[i] = Rzr[i]; }
            double Rxy[9], Ryz[9], t1[9], t2[9], gRx[9], gRy[9], gRz[9];
            mm_d(3, 3, 3, Rx, Ry, Rxy); mm_d(3, 3, 3, Ry, Rz, Ryz);
            tr_d(3, 3, Ryz, t1); mm_d(3, 3, 3, gR, t1, gRx);                    /* gRx = gR (Ry Rz)^T */
            tr_d(3, 3, Rx, t1); mm_d(3, 3, 3, t1, gR, t2); tr_d(3, 3, Rz, t1); mm_d(3, 3, 3, t2, t1, gRy); /* Rx^T gR Rz^T */
            tr_d(3, 3, Rxy, t1); mm_d(3, 3, 3, t1, gR, gRz);                    /* (Rx Ry)^T gR */
            double cx = Rx[4], sx = Rx[7], cy = Ry[0], sy = Ry[2], cz = Rz[0], sz = Rz[3];
            gv[3] += gRx[4] * (-sx) + gRx[5] * (-cx) + gRx[7] * cx + gRx[8] * (-sx);
            gv[4] += gRy[0] * (-sy) + gRy[2] * cy + gRy[6] * (-cy) + gRy[8] * (-sy);
            gv[5] += gRz[0] * (-sz) + gRz[1] * (-cz) + gRz[3] * cz + gRz[4] * (-sz);
        }
        if (gK) for (int i = 0; i < 9; ++i) gK[9 * b + i] += gKb[i];
    }
    return 0;
}

/* ---------------------------------------------------------------------------
 * compute_rigid_flow, per-pixel part.  reference: utils/utils_edited.py:331-362
 * (meshgrid :291, pixel2cam :216-239, cam2pixel :241-266), given the matrices
 * from snde_oracle_pose_prep.  depth (B,h,w), Kinv (B,3,3), P (B,3,4) -> flow (B,h,w,2)
 * ------------------------------------------------------------------------- */
typedef struct { real ray[3], cam[3], u[3], z, xn, yn; } reproj_t;

static inline void reproject(const real *Kinv, const real *P, real xt, real yt, real d, reproj_t *o) {
    for (int j = 0; j < 3; ++j) {            /* matmul(inverse(K), pix): FMA over k = 0,1,2 */
        real acc = Kinv[j * 3 + 0] * xt;
        acc = R_FMA(Kinv[j * 3 + 1], yt, acc);
        acc = R_FMA(Kinv[j * 3 + 2], (real)1, acc);
        o->ray[j] = acc;
        o->cam[j] = acc * d;                 /* ... * depth */
    }
    for (int r = 0; r < 3; ++r) {            /* matmul(proj, cam): FMA over k = 0..3, cam3 = 1 */
        real acc = P[r * 4 + 0] * o->cam[0];
        acc = R_FMA(P[r * 4 + 1], o->cam[1], acc);
        acc = R_FMA(P[r * 4 + 2], o->cam[2], acc);
        acc = R_FMA(P[r * 4 + 3], (real)1, acc);
        o->u[r] = acc;
    }
    o->z = o->u[2] + (real)1e-10;
    o->xn = o->u[0] / o->z;
    o->yn = o->u[1] / o->z;
}

API void snde_oracle_rigid_flow(int B, int h, int w, const real *depth, const real *Kinv, const real *P, real *flow) {
    real *xt = (real *)malloc(sizeof(real) * (size_t)w), *yt = (real *)malloc(sizeof(real) * (size_t)h);
    snde_oracle_grid_table(w, xt); snde_oracle_grid_table(h, yt);
#pragma omp parallel for collapse(2)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j) {
                reproj_t r;
                size_t p = ((size_t)b * h + i) * w + j;
                reproject(Kinv + 9 * b, P + 12 * b, xt[j], yt[i], depth[p], &r);
                flow[2 * p + 0] = r.xn - xt[j];
                flow[2 * p + 1] = r.yn - yt[i];
            }
    free(xt); free(yt);
}

/* backward: gflow (B,h,w,2) -> gdepth (B,h,w) [written], gP (B,3,4) double [+=], gKinv (B,3,3) double [+=, may be NULL] */
API void snde_oracle_rigid_flow_bwd(int B, int h, int w, const real *depth, const real *Kinv, const real *P,
                                    const real *gflow, real *gdepth, double *gP, double *gKinv) {
    real *xt = (real *)malloc(sizeof(real) * (size_t)w), *yt = (real *)malloc(sizeof(real) * (size_t)h);
    snde_oracle_grid_table(w, xt); snde_oracle_grid_table(h, yt);
    for (int b = 0; b < B; ++b) {
        const real *Pb = P + 12 * b;
        double aP[12] = {0}, aK[9] = {0};
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j) {
                reproj_t r;
                size_t p = ((size_t)b * h + i) * w + j;
                real d = depth[p];
                reproject(Kinv + 9 * b, Pb, xt[j], yt[i], d, &r);
                real gx = gflow[2 * p], gy = gflow[2 * p + 1];
                real gu[3];
                gu[0] = gx / r.z;
                gu[1] = gy / r.z;
                gu[2] = -(gx * r.u[0]) / (r.z * r.z) + -(gy * r.u[1]) / (r.z * r.z);
                real cam4[4] = {r.cam[0], r.cam[1], r.cam[2], 1};
                for (int q = 0; q < 3; ++q) for (int c = 0; c < 4; ++c) aP[q * 4 + c] += (double)gu[q] * cam4[c];
                real gd = 0;
                real pix[3] = {xt[j], yt[i], 1};
                for (int c = 0; c < 3; ++c) {
                    real gcam = Pb[0 * 4 + c] * gu[0];
                    gcam = R_FMA(Pb[1 * 4 + c], gu[1], gcam);
                    gcam = R_FMA(Pb[2 * 4 + c], gu[2], gcam);
                    gd = gd + gcam * r.ray[c];
                    if (gKinv) { real gray = gcam * d; for (int k = 0; k < 3; ++k) aK[c * 3 + k] += (double)gray * pix[k]; }
                }
                gdepth[p] = gd;
            }
        for (int q = 0; q < 12; ++q) gP[12 * b + q] += aP[q];
        if (gKinv) for (int q = 0; q < 9; ++q) gKinv[9 * b + q] += aK[q];
    }
    free(xt); free(yt);
}

/* flow -> absolute coords as flow_warp does.  reference: utils/utils_edited.py:386-392 */
API void snde_oracle_flow_to_coords(int N, int h, int w, const real *flow, real *coords) {
    real *xt = (real *)malloc(sizeof(real) * (size_t)w), *yt = (real *)malloc(sizeof(real) * (size_t)h);
    snde_oracle_grid_table(w, xt); snde_oracle_grid_table(h, yt);
    for (int n = 0; n < N; ++n) for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) {
        size_t p = ((size_t)n * h + i) * w + j;
        coords[2 * p] = xt[j] + flow[2 * p];
        coords[2 * p + 1] = yt[i] + flow[2 * p + 1];
    }
    free(xt); free(yt);
}

/* ---------------------------------------------------------------------------
 * bilinear_sampler.  reference: utils/utils_edited.py:402-486
 * imgs (N,hs,ws,C), coords (N,ht,wt,2) -> out (N,ht,wt,C).  Weights come from the
 * CLAMPED corners, so the result is ~0 outside [0,ws-1) x [0,hs-1).  Gather
 * indices are formed in integers here: the reference forms them in float32,
 * which is wrong once N*hs*ws > 2^24 -- a reference
 * bug that is not part of the specification.
 * ------------------------------------------------------------------------- */
typedef struct { int x0, x1, y0, y1; real wx0, wx1, wy0, wy1; } taps_t;

static inline real clampr(real v, real lo, real hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline void make_taps(real x, real y, int hs, int ws, taps_t *t) {
    real x0 = R_FLOOR(x), y0 = R_FLOOR(y);
    real x1 = x0 + (real)1, y1 = y0 + (real)1;
    real xmax = (real)(ws - 1), ymax = (real)(hs - 1);
    real x0s = clampr(x0, (real)0, xmax), x1s = clampr(x1, (real)0, xmax);
    real y0s = clampr(y0, (real)0, ymax), y1s = clampr(y1, (real)0, ymax);
    t->wx0 = x1s - x; t->wx1 = x - x0s; t->wy0 = y1s - y; t->wy1 = y - y0s;
    t->x0 = (int)x0s; t->x1 = (int)x1s; t->y0 = (int)y0s; t->y1 = (int)y1s;
}

API void snde_oracle_bilinear_sampler(int N, int hs, int ws, int C, const real *img, int ht, int wt,
                                      const real *coords, real *out) {
#pragma omp parallel for
    for (int n = 0; n < N; ++n) {
        const real *I = img + (size_t)n * hs * ws * C;
        for (int i = 0; i < ht; ++i) for (int j = 0; j < wt; ++j) {
            size_t p = ((size_t)n * ht + i) * wt + j;
            taps_t t; make_taps(coords[2 * p], coords[2 * p + 1], hs, ws, &t);
            real w00 = t.wx0 * t.wy0, w01 = t.wx0 * t.wy1, w10 = t.wx1 * t.wy0, w11 = t.wx1 * t.wy1;
            for (int c = 0; c < C; ++c) {
                real i00 = I[((size_t)t.y0 * ws + t.x0) * C + c], i01 = I[((size_t)t.y1 * ws + t.x0) * C + c];
                real i10 = I[((size_t)t.y0 * ws + t.x1) * C + c], i11 = I[((size_t)t.y1 * ws + t.x1) * C + c];
                out[p * C + c] = ((w00 * i00 + w01 * i01) + w10 * i10) + w11 * i11;
            }
        }
    }
}

/* backward: gout (N,ht,wt,C) -> gcoords (N,ht,wt,2) [written]; gimg (N,hs,ws,C) [+=, may be NULL] */
API void snde_oracle_bilinear_sampler_bwd(int N, int hs, int ws, int C, const real *img, int ht, int wt,
                                          const real *coords, const real *gout, real *gcoords, real *gimg) {
    for (int n = 0; n < N; ++n) {
        const real *I = img + (size_t)n * hs * ws * C;
        real *GI = gimg ? gimg + (size_t)n * hs * ws * C : NULL;
        for (int i = 0; i < ht; ++i) for (int j = 0; j < wt; ++j) {
            size_t p = ((size_t)n * ht + i) * wt + j;
            taps_t t; make_taps(coords[2 * p], coords[2 * p + 1], hs, ws, &t);
            real g00 = 0, g01 = 0, g10 = 0, g11 = 0;  /* grads of w00.. = sum_c gout_c * im */
            real w00 = t.wx0 * t.wy0, w01 = t.wx0 * t.wy1, w10 = t.wx1 * t.wy0, w11 = t.wx1 * t.wy1;
            for (int c = 0; c < C; ++c) {
                size_t a00 = ((size_t)t.y0 * ws + t.x0) * C + c, a01 = ((size_t)t.y1 * ws + t.x0) * C + c;
                size_t a10 = ((size_t)t.y0 * ws + t.x1) * C + c, a11 = ((size_t)t.y1 * ws + t.x1) * C + c;
                real g = gout[p * C + c];
                g00 += g * I[a00]; g01 += g * I[a01]; g10 += g * I[a10]; g11 += g * I[a11];
                if (GI) { GI[a00] += w00 * g; GI[a01] += w01 * g; GI[a10] += w10 * g; GI[a11] += w11 * g; }
            }
            real gwx0 = g00 * t.wy0 + g01 * t.wy1, gwx1 = g10 * t.wy0 + g11 * t.wy1;
            real gwy0 = g00 * t.wx0 + g10 * t.wx1, gwy1 = g01 * t.wx0 + g11 * t.wx1;
            gcoords[2 * p] = gwx1 - gwx0;       /* wt_x0 = x1s - x, wt_x1 = x - x0s */
            gcoords[2 * p + 1] = gwy1 - gwy0;
        }
    }
}

/* ---------------------------------------------------------------------------
 * image_similarity = alpha*DSSIM + (1-alpha)*|x-y|.  reference:
 * models/loss_functions.py:4-6 and DSSIM utils/utils_edited.py:121-141
 * (AvgPool2d(3,1,padding=1): zero padding, always /9; sum in window row-major order)
 * x,y (N,h,w,C) -> err (N,h,w,C)
 * ------------------------------------------------------------------------- */
typedef struct { real mx, my, exx, eyy, exy; } pooled_t;

static inline void pool5(const real *X, const real *Y, int h, int w, int C, int i, int j, int c, pooled_t *o) {
    real sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int di = -1; di <= 1; ++di) {
        int ii = i + di; if (ii < 0 || ii >= h) continue;
        for (int dj = -1; dj <= 1; ++dj) {
            int jj = j + dj; if (jj < 0 || jj >= w) continue;
            real xv = X[((size_t)ii * w + jj) * C + c], yv = Y[((size_t)ii * w + jj) * C + c];
            sx = sx + xv; sy = sy + yv; sxx = sxx + xv * xv; syy = syy + yv * yv; sxy = sxy + xv * yv;
        }
    }
    o->mx = sx / (real)9; o->my = sy / (real)9; o->exx = sxx / (real)9; o->eyy = syy / (real)9; o->exy = sxy / (real)9;
}

typedef struct { real S, A1, A2, B1, B2, mx, my; } ssim_t;

static inline void ssim_eval(const pooled_t *p, ssim_t *s) {
    const real C1 = (real)(0.01 * 0.01), C2 = (real)(0.03 * 0.03);
    real sig_x = p->exx - p->mx * p->mx;
    real sig_y = p->eyy - p->my * p->my;
    real sig_xy = p->exy - p->mx * p->my;
    s->A1 = ((real)2 * p->mx) * p->my + C1;
    s->A2 = (real)2 * sig_xy + C2;
    s->B1 = (p->mx * p->mx + p->my * p->my) + C1;
    s->B2 = (sig_x + sig_y) + C2;
    s->S = (s->A1 * s->A2) / (s->B1 * s->B2);
    s->mx = p->mx; s->my = p->my;
}

API void snde_oracle_image_similarity(int N, int h, int w, int C, real alpha, const real *x, const real *y, real *err) {
    const real oma = (real)(1.0 - (double)alpha);
#pragma omp parallel for
    for (int n = 0; n < N; ++n) {
        const real *X = x + (size_t)n * h * w * C, *Y = y + (size_t)n * h * w * C;
        real *E = err + (size_t)n * h * w * C;
        for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) for (int c = 0; c < C; ++c) {
            pooled_t p; ssim_t s;
            pool5(X, Y, h, w, C, i, j, c, &p); ssim_eval(&p, &s);
            real v = ((real)1 - s.S) / (real)2;
            v = clampr(v, (real)0, (real)1);
            size_t a = ((size_t)i * w + j) * C + c;
            E[a] = alpha * v + oma * R_FABS(X[a] - Y[a]);
        }
    }
}

/* backward: gerr (N,h,w,C) -> gx, gy (N,h,w,C) [written; either may be NULL] */
API void snde_oracle_image_similarity_bwd(int N, int h, int w, int C, real alpha, const real *x, const real *y,
                                          const real *gerr, real *gx, real *gy) {
    const real oma = (real)(1.0 - (double)alpha);
#pragma omp parallel for
    for (int n = 0; n < N; ++n) {
        size_t sz = (size_t)h * w * C;
        const real *X = x + n * sz, *Y = y + n * sz, *G = gerr + n * sz;
        /* coefficient maps: dL/dmu_x, dL/dExx, dL/dmu_y, dL/dEyy, dL/dExy at every window centre */
        real *cmx = (real *)calloc(sz * 5, sizeof(real));
        real *cexx = cmx + sz, *cmy = cexx + sz, *ceyy = cmy + sz, *cexy = ceyy + sz;
        for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) for (int c = 0; c < C; ++c) {
            pooled_t p; ssim_t s;
            pool5(X, Y, h, w, C, i, j, c, &p); ssim_eval(&p, &s);
            size_t a = ((size_t)i * w + j) * C + c;
            real v = ((real)1 - s.S) / (real)2;
            real gS = (v >= (real)0 && v <= (real)1) ? (real)-0.5 * alpha * G[a] : (real)0;
            real rB = (real)1 / (s.B1 * s.B2);
            real k = s.S * ((real)1 / s.B1 - (real)1 / s.B2);
            cmx[a] = gS * ((real)2 * s.my * (s.A2 - s.A1) * rB - (real)2 * s.mx * k);
            cmy[a] = gS * ((real)2 * s.mx * (s.A2 - s.A1) * rB - (real)2 * s.my * k);
            cexx[a] = gS * (-s.S / s.B2);
            ceyy[a] = cexx[a];
            cexy[a] = gS * ((real)2 * s.A1 * rB);
        }
        for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) for (int c = 0; c < C; ++c) {
            real bmx = 0, bexx = 0, bmy = 0, beyy = 0, bexy = 0;
            for (int di = -1; di <= 1; ++di) { int ii = i + di; if (ii < 0 || ii >= h) continue;
                for (int dj = -1; dj <= 1; ++dj) { int jj = j + dj; if (jj < 0 || jj >= w) continue;
                    size_t q = ((size_t)ii * w + jj) * C + c;
                    bmx += cmx[q]; bexx += cexx[q]; bmy += cmy[q]; beyy += ceyy[q]; bexy += cexy[q]; } }
            size_t a = ((size_t)i * w + j) * C + c;
            real d = X[a] - Y[a];
            real sg = d > (real)0 ? (real)1 : (d < (real)0 ? (real)-1 : (real)0);
            real l1 = oma * G[a] * sg;
            if (gx) gx[n * sz + a] = (bmx + (real)2 * X[a] * bexx + Y[a] * bexy) / (real)9 + l1;
            if (gy) gy[n * sz + a] = (bmy + (real)2 * Y[a] * beyy + X[a] * bexy) / (real)9 - l1;
        }
        free(cmx);
    }
}

/* ---------------------------------------------------------------------------
 * smooth_loss(depth, image).  reference: models/loss_functions.py:8-24,
 * gradient_x/y utils/utils_edited.py:143-147.  depth (M,h,w[,1]), image (M,h,w,C)
 * returns mean|dx*wx| + mean|dy*wy| (sums carried in double)
 * ------------------------------------------------------------------------- */
static inline real edge_weight(const real *I, size_t a, size_t b, int C) {
    real s = 0;
    for (int c = 0; c < C; ++c) s = s + R_FABS(I[a * C + c] - I[b * C + c]);
    return R_EXP(-(s / (real)C));
}

API double snde_oracle_smooth_loss(int M, int h, int w, int C, const real *depth, const real *img) {
    double sx = 0, sy = 0;
    for (int m = 0; m < M; ++m) {
        const real *D = depth + (size_t)m * h * w, *I = img + (size_t)m * h * w * C;
        for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) {
            size_t a = (size_t)i * w + j;
            if (j + 1 < w) sx += R_FABS((D[a] - D[a + 1]) * edge_weight(I, a, a + 1, C));
            if (i + 1 < h) sy += R_FABS((D[a] - D[a + w]) * edge_weight(I, a, a + w, C));
        }
    }
    double nx = (double)M * h * (w - 1), ny = (double)M * (h - 1) * w;
    return (nx > 0 ? sx / nx : NAN) + (ny > 0 ? sy / ny : NAN);
}

/* backward: gdepth (M,h,w) += gscalar * d loss / d depth */
API void snde_oracle_smooth_loss_bwd(int M, int h, int w, int C, const real *depth, const real *img, real gscalar, real *gdepth) {
    double nx = (double)M * h * (w - 1), ny = (double)M * (h - 1) * w;
    real cx = (real)((double)gscalar / nx), cy = (real)((double)gscalar / ny);
    for (int m = 0; m < M; ++m) {
        const real *D = depth + (size_t)m * h * w, *I = img + (size_t)m * h * w * C;
        real *G = gdepth + (size_t)m * h * w;
        for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) {
            size_t a = (size_t)i * w + j;
            if (j + 1 < w) {
                real wgt = edge_weight(I, a, a + 1, C), v = (D[a] - D[a + 1]) * wgt;
                real sg = v > (real)0 ? (real)1 : (v < (real)0 ? (real)-1 : (real)0);
                G[a] += cx * sg * wgt; G[a + 1] -= cx * sg * wgt;
            }
            if (i + 1 < h) {
                real wgt = edge_weight(I, a, a + w, C), v = (D[a] - D[a + w]) * wgt;
                real sg = v > (real)0 ? (real)1 : (v < (real)0 ? (real)-1 : (real)0);
                G[a] += cy * sg * wgt; G[a + w] -= cy * sg * wgt;
            }
        }
    }
}

/* mean over all elements, double accumulation (torch.mean, models/baseline.py:1128-1129) */
API double snde_oracle_mean(size_t n, const real *v) {
    double s = 0; for (size_t i = 0; i < n; ++i) s += v[i];
    return s / (double)n;
}
