// snde_tile.cuh -- the fused loss tile program (one CTA = one tile of one "owner" image).
//
// Work decomposition.  Every image of the step owns exactly one depth map:
//   class F item: owner = target image b; "others" = its S source images, projected with
//                 P = K4 @ T_k               (the reference's fwd direction, baseline.py:903-906)
//   class R item: owner = source image n = k*B + b; other = the target image b, projected with
//                 P = K4 @ inverse(T_k)      (the bwd direction, baseline.py:909-913)
// For each other image the CTA reprojects the owner's pixels with the owner's depth, samples
// the other image (bilinear_sampler semantics), scores it against the owner with
// alpha*DSSIM + (1-alpha)*L1, and -- in the same pass -- back-propagates through SSIM's 3x3
// pools, the sampler, the projection and the back-projection into d/d(depth), d/dP, d/dKinv.
// The owner's edge-aware smoothness term (depth, image already in shared memory) rides along.
// Each depth gradient element is written exactly once, by the CTA whose inner tile holds it.
//
// Region = inner tile (60 x 16) + 2-pixel halo = 64 x 20: SSIM at p needs a 3x3 window of warped
// values, and dL/d(warped)(p) needs the SSIM coefficients of the 3x3 neighbours of p.
//
// Phases (barrier separated):
//   0  owner image + depth -> shared memory
//   A  per region pixel: reproject, 4-tap gather of the other image -> Y (+ tap derivatives G)
//   B  9 warp tasks = 3 channels x 3 row bands; each lane owns two adjacent columns and slides
//      a 3-row window down its band: row sums are formed once and reused by three output rows.
//      SSIM, the error value and the three backward coefficient maps come out of this phase.
//   C  same task shape on the coefficient maps: dL/d(warped_ch) -> written over Y in place
//   G  per inner pixel: dot with the tap derivatives, back through the projection
//   S  smoothness pairs, then the depth-gradient tile is written
//
// The code is written as phases over "all threads of the CTA" (SNDE_THREADS_BEGIN/END).  On the
// GPU that is `tid = threadIdx.x` + __syncthreads(); under the host-emulation test harness (g++,
// tests/host_emul) it is a loop over tid, which lets the tile logic be checked against the
// oracle without a GPU.  The product never runs the host variant.
#pragma once
#include "snde_math.cuh"

#ifndef SNDE_MAX_SCALES
#define SNDE_MAX_SCALES 6
#endif

namespace snde {

struct TileCfg {
    static constexpr int RW = 64;        // region width  (32 lanes x 2 columns)
    static constexpr int TH = 16;        // inner tile height
    static constexpr int RH = TH + 4;    // region height
    static constexpr int TW = RW - 4;    // inner tile width (60)
    static constexpr int NWARP = 9;      // 3 channels x 3 row bands
    static constexpr int NT = NWARP * 32;
    static constexpr int PPT = (TH * RW + NT - 1) / NT;   // inner-pixel slots per thread (4)
    static constexpr int BAND = 6;       // rows per band (phase B: 18 rows, phase C: 16 rows)
    static constexpr int MAX_S = 8;      // most source views per target (SNDE_MAX_SOURCES)
};

struct DevImage { const float *p; long long sn; int sc, sh, sw, pad_; };
struct DevImageOut { float *p; long long sn; int sc, sh, sw, pad_; };
struct DevMap { const float *p; long long sn; int sh, sw; };
struct DevMapOut { float *p; long long sn; int sh, sw; };

struct DevScale {
    int h, w, tiles_x, tiles_y;
    DevImage tgt, src;
    DevMap depth;
    DevMapOut gdepth;
    const float *xtab, *ytab;
    const float *P;        // workspace [N][2][12]  (dir 0: K4@T, dir 1: K4@inv(T))
    const float *Kinv;     // workspace [B][9]
    double *gP;            // workspace [N][2][12]
    double *gKinv;         // workspace [B][9]
    double *loss;          // workspace [4]: sum fwd err, sum bwd err, sum |dx wx|, sum |dy wy|
    float gphoto;          // photo_weight / (N*h*w*3): upstream gradient of every error-map element
    float gsx, gsy;        // smooth_weight / (M*h*(w-1)),  / (M*(h-1)*w)
    int pad_;
    DevMap mask_fwd, mask_bwd;
    DevImageOut warped_fwd, warped_bwd, err_fwd, err_bwd;
};

struct DevEntry { int begin, count, scale, cls; };   // cls 0 = F items, 1 = R items

struct DevLoss {
    int B, S, num_scales, flags, padding, num_entries;
    float alpha, oma;
    DevEntry entry[2 * SNDE_MAX_SCALES];
    DevScale sc[SNDE_MAX_SCALES];
};

enum { FLAG_GRAD = 1, FLAG_GRAD_K = 2, FLAG_NO_PHOTO = 4, FLAG_NO_SMOOTH = 8 };   // == SNDE_LOSS_* in include/snde.h
// acc[0..11] = dP of the current other image; tail[0..8] = dKinv, tail[9] = err, tail[10], tail[11] = smoothness sums
enum { T_DK = 0, T_ERR = 9, T_SX = 10, T_SY = 11 };

// per-thread state that lives across phases (registers on the GPU)
struct Regs {
    float gd[TileCfg::PPT];
    float dP[12];
    float tail[12];
};

struct TileSmem {
    float X[3][TileCfg::RH][TileCfg::RW];         // owner image, 0 outside the image (= the pools' zero padding)
    float Y[3][TileCfg::RH][TileCfg::RW];         // warped other image, 0 outside the image; phase C overwrites inner pixels with dL/dY
    float D[TileCfg::RH][TileCfg::RW];            // owner depth
    float Co[9][TileCfg::RH - 2][TileCfg::RW];    // [ch*3+m]: SSIM backward coefficient maps (rows = region rows 1..RH-2); reused for the smoothness pair weights
    float G[6][TileCfg::TH][TileCfg::RW];         // d(warped_c)/dx, /dy of the inner pixels
    float red[TileCfg::MAX_S + 1][TileCfg::NWARP][16];   // per-warp partial sums: slot k = dP of other k; last used slot = dKinv, err, smooth
    float PK[2][24];                              // P (12) and Kinv (9) of the current / next other image
    float xt[TileCfg::RW], yt[TileCfg::RH];       // meshgrid tables of the region's columns / rows
};

#if defined(__CUDA_ARCH__)
#define SNDE_THREADS_BEGIN { const int tid = (int)threadIdx.x; Regs &rg = regs[0]; (void)rg;
#define SNDE_THREADS_END } __syncthreads();
#else
#define SNDE_THREADS_BEGIN for (int tid = 0; tid < TileCfg::NT; ++tid) { Regs &rg = regs[tid]; (void)rg;
#define SNDE_THREADS_END }
#endif

// Pin a pointer in a register pair: stops the compiler from re-deriving it from (struct field + 64-bit offset) at
// every use, so that base[int_offset] costs one IMAD.WIDE instead of a four-instruction 64-bit add-and-shift.
#if defined(__CUDA_ARCH__)
#define SNDE_PIN_PTR(p) asm volatile("" : "+l"(p))
#else
#define SNDE_PIN_PTR(p) ((void)0)
#endif

#if defined(__CUDACC__)
#define SNDE_TILE_FN __device__ __forceinline__
#else
#define SNDE_TILE_FN inline
#endif

// Warp-level sum of up to 16 per-lane values into dst[0..15] (dst = this warp's row of sm.red).  On the GPU a
// transposing butterfly: every xor step halves the number of live values, 16 shuffles in all instead of 5 per value.
SNDE_TILE_FN void warp_sum16(const float *v12, float *dst, int lane) {
#if defined(__CUDA_ARCH__)
    float v[16];
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = v12[i];
#pragma unroll
    for (int i = 12; i < 16; ++i) v[i] = 0.f;
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const int mask = half * 2;                      // lane bit consumed at this step: 16, 8, 4, 2
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    if ((lane & 1) == 0) dst[((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = v[0];
#else
    // host emulation: called once per thread; lane 0 zeroes the row, every lane adds its values
    if (lane == 0) for (int i = 0; i < 16; ++i) dst[i] = 0.f;
    for (int i = 0; i < 12; ++i) dst[i] += v12[i];
#endif
}

// horizontal 3-sums of one region row for the two columns (c0, c0+1) a lane owns
struct RowSums { float x[2], y[2], xx[2], yy[2], xy[2], cx[2], cy[2]; };

SNDE_HD void row_sums(const float *Xrow, const float *Yrow, int c0, int cl, int cr, RowSums &o) {
    const F2 xv = *reinterpret_cast<const F2 *>(Xrow + c0), yv = *reinterpret_cast<const F2 *>(Yrow + c0);
    const float xm = Xrow[cl], xp = Xrow[cr], ym = Yrow[cl], yp = Yrow[cr];
    const float tx = xv.x + xv.y, ty = yv.x + yv.y;
    const float txx = fmaf(xv.y, xv.y, xv.x * xv.x), tyy = fmaf(yv.y, yv.y, yv.x * yv.x), txy = fmaf(xv.y, yv.y, xv.x * yv.x);
    o.x[0] = xm + tx;             o.x[1] = tx + xp;
    o.y[0] = ym + ty;             o.y[1] = ty + yp;
    o.xx[0] = fmaf(xm, xm, txx);  o.xx[1] = fmaf(xp, xp, txx);
    o.yy[0] = fmaf(ym, ym, tyy);  o.yy[1] = fmaf(yp, yp, tyy);
    o.xy[0] = fmaf(xm, ym, txy);  o.xy[1] = fmaf(xp, yp, txy);
    o.cx[0] = xv.x; o.cx[1] = xv.y; o.cy[0] = yv.x; o.cy[1] = yv.y;
}

struct CoSums { float a[2], b[2], c[2]; };

SNDE_HD void co_row_sums(const float *A, const float *Bm, const float *Cm, int c0, int cl, int cr, CoSums &o) {
    const F2 av = *reinterpret_cast<const F2 *>(A + c0), bv = *reinterpret_cast<const F2 *>(Bm + c0), cv = *reinterpret_cast<const F2 *>(Cm + c0);
    const float ta = av.x + av.y, tb = bv.x + bv.y, tc = cv.x + cv.y;
    o.a[0] = A[cl] + ta;  o.a[1] = ta + A[cr];
    o.b[0] = Bm[cl] + tb; o.b[1] = tb + Bm[cr];
    o.c[0] = Cm[cl] + tc; o.c[1] = tc + Cm[cr];
}

// The tile program.  GRAD: also back-propagate.  GRADK: also accumulate d/dKinv.
// EXTRA: honour the optional explainability masks and warped / error-map outputs.
template <bool GRAD, bool GRADK, bool EXTRA>
SNDE_TILE_FN void loss_tile(const DevLoss &p, int block_id, TileSmem &sm, Regs *regs) {
    using C = TileCfg;
    // ---- decode the work item (uniform over the CTA) ----
    int e = 0;
    while (e + 1 < p.num_entries && block_id >= p.entry[e + 1].begin) ++e;
    const DevEntry ent = p.entry[e];
    const DevScale &sc = p.sc[ent.scale];
    const int local = block_id - ent.begin;
    const int tiles = sc.tiles_x * sc.tiles_y;
    const int m = local / tiles;                 // owner index inside its class
    const int t = local - m * tiles;
    const int ty = t / sc.tiles_x, tx = t - ty * sc.tiles_x;
    const int ry0 = ty * C::TH - 2, rx0 = tx * C::TW - 2;
    const int h = sc.h, w = sc.w;
    const int B = p.B, S = p.S;
    const bool clsF = ent.cls == 0;
    const int b = clsF ? m : (m % B);            // batch element (selects K)
    const DevImage &own = clsF ? sc.tgt : sc.src;
    const float *own_base = own.p + (long long)m * own.sn;
    const long long depth_n = clsF ? m : (long long)B + m;
    const float *depth_base = sc.depth.p + depth_n * sc.depth.sn;
    const int n_others = (p.flags & FLAG_NO_PHOTO) ? 0 : (clsF ? S : 1);
    const bool do_smooth = !(p.flags & FLAG_NO_SMOOTH);
    const float alpha = p.alpha, oma = p.oma;
    const float *Kinv = sc.Kinv + 9 * b;

    // ---- phase 0: owner image + depth into shared memory, zero the per-thread state ----
    SNDE_THREADS_BEGIN
#pragma unroll
        for (int i = 0; i < C::PPT; ++i) rg.gd[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) { rg.dP[i] = 0.f; rg.tail[i] = 0.f; }
        if (n_others > 0 && tid < 21) {           // stage P / Kinv of the first other image
            const long long n0 = clsF ? (long long)b : m;
            sm.PK[0][tid] = tid < 12 ? ldg(sc.P + (n0 * 2 + (clsF ? 0 : 1)) * 12 + tid) : ldg(Kinv + tid - 12);
        }
        if (tid < C::RW) { const int ix = rx0 + tid; sm.xt[tid] = (ix >= 0 && ix < w) ? ldg(sc.xtab + ix) : 0.f; }
        else if (tid < C::RW + C::RH) { const int iy = ry0 + tid - C::RW; sm.yt[tid - C::RW] = (iy >= 0 && iy < h) ? ldg(sc.ytab + iy) : 0.f; }
        const float *ob = own_base, *db = depth_base;
        SNDE_PIN_PTR(ob); SNDE_PIN_PTR(db);
        const int osc = own.sc, osh = own.sh, osw = own.sw, dsh = sc.depth.sh, dsw = sc.depth.sw;
        for (int idx = tid; idx < C::RH * C::RW; idx += C::NT) {
            const int r = idx >> 6, c = idx & 63;
            const int iy = ry0 + r, ix = rx0 + c;
            const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f, d = 0.f;
            if (in) {
                const int q = iy * osh + ix * osw;
                x0 = ldg(ob + q); x1 = ldg(ob + (q + osc)); x2 = ldg(ob + (q + 2 * osc));
                d = ldg(db + (iy * dsh + ix * dsw));
            }
            sm.X[0][r][c] = x0; sm.X[1][r][c] = x1; sm.X[2][r][c] = x2; sm.D[r][c] = d;
        }
    SNDE_THREADS_END

    for (int k = 0; k < n_others; ++k) {
        const long long n = clsF ? (long long)k * B + b : m;      // index into the N-sized tensors
        const DevImage &oth = clsF ? sc.src : sc.tgt;
        const float *oth_base = oth.p + (clsF ? n : (long long)b) * oth.sn;
        const DevMap &mask = clsF ? sc.mask_fwd : sc.mask_bwd;
        const DevImageOut &wout = clsF ? sc.warped_fwd : sc.warped_bwd;
        const DevImageOut &eout = clsF ? sc.err_fwd : sc.err_bwd;
        const bool has_mask = EXTRA && mask.p != nullptr;
        const float *PKs = sm.PK[k & 1];
        const int oth_sc = oth.sc, oth_sh = oth.sh, oth_sw = oth.sw;

        // ---- phase A: reproject + sample every in-image region pixel ----
        SNDE_THREADS_BEGIN
            float P[12], Ki[9];
#pragma unroll
            for (int i = 0; i < 12; ++i) P[i] = PKs[i];
#pragma unroll
            for (int i = 0; i < 9; ++i) Ki[i] = PKs[12 + i];
            if (k + 1 < n_others && tid < 21) {       // stage the next other image's matrices (read after two more barriers)
                const long long n1 = (long long)(k + 1) * B + b;
                sm.PK[(k + 1) & 1][tid] = tid < 12 ? ldg(sc.P + (n1 * 2) * 12 + tid) : ldg(Kinv + tid - 12);
            }
            const float *obase = oth_base;
            SNDE_PIN_PTR(obase);
            // Two region pixels per iteration: both coordinate chains first, then all 24 gathers back to back (memory-level
            // parallelism: the gathers are the long-latency part of the tile), then the blends.
            for (int idx0 = tid; idx0 < C::RH * C::RW; idx0 += 2 * C::NT) {
                int rr[2], cc[2], o00[2], o01[2], o10[2], o11[2];
                bool in[2], valid[2], inner[2];
                Taps tp[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int idx = idx0 + j * C::NT;
                    const int r = idx >> 6, c = idx & 63;
                    rr[j] = r; cc[j] = c;
                    const int iy = ry0 + r, ix = rx0 + c;
                    in[j] = idx < C::RH * C::RW && iy >= 0 && iy < h && ix >= 0 && ix < w;
                    valid[j] = false; inner[j] = false;
                    o00[j] = o01[j] = o10[j] = o11[j] = 0;
                    if (in[j]) {
                        const float xt = sm.xt[c], yt = sm.yt[r];
                        Reproj rp; reproject_fast(Ki, P, xt, yt, sm.D[r][c], rp);
                        float x, y; flow_coords(rp, xt, yt, x, y);
                        // Outside [0,w-1) x [0,h-1) the reference's clamped-corner weights cancel to zero (up to rounding
                        // residue that grows with the distance from the image); the fused path writes the exact zero.
                        if (EXTRA && p.padding) { make_taps(x, y, h, w, p.padding, tp[j]); valid[j] = true; }   // border mode samples everywhere
                        else { make_taps_valid(x, y, tp[j]); valid[j] = tap_valid(x, y, h, w, 0) != 0.f; }
                        inner[j] = r >= 2 && r < C::TH + 2 && c >= 2 && c < C::RW - 2;
                        if (valid[j]) {
                            o00[j] = tp[j].y0 * oth_sh + tp[j].x0 * oth_sw; o01[j] = tp[j].y1 * oth_sh + tp[j].x0 * oth_sw;
                            o10[j] = tp[j].y0 * oth_sh + tp[j].x1 * oth_sw; o11[j] = tp[j].y1 * oth_sh + tp[j].x1 * oth_sw;
                        }
                    }
                }
                float v[2][12];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const int oc = ch * oth_sc;          // 32-bit offsets: one IMAD.WIDE per gather
                        v[j][4 * ch + 0] = valid[j] ? ldg(obase + (o00[j] + oc)) : 0.f;
                        v[j][4 * ch + 1] = valid[j] ? ldg(obase + (o01[j] + oc)) : 0.f;
                        v[j][4 * ch + 2] = valid[j] ? ldg(obase + (o10[j] + oc)) : 0.f;
                        v[j][4 * ch + 3] = valid[j] ? ldg(obase + (o11[j] + oc)) : 0.f;
                    }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (idx0 + j * C::NT >= C::RH * C::RW) continue;
                    const int r = rr[j], c = cc[j];
                    float yv[3] = {0.f, 0.f, 0.f};
                    if (valid[j]) {
                        const float w00 = tp[j].wx0 * tp[j].wy0, w01 = tp[j].wx0 * tp[j].wy1, w10 = tp[j].wx1 * tp[j].wy0, w11 = tp[j].wx1 * tp[j].wy1;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float i00 = v[j][4 * ch], i01 = v[j][4 * ch + 1], i10 = v[j][4 * ch + 2], i11 = v[j][4 * ch + 3];
                            yv[ch] = tap_blend_fma(w00, w01, w10, w11, i00, i01, i10, i11);
                            if (GRAD && inner[j]) {
                                float ddx, ddy; tap_derivs_fma(tp[j], i00, i01, i10, i11, ddx, ddy);
                                sm.G[2 * ch][r - 2][c] = ddx; sm.G[2 * ch + 1][r - 2][c] = ddy;
                            }
                        }
                    } else if (GRAD && inner[j]) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) sm.G[q][r - 2][c] = 0.f;
                    }
                    if (EXTRA && inner[j] && wout.p) {
                        float *o = wout.p + n * wout.sn + (ry0 + r) * wout.sh + (rx0 + c) * wout.sw;
                        o[0] = yv[0]; o[wout.sc] = yv[1]; o[2 * wout.sc] = yv[2];
                    }
                    sm.Y[0][r][c] = yv[0]; sm.Y[1][r][c] = yv[1]; sm.Y[2][r][c] = yv[2];
                }
            }
        SNDE_THREADS_END

        // ---- phase B: SSIM + error value + backward coefficients; warp task = (channel, band of 6 rows) ----
        SNDE_THREADS_BEGIN
            const int warp = tid >> 5, lane = tid & 31;
            const int ch = warp % 3, band = warp / 3;
            const int ra = 1 + C::BAND * band;                 // output rows ra .. ra+5 (region rows 1..18)
            const int c0 = 2 * lane, cl = c0 > 0 ? c0 - 1 : 0, cr = c0 + 2 < C::RW ? c0 + 2 : C::RW - 1;
            const float gS0 = -0.5f * alpha * sc.gphoto;
            RowSums ra_, rb_, rc_;
            row_sums(sm.X[ch][ra - 1], sm.Y[ch][ra - 1], c0, cl, cr, ra_);
            row_sums(sm.X[ch][ra], sm.Y[ch][ra], c0, cl, cr, rb_);
            float err_acc = 0.f;
#pragma unroll
            for (int j = 0; j < C::BAND; ++j) {
                const int r = ra + j;
                row_sums(sm.X[ch][r + 1], sm.Y[ch][r + 1], c0, cl, cr, rc_);
                const int iy = ry0 + r;
                const bool row_in = iy >= 0 && iy < h;
                const bool row_inner = r >= 2 && r < C::TH + 2;
                float oa[2], ob[2], oc[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = c0 + q, ix = rx0 + c;
                    const bool valid = row_in && c >= 1 && c <= C::RW - 2 && ix >= 0 && ix < w;
                    const float sx = (ra_.x[q] + rb_.x[q]) + rc_.x[q], sy = (ra_.y[q] + rb_.y[q]) + rc_.y[q];
                    const float sxx = (ra_.xx[q] + rb_.xx[q]) + rc_.xx[q], syy = (ra_.yy[q] + rb_.yy[q]) + rc_.yy[q];
                    const float sxy = (ra_.xy[q] + rb_.xy[q]) + rc_.xy[q];
                    SsimS ss; ssim_sums(sx, sy, sxx, syy, sxy, ss);
                    const float v = fmaf(-0.5f, ss.S, 0.5f);
                    const float err = fmaf(alpha, clampf(v, 0.f, 1.f), oma * fabsf(rb_.cx[q] - rb_.cy[q]));
                    float mk = 1.f;
                    if (has_mask && valid) mk = ldg(mask.p + n * mask.sn + iy * mask.sh + ix * mask.sw);
                    if (valid && row_inner && c >= 2 && c < C::RW - 2) {
                        err_acc += EXTRA ? err * mk : err;
                        if (EXTRA && eout.p) eout.p[n * eout.sn + ch * eout.sc + iy * eout.sh + ix * eout.sw] = err;
                    }
                    float ca = 0.f, cb = 0.f, cc = 0.f;
                    if (GRAD) {
                        const float gS = (valid && v >= 0.f && v <= 1.f) ? (EXTRA ? gS0 * mk : gS0) : 0.f;
                        ssim_sums_bwd(ss, sx, sy, gS, ca, cb, cc);
                    }
                    oa[q] = ca; ob[q] = cb; oc[q] = cc;
                }
                if (GRAD) {
                    *reinterpret_cast<F2 *>(&sm.Co[ch * 3 + 0][r - 1][c0]) = F2{oa[0], oa[1]};
                    *reinterpret_cast<F2 *>(&sm.Co[ch * 3 + 1][r - 1][c0]) = F2{ob[0], ob[1]};
                    *reinterpret_cast<F2 *>(&sm.Co[ch * 3 + 2][r - 1][c0]) = F2{oc[0], oc[1]};
                }
                ra_ = rb_; rb_ = rc_;
            }
            rg.tail[T_ERR] += err_acc;
        SNDE_THREADS_END

        if (GRAD) {
            // ---- phase C: dL/d(warped_ch) at the inner pixels, written over Y in place ----
            SNDE_THREADS_BEGIN
                const int warp = tid >> 5, lane = tid & 31;
                const int ch = warp % 3, band = warp / 3;
                const int ra = 2 + C::BAND * band;             // output rows ra .. min(ra+5, TH+1)  (region rows 2..17)
                const int c0 = 2 * lane, cl = c0 > 0 ? c0 - 1 : 0, cr = c0 + 2 < C::RW ? c0 + 2 : C::RW - 1;
                const float g1 = oma * sc.gphoto;
                const float *A = &sm.Co[ch * 3 + 0][0][0], *Bm = &sm.Co[ch * 3 + 1][0][0], *Cm = &sm.Co[ch * 3 + 2][0][0];
                CoSums sa, sb, sc_;
                co_row_sums(A + (ra - 2) * C::RW, Bm + (ra - 2) * C::RW, Cm + (ra - 2) * C::RW, c0, cl, cr, sa);   // Co row index = region row - 1
                co_row_sums(A + (ra - 1) * C::RW, Bm + (ra - 1) * C::RW, Cm + (ra - 1) * C::RW, c0, cl, cr, sb);
#pragma unroll
                for (int j = 0; j < C::BAND; ++j) {
                    const int r = ra + j;
                    if (r < C::TH + 2) {
                        co_row_sums(A + r * C::RW, Bm + r * C::RW, Cm + r * C::RW, c0, cl, cr, sc_);
                        const F2 xv = *reinterpret_cast<const F2 *>(&sm.X[ch][r][c0]), yv = *reinterpret_cast<const F2 *>(&sm.Y[ch][r][c0]);
                        const float xs[2] = {xv.x, xv.y}, ys[2] = {yv.x, yv.y};
                        const int iy = ry0 + r;
                        float po[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float ba = (sa.a[q] + sb.a[q]) + sc_.a[q], bb = (sa.b[q] + sb.b[q]) + sc_.b[q], bc = (sa.c[q] + sb.c[q]) + sc_.c[q];
                            float gl = g1;
                            if (has_mask) {
                                const int ix = rx0 + c0 + q;
                                if (iy < h && ix >= 0 && ix < w) gl = g1 * ldg(mask.p + n * mask.sn + iy * mask.sh + ix * mask.sw);
                            }
                            const float d = xs[q] - ys[q];
                            const float l1 = d > 0.f ? gl : (d < 0.f ? -gl : 0.f);
                            po[q] = fmaf(ys[q] + ys[q], bb, fmaf(xs[q], bc, ba)) - l1;
                        }
                        *reinterpret_cast<F2 *>(&sm.Y[ch][r][c0]) = F2{po[0], po[1]};
                        sa = sb; sb = sc_;
                    }
                }
            SNDE_THREADS_END

            // ---- phase G: dot with the tap derivatives, back through the projection (inner pixels) ----
            SNDE_THREADS_BEGIN
                float P[12], Ki[9];
#pragma unroll
                for (int i = 0; i < 12; ++i) P[i] = PKs[i];
#pragma unroll
                for (int i = 0; i < 9; ++i) Ki[i] = PKs[12 + i];
#pragma unroll
                for (int i = 0; i < 12; ++i) rg.dP[i] = 0.f;
#pragma unroll
                for (int i = 0; i < C::PPT; ++i) {
                    const int idx = tid + i * C::NT;
                    const int row = idx >> 6, c = idx & 63;
                    const int r = row + 2, iy = ry0 + r, ix = rx0 + c;
                    if (idx >= C::TH * C::RW || c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;   // iy, ix >= 0 for inner pixels
                    const float g0 = sm.Y[0][r][c], g1_ = sm.Y[1][r][c], g2 = sm.Y[2][r][c];
                    const float gx = fmaf(g2, sm.G[4][row][c], fmaf(g1_, sm.G[2][row][c], g0 * sm.G[0][row][c]));
                    const float gy = fmaf(g2, sm.G[5][row][c], fmaf(g1_, sm.G[3][row][c], g0 * sm.G[1][row][c]));
                    const float xt = sm.xt[c], yt = sm.yt[r], d = sm.D[r][c];
                    Reproj rp; reproject_nodiv(Ki, P, xt, yt, d, rp);
                    float gu[3], gcam[3];
                    rg.gd[i] += reproject_bwd_fast(P, rp, gx, gy, gu, gcam);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        rg.dP[q * 4 + 0] = fmaf(gu[q], rp.cam[0], rg.dP[q * 4 + 0]);
                        rg.dP[q * 4 + 1] = fmaf(gu[q], rp.cam[1], rg.dP[q * 4 + 1]);
                        rg.dP[q * 4 + 2] = fmaf(gu[q], rp.cam[2], rg.dP[q * 4 + 2]);
                        rg.dP[q * 4 + 3] += gu[q];
                    }
                    if (GRADK) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float gray = gcam[q] * d;
                            rg.tail[T_DK + q * 3 + 0] = fmaf(gray, xt, rg.tail[T_DK + q * 3 + 0]);
                            rg.tail[T_DK + q * 3 + 1] = fmaf(gray, yt, rg.tail[T_DK + q * 3 + 1]);
                            rg.tail[T_DK + q * 3 + 2] += gray;
                        }
                    }
                }
                warp_sum16(rg.dP, sm.red[k][tid >> 5], tid & 31);     // combined across warps at the end of the tile
            SNDE_THREADS_END
        }
    }

    if (do_smooth) {
        // ---- phase S1: signed pair weights sign(dx*wx)*wx of the pairs starting at region rows 1..TH+1 ----
        float (*SXw)[C::RW] = sm.Co[0];                       // [TH+1][RW]
        float (*SYw)[C::RW] = (float (*)[C::RW])(&sm.Co[0][0][0] + (C::TH + 1) * C::RW);
        SNDE_THREADS_BEGIN
            for (int idx = tid; idx < (C::TH + 1) * C::RW; idx += C::NT) {
                const int pr = idx >> 6, c = idx & 63;
                const int r = pr + 1, iy = ry0 + r, ix = rx0 + c;
                float vx = 0.f, vy = 0.f;
                if (c >= 1 && c <= C::RW - 2 && iy >= 0 && iy < h && ix >= 0 && ix < w) {
                    const bool inner = r >= 2 && r < C::TH + 2 && c >= 2 && c < C::RW - 2;
                    const float d = sm.D[r][c];
                    if (ix + 1 < w) {
                        const float wgt = edge_weight3_fast(sm.X[0][r][c], sm.X[1][r][c], sm.X[2][r][c],
                                                       sm.X[0][r][c + 1], sm.X[1][r][c + 1], sm.X[2][r][c + 1]);
                        const float v = (d - sm.D[r][c + 1]) * wgt;
                        vx = v > 0.f ? wgt : (v < 0.f ? -wgt : 0.f);
                        if (inner) rg.tail[T_SX] += fabsf(v);
                    }
                    if (iy + 1 < h) {
                        const float wgt = edge_weight3_fast(sm.X[0][r][c], sm.X[1][r][c], sm.X[2][r][c],
                                                       sm.X[0][r + 1][c], sm.X[1][r + 1][c], sm.X[2][r + 1][c]);
                        const float v = (d - sm.D[r + 1][c]) * wgt;
                        vy = v > 0.f ? wgt : (v < 0.f ? -wgt : 0.f);
                        if (inner) rg.tail[T_SY] += fabsf(v);
                    }
                }
                SXw[pr][c] = vx; SYw[pr][c] = vy;
            }
        SNDE_THREADS_END
        if (GRAD) {
            // ---- phase S2: gather the (up to) four pair terms that touch each inner pixel ----
            SNDE_THREADS_BEGIN
#pragma unroll
                for (int i = 0; i < C::PPT; ++i) {
                    const int idx = tid + i * C::NT;
                    const int row = idx >> 6, c = idx & 63;
                    const int r = row + 2, iy = ry0 + r, ix = rx0 + c;
                    if (idx >= C::TH * C::RW || c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;
                    const int pr = r - 1;
                    rg.gd[i] += sc.gsx * (SXw[pr][c] - SXw[pr][c - 1]) + sc.gsy * (SYw[pr][c] - SYw[pr - 1][c]);
                }
            SNDE_THREADS_END
        }
    }

    // ---- write the depth gradient, publish the partial sums ----
    if (GRAD) {
        SNDE_THREADS_BEGIN
            float *gbase = sc.gdepth.p + depth_n * sc.gdepth.sn;
#pragma unroll
            for (int i = 0; i < C::PPT; ++i) {
                const int idx = tid + i * C::NT;
                const int row = idx >> 6, c = idx & 63;
                const int iy = ry0 + row + 2, ix = rx0 + c;
                if (idx >= C::TH * C::RW || c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;
                gbase[iy * sc.gdepth.sh + ix * sc.gdepth.sw] = rg.gd[i];
            }
        SNDE_THREADS_END
    }
    // ---- publish: per-warp partials -> shared memory -> one double atomicAdd per value per CTA ----
    const int nslots = (GRAD ? n_others : 0) + 1;
    SNDE_THREADS_BEGIN
        warp_sum16(rg.tail, sm.red[nslots - 1][tid >> 5], tid & 31);
    SNDE_THREADS_END
    SNDE_THREADS_BEGIN
        if (tid < 16 * nslots) {
            const int slot = tid >> 4, i = tid & 15;
            double sum = 0.0;
#pragma unroll
            for (int wq = 0; wq < C::NWARP; ++wq) sum += (double)sm.red[slot][wq][i];
            double *dst = nullptr;
            if (slot < nslots - 1) {
                const long long nk = clsF ? (long long)slot * B + b : m;
                if (i < 12) dst = sc.gP + (nk * 2 + (clsF ? 0 : 1)) * 12 + i;
            } else if (i < 9) { if (GRADK) dst = sc.gKinv + 9 * b + i; }
            else if (i == T_ERR) dst = sc.loss + (clsF ? 0 : 1);
            else if (i == T_SX || i == T_SY) dst = sc.loss + 2 + (i - T_SX);
            if (dst && sum != 0.0) {
#if defined(__CUDA_ARCH__)
                atomicAdd(dst, sum);
#else
                *dst += sum;
#endif
            }
        }
    SNDE_THREADS_END
}

}  // namespace snde
