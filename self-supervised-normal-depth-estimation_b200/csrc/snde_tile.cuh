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
// Region = inner tile + 2-pixel halo: SSIM at p needs a 3x3 window of warped values, and
// dL/d(warped)(p) needs the SSIM coefficients of the 3x3 neighbours of p.
//
// The code is written as barrier-separated phases over "all threads of the CTA"
// (SNDE_THREADS_BEGIN/END).  On the GPU that is `tid = threadIdx.x` + __syncthreads(); under the
// host-emulation test harness (g++, tests/host_emul) it is a loop over tid, which lets the tile
// logic be checked against the oracle without a GPU.  The product never runs the host variant.
#pragma once
#include "snde_math.cuh"

#ifndef SNDE_MAX_SCALES
#define SNDE_MAX_SCALES 6
#endif

namespace snde {

struct TileCfg {
    static constexpr int RW = 64;        // region width  (2 warps of columns)
    static constexpr int TH = 16;        // inner tile height
    static constexpr int RH = TH + 4;    // region height
    static constexpr int TW = RW - 4;    // inner tile width (60)
    static constexpr int NT = 256;       // threads per CTA
    static constexpr int PPT = TH * RW / NT;   // inner pixels owned per thread (4)
    static constexpr int NWARP = NT / 32;
};

struct DevImage { const float *p; long long sn, sc, sh, sw; };
struct DevImageOut { float *p; long long sn, sc, sh, sw; };
struct DevMap { const float *p; long long sn, sh, sw; };
struct DevMapOut { float *p; long long sn, sh, sw; };

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
enum { ACC_DP = 0, ACC_DK = 12, ACC_ERR = 21, ACC_SX = 22, ACC_SY = 23, ACC_N = 24 };

// per-thread state that lives across phases (registers on the GPU)
struct Regs {
    float gx[TileCfg::PPT], gy[TileCfg::PPT], gd[TileCfg::PPT];
    float acc[ACC_N];
};

struct TileSmem {
    float X[3][TileCfg::RH][TileCfg::RW];         // owner image, 0 outside the image (= the pools' zero padding)
    float Y[3][TileCfg::RH][TileCfg::RW];         // warped other image, 0 outside the image
    float D[TileCfg::RH][TileCfg::RW];            // owner depth
    float Co[3][TileCfg::RH - 2][TileCfg::RW];    // SSIM backward coefficient maps of the current channel; reused for the smoothness pair weights
    float G[6][TileCfg::TH][TileCfg::RW];         // d(warped_c)/dx, /dy of the inner pixels
    float red[TileCfg::NWARP][ACC_N];             // block-reduction scratch
};

#if defined(__CUDA_ARCH__)
#define SNDE_THREADS_BEGIN { const int tid = (int)threadIdx.x; Regs &rg = regs[0];
#define SNDE_THREADS_END } __syncthreads();
#else
#define SNDE_THREADS_BEGIN for (int tid = 0; tid < TileCfg::NT; ++tid) { Regs &rg = regs[tid];
#define SNDE_THREADS_END }
#endif

#if defined(__CUDACC__)
#define SNDE_TILE_FN __device__ __forceinline__
#else
#define SNDE_TILE_FN inline
#endif

// Sum acc[first..first+count) over the CTA and add each to dst[i*stride] (double atomics).
template <int first, int count>
SNDE_TILE_FN void block_reduce_add(TileSmem &sm, Regs *regs, double *dst) {
#if defined(__CUDA_ARCH__)
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < count; ++i) {
        float v = regs[0].acc[first + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm.red[warp][i] = v;
    }
    __syncthreads();
    if (tid < count) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < TileCfg::NWARP; ++w) s += (double)sm.red[w][tid];
        if (s != 0.0) atomicAdd(dst + tid, s);
    }
    __syncthreads();
#else
    (void)sm;
    for (int i = 0; i < count; ++i) {
        double s = 0.0;
        for (int t = 0; t < TileCfg::NT; ++t) s += (double)regs[t].acc[first + i];
        dst[i] += s;
    }
#endif
}

SNDE_HD float ld_img(const DevImage &im, long long n, int c, int y, int x) {
    return ldg(im.p + n * im.sn + c * im.sc + (long long)y * im.sh + (long long)x * im.sw);
}
SNDE_HD float ld_map(const DevMap &m, long long n, int y, int x) {
    return ldg(m.p + n * m.sn + (long long)y * m.sh + (long long)x * m.sw);
}

// The tile program.  GRAD: also back-propagate.  GRADK: also accumulate d/dKinv.
template <bool GRAD, bool GRADK>
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
    const long long own_n = m;
    const long long depth_n = clsF ? m : (long long)B + m;
    const int n_others = (p.flags & FLAG_NO_PHOTO) ? 0 : (clsF ? S : 1);
    const bool do_smooth = !(p.flags & FLAG_NO_SMOOTH);
    const float alpha = p.alpha, oma = p.oma;
    const float *Kinv = sc.Kinv + 9 * b;

    // ---- phase 0: owner image + depth into shared memory, zero the per-thread state ----
    SNDE_THREADS_BEGIN
#pragma unroll
        for (int i = 0; i < C::PPT; ++i) { rg.gx[i] = 0.f; rg.gy[i] = 0.f; rg.gd[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < ACC_N; ++i) rg.acc[i] = 0.f;
        for (int idx = tid; idx < C::RH * C::RW; idx += C::NT) {
            const int r = idx / C::RW, c = idx - r * C::RW;
            const int iy = ry0 + r, ix = rx0 + c;
            const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
            sm.X[0][r][c] = in ? ld_img(own, own_n, 0, iy, ix) : 0.f;
            sm.X[1][r][c] = in ? ld_img(own, own_n, 1, iy, ix) : 0.f;
            sm.X[2][r][c] = in ? ld_img(own, own_n, 2, iy, ix) : 0.f;
            sm.D[r][c] = in ? ld_map(sc.depth, depth_n, iy, ix) : 0.f;
        }
    SNDE_THREADS_END

    for (int k = 0; k < n_others; ++k) {
        const long long n = clsF ? (long long)k * B + b : m;      // index into the N-sized tensors
        const DevImage &oth = clsF ? sc.src : sc.tgt;
        const long long oth_n = clsF ? n : b;
        const float *Pg = sc.P + (n * 2 + (clsF ? 0 : 1)) * 12;
        const DevMap &mask = clsF ? sc.mask_fwd : sc.mask_bwd;
        const DevImageOut &wout = clsF ? sc.warped_fwd : sc.warped_bwd;
        const DevImageOut &eout = clsF ? sc.err_fwd : sc.err_bwd;
        float P[12], Ki[9];
        for (int i = 0; i < 12; ++i) P[i] = ldg(Pg + i);
        for (int i = 0; i < 9; ++i) Ki[i] = ldg(Kinv + i);

        // ---- phase A: reproject + sample every in-image region pixel ----
        SNDE_THREADS_BEGIN
            for (int idx = tid; idx < C::RH * C::RW; idx += C::NT) {
                const int r = idx / C::RW, c = idx - r * C::RW;
                const int iy = ry0 + r, ix = rx0 + c;
                const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
                float y0 = 0.f, y1 = 0.f, y2 = 0.f;
                if (in) {
                    const float xt = ldg(sc.xtab + ix), yt = ldg(sc.ytab + iy);
                    Reproj rp; reproject(Ki, P, xt, yt, sm.D[r][c], rp);
                    float x, y; flow_coords(rp, xt, yt, x, y);
                    Taps tp; make_taps(x, y, h, w, p.padding, tp);
                    const float *base = oth.p + oth_n * oth.sn;
                    const long long o00 = (long long)tp.y0 * oth.sh + (long long)tp.x0 * oth.sw;
                    const long long o01 = (long long)tp.y1 * oth.sh + (long long)tp.x0 * oth.sw;
                    const long long o10 = (long long)tp.y0 * oth.sh + (long long)tp.x1 * oth.sw;
                    const long long o11 = (long long)tp.y1 * oth.sh + (long long)tp.x1 * oth.sw;
                    const bool inner = r >= 2 && r < C::TH + 2 && c >= 2 && c < C::RW - 2;
                    float yv[3];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float *bc = base + ch * oth.sc;
                        const float i00 = ldg(bc + o00), i01 = ldg(bc + o01), i10 = ldg(bc + o10), i11 = ldg(bc + o11);
                        yv[ch] = tap_blend(tp, i00, i01, i10, i11);
                        if (GRAD && inner) {
                            float ddx, ddy; tap_derivs(tp, i00, i01, i10, i11, ddx, ddy);
                            sm.G[2 * ch][r - 2][c] = ddx; sm.G[2 * ch + 1][r - 2][c] = ddy;
                        }
                    }
                    y0 = yv[0]; y1 = yv[1]; y2 = yv[2];
                    if (inner && wout.p) {
                        float *o = wout.p + n * wout.sn + (long long)iy * wout.sh + (long long)ix * wout.sw;
                        o[0] = y0; o[wout.sc] = y1; o[2 * wout.sc] = y2;
                    }
                }
                sm.Y[0][r][c] = y0; sm.Y[1][r][c] = y1; sm.Y[2][r][c] = y2;
            }
        SNDE_THREADS_END

        for (int ch = 0; ch < 3; ++ch) {
            // ---- phase B: SSIM + error at every window centre of region rows 1..RH-2, cols 1..RW-2 ----
            SNDE_THREADS_BEGIN
                for (int idx = tid; idx < (C::RH - 2) * C::RW; idx += C::NT) {
                    const int r1 = idx / C::RW, c = idx - r1 * C::RW;
                    const int r = r1 + 1;
                    if (c == 0 || c == C::RW - 1) continue;
                    const int iy = ry0 + r, ix = rx0 + c;
                    const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
                    const bool inner = r >= 2 && r < C::TH + 2 && c >= 2 && c < C::RW - 2;
                    float ca = 0.f, cb = 0.f, cc = 0.f;
                    if (in && (GRAD || inner)) {
                        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
                        for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                            for (int dc = -1; dc <= 1; ++dc) {
                                const float xv = sm.X[ch][r + dr][c + dc], yv = sm.Y[ch][r + dr][c + dc];
                                sx = sx + xv; sy = sy + yv; sxx = sxx + xv * xv; syy = syy + yv * yv; sxy = sxy + xv * yv;
                            }
                        Ssim ss; ssim_from_sums(sx, sy, sxx, syy, sxy, ss);
                        float v;
                        const float err = similarity_value(alpha, oma, ss.S, sm.X[ch][r][c], sm.Y[ch][r][c], v);
                        const float mk = mask.p ? ld_map(mask, n, iy, ix) : 1.f;
                        if (inner) {
                            rg.acc[ACC_ERR] += err * mk;
                            if (eout.p) eout.p[n * eout.sn + ch * eout.sc + (long long)iy * eout.sh + (long long)ix * eout.sw] = err;
                        }
                        if (GRAD) {
                            const float g = sc.gphoto * mk;
                            const float gS = (v >= 0.f && v <= 1.f) ? -0.5f * alpha * g : 0.f;
                            float cmx;
                            ssim_bwd_coeffs(ss, gS, cmx, ca, cb, cc);
                        }
                    }
                    if (GRAD) { sm.Co[0][r1][c] = ca; sm.Co[1][r1][c] = cb; sm.Co[2][r1][c] = cc; }
                }
            SNDE_THREADS_END
            if (GRAD) {
                // ---- phase C: dL/d(warped_ch) at the inner pixels -> accumulate dL/dx, dL/dy ----
                SNDE_THREADS_BEGIN
#pragma unroll
                    for (int i = 0; i < C::PPT; ++i) {
                        const int idx = tid + i * C::NT;
                        const int row = idx / C::RW, c = idx - row * C::RW;
                        const int r = row + 2, iy = ry0 + r, ix = rx0 + c;
                        if (c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;   // iy, ix >= 0 for inner pixels
                        float ba = 0.f, bb = 0.f, bc = 0.f;
#pragma unroll
                        for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                            for (int dc = -1; dc <= 1; ++dc) {
                                ba += sm.Co[0][r - 1 + dr][c + dc]; bb += sm.Co[1][r - 1 + dr][c + dc]; bc += sm.Co[2][r - 1 + dr][c + dc];
                            }
                        const float xv = sm.X[ch][r][c], yv = sm.Y[ch][r][c];
                        const float mk = mask.p ? ld_map(mask, n, iy, ix) : 1.f;
                        const float l1 = oma * (sc.gphoto * mk) * signf(xv - yv);
                        const float gyh = (ba + 2.f * yv * bb + xv * bc) / 9.f - l1;
                        rg.gx[i] += gyh * sm.G[2 * ch][row][c];
                        rg.gy[i] += gyh * sm.G[2 * ch + 1][row][c];
                    }
                SNDE_THREADS_END
            }
        }

        if (GRAD) {
            // ---- phase G: back through the projection for the inner pixels ----
            SNDE_THREADS_BEGIN
#pragma unroll
                for (int i = 0; i < C::PPT; ++i) {
                    const int idx = tid + i * C::NT;
                    const int row = idx / C::RW, c = idx - row * C::RW;
                    const int r = row + 2, iy = ry0 + r, ix = rx0 + c;
                    const float gx = rg.gx[i], gy = rg.gy[i];
                    rg.gx[i] = 0.f; rg.gy[i] = 0.f;
                    if (c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;
                    const float xt = ldg(sc.xtab + ix), yt = ldg(sc.ytab + iy), d = sm.D[r][c];
                    Reproj rp; reproject(Ki, P, xt, yt, d, rp);
                    float gu[3], gcam[3];
                    rg.gd[i] += reproject_bwd(P, rp, gx, gy, gu, gcam);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        rg.acc[ACC_DP + q * 4 + 0] += gu[q] * rp.cam[0];
                        rg.acc[ACC_DP + q * 4 + 1] += gu[q] * rp.cam[1];
                        rg.acc[ACC_DP + q * 4 + 2] += gu[q] * rp.cam[2];
                        rg.acc[ACC_DP + q * 4 + 3] += gu[q];
                    }
                    if (GRADK) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float gray = gcam[q] * d;
                            rg.acc[ACC_DK + q * 3 + 0] += gray * xt;
                            rg.acc[ACC_DK + q * 3 + 1] += gray * yt;
                            rg.acc[ACC_DK + q * 3 + 2] += gray;
                        }
                    }
                }
            SNDE_THREADS_END
            block_reduce_add<ACC_DP, 12>(sm, regs, sc.gP + (n * 2 + (clsF ? 0 : 1)) * 12);
            SNDE_THREADS_BEGIN
#pragma unroll
                for (int i = 0; i < 12; ++i) rg.acc[ACC_DP + i] = 0.f;
                (void)tid;
            SNDE_THREADS_END
        }
    }

    if (do_smooth) {
        // ---- phase S1: signed pair weights sign(dx*wx)*wx of the pairs starting at region rows 1..TH+1 ----
        float (*SXw)[C::RW] = sm.Co[0];                       // [TH+1][RW]
        float (*SYw)[C::RW] = (float (*)[C::RW])(&sm.Co[0][0][0] + (C::TH + 1) * C::RW);
        SNDE_THREADS_BEGIN
            for (int idx = tid; idx < (C::TH + 1) * C::RW; idx += C::NT) {
                const int pr = idx / C::RW, c = idx - pr * C::RW;
                const int r = pr + 1, iy = ry0 + r, ix = rx0 + c;
                float vx = 0.f, vy = 0.f;
                if (c >= 1 && c <= C::RW - 2 && iy >= 0 && iy < h && ix >= 0 && ix < w) {
                    const bool inner = r >= 2 && r < C::TH + 2 && c >= 2 && c < C::RW - 2;
                    const float d = sm.D[r][c];
                    if (ix + 1 < w) {
                        const float wgt = edge_weight3(sm.X[0][r][c], sm.X[1][r][c], sm.X[2][r][c],
                                                       sm.X[0][r][c + 1], sm.X[1][r][c + 1], sm.X[2][r][c + 1]);
                        const float v = (d - sm.D[r][c + 1]) * wgt;
                        vx = signf(v) * wgt;
                        if (inner) rg.acc[ACC_SX] += fabsf(v);
                    }
                    if (iy + 1 < h) {
                        const float wgt = edge_weight3(sm.X[0][r][c], sm.X[1][r][c], sm.X[2][r][c],
                                                       sm.X[0][r + 1][c], sm.X[1][r + 1][c], sm.X[2][r + 1][c]);
                        const float v = (d - sm.D[r + 1][c]) * wgt;
                        vy = signf(v) * wgt;
                        if (inner) rg.acc[ACC_SY] += fabsf(v);
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
                    const int row = idx / C::RW, c = idx - row * C::RW;
                    const int r = row + 2, iy = ry0 + r, ix = rx0 + c;
                    if (c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;
                    const int pr = r - 1;
                    rg.gd[i] += sc.gsx * (SXw[pr][c] - SXw[pr][c - 1]) + sc.gsy * (SYw[pr][c] - SYw[pr - 1][c]);
                }
            SNDE_THREADS_END
        }
    }

    // ---- write the depth gradient, publish the partial sums ----
    if (GRAD) {
        SNDE_THREADS_BEGIN
#pragma unroll
            for (int i = 0; i < C::PPT; ++i) {
                const int idx = tid + i * C::NT;
                const int row = idx / C::RW, c = idx - row * C::RW;
                const int iy = ry0 + row + 2, ix = rx0 + c;
                if (c < 2 || c >= C::RW - 2 || iy >= h || ix >= w) continue;
                sc.gdepth.p[depth_n * sc.gdepth.sn + (long long)iy * sc.gdepth.sh + (long long)ix * sc.gdepth.sw] = rg.gd[i];
            }
        SNDE_THREADS_END
        if (GRADK) block_reduce_add<ACC_DK, 9>(sm, regs, sc.gKinv + 9 * b);
    }
    // acc[ERR] -> loss[0|1]; acc[SX], acc[SY] -> loss[2], loss[3]
    block_reduce_add<ACC_ERR, 1>(sm, regs, sc.loss + (clsF ? 0 : 1));
    block_reduce_add<ACC_SX, 2>(sm, regs, sc.loss + 2);
}

}  // namespace snde
