// snde_loss_setup.cuh -- host-side setup of the fused loss (argument validation, workspace layout,
// task table) and the per-thread bodies of its two tiny companion kernels (matrix prep, finalize).
// Shared by snde_kernels.cu (the product) and the host-emulation test harness.
#pragma once
#include <string.h>

#include "snde.h"
#include "snde_strip.cuh"

namespace snde {

// in-image strides are 32-bit on the device (validated in build_dev_loss); the batch stride stays 64-bit
static inline DevImage dev(const SndeImage &v) { return DevImage{v.ptr, v.sn, (int)v.sc, (int)v.sh, (int)v.sw, 0}; }
static inline DevImageOut dev(const SndeImageOut &v) { return DevImageOut{v.ptr, v.sn, (int)v.sc, (int)v.sh, (int)v.sw, 0}; }
static inline DevMap dev(const SndeMap &v) { return DevMap{v.ptr, v.sn, (int)v.sh, (int)v.sw}; }
static inline DevMapOut dev(const SndeMapOut &v) { return DevMapOut{v.ptr, v.sn, (int)v.sh, (int)v.sw}; }
static inline bool fits32(int64_t sc, int64_t sh, int64_t sw, int h, int w) {
    const int64_t lim = (int64_t)1 << 30;
    auto ab = [](int64_t v) { return v < 0 ? -v : v; };
    return 2 * ab(sc) + ab(sh) * h + ab(sw) * w < lim;
}

// =============================================================================================
// Fused loss
// =============================================================================================
struct LossWorkspace {   // offsets in bytes; [gP, end) is zeroed at the start of every call
    size_t P, Kinv, gP, gKinv, loss, counter, end;
};
static inline LossWorkspace loss_ws_layout(int B, int S, int ns) {
    LossWorkspace o;
    auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t N = (size_t)B * S;
    o.P = 0;
    o.Kinv = up(o.P + sizeof(float) * ns * N * 24);
    o.gP = up(o.Kinv + sizeof(float) * ns * B * 9);
    o.gKinv = up(o.gP + sizeof(double) * ns * N * 24);
    o.loss = up(o.gKinv + sizeof(double) * ns * B * 9);
    o.counter = up(o.loss + sizeof(double) * ns * 4);
    o.end = up(o.counter + sizeof(int) * 4);
    return o;
}


// one thread per (scale, n): P_fwd = K_s @ T, P_bwd = K_s @ inv(T); threads with k == 0 also write Kinv
struct PrepParams { const float *K[SNDE_MAX_SCALES]; };

SNDE_HD void loss_prep_thread(int i, int B, int S, int ns, const float *poses, PrepParams kp, float *P, float *Kinv) {
    const int N = B * S;
    if (i >= ns * N) return;
    const int s = i / N, n = i - s * N;
    const int k = n / B, b = n - k * B;
    float vec[6], K[9], T[12], Ti[12], Pm[12];
    for (int j = 0; j < 6; ++j) vec[j] = poses[(b * S + k) * 6 + j];
    for (int j = 0; j < 9; ++j) K[j] = kp.K[s][b * 9 + j];
    pose_to_T(vec, T, Ti);
    float *dst = P + ((size_t)s * N + n) * 24;
    k_times_T(K, T, Pm);  for (int j = 0; j < 12; ++j) dst[j] = Pm[j];
    k_times_T(K, Ti, Pm); for (int j = 0; j < 12; ++j) dst[12 + j] = Pm[j];
    if (k == 0) {
        float Ki[9];
        if (!inverse3(K, Ki)) for (int j = 0; j < 9; ++j) Ki[j] = nanf("");
        for (int j = 0; j < 9; ++j) Kinv[((size_t)s * B + b) * 9 + j] = Ki[j];
    }
}

struct FinalParams {
    const float *K[SNDE_MAX_SCALES];
    float photo_weight[SNDE_MAX_SCALES], smooth_weight[SNDE_MAX_SCALES];
    double inv_count[SNDE_MAX_SCALES], inv_nx[SNDE_MAX_SCALES], inv_ny[SNDE_MAX_SCALES];
};

// Finalize, split so that the double-precision chains run in parallel:
//   loss_final_pair(b, k, s): backward of the matrix prep for one (batch element, source, scale) -> gv[6], gK[9]
//   loss_final_scalars: the loss scalars from the per-scale sums
SNDE_HD void loss_final_pair(int b, int k, int s, int B, int S, int flags, const float *poses, const FinalParams &fp,
                             const double *gP, const double *gKinv, double *gv /*6*/, double *gK /*9*/) {
    const int N = B * S, n = k * B + b;
    const bool wantK = (flags & FLAG_GRAD_K) != 0;
    float vec[6], K[9];
    for (int j = 0; j < 6; ++j) { vec[j] = poses[(b * S + k) * 6 + j]; gv[j] = 0.0; }
    for (int j = 0; j < 9; ++j) { K[j] = fp.K[s][b * 9 + j]; gK[j] = 0.0; }
    const double *g = gP + ((size_t)s * N + n) * 24;
    pose_prep_bwd(vec, K, g, g + 12, (wantK && k == 0) ? gKinv + ((size_t)s * B + b) * 9 : nullptr, gv, wantK ? gK : nullptr);
}

SNDE_HD void loss_final_scalars(int ns, const FinalParams &fp, const double *loss, float *loss_out) {
    double rigid = 0.0, smooth = 0.0;
    for (int s = 0; s < ns; ++s) {
        const double fm = loss[s * 4 + 0] * fp.inv_count[s], bm = loss[s * 4 + 1] * fp.inv_count[s];
        const double sm = loss[s * 4 + 2] * fp.inv_nx[s] + loss[s * 4 + 3] * fp.inv_ny[s];
        rigid += (double)fp.photo_weight[s] * (fm + bm);
        smooth += (double)fp.smooth_weight[s] * sm;
        loss_out[3 + 3 * s + 0] = (float)fm; loss_out[3 + 3 * s + 1] = (float)bm; loss_out[3 + 3 * s + 2] = (float)sm;
    }
    loss_out[0] = (float)(rigid + smooth); loss_out[1] = (float)rigid; loss_out[2] = (float)smooth;
}

// serial composition (host emulation; the GPU kernel spreads loss_final_pair over threads)
static inline void loss_final_host(int B, int S, int ns, int flags, const float *poses, const FinalParams &fp, const double *gP,
                                   const double *gKinv, const double *loss, float *loss_out, float *grad_poses, float *grad_K) {
    loss_final_scalars(ns, fp, loss, loss_out);
    if (!(flags & FLAG_GRAD)) return;
    for (int b = 0; b < B; ++b) {
        double gKs[SNDE_MAX_SCALES][9] = {};
        for (int k = 0; k < S; ++k) {
            double gvs[6] = {0, 0, 0, 0, 0, 0};
            for (int s = 0; s < ns; ++s) {
                double gv[6], gK[9];
                loss_final_pair(b, k, s, B, S, flags, poses, fp, gP, gKinv, gv, gK);
                for (int j = 0; j < 6; ++j) gvs[j] += gv[j];
                for (int j = 0; j < 9; ++j) gKs[s][j] += gK[j];
            }
            for (int j = 0; j < 6; ++j) grad_poses[(b * S + k) * 6 + j] = (float)gvs[j];
        }
        if (flags & FLAG_GRAD_K) for (int s = 0; s < ns; ++s) for (int j = 0; j < 9; ++j) grad_K[((size_t)b * ns + s) * 9 + j] = (float)gKs[s][j];
    }
}

static inline int build_dev_loss(const SndeLossArgs *a, DevLoss &d, const LossWorkspace *lay, int64_t *num_tiles, int seg_rows = 0,
                                 int fine_rows = 0, int fine_permille = 0) {
    if (!a) return SNDE_ERR_NULL;
    if (a->B <= 0 || a->S <= 0 || a->num_scales <= 0 || a->num_scales > SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
    if (a->S > StripCfg::MAX_S) return SNDE_ERR_SHAPE;       // SNDE_MAX_SOURCES
    if (a->padding_mode != SNDE_PAD_ZEROS && a->padding_mode != SNDE_PAD_BORDER) return SNDE_ERR_ARG;
    if (a->flags & ~31) return SNDE_ERR_ARG;
    if ((a->flags & SNDE_LOSS_GRAD_K) && !(a->flags & SNDE_LOSS_GRAD)) return SNDE_ERR_ARG;
    if (!(a->alpha >= 0.f && a->alpha <= 1.f)) return SNDE_ERR_ARG;
    memset(&d, 0, sizeof(d));
    const int B = a->B, S = a->S, ns = a->num_scales;
    const long long N = (long long)B * S, M = (long long)B * (1 + S);
    d.B = B; d.S = S; d.num_scales = ns; d.flags = a->flags; d.padding = a->padding_mode;
    d.alpha = a->alpha; d.oma = (float)(1.0 - (double)a->alpha);
    unsigned char *ws = (unsigned char *)a->workspace;
    for (int s = 0; s < ns; ++s) {
        const SndeLossScale &in = a->scale[s];
        DevScale &o = d.sc[s];
        if (in.h < 2 || in.w < 2) return SNDE_ERR_SHAPE;
        if (M * in.h * in.w >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
        if (!in.tgt.ptr || !in.src.ptr || !in.depth.ptr || !in.K || !in.xtab || !in.ytab) return SNDE_ERR_NULL;
        if ((a->flags & SNDE_LOSS_GRAD) && !in.grad_depth.ptr) return SNDE_ERR_NULL;
        if (!fits32(in.tgt.sc, in.tgt.sh, in.tgt.sw, in.h, in.w) || !fits32(in.src.sc, in.src.sh, in.src.sw, in.h, in.w) ||
            !fits32(0, in.depth.sh, in.depth.sw, in.h, in.w) || !fits32(0, in.grad_depth.sh, in.grad_depth.sw, in.h, in.w) ||
            !fits32(0, in.mask_fwd.sh, in.mask_fwd.sw, in.h, in.w) || !fits32(0, in.mask_bwd.sh, in.mask_bwd.sw, in.h, in.w) ||
            !fits32(in.warped_fwd.sc, in.warped_fwd.sh, in.warped_fwd.sw, in.h, in.w) || !fits32(in.warped_bwd.sc, in.warped_bwd.sh, in.warped_bwd.sw, in.h, in.w) ||
            !fits32(in.err_fwd.sc, in.err_fwd.sh, in.err_fwd.sw, in.h, in.w) || !fits32(in.err_bwd.sc, in.err_bwd.sh, in.err_bwd.sw, in.h, in.w) ||
            !fits32(in.grad_tgt_img.sc, in.grad_tgt_img.sh, in.grad_tgt_img.sw, in.h, in.w) ||
            !fits32(in.grad_src_img.sc, in.grad_src_img.sh, in.grad_src_img.sw, in.h, in.w))
            return SNDE_ERR_TOO_LARGE;
        if ((in.grad_tgt_img.ptr || in.grad_src_img.ptr) && !(a->flags & SNDE_LOSS_GRAD)) return SNDE_ERR_ARG;
        o.h = in.h; o.w = in.w;
        o.tiles_x = (in.w + StripCfg::TW - 1) / StripCfg::TW;
        // fast path: planar tensors, unit x-stride, even row / plane / batch strides and 8-byte aligned bases, even width:
        // the lane's pixel pair is then one aligned 64-bit access everywhere
        {
            auto even = [](int64_t v) { return (v & 1) == 0; };
            auto al8 = [](const void *q) { return ((uintptr_t)q & 7) == 0; };
            bool f = (in.w & 1) == 0;
            f = f && in.tgt.sw == 1 && even(in.tgt.sh) && even(in.tgt.sc) && even(in.tgt.sn) && al8(in.tgt.ptr);
            f = f && in.src.sw == 1 && even(in.src.sh) && even(in.src.sc) && even(in.src.sn) && al8(in.src.ptr);
            f = f && in.depth.sw == 1 && even(in.depth.sh) && even(in.depth.sn) && al8(in.depth.ptr);
            if (a->flags & SNDE_LOSS_GRAD) f = f && in.grad_depth.sw == 1 && even(in.grad_depth.sh) && even(in.grad_depth.sn) && al8(in.grad_depth.ptr);
            o.fast = f ? 1 : 0;
        }
        o.tgt = dev(in.tgt); o.src = dev(in.src); o.depth = dev(in.depth); o.gdepth = dev(in.grad_depth);
        o.xtab = in.xtab; o.ytab = in.ytab;
        if (lay) {
            o.P = (const float *)(ws + lay->P) + (size_t)s * N * 24;
            o.Kinv = (const float *)(ws + lay->Kinv) + (size_t)s * B * 9;
            o.gP = (double *)(ws + lay->gP) + (size_t)s * N * 24;
            o.gKinv = (double *)(ws + lay->gKinv) + (size_t)s * B * 9;
            o.loss = (double *)(ws + lay->loss) + (size_t)s * 4;
        }
        o.gphoto = (float)((double)in.photo_weight / ((double)N * in.h * in.w * 3.0));
        o.gsx = (float)((double)in.smooth_weight / ((double)M * in.h * (in.w - 1)));
        o.gsy = (float)((double)in.smooth_weight / ((double)M * (in.h - 1) * in.w));
        o.mask_fwd = dev(in.mask_fwd); o.mask_bwd = dev(in.mask_bwd);
        o.warped_fwd = dev(in.warped_fwd); o.warped_bwd = dev(in.warped_bwd);
        o.err_fwd = dev(in.err_fwd); o.err_bwd = dev(in.err_bwd);
        o.gimg_tgt = dev(in.grad_tgt_img); o.gimg_src = dev(in.grad_src_img);
    }
    if (lay) d.counter = (int *)(ws + lay->counter);
    // Task table.  Most owners are cut into long segments (seg_rows rows: few halo rows recomputed); the last fine_permille
    // of the owners of every (class, scale) into short ones (fine_rows; tuning builds only).  A task costs passes x (rows + 6)
    // row steps (class F: S passes); the blocks are handed out in order of decreasing task cost -- longest processing time
    // first, the list-scheduling rule -- and the short last segment of a height that is not a multiple of the segment is its
    // own block, so that it queues with the tasks of its cost (NYU and 376-row shapes: 2 % off the simulated makespan).
    struct Block { int scale, cls, first, cnt_owners, seg, ty, y_base, cost; };
    Block blk[8 * SNDE_MAX_SCALES];
    int nb = 0;
    const int coarse = seg_rows > 0 ? seg_rows : StripCfg::SEG;
    for (int grp = 0; grp < 2; ++grp)
        for (int cls = 0; cls < 2; ++cls)
            for (int s = 0; s < ns; ++s) {
                const int owners = cls == 0 ? B : (int)N, h = d.sc[s].h;
                int n_fine = (fine_rows > 0 && fine_rows < coarse && h > fine_rows) ? (int)(((long long)owners * fine_permille + 999) / 1000) : 0;
                if (n_fine > owners) n_fine = owners;
                const int first = grp == 0 ? 0 : owners - n_fine, cnt_owners = grp == 0 ? owners - n_fine : n_fine;
                if (cnt_owners <= 0) continue;
                int seg = grp == 0 ? coarse : fine_rows;
                if (seg > h) seg = h;
                const int passes = cls == 0 ? S : 1;
                const int full = h / seg, rem = h - full * seg;
                if (nb + 2 > 8 * SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
                if (full > 0) blk[nb++] = Block{s, cls, first, cnt_owners, seg, full, 0, passes * (seg + 6)};
                if (rem > 0) blk[nb++] = Block{s, cls, first, cnt_owners, rem, 1, full * seg, passes * (rem + 6)};
            }
    for (int i = 1; i < nb; ++i) {                     // stable insertion sort by decreasing cost
        const Block key = blk[i];
        int j = i - 1;
        while (j >= 0 && blk[j].cost < key.cost) { blk[j + 1] = blk[j]; --j; }
        blk[j + 1] = key;
    }
    long long total = 0;
    int ne = 0;
    for (int i = 0; i < nb; ++i) {
        const long long cnt = (long long)d.sc[blk[i].scale].tiles_x * blk[i].ty * blk[i].cnt_owners;
        if (total + cnt >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
        d.entry[ne++] = DevEntry{(int)total, (int)cnt, blk[i].scale, blk[i].cls, blk[i].first, blk[i].seg, blk[i].ty, blk[i].y_base};
        total += cnt;
    }
    d.num_entries = ne;
    d.total_tasks = (int)total;
    if (num_tiles) *num_tiles = total;
    return SNDE_OK;
}


// fill the per-scale parameter blocks of the prep / finalize kernels
static inline void build_prep_final(const SndeLossArgs *a, PrepParams &kp, FinalParams &fp) {
    memset(&kp, 0, sizeof kp); memset(&fp, 0, sizeof fp);
    const int B = a->B, S = a->S, N = B * S;
    for (int s = 0; s < a->num_scales; ++s) {
        const SndeLossScale &in = a->scale[s];
        kp.K[s] = in.K; fp.K[s] = in.K;
        fp.photo_weight[s] = in.photo_weight; fp.smooth_weight[s] = in.smooth_weight;
        fp.inv_count[s] = 1.0 / ((double)N * in.h * in.w * 3.0);
        fp.inv_nx[s] = 1.0 / ((double)B * (1 + S) * in.h * (in.w - 1));
        fp.inv_ny[s] = 1.0 / ((double)B * (1 + S) * (in.h - 1) * in.w);
    }
}

}  // namespace snde
