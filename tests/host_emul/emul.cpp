// Host emulation of the fused loss tile program -- TEST HARNESS ONLY (never shipped, never a
// fallback).  Compiles csrc/snde_tile.cuh with g++: every barrier-separated phase becomes a loop
// over the CTA's threads, so the tile logic (halo indexing, zero padding, work table, gradient
// ownership) can be checked against the oracle in the GPU-less authoring container.
// All pointers in SndeLossArgs are HOST pointers here.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "snde_loss_setup.cuh"

using namespace snde;

extern "C" __attribute__((visibility("default"))) size_t emul_sizeof_loss_args(void) { return sizeof(SndeLossArgs); }

extern "C" __attribute__((visibility("default"))) int emul_loss_fwdbwd(const SndeLossArgs *a) {
    if (!a || !a->poses || !a->loss_out) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD) && !a->grad_poses) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD_K) && !a->grad_K) return SNDE_ERR_NULL;
    if (a->B <= 0 || a->S <= 0 || a->num_scales <= 0 || a->num_scales > SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
    const LossWorkspace lay = loss_ws_layout(a->B, a->S, a->num_scales);
    if (!a->workspace || a->workspace_bytes < lay.end || ((uintptr_t)a->workspace & 15)) return SNDE_ERR_WORKSPACE;
    DevLoss d; int64_t tiles = 0;
    int rc = build_dev_loss(a, d, &lay, &tiles);
    if (rc) return rc;
    const int B = a->B, S = a->S, ns = a->num_scales, N = B * S;
    unsigned char *ws = (unsigned char *)a->workspace;
    memset(ws + lay.gP, 0, lay.end - lay.gP);
    PrepParams kp; FinalParams fp;
    build_prep_final(a, kp, fp);
    for (int i = 0; i < ns * N; ++i) loss_prep_thread(i, B, S, ns, a->poses, kp, (float *)(ws + lay.P), (float *)(ws + lay.Kinv));
    TileSmem *sm = new TileSmem;
    std::vector<Regs> regs(TileCfg::NT);
    const bool grad = a->flags & SNDE_LOSS_GRAD, gradk = a->flags & SNDE_LOSS_GRAD_K;
    bool extra = gradk || a->padding_mode != SNDE_PAD_ZEROS;      // same variant selection as snde_loss_fwdbwd
    for (int s = 0; s < ns; ++s) {
        const SndeLossScale &in = a->scale[s];
        extra = extra || in.mask_fwd.ptr || in.mask_bwd.ptr || in.warped_fwd.ptr || in.warped_bwd.ptr || in.err_fwd.ptr || in.err_bwd.ptr;
    }
    for (int64_t blk = 0; blk < tiles; ++blk) {
        memset(sm, 0xff, sizeof(TileSmem));   // NaN-fill: catches reads of unwritten shared memory
        if (!grad) { if (extra) loss_tile<false, false, true>(d, (int)blk, *sm, regs.data()); else loss_tile<false, false, false>(d, (int)blk, *sm, regs.data()); }
        else if (extra) loss_tile<true, true, true>(d, (int)blk, *sm, regs.data());
        else loss_tile<true, false, false>(d, (int)blk, *sm, regs.data());
    }
    delete sm;
    loss_final_host(B, S, ns, a->flags, a->poses, fp, (const double *)(ws + lay.gP), (const double *)(ws + lay.gKinv),
                    (const double *)(ws + lay.loss), a->loss_out, a->grad_poses, a->grad_K);
    return SNDE_OK;
}
