// Host emulation of the fused loss strip program -- TEST HARNESS ONLY (never shipped, never a
// fallback).  Compiles csrc/snde_strip.cuh with g++: the 32 lanes of a warp run as 32 fibers
// (ucontext) that are switched round-robin at every warp exchange, so the per-lane device code
// -- halo columns, rolling row sums, zero padding, task table, gradient ownership -- can be
// checked against the oracle in the GPU-less authoring container.
// All pointers in SndeLossArgs are HOST pointers here.
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <vector>

#include "snde_loss_setup.cuh"

using namespace snde;

// ---------------------------------------------------------------------------------------------
// the emulated warp
// ---------------------------------------------------------------------------------------------
namespace {
struct HostWarp {
    static constexpr int STACK = 1 << 18;
    ucontext_t main_ctx, lane_ctx[32];
    std::vector<char> stacks;
    float slots[2][32][64];
    int parity[32];
    bool done[32];
    int current = -1;
    // the task being run
    const DevLoss *p = nullptr; int task = 0; int variant = 0; StripSmem *sm = nullptr;
} g_warp;

void lane_main(int lane) {
    HostWarp &w = g_warp;
    switch (w.variant) {
        case 0: strip_task<false, false, true, false>(*w.p, w.task, lane, *w.sm); break;
        case 1: strip_task<false, false, false, true>(*w.p, w.task, lane, *w.sm); break;
        case 2: strip_task<false, false, false, false>(*w.p, w.task, lane, *w.sm); break;
        case 3: strip_task<true, true, true, false>(*w.p, w.task, lane, *w.sm); break;
        case 4: strip_task<true, false, false, true>(*w.p, w.task, lane, *w.sm); break;
        default: strip_task<true, false, false, false>(*w.p, w.task, lane, *w.sm); break;
    }
    w.done[lane] = true;
}

void run_warp_task(const DevLoss &p, int task, int variant, StripSmem &sm) {
    HostWarp &w = g_warp;
    if (w.stacks.empty()) w.stacks.resize((size_t)32 * HostWarp::STACK);
    w.p = &p; w.task = task; w.variant = variant; w.sm = &sm;
    for (int l = 0; l < 32; ++l) {
        w.parity[l] = 0; w.done[l] = false;
        getcontext(&w.lane_ctx[l]);
        w.lane_ctx[l].uc_stack.ss_sp = w.stacks.data() + (size_t)l * HostWarp::STACK;
        w.lane_ctx[l].uc_stack.ss_size = HostWarp::STACK;
        w.lane_ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.lane_ctx[l], (void (*)())lane_main, 1, l);
    }
    // lock-step at exchange granularity: every live lane runs up to its next exchange, then the next lane
    for (bool any = true; any;) {
        any = false;
        for (int l = 0; l < 32; ++l) {
            if (w.done[l]) continue;
            w.current = l;
            swapcontext(&w.main_ctx, &w.lane_ctx[l]);
            any = any || !w.done[l];
        }
    }
}
}  // namespace

namespace snde {
float *host_warp_slots(int lane, int parity) { return g_warp.slots[parity][lane]; }
int &host_warp_parity(int lane) { return g_warp.parity[lane]; }
void host_warp_yield() { const int l = g_warp.current; swapcontext(&g_warp.lane_ctx[l], &g_warp.main_ctx); }
}  // namespace snde

extern "C" __attribute__((visibility("default"))) size_t emul_sizeof_loss_args(void) { return sizeof(SndeLossArgs); }

// seg_rows: rows per strip segment (0 = the library's default); lets the tests exercise segment seams on small images
extern "C" __attribute__((visibility("default"))) int emul_loss_fwdbwd(const SndeLossArgs *a, int seg_rows) {
    if (!a || !a->poses || !a->loss_out) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD) && !a->grad_poses) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD_K) && !a->grad_K) return SNDE_ERR_NULL;
    if (a->B <= 0 || a->S <= 0 || a->num_scales <= 0 || a->num_scales > SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
    const LossWorkspace lay = loss_ws_layout(a->B, a->S, a->num_scales);
    if (!a->workspace || a->workspace_bytes < lay.end || ((uintptr_t)a->workspace & 15)) return SNDE_ERR_WORKSPACE;
    DevLoss d; int64_t tasks = 0;
    int rc = build_dev_loss(a, d, &lay, &tasks, seg_rows);
    if (rc) return rc;
    const int B = a->B, S = a->S, ns = a->num_scales, N = B * S;
    unsigned char *ws = (unsigned char *)a->workspace;
    memset(ws + lay.gP, 0, lay.end - lay.gP);
    PrepParams kp; FinalParams fp;
    build_prep_final(a, kp, fp);
    for (int i = 0; i < ns * N; ++i) loss_prep_thread(i, B, S, ns, a->poses, kp, (float *)(ws + lay.P), (float *)(ws + lay.Kinv));
    const bool grad = a->flags & SNDE_LOSS_GRAD, gradk = a->flags & SNDE_LOSS_GRAD_K;
    bool extra = gradk || a->padding_mode != SNDE_PAD_ZEROS, fast = true;      // same variant selection as snde_loss_fwdbwd
    for (int s = 0; s < ns; ++s) {
        const SndeLossScale &in = a->scale[s];
        extra = extra || in.mask_fwd.ptr || in.mask_bwd.ptr || in.warped_fwd.ptr || in.warped_bwd.ptr || in.err_fwd.ptr || in.err_bwd.ptr;
        fast = fast && d.sc[s].fast;
    }
    const int variant = !grad ? (extra ? 0 : (fast ? 1 : 2)) : (extra ? 3 : (fast ? 4 : 5));
    StripSmem *sm = new StripSmem;
    for (int64_t t = 0; t < tasks; ++t) {
        memset(sm, 0xff, sizeof(StripSmem));   // NaN-fill: catches reads of unwritten shared memory
        run_warp_task(d, (int)t, variant, *sm);
    }
    delete sm;
    loss_final_host(B, S, ns, a->flags, a->poses, fp, (const double *)(ws + lay.gP), (const double *)(ws + lay.gKinv),
                    (const double *)(ws + lay.loss), a->loss_out, a->grad_poses, a->grad_K);
    return variant;
}
