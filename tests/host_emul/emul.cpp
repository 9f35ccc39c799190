// Host emulation of the fused loss strip program -- TEST HARNESS ONLY (never shipped, never a
// fallback).  Compiles csrc/snde_strip.cuh with g++: the 64 lanes of a warp pair run as 64 fibers
// (ucontext) that block on counting barriers at every warp exchange / pair barrier, so the per-lane device code
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
// the emulated warp pair: 64 fibers (0..31 geometry warp, 32..63 SSIM warp), cooperative round-robin
// ---------------------------------------------------------------------------------------------
namespace {
struct SyncObj { int count = 0; int gen = 0; };
struct HostPair {
    static constexpr int STACK = 1 << 18, NF = 64;
    ucontext_t main_ctx, ctx[NF];
    std::vector<char> stacks;
    float slots[2][NF][64];
    int parity[NF];
    bool done[NF];
    int current = -1;
    SyncObj warp_sync[2], pair_sync;
    // the task being run
    const DevLoss *p = nullptr; int task = 0; int variant = 0; PairSmem *sm = nullptr;
} g_pair;

void yield_to_scheduler() { const int f = g_pair.current; swapcontext(&g_pair.ctx[f], &g_pair.main_ctx); }

void sync_on(SyncObj &s, int n) {
    const int g = s.gen;
    if (++s.count == n) { s.count = 0; ++s.gen; return; }
    while (s.gen == g) yield_to_scheduler();
}

template <bool GRAD, bool GRADK, bool EXTRA, bool FAST, bool DISP = false, bool GIMG = false>
void run_role(int who) {
    HostPair &w = g_pair;
    if (who < 32) geo_role<GRAD, GRADK, EXTRA, FAST, DISP, GIMG>(*w.p, w.task, who, *w.sm, 1);
    else ssim_role<GRAD, EXTRA, FAST, DISP, GIMG>(*w.p, w.task, who, *w.sm, 1);
}

void fiber_main(int who) {
    HostPair &w = g_pair;
    switch (w.variant) {
        case 0: run_role<false, false, true, false>(who); break;
        case 1: run_role<false, false, false, true>(who); break;
        case 2: run_role<false, false, false, false>(who); break;
        case 3: run_role<true, true, true, false>(who); break;
        case 4: run_role<true, false, false, true>(who); break;
        case 6: run_role<true, false, false, true, true>(who); break;
        case 7: run_role<true, true, true, true>(who); break;
        case 8: run_role<true, true, true, true, false, true>(who); break;
        case 9: run_role<true, true, true, false, false, true>(who); break;
        default: run_role<true, false, false, false>(who); break;
    }
    w.done[who] = true;
}

void run_pair_task(const DevLoss &p, int task, int variant, PairSmem &sm) {
    HostPair &w = g_pair;
    if (w.stacks.empty()) w.stacks.resize((size_t)HostPair::NF * HostPair::STACK);
    w.p = &p; w.task = task; w.variant = variant; w.sm = &sm;
    w.warp_sync[0] = SyncObj(); w.warp_sync[1] = SyncObj(); w.pair_sync = SyncObj();
    for (int l = 0; l < HostPair::NF; ++l) {
        w.parity[l] = 0; w.done[l] = false;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks.data() + (size_t)l * HostPair::STACK;
        w.ctx[l].uc_stack.ss_size = HostPair::STACK;
        w.ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.ctx[l], (void (*)())fiber_main, 1, l);
    }
    // every live fiber runs until it blocks on a barrier (or finishes), then the next one
    for (bool any = true; any;) {
        any = false;
        for (int l = 0; l < HostPair::NF; ++l) {
            if (w.done[l]) continue;
            w.current = l;
            swapcontext(&w.main_ctx, &w.ctx[l]);
            any = any || !w.done[l];
        }
    }
}
}  // namespace

namespace snde {
float *host_slots(int who, int parity) { return g_pair.slots[parity][who]; }
int &host_parity(int who) { return g_pair.parity[who]; }
void host_warp_sync(int who) { sync_on(g_pair.warp_sync[who >> 5], 32); }
void host_pair_sync(int) { sync_on(g_pair.pair_sync, 64); }
}  // namespace snde

extern "C" __attribute__((visibility("default"))) size_t emul_sizeof_loss_args(void) { return sizeof(SndeLossArgs); }

// seg_rows: rows per strip segment (0 = the library's default), fine_rows / fine_permille: the tail-filling short segments;
// they let the tests exercise segment seams and mixed task tables on small images
extern "C" __attribute__((visibility("default"))) int emul_loss_fwdbwd(const SndeLossArgs *a, int seg_rows, int fine_rows, int fine_permille) {
    if (!a || !a->poses || !a->loss_out) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD) && !a->grad_poses) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD_K) && !a->grad_K) return SNDE_ERR_NULL;
    if (a->B <= 0 || a->S <= 0 || a->num_scales <= 0 || a->num_scales > SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
    const LossWorkspace lay = loss_ws_layout(a->B, a->S, a->num_scales);
    if (!a->workspace || a->workspace_bytes < lay.end || ((uintptr_t)a->workspace & 15)) return SNDE_ERR_WORKSPACE;
    DevLoss d; int64_t tasks = 0;
    int rc = build_dev_loss(a, d, &lay, &tasks, seg_rows, fine_rows, fine_permille);
    if (rc) return rc;
    const int B = a->B, S = a->S, ns = a->num_scales, N = B * S;
    unsigned char *ws = (unsigned char *)a->workspace;
    memset(ws + lay.gP, 0, lay.end - lay.gP);
    PrepParams kp; FinalParams fp;
    build_prep_final(a, kp, fp);
    for (int i = 0; i < ns * N; ++i) loss_prep_thread(i, B, S, ns, a->poses, kp, (float *)(ws + lay.P), (float *)(ws + lay.Kinv));
    const bool grad = a->flags & SNDE_LOSS_GRAD, gradk = a->flags & SNDE_LOSS_GRAD_K;
    bool extra = gradk || a->padding_mode != SNDE_PAD_ZEROS, fast = true, gimg = false;      // same variant selection as snde_loss_fwdbwd
    for (int s = 0; s < ns; ++s) {
        const SndeLossScale &in = a->scale[s];
        extra = extra || in.mask_fwd.ptr || in.mask_bwd.ptr || in.warped_fwd.ptr || in.warped_bwd.ptr || in.err_fwd.ptr || in.err_bwd.ptr;
        gimg = gimg || in.grad_tgt_img.ptr || in.grad_src_img.ptr;
        fast = fast && d.sc[s].fast;
    }
    if (gimg) extra = true;
    const bool disp = (a->flags & SNDE_LOSS_DISPARITY) != 0;
    if (disp && !(grad && fast && !extra)) extra = true;
    const int variant = !grad ? (extra ? 0 : (fast ? 1 : 2)) : (gimg ? (fast ? 8 : 9) : (extra ? (fast ? 7 : 3) : (disp ? 6 : (fast ? 4 : 5))));
    PairSmem *sm = new PairSmem;
    for (int64_t t = 0; t < tasks; ++t) {
        memset(sm, 0xff, sizeof(PairSmem));    // NaN-fill: catches reads of unwritten shared memory
        run_pair_task(d, (int)t, variant, *sm);
    }
    delete sm;
    loss_final_host(B, S, ns, a->flags, a->poses, fp, (const double *)(ws + lay.gP), (const double *)(ws + lay.gKinv),
                    (const double *)(ws + lay.loss), a->loss_out, a->grad_poses, a->grad_K);
    return variant;
}
