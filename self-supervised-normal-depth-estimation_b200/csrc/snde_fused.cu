// snde_fused.cu -- the fused loss (matrix prep, strip kernel, finalize) and the small host helpers of the C ABI
// declared in include/snde.h.  Built with -fmad=true: the strip program states its exact-order arithmetic explicitly
// (fmaf / SNDE_X* ops in snde_math.cuh), everything else may be contracted.  See build.py.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "snde.h"
#include "snde_strip.cuh"
#include "snde_loss_setup.cuh"

using namespace snde;

#ifndef SNDE_STRIP_CTAS_PER_SM
#define SNDE_STRIP_CTAS_PER_SM 3
#endif

// (SNDE_DEBUG in the environment: say which CUDA error it was)
static inline bool snde_cuda_failed(cudaError_t e, const char *what, int line) {
    if (e == cudaSuccess) return false;
    if (getenv("SNDE_DEBUG")) fprintf(stderr, "snde: %s: %s (snde_fused.cu:%d)\n", what, cudaGetErrorString(e), line);
    return true;
}
#define SNDE_CHECK_LAUNCH() do { if (snde_cuda_failed(cudaGetLastError(), "launch", __LINE__)) return SNDE_ERR_CUDA; } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// =============================================================================================
// small helpers
// =============================================================================================
extern "C" int snde_version(void) { return SNDE_VERSION; }

extern "C" const char *snde_strerror(int status) {
    switch (status) {
        case SNDE_OK: return "ok";
        case SNDE_ERR_NULL: return "snde: required pointer is NULL";
        case SNDE_ERR_SHAPE: return "snde: bad shape (non-positive dimension, C != 3, or h,w < 2 for smoothness)";
        case SNDE_ERR_TOO_LARGE: return "snde: tensor too large (N*h*w must stay below 2^31)";
        case SNDE_ERR_WORKSPACE: return "snde: workspace missing, too small or not 16-byte aligned";
        case SNDE_ERR_ARG: return "snde: bad argument (enum, flag or scalar out of range)";
        case SNDE_ERR_CUDA: return "snde: CUDA launch failed";
        default: return "snde: unknown status";
    }
}

// reference: utils/utils_edited.py:307-317 -- (linspace(-1,1,n)+1)*0.5*(n-1); torch.linspace rounds
// start + step*i (first half) / end - step*(n-1-i) (second half) once, i.e. as an FMA.
extern "C" int snde_fill_grid_table(int n, float *out) {
    if (!out) return SNDE_ERR_NULL;
    if (n <= 0) return SNDE_ERR_SHAPE;
    if (n == 1) { out[0] = 0.f; return SNDE_OK; }
    const float step = 2.f / (float)(n - 1);
    const int half = n / 2;
    for (int i = 0; i < n; ++i) {
        volatile float lin = (i < half) ? fmaf(step, (float)i, -1.f) : fmaf(-step, (float)(n - i - 1), 1.f);
        volatile float t = lin + 1.f;
        t = t * 0.5f;
        out[i] = t * (float)(n - 1);
    }
    return SNDE_OK;
}


// Programmatic dependent launch (sm_90+): the three kernels of a call are launched with the programmatic-stream-serialization
// attribute, so the next one is scheduled while its predecessor drains and its launch latency disappears from the step; each
// waits for the predecessor's memory with griddepcontrol.wait before touching anything the predecessor writes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// (also zeroes the accumulators and the task counter of the call, [zero, zero + zero_words) in 16-byte words: one graph node
// less than a separate memset)
__global__ void __launch_bounds__(128) loss_prep_kernel(int B, int S, int ns, const float *__restrict__ poses, PrepParams kp,
                                                        float *__restrict__ P, float *__restrict__ Kinv, uint4 *__restrict__ zero,
                                                        int zero_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();                  // the strip kernel may be scheduled (it waits for this grid's memory itself)
    pdl_wait();                               // whatever ran before on the stream may still be reading the workspace
    for (int j = i; j < zero_words; j += gridDim.x * blockDim.x) zero[j] = make_uint4(0u, 0u, 0u, 0u);
    loss_prep_thread(i, B, S, ns, poses, kp, P, Kinv);
}

// block b < B: one thread per (source k, scale s) runs the double-precision backward of the matrix prep, then the
// block reduces over scales (pose gradient) and over sources (intrinsics gradient) in a fixed order.  Block B: scalars.
__global__ void __launch_bounds__(64) loss_final_kernel(int B, int S, int ns, int flags, const float *__restrict__ poses,
                                                        FinalParams fp, const double *__restrict__ gP,
                                                        const double *__restrict__ gKinv, const double *__restrict__ loss,
                                                        float *__restrict__ loss_out, float *__restrict__ grad_poses,
                                                        float *__restrict__ grad_K) {
    __shared__ double sv[64][6];
    __shared__ double sk[64][9];
    const int b = blockIdx.x, t = threadIdx.x;
    pdl_wait();                               // the strip kernel's sums
    if (b == B) { if (t == 0) loss_final_scalars(ns, fp, loss, loss_out); return; }
    if (!(flags & FLAG_GRAD)) return;
    const bool wantK = (flags & FLAG_GRAD_K) != 0;
    for (int base = 0; base < S * ns; base += 64) {          // S*ns <= 64 in practice: one pass
        const int i = base + t;
        const int k = i / ns, s = i - k * ns;
        if (i < S * ns) loss_final_pair(b, k, s, B, S, flags, poses, fp, gP, gKinv, sv[t], sk[t]);
        __syncthreads();
        const int cnt = min(64, S * ns - base);
        // pose gradient: sum over the scales of source k (threads base+k*ns .. base+k*ns+ns-1); requires ns | 64-chunk alignment
        if (t < cnt && (base + t) % ns == 0) {
            const int kk = (base + t) / ns;
            for (int j = 0; j < 6; ++j) {
                double a = 0.0;
                for (int q = 0; q < ns && t + q < cnt; ++q) a += sv[t + q][j];
                grad_poses[(b * S + kk) * 6 + j] = (float)a;
            }
        }
        if (wantK && t < ns && base == 0) {
            for (int j = 0; j < 9; ++j) {
                double a = 0.0;
                for (int kk = 0; kk * ns + t < cnt; ++kk) a += sk[kk * ns + t][j];
                grad_K[((size_t)b * ns + t) * 9 + j] = (float)a;
            }
        }
        __syncthreads();
    }
}

// Persistent grid: every warp PAIR pulls strip tasks from a counter until none is left.  The pairs of a CTA share
// nothing, so there is no CTA-wide barrier in the kernel; the two warps of a pair meet at a 64-thread named barrier.
// Warps w and w + PAIRS form a pair; which of the two is the geometry warp alternates with the CTA index, so every
// scheduler (warp index mod 4) hosts both kinds of warp.
template <bool GRAD, bool GRADK, bool EXTRA, bool FAST, bool DISP = false, bool GIMG = false>
// (register budget: each of the four schedulers of an SM has its own 16 K registers, so 3 warps per scheduler may use 168
// registers and 4 warps 128 -- nothing in between: 14 warps of 144 registers do not launch)
__global__ void __launch_bounds__(StripCfg::NT, SNDE_STRIP_CTAS_PER_SM) loss_strip_kernel(const __grid_constant__ DevLoss p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
    const int pair = warp % StripCfg::PAIRS;
    const int role = ((warp / StripCfg::PAIRS) + (int)blockIdx.x) & 1;      // 0 = geometry warp, 1 = SSIM warp
    PairSmem &sm = reinterpret_cast<PairSmem *>(smem_raw)[pair];
    const int bar = 1 + pair;                                               // barrier 0 is __syncthreads'
    pdl_launch_dependents();                  // the finalize kernel may be scheduled as CTAs of this grid retire
    pdl_wait();                               // the prep kernel's matrices, zeroed accumulators and task counter
    for (;;) {
        if (role == 0 && lane == 0) sm.task = atomicAdd(p.counter, 1);
        pair_barrier(lane, bar);
        // (rewritten only after both warps went through the barriers of a whole task; the shuffle tells the compiler the
        // value -- and every row index derived from it -- is warp-uniform: uniform branches, no reconvergence stack)
        const int task = __shfl_sync(0xffffffffu, sm.task, 0);
        if (task >= p.total_tasks) break;
#if defined(SNDE_ONLY_ROLE)            // register-budget probe of one role (never shipped)
        if (role == SNDE_ONLY_ROLE) { if (SNDE_ONLY_ROLE == 0) geo_role<GRAD, GRADK, EXTRA, FAST, DISP, GIMG>(p, task, lane, sm, bar); else ssim_role<GRAD, EXTRA, FAST, DISP, GIMG>(p, task, lane, sm, bar); }
#else
        if (role == 0) geo_role<GRAD, GRADK, EXTRA, FAST, DISP, GIMG>(p, task, lane, sm, bar);
        else ssim_role<GRAD, EXTRA, FAST, DISP, GIMG>(p, task, lane, sm, bar);
#endif
    }
}

extern "C" size_t snde_sizeof_loss_args(void) { return sizeof(SndeLossArgs); }

extern "C" size_t snde_loss_workspace_bytes(int B, int S, int num_scales) {
    if (B <= 0 || S <= 0 || num_scales <= 0) return 0;
    return loss_ws_layout(B, S, num_scales).end;
}

// ---- per-process caches, written once each under std::call_once / atomics (the entry points may be called from any
//      number of host threads; nothing else in the library is shared between calls) ----
#include <atomic>
#include <mutex>
namespace {
// Tuning overrides of the task decomposition exist only in builds made with -DSNDE_TUNING (scripts/build_variants.sh);
// the shipped library ignores the environment.
struct SegOverrides { int seg_rows = 0, fine_rows = 0, fine_permille = 0; };
static const SegOverrides &seg_overrides() {
    static SegOverrides o;
#ifdef SNDE_TUNING
    static std::once_flag once;
    std::call_once(once, [] {
        auto env_int = [](const char *name) { const char *e = getenv(name); return e ? atoi(e) : 0; };
        const int s = env_int("SNDE_SEG_ROWS");
        o.seg_rows = (s == 16 || s == 32 || s == 64 || s == 128) ? s : 0;          // anything else is ignored
        const int f = env_int("SNDE_FINE_ROWS");
        o.fine_rows = (f == 8 || f == 16 || f == 32) ? f : 0;
        const int pm = env_int("SNDE_FINE_PERMILLE");
        o.fine_permille = pm > 0 && pm <= 1000 ? pm : 0;
    });
#endif
    return o;
}
static std::atomic<int> g_sm_count[64];
static int sm_count_of(int dev) {            // 0 on failure
    int n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 0;
    g_sm_count[dev].store(n, std::memory_order_relaxed);
    return n;
}
static std::atomic<bool> g_attr_done[12][64];
}  // namespace

// Rows per strip segment, shared by snde_loss_fwdbwd and snde_loss_num_tiles.  Longer segments recompute fewer halo rows (a
// segment of s rows costs s + 6 row steps) but give fewer tasks to balance over the resident warp pairs: 64 rows when that still
// leaves two tasks per pair, else 32, else (fewer tasks than pairs: small batches such as the reference's default B = 4) 16.
// Measured at config 2: 16 -> 344 us, 32 -> 318, 64 -> 308, 128 -> 416.
static int build_dev_loss_for(const SndeLossArgs *a, DevLoss &d, const LossWorkspace *lay, int64_t *tiles, int sms) {
    const SegOverrides &ov = seg_overrides();
    int rc = build_dev_loss(a, d, lay, tiles, ov.seg_rows ? ov.seg_rows : 64, ov.fine_rows, ov.fine_permille);   // validates every argument
    if (rc || ov.seg_rows) return rc;
    const long long resident_pairs = (long long)sms * SNDE_STRIP_CTAS_PER_SM * StripCfg::PAIRS;
    if (*tiles < 2 * resident_pairs) {
        rc = build_dev_loss(a, d, lay, tiles, 32, ov.fine_rows, ov.fine_permille);
        if (!rc && *tiles < resident_pairs) rc = build_dev_loss(a, d, lay, tiles, 16, ov.fine_rows, ov.fine_permille);
    }
    return rc;
}

// host-only preview of the task count of snde_loss_fwdbwd on the current device (148 SMs when no device is available)
extern "C" int64_t snde_loss_num_tiles(const SndeLossArgs *args) {
    DevLoss d; int64_t n = 0;
    if (!args) return SNDE_ERR_NULL;
    int devi = 0, sms = 0;
    if (cudaGetDevice(&devi) == cudaSuccess && devi >= 0 && devi < 64) sms = sm_count_of(devi);
    const int rc = build_dev_loss_for(args, d, nullptr, &n, sms > 0 ? sms : 148);
    return rc ? rc : n;
}

extern "C" int snde_loss_fwdbwd(const SndeLossArgs *a, void *stream_) {
    if (!a) return SNDE_ERR_NULL;
    if (!a->poses || !a->loss_out) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD) && !a->grad_poses) return SNDE_ERR_NULL;
    if ((a->flags & SNDE_LOSS_GRAD_K) && !a->grad_K) return SNDE_ERR_NULL;
    if (a->B <= 0 || a->S <= 0 || a->num_scales <= 0 || a->num_scales > SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
    if (a->S * a->num_scales > 64) return SNDE_ERR_SHAPE;     // finalize kernel: one thread per (source, scale)
    const LossWorkspace lay = loss_ws_layout(a->B, a->S, a->num_scales);
    if (!a->workspace || a->workspace_bytes < lay.end || ((uintptr_t)a->workspace & 15)) return SNDE_ERR_WORKSPACE;
    DevLoss d; int64_t tiles = 0;
    int rc = build_dev_loss(a, d, &lay, &tiles, 64);            // validates every argument before any CUDA call
    if (rc) return rc;
    int devi = 0;
    if (cudaGetDevice(&devi) != cudaSuccess || devi < 0 || devi >= 64) return SNDE_ERR_CUDA;
    const int sms = sm_count_of(devi);
    if (sms <= 0) return SNDE_ERR_CUDA;
    rc = build_dev_loss_for(a, d, &lay, &tiles, sms);           // the task table for this device
    if (rc) return rc;
    cudaStream_t st = as_stream(stream_);
    const int B = a->B, S = a->S, ns = a->num_scales, N = B * S;
    unsigned char *ws = (unsigned char *)a->workspace;
    PrepParams kp; FinalParams fp;
    build_prep_final(a, kp, fp);
    // [gP, end) is a whole number of 16-byte words (every offset of the layout is rounded up to 16)
    if (snde_cuda_failed(launch_pdl(loss_prep_kernel, dim3((ns * N + 127) / 128), dim3(128), 0, st, B, S, ns, a->poses, kp, (float *)(ws + lay.P),
                                    (float *)(ws + lay.Kinv), (uint4 *)(ws + lay.gP), (int)((lay.end - lay.gP) / 16)), "prep launch", __LINE__))
        return SNDE_ERR_CUDA;
    const size_t smem = sizeof(PairSmem) * StripCfg::PAIRS;
    const bool grad = a->flags & SNDE_LOSS_GRAD, gradk = a->flags & SNDE_LOSS_GRAD_K;
    // opt in to > 48 KB dynamic shared memory once per kernel variant and device (not a stream operation; kept out of
    // the per-call path so the call can be captured into a CUDA graph after one eager warm-up).  Setting it twice from two
    // racing threads is harmless: same value.
    // persistent grid: as many CTAs as fit, but no more warp pairs than tasks
    long long ctas = (long long)sms * SNDE_STRIP_CTAS_PER_SM;
    const long long need = (tiles + StripCfg::PAIRS - 1) / StripCfg::PAIRS;
    if (ctas > need) ctas = need;
    auto launch = [&](auto kern, int variant) -> int {
        if (!g_attr_done[variant][devi].load(std::memory_order_acquire)) {
            if (snde_cuda_failed(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "shared-memory opt-in", __LINE__)) return SNDE_ERR_CUDA;
            g_attr_done[variant][devi].store(true, std::memory_order_release);
        }
        if (snde_cuda_failed(launch_pdl(kern, dim3((unsigned)ctas), dim3(StripCfg::NT), smem, st, d), "strip launch", __LINE__)) return SNDE_ERR_CUDA;
        return SNDE_OK;
    };
    // EXTRA variants honour masks / warped+error outputs / border padding / intrinsics gradients; FAST ones need planar,
    // aligned tensors at every scale (the optional mask / output tensors keep their own strides).  The lean fast one is
    // the hot path.
    bool extra = gradk || a->padding_mode != SNDE_PAD_ZEROS, fast = true, gimg = false;
    const bool disp = (a->flags & SNDE_LOSS_DISPARITY) != 0;
    for (int s = 0; s < ns; ++s) {
        const SndeLossScale &in = a->scale[s];
        extra = extra || in.mask_fwd.ptr || in.mask_bwd.ptr || in.warped_fwd.ptr || in.warped_bwd.ptr || in.err_fwd.ptr || in.err_bwd.ptr;
        gimg = gimg || in.grad_tgt_img.ptr || in.grad_src_img.ptr;
        fast = fast && d.sc[s].fast;
    }
    if (gimg) extra = true;                      // image gradients ride on the general instance
    // disparity input: one more lean instance (<GRAD, FAST, DISP>); every other combination goes through the general one
    if (disp && !(grad && fast && !extra)) extra = true;
    if (!grad) {
        if (extra) rc = launch(loss_strip_kernel<false, false, true, false>, 0);
        else rc = fast ? launch(loss_strip_kernel<false, false, false, true>, 1) : launch(loss_strip_kernel<false, false, false, false>, 2);
    } else if (gimg) rc = fast ? launch(loss_strip_kernel<true, true, true, true, false, true>, 8) : launch(loss_strip_kernel<true, true, true, false, false, true>, 9);
    else if (extra) rc = fast ? launch(loss_strip_kernel<true, true, true, true>, 7) : launch(loss_strip_kernel<true, true, true, false>, 3);
    else if (disp) rc = launch(loss_strip_kernel<true, false, false, true, true>, 6);
    else rc = fast ? launch(loss_strip_kernel<true, false, false, true>, 4) : launch(loss_strip_kernel<true, false, false, false>, 5);
    if (rc) return rc;
    SNDE_CHECK_LAUNCH();
    if (snde_cuda_failed(launch_pdl(loss_final_kernel, dim3(B + 1), dim3(64), 0, st, B, S, ns, (int)a->flags, a->poses, fp, (const double *)(ws + lay.gP),
                                    (const double *)(ws + lay.gKinv), (const double *)(ws + lay.loss), a->loss_out, a->grad_poses, a->grad_K),
                         "finalize launch", __LINE__))
        return SNDE_ERR_CUDA;
    return SNDE_OK;
}

