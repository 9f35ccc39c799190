// snde_strip.cuh -- the fused loss as a STREAMING STRIP PROGRAM run by warp PAIRS.
//
// Work decomposition.  Every image of the step owns exactly one depth map:
//   class F task: owner = target image b; "others" = its S source images, projected with
//                 P = K4 @ T_k               (the reference's fwd direction, baseline.py:903-906)
//   class R task: owner = source image n = k*B + b; other = the target image b, projected with
//                 P = K4 @ inverse(T_k)      (the bwd direction, baseline.py:909-913)
// A task is (owner, strip of 60 columns, segment of seg_h rows).  A pair of warps walks DOWN the strip one row per
// step; in both warps lane l owns the two adjacent region columns 2l, 2l+1 (region = the 60 inner columns + a
// 2-column halo each side).  The two warps are specialised and hand rows to each other through shared memory, with
// ONE 64-thread named barrier per row and no CTA-wide barrier anywhere:
//   geometry warp           A  row t+1: reproject the lane's two pixels with the owner's depth and start the 4-tap
//                              gathers of the other image (software prefetch: the loads land one step later)
//                           A' row t  : blend the gathered taps -> warped row (shared) + tap derivatives (private)
//                           G  row t-3: dot dL/d(warped) with the tap derivatives, back through the projection
//                                       -> d/d(depth), d/dP, d/dKinv
//                           S  row t-3: (first pass only) the owner's edge-aware smoothness pairs + depth gradient
//   SSIM warp               B  row t-1: SSIM over the 3x3 window from rolling row sums (horizontal neighbours by
//                                       shuffle, vertical by two carried rows), error value, backward coefficients
//                           C  row t-2: 3x3 box sums of the coefficient maps (same rolling scheme)
//                                       -> dL/d(warped) row (shared)
// Splitting the work this way halves the state a warp has to carry (the rolling sums live only in the SSIM warp; the
// projection matrices, gather registers and gradient accumulators only in the geometry warp), which is what lets
// enough warps be resident to hide the gather and shuffle latencies.
// Each depth-gradient element is written by exactly one lane (passes over the S others of a class-F owner add to
// it in a fixed order), so the gradient is deterministic; only the per-sample 3x4 / 3x3 sums and the loss scalars
// go through (double-precision) atomics.
//
// The code is per-lane ("SIMT") code.  On the GPU the neighbour exchange is __shfl_*_sync and the pair barrier is
// bar.sync; under the host-emulation test harness (g++, tests/host_emul) the 64 lanes of a pair are fibers and both
// go through counting barriers, which lets the strip logic be checked against the oracle without a GPU.  The
// product never runs the host variant.
#pragma once
#include <string.h>

#include "snde_math.cuh"

#ifndef SNDE_MAX_SCALES
#define SNDE_MAX_SCALES 6
#endif
#ifndef SNDE_STRIP_PAIRS
#define SNDE_STRIP_PAIRS 2
#endif
// 1: the SSIM warp's row step is one basic block (phases B and C run on every row and are masked), so that the
// scheduler can interleave the three channels; costs registers (tuning switch, measured in profiles/README.md)
// 1: the geometry warp's row step runs reprojection + gathers of row t+1 and the back-projection of row t-3 as ONE basic
// block (both unconditionally, idle rows / lanes masked), so that the scheduler can interleave the two dependency chains
#ifndef SNDE_GEO_STRAIGHT
#define SNDE_GEO_STRAIGHT 1
#endif
// 1: two copies of the geometry warp's row loop, with and without the smoothness term (no run-time branch around it);
// measured slower than one copy with a uniform branch (more hot code than the instruction cache likes)
#ifndef SNDE_SMOOTH_SPLIT
#define SNDE_SMOOTH_SPLIT 0
#endif
#ifndef SNDE_GATHER_WALK
#define SNDE_GATHER_WALK 1       // tap rows walked channel by channel as pinned pointers (tuning switch)
#endif
#ifndef SNDE_SSIM_STRAIGHT
#define SNDE_SSIM_STRAIGHT 1
#endif
// 1: the lean instances (zero padding, no optional outputs) blend the four taps as two nested linear interpolations, which
// yields the warped value and both tap derivatives from the same differences (27 packed operations per row instead of 46);
// an invalid pixel is switched off by one 0/1 factor on the warped value and by a validity bit on the way back (the bit
// travels in a per-lane shift register: row t-3's bits are 4 pushes old when the row is taken back through the projection)
#ifndef SNDE_LERP_BLEND
#define SNDE_LERP_BLEND 1
#endif
// 1: the lean instances add the L1 part of the error sum when a row arrives (phase H) instead of one row later (phase B
// re-read x and y of row t-1 for nothing else: 6 shared-memory loads per row less), and form the three backward
// coefficients from one shared factor ngS * rB (12 packed operations instead of 16)
#ifndef SNDE_SSIM_LEAN
#define SNDE_SSIM_LEAN 1
#endif
// 1: the SSIM warp's row step runs phase-major over the three colour channels (see the row step)
#ifndef SNDE_SSIM_PHASED
#define SNDE_SSIM_PHASED 1
#endif
// 1: the geometry warp runs the smoothness term on every pass (masked to zero after the first) instead of branching around it
#ifndef SNDE_SMOOTH_ALWAYS
#define SNDE_SMOOTH_ALWAYS 0
#endif

namespace snde {

struct StripCfg {
    static constexpr int RW = 64;        // region columns of a strip (32 lanes x 2 columns)
    static constexpr int TW = RW - 4;    // inner columns (60)
    static constexpr int SEG = 32;       // default rows per segment
    static constexpr int PAIRS = SNDE_STRIP_PAIRS;   // warp pairs per CTA (independent of each other)
    static constexpr int NT = PAIRS * 64;
    static constexpr int MAX_S = 8;      // most source views per target (SNDE_MAX_SOURCES)
};

struct DevImage { const float *p; long long sn; int sc, sh, sw, pad_; };
struct DevImageOut { float *p; long long sn; int sc, sh, sw, pad_; };
struct DevMap { const float *p; long long sn; int sh, sw; };
struct DevMapOut { float *p; long long sn; int sh, sw; };

struct DevScale {
    int h, w, tiles_x, fast;         // tiles_x = strips per row; fast != 0: every tensor of this scale is planar, unit x-stride, 8-byte aligned rows
    int pad0_, pad1_;
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
    DevImageOut gimg_tgt, gimg_src;  // optional: d loss_rigid_warp / d image, accumulated (section 8(f3))
};

// one block of the task table: `count` tasks starting at number `begin`; owners first_owner.. of class cls (0 = F, 1 = R)
// at `scale`, cut into tiles_y segments of seg_h rows
// (y_base: first row of the block's segments; a scale whose height is not a multiple of seg_h gets a second block for the
//  short last segment, so that the table can be ordered by task cost)
struct DevEntry { int begin, count, scale, cls, first_owner, seg_h, tiles_y, y_base; };

struct DevLoss {
    int B, S, num_scales, flags, padding, num_entries;
    float alpha, oma;
    int total_tasks, pad_;
    int *counter;                    // workspace: next task to hand out (zeroed per call)
    DevEntry entry[8 * SNDE_MAX_SCALES];
    DevScale sc[SNDE_MAX_SCALES];
};

enum { FLAG_GRAD = 1, FLAG_GRAD_K = 2, FLAG_NO_PHOTO = 4, FLAG_NO_SMOOTH = 8, FLAG_DISP = 16 };   // == SNDE_LOSS_* in include/snde.h


struct alignas(16) F4 { float x, y, z, w; };
struct SmoothOn { static constexpr bool value = true; };
struct SmoothOff { static constexpr bool value = false; };

// Shared memory of one warp pair.  Every element is written by one lane and read by the same lane index of either
// warp (neighbouring columns are exchanged by shuffles inside a warp), so the only synchronisation is the pair barrier.
struct PairSmem {
    F2 X[8][3][32];                  // owner image rows (slot = row & 7; rows t-3 .. t+3 are live), 0 outside the image (= the pools' zero padding)
    F2 D[8][32];                     // owner depth rows
    F2 Y[4][3][32];                  // warped other image rows, 0 outside the image          geometry -> SSIM
    F2 PO[2][3][32];                 // dL/d(warped) rows (slot = row & 1)                     SSIM -> geometry
    F2 G[4][6][32];                  // d(warped_c)/dx, /dy (geometry warp only)
    float YT[8];                     // meshgrid y of the ring rows (copied with the row: no global load in the row loop)
    int task; int pad_[3];
};

#if defined(__CUDACC__)
#define SNDE_STRIP_FN __device__ __forceinline__
#else
#define SNDE_STRIP_FN inline
#endif

// ---------------------------------------------------------------------------------------------------------------
// warp / pair primitives.  `who` is the lane on the GPU; under host emulation it is role*32 + lane.
// ---------------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
// from_left[i] = to_right[i] of lane-1, from_right[i] = to_left[i] of lane+1 (own value at the warp's edges)
template <int N>
SNDE_STRIP_FN void neighbor_exchange(int, const float (&to_right)[N], const float (&to_left)[N], float (&from_left)[N], float (&from_right)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        from_left[i] = __shfl_up_sync(0xffffffffu, to_right[i], 1);
        from_right[i] = __shfl_down_sync(0xffffffffu, to_left[i], 1);
    }
}
template <int N>
SNDE_STRIP_FN void take_from_right(int, const float (&to_left)[N], float (&from_right)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) from_right[i] = __shfl_down_sync(0xffffffffu, to_left[i], 1);
}
template <int N>
SNDE_STRIP_FN void take_from_left(int, const float (&to_right)[N], float (&from_left)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) from_left[i] = __shfl_up_sync(0xffffffffu, to_right[i], 1);
}
// sum over the warp of N per-lane values; the totals come back in every lane
template <int N>
SNDE_STRIP_FN void warp_sum(int, float (&v)[N]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}
// Warp sums of N <= 16 per-lane values by a transposing butterfly: every xor step halves the number of live values
// (16 shuffles in all instead of 5 per value).  The sum of value i comes back in the even lane whose index is
// warp_sum_owner(i); other lanes get unspecified values.
SNDE_STRIP_FN int warp_sum_slot(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
template <int N>
SNDE_STRIP_FN float warp_sum_scatter(int lane, const float (&in)[N]) {
    static_assert(N <= 16, "at most 16 values");
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = i < N ? in[i] : 0.f;
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
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
SNDE_STRIP_FN void atomic_add_double(double *dst, double v) { atomicAdd(dst, v); }
// dst[0..1] += v without reading anything back (no scoreboard to wait on).  Only ever used on elements that this very
// lane wrote earlier and nobody else touches: one add per element and pass, so the result does not depend on timing.
SNDE_STRIP_FN void add_pair_noreturn(float *dst, F2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(v.x), "f"(v.y) : "memory");
}
SNDE_STRIP_FN void atomic_add_float(float *dst, float v) { atomicAdd(dst, v); }
SNDE_STRIP_FN void add_pair_atomic(float *dst, F2 v) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(v.x), "f"(v.y) : "memory"); }
// Warp-aggregated scatter: all 32 lanes call this together.  Lanes with the same `key` (a pixel of the image being
// scattered into) form a group; the lowest lane of a group gathers the others' values in lane order and alone issues the
// atomic.  `above` = the group's lanes above this one, `max_group` = the size of the largest group of the warp (the trip
// count of the gathering loop; 1 when no two lanes collide).
struct TapGroup { unsigned above; int max_group; bool leader; };
SNDE_STRIP_FN TapGroup tap_group(int lane, long long key) {
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    TapGroup g;
    g.above = lane == 31 ? 0u : peers & ~((2u << lane) - 1u);
    g.leader = (peers & ((1u << lane) - 1u)) == 0u;
    g.max_group = (int)__reduce_max_sync(0xffffffffu, (unsigned)__popc(peers));
    return g;
}
SNDE_STRIP_FN float group_sum(int lane, const TapGroup &g, float v) {
    float sum = v;
    unsigned rest = g.above;
    for (int it = 1; it < g.max_group; ++it) {
        const int src = rest ? __ffs((int)rest) - 1 : lane;
        const float o = __shfl_sync(0xffffffffu, v, src);
        if (rest) { sum += o; rest &= rest - 1u; }
    }
    return sum;
}
SNDE_STRIP_FN void warp_fence(int) { __syncwarp(); }
// rendezvous of the two warps of a pair (named barrier `bar`, 64 threads); orders their shared-memory traffic
SNDE_STRIP_FN void pair_barrier(int, int bar) { asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory"); }
// asynchronous global -> shared copy of BYTES (4 or 8) bytes; `take` == false writes zeros instead (LDGSTS with src-size 0)
template <int BYTES>
SNDE_STRIP_FN void row_copy_async(void *smem_dst, const void *src, bool take) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = take ? BYTES : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(src), "n"(BYTES), "r"(n) : "memory");
}
SNDE_STRIP_FN void row_copy_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
SNDE_STRIP_FN void row_copy_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
SNDE_STRIP_FN F2 ldg2(const float *p) { const float2 v = __ldg(reinterpret_cast<const float2 *>(p)); return F2{v.x, v.y}; }
SNDE_STRIP_FN F2 ldcg2(const float *p) { const float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); return F2{v.x, v.y}; }
// Pin a pointer in a register pair: stops the compiler from re-deriving it from (struct field + 64-bit offset) at
// every use, so that base[int_offset] costs one IMAD.WIDE instead of a four-instruction 64-bit add-and-shift.
#define SNDE_PIN_PTR(p) asm volatile("" : "+l"(p))
#define SNDE_SCHED_FENCE() asm volatile("" ::: "memory")
// keep a per-task float in a register (the compiler would re-load it from the constant bank through two dependent
// LDCs at every use)
#define SNDE_PIN_F32(x) asm volatile("" : "+f"(x))
#else
// host emulation (tests/host_emul/emul.cpp): 64 fibers per pair; exchange slots are double-buffered
float *host_slots(int who, int parity);             // 64 floats per fiber and parity
int &host_parity(int who);
void host_warp_sync(int who);                       // barrier over the 32 fibers of who's warp
void host_pair_sync(int who);                       // barrier over the 64 fibers of the pair
template <int N>
SNDE_STRIP_FN void neighbor_exchange(int who, const float (&to_right)[N], const float (&to_left)[N], float (&from_left)[N], float (&from_right)[N]) {
    static_assert(2 * N <= 64, "slot array too small");
    const int par = host_parity(who), lane = who & 31;
    float *mine = host_slots(who, par);
    for (int i = 0; i < N; ++i) { mine[i] = to_right[i]; mine[N + i] = to_left[i]; }
    host_warp_sync(who);
    const float *l = host_slots(lane > 0 ? who - 1 : who, par), *r = host_slots(lane < 31 ? who + 1 : who, par);
    for (int i = 0; i < N; ++i) { from_left[i] = l[i]; from_right[i] = r[N + i]; }
    host_parity(who) = par ^ 1;
}
template <int N>
SNDE_STRIP_FN void take_from_right(int who, const float (&to_left)[N], float (&from_right)[N]) {
    float zr[N], fl[N];
    for (int i = 0; i < N; ++i) zr[i] = 0.f;
    neighbor_exchange<N>(who, zr, to_left, fl, from_right);
}
template <int N>
SNDE_STRIP_FN void take_from_left(int who, const float (&to_right)[N], float (&from_left)[N]) {
    float zl[N], fr[N];
    for (int i = 0; i < N; ++i) zl[i] = 0.f;
    neighbor_exchange<N>(who, to_right, zl, from_left, fr);
}
template <int N>
SNDE_STRIP_FN void warp_sum(int who, float (&v)[N]) {
    static_assert(N <= 64, "slot array too small");
    const int par = host_parity(who), base = who & ~31;
    float *mine = host_slots(who, par);
    for (int i = 0; i < N; ++i) mine[i] = v[i];
    host_warp_sync(who);
    for (int i = 0; i < N; ++i) {
        float part[32];
        for (int l = 0; l < 32; ++l) part[l] = host_slots(base + l, par)[i];
        for (int o = 16; o > 0; o >>= 1) for (int l = 0; l < o; ++l) part[l] += part[l + o];   // same pairing as the xor butterfly
        v[i] = part[0];
    }
    host_parity(who) = par ^ 1;
}
SNDE_STRIP_FN int warp_sum_slot(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
template <int N>
SNDE_STRIP_FN float warp_sum_scatter(int who, const float (&in)[N]) {
    float v[N];
    for (int i = 0; i < N; ++i) v[i] = in[i];
    warp_sum<N>(who, v);
    const int slot = warp_sum_slot(who & 31);
    return slot < N ? v[slot] : 0.f;
}
SNDE_STRIP_FN void atomic_add_double(double *dst, double v) { *dst += v; }
SNDE_STRIP_FN void add_pair_noreturn(float *dst, F2 v) { dst[0] += v.x; dst[1] += v.y; }
SNDE_STRIP_FN void atomic_add_float(float *dst, float v) { *dst += v; }
SNDE_STRIP_FN void add_pair_atomic(float *dst, F2 v) { dst[0] += v.x; dst[1] += v.y; }
struct TapGroup { unsigned above; int max_group; bool leader; };
SNDE_STRIP_FN TapGroup tap_group(int, long long) { return TapGroup{0u, 1, true}; }        // emulation: every fiber adds its own value
SNDE_STRIP_FN float group_sum(int, const TapGroup &, float v) { return v; }
SNDE_STRIP_FN void warp_fence(int who) { host_warp_sync(who); }
SNDE_STRIP_FN void pair_barrier(int who, int) { host_pair_sync(who); }
template <int BYTES>
SNDE_STRIP_FN void row_copy_async(void *smem_dst, const void *src, bool take) {
    if (take) memcpy(smem_dst, src, BYTES); else memset(smem_dst, 0, BYTES);
}
SNDE_STRIP_FN void row_copy_commit() {}
template <int PENDING>
SNDE_STRIP_FN void row_copy_wait() {}
SNDE_STRIP_FN F2 ldg2(const float *p) { return F2{p[0], p[1]}; }
SNDE_STRIP_FN F2 ldcg2(const float *p) { return F2{p[0], p[1]}; }
#define SNDE_PIN_PTR(p) ((void)0)
#define SNDE_SCHED_FENCE() ((void)0)
#define SNDE_PIN_F32(x) ((void)0)
#endif

// g * sign(d) (0 for d == 0) without a branch: the sign bit of d is xor-ed into g
SNDE_STRIP_FN float times_sign(float g, float d) {
#if defined(__CUDA_ARCH__)
    const float s = __int_as_float(__float_as_int(g) ^ (__float_as_int(d) & (int)0x80000000u));
    return d != 0.f ? s : 0.f;
#else
    return d > 0.f ? g : (d < 0.f ? -g : 0.f);
#endif
}

// horizontal 3-sums of one row for the lane's two columns (left column = left neighbour + (a + b), right column =
// (a + b) + right neighbour: the association the golden-vector tolerances were measured with)
// (xx + yy is carried as ONE sum q: SSIM and its backward use the two second moments only through their sum)
struct HS { P2 x, y, q, xy; };
struct CS { P2 a, b, c; };
struct CSX { P2 d; };           // box sums of the x-side coefficient (image gradients only)

SNDE_STRIP_FN void hsum_row(const F2 xv, const F2 yv, float xm, float xp, float ym, float yp, HS &o) {
    const float tx = xv.x + xv.y, ty = yv.x + yv.y;
    const float tq = fmaf(yv.y, yv.y, fmaf(yv.x, yv.x, fmaf(xv.y, xv.y, xv.x * xv.x))), txy = fmaf(xv.y, yv.y, xv.x * yv.x);
    // the two neighbour columns as one pair (left, right): one packed operation per output, the lane's own part a broadcast
    // operand.  Lane-wise this is (xm + tx, tx + xp), (fma(ym, ym, fma(xm, xm, tq)), ...) as before: addition commutes exactly.
    const P2 mx = p2(xm, xp), my = p2(ym, yp);
    o.x = p2add(mx, p2s(tx));
    o.y = p2add(my, p2s(ty));
    o.q = p2fma(my, my, p2fma(mx, mx, p2s(tq)));
    o.xy = p2fma(mx, my, p2s(txy));
}
SNDE_STRIP_FN void hs_add(const HS &a, const HS &b, HS &o) {
    o.x = p2add(a.x, b.x); o.y = p2add(a.y, b.y); o.q = p2add(a.q, b.q); o.xy = p2add(a.xy, b.xy);
}
SNDE_STRIP_FN void cs_add(const CS &a, const CS &b, CS &o) { o.a = p2add(a.a, b.a); o.b = p2add(a.b, b.b); o.c = p2add(a.c, b.c); }
SNDE_STRIP_FN void hs_zero(HS &o) { o.x = o.y = o.q = o.xy = p2s(0.f); }
SNDE_STRIP_FN void cs_zero(CS &o) { o.a = o.b = o.c = p2s(0.f); }

// ---------------------------------------------------------------------------------------------------------------
// what both warps of a pair derive from the task number (uniform over the pair, except the column flags)
// ---------------------------------------------------------------------------------------------------------------
struct StripTask {
    int scale, m, b, h, w, y0, y1, ix0, t_begin, t_end, n_others, n_pass;
    bool clsF, do_smooth, do_photo;
    bool in[2], colB[2], colI[2];
    float mI[2];              // 1 for an inner in-image column, else 0
};

SNDE_STRIP_FN void decode_task(const DevLoss &p, int task, int lane, StripTask &T) {
    using C = StripCfg;
    int e = 0;
    while (e + 1 < p.num_entries && task >= p.entry[e + 1].begin) ++e;
    const DevEntry ent = p.entry[e];
    const DevScale &sc = p.sc[ent.scale];
    const int local = task - ent.begin;
    const int per_owner = sc.tiles_x * ent.tiles_y;
    T.scale = ent.scale;
    const int mo = local / per_owner;
    T.m = ent.first_owner + mo;                      // owner index inside its class
    const int tt = local - mo * per_owner;
    const int sy = tt / sc.tiles_x, sx = tt - sy * sc.tiles_x;
    T.h = sc.h; T.w = sc.w;
    T.y0 = ent.y_base + sy * ent.seg_h; T.y1 = T.y0 + ent.seg_h < sc.h ? T.y0 + ent.seg_h : sc.h;
    T.ix0 = sx * C::TW - 2 + 2 * lane;
    T.clsF = ent.cls == 0;
    T.b = T.clsF ? T.m : (T.m % p.B);                // batch element (selects K)
    T.do_smooth = !(p.flags & FLAG_NO_SMOOTH);
    T.do_photo = !(p.flags & FLAG_NO_PHOTO);
    T.n_others = T.do_photo ? (T.clsF ? p.S : 1) : 0;
    T.n_pass = T.n_others > 0 ? T.n_others : 1;      // smoothness alone still needs one pass over the owner
    T.t_begin = T.y0 - 2 > 0 ? T.y0 - 2 : 0;         // rows above the image are zero rows == the initial state
    T.t_end = T.y1 + 1;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int c = 2 * lane + q, ix = T.ix0 + q;
        T.in[q] = ix >= 0 && ix < T.w;
        T.colB[q] = T.in[q] && c >= 1 && c <= C::RW - 2;
        T.colI[q] = T.in[q] && c >= 2 && c <= C::RW - 3;
        T.mI[q] = T.colI[q] ? 1.f : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Owner rows (image + depth + meshgrid y) on their way into the pair's ring.  Run by the SSIM warp, which has the slack
// (the geometry warp is the longer of the two): one commit group per row, zeros outside the image.
// ---------------------------------------------------------------------------------------------------------------
template <bool FAST>
struct OwnerRows {
    const float *own_base; int own_sc, own_sh, own_sw;
    const float *dep_base; int dep_sh, dep_sw;
    const float *ytab;

    SNDE_STRIP_FN void init(const DevLoss &p, const DevScale &sc, const StripTask &T) {
        const DevImage &own = T.clsF ? sc.tgt : sc.src;
        own_base = own.p + (long long)T.m * own.sn; own_sc = own.sc; own_sh = own.sh; own_sw = own.sw;
        const long long depth_n = T.clsF ? T.m : (long long)p.B + T.m;
        dep_base = sc.depth.p + depth_n * sc.depth.sn; dep_sh = sc.depth.sh; dep_sw = sc.depth.sw;
        ytab = sc.ytab;
        SNDE_PIN_PTR(own_base); SNDE_PIN_PTR(dep_base);
    }
    // owner row iy of the lane's two columns -> ring slot iy & 7, asynchronously (no registers, no copy instruction waits
    // for it)
    SNDE_STRIP_FN void prefetch_row(PairSmem &sm, const StripTask &T, int lane, int iy) const {
        const bool row_in = iy >= 0 && iy < T.h;
        const int y = row_in ? iy : 0;                        // any valid address when nothing is taken
        F2 (*xs)[32] = sm.X[iy & 7];
        if (FAST) {
            const bool take = row_in && T.in[0];              // w and ix0 are even: the pair is inside or outside together
            const int q = y * own_sh + (take ? T.ix0 : 0);
            row_copy_async<8>(&xs[0][lane], own_base + q, take);
            row_copy_async<8>(&xs[1][lane], own_base + (q + own_sc), take);
            row_copy_async<8>(&xs[2][lane], own_base + (q + 2 * own_sc), take);
            row_copy_async<8>(&sm.D[iy & 7][lane], dep_base + (y * dep_sh + (take ? T.ix0 : 0)), take);
        } else {
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
                const bool take = row_in && T.in[k2];
                const int ix = take ? T.ix0 + k2 : 0;
                const int q = y * own_sh + ix * own_sw;
                row_copy_async<4>(&xs[0][lane].x + k2, own_base + q, take);
                row_copy_async<4>(&xs[1][lane].x + k2, own_base + (q + own_sc), take);
                row_copy_async<4>(&xs[2][lane].x + k2, own_base + (q + 2 * own_sc), take);
                row_copy_async<4>(&sm.D[iy & 7][lane].x + k2, dep_base + (y * dep_sh + ix * dep_sw), take);
            }
        }
        // the row's meshgrid y rides along (lane 0): a global load inside the geometry warp's row loop would share a
        // scoreboard with the gathers and make the warp wait for all of them
        if (lane == 0) row_copy_async<4>(&sm.YT[iy & 7], ytab + y, row_in);
        row_copy_commit();
    }
    // disparity input (SNDE_LOSS_DISPARITY): the landed row iy of the ring becomes depth = 1/disp (IEEE division, as
    // torch.reciprocal; models/baseline.py:835); slots outside the image stay 0
    SNDE_STRIP_FN void disparity_to_depth(PairSmem &sm, const StripTask &T, int lane, int iy) const {
        if (iy < 0 || iy >= T.h) return;
        F2 v = sm.D[iy & 7][lane];
        v.x = T.in[0] ? SNDE_XDIV(1.f, v.x) : 0.f;
        v.y = T.in[1] ? SNDE_XDIV(1.f, v.y) : 0.f;
        sm.D[iy & 7][lane] = v;
    }
};

// ===============================================================================================================
// SSIM warp
// ===============================================================================================================
template <bool GRAD, bool EXTRA, bool FAST, bool DISP, bool GIMG = false>
SNDE_STRIP_FN void ssim_role(const DevLoss &p, int task, int who, PairSmem &sm, int bar) {
    const int lane = who & 31;
    StripTask T;
    decode_task(p, task, lane, T);
    const DevScale &sc = p.sc[T.scale];
    OwnerRows<FAST> rows; rows.init(p, sc, T);
    const bool disp_in = DISP || (EXTRA && (p.flags & FLAG_DISP) != 0);
    float alpha = p.alpha, oma = p.oma, g1 = p.oma * sc.gphoto;
    SNDE_PIN_F32(alpha); SNDE_PIN_F32(oma); SNDE_PIN_F32(g1);
    float ngSm[2];            // -dL/dS = 0.5*alpha*gphoto where the column takes part in phase B, else 0
    constexpr bool LEAN = SNDE_SSIM_LEAN && !EXTRA;
#pragma unroll
    for (int q = 0; q < 2; ++q) ngSm[q] = T.colB[q] ? -(-0.5f * p.alpha * sc.gphoto) * (LEAN ? 2.f : 1.f) : 0.f;      // (LEAN carries 2 ngS)
    float alI[2] = {alpha * T.mI[0], alpha * T.mI[1]}, omI[2] = {oma * T.mI[0], oma * T.mI[1]};      // LEAN: the inner-column gate folded in
    const int y0 = T.y0, y1 = T.y1, h = T.h;
    for (int k = 0; k < T.n_pass; ++k) {
        const long long n = T.clsF ? (long long)k * p.B + T.b : T.m;      // index into the N-sized tensors
        const float *mask_base = nullptr; int mask_sh = 0, mask_sw = 0;
        float *eout_base = nullptr; int eout_sc = 0, eout_sh = 0, eout_sw = 0;
        if (EXTRA) {
            const DevMap &mask = T.clsF ? sc.mask_fwd : sc.mask_bwd;
            const DevImageOut &eout = T.clsF ? sc.err_fwd : sc.err_bwd;
            mask_base = mask.p ? mask.p + n * mask.sn : nullptr; mask_sh = mask.sh; mask_sw = mask.sw;
            eout_base = eout.p ? eout.p + n * eout.sn : nullptr; eout_sc = eout.sc; eout_sh = eout.sh; eout_sw = eout.sw;
        }
        const bool has_mask = EXTRA && mask_base != nullptr;
        // image gradients (GIMG): the dense part -- d loss / d(owner image) through the x argument of image_similarity -- is
        // formed here next to d loss / d(warped) and added to the owner's gradient image
        float *gown_base = nullptr; int gown_sc = 0, gown_sh = 0, gown_sw = 0;
        if (GIMG) {
            const DevImageOut &go = T.clsF ? sc.gimg_tgt : sc.gimg_src;
            gown_base = go.p ? go.p + (long long)T.m * go.sn : nullptr; gown_sc = go.sc; gown_sh = go.sh; gown_sw = go.sw;
        }
        // rolling sums: the two carried rows start as zero rows
        HS s1[3], s2[3];          // H(t-2) + H(t-1), H(t-1)
        CS c1[3], c2[3];          // Hc(t-3) + Hc(t-2), Hc(t-2)
        CSX xa1[3], xa2[3];       // the same for the x-side coefficient (GIMG)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { hs_zero(s1[ch]); hs_zero(s2[ch]); cs_zero(c1[ch]); cs_zero(c2[ch]); xa1[ch].d = xa2[ch].d = p2s(0.f); }
        float err_acc = 0.f;

        // prologue: owner rows t_begin .. t_begin+2 on their way into the ring, the first two landed and published
        rows.prefetch_row(sm, T, lane, T.t_begin); rows.prefetch_row(sm, T, lane, T.t_begin + 1); rows.prefetch_row(sm, T, lane, T.t_begin + 2);
        row_copy_wait<1>();
        if (disp_in) { rows.disparity_to_depth(sm, T, lane, T.t_begin); rows.disparity_to_depth(sm, T, lane, T.t_begin + 1); }
        pair_barrier(who, bar);

        auto row_step = [&](const int t) {
            pair_barrier(who, bar);              // row t of the warped image is in shared memory
            if (t <= T.t_end && T.do_photo) {
            const int rb = t - 1, rc = t - 2;
            const bool doB = t >= y0, doC = t >= y0 + 2;
            const bool rowB = doB && rb >= 0 && rb < h;
            const bool row_inner = rb >= y0 && rb < y1;
            float mk[2] = {1.f, 1.f}, gl[2] = {g1, g1};
            if (has_mask) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (rowB && T.colB[q]) mk[q] = ldg(mask_base + rb * mask_sh + (T.ix0 + q) * mask_sw);
                    if (GRAD && doC && T.colI[q]) gl[q] = g1 * ldg(mask_base + rc * mask_sh + (T.ix0 + q) * mask_sw);
                }
            }
            // The row step as four phases per colour channel.  Warp shuffles keep their program order, so the ORDER OF THE PHASES
            // IN THE SOURCE decides what can overlap: channel-major (phases A..C2 of channel 0, then channel 1, ...) runs the
            // three channels strictly one after the other; phase-major (SNDE_SSIM_PHASED: phase A of all channels, then B of
            // all, ...) lets the loads, shuffles and reciprocals of the three channels overlap, at the price of holding the
            // three channels' intermediate rows at once.
            F2 xr3[3], yr3[3]; HS H3[3]; CS co3[3]; P2 cod3[3]; CS Hc3[3]; P2 Hd3[3]; F2 po3[3];
            auto phase_load = [&](const int ch) { xr3[ch] = sm.X[t & 7][ch][lane]; yr3[ch] = sm.Y[t & 3][ch][lane]; };
            auto phase_a = [&](const int ch) {
                // ---------------- horizontal sums of row t ----------------
                const F2 xr = xr3[ch], yr = yr3[ch];
                {
                    const float to_right[2] = {xr.y, yr.y}, to_left[2] = {xr.x, yr.x};
                    float from_left[2], from_right[2];
                    neighbor_exchange<2>(who, to_right, to_left, from_left, from_right);
                    hsum_row(xr, yr, from_left[0], from_right[0], from_left[1], from_right[1], H3[ch]);
                }
                if (LEAN && t >= y0 && t < y1) {      // L1 part of the error of row t (an inner row)
                    err_acc = fmaf(fabsf(xr.x - yr.x), omI[0], err_acc);
                    err_acc = fmaf(fabsf(xr.y - yr.y), omI[1], err_acc);
                }
            };
            auto phase_b = [&](const int ch) {
                // ---------------- phase B (row t-1) ----------------
                // (SNDE_SSIM_STRAIGHT: a row that is not a phase-B row runs the same arithmetic on finite sums and contributes
                // exact zeros (ngq = 0) and nothing to the error sum)
                const HS H = H3[ch];
                CS co; cs_zero(co);
                P2 cod = p2s(0.f);
                if (SNDE_SSIM_STRAIGHT || rowB) {
                    HS tot; hs_add(s1[ch], H, tot);
                    F2 xc = F2{0.f, 0.f}, yc = F2{0.f, 0.f};
                    if (!LEAN) { xc = sm.X[rb & 7][ch][lane]; yc = sm.Y[rb & 3][ch][lane]; }
                    SsimP ss; ssim_sums_p(tot.x, tot.y, tot.q, tot.xy, ss);
                    const P2 v = p2fma(p2s(-0.5f), ss.S, p2s(0.5f));
                    const float vq[2] = {p2lo(v), p2hi(v)};
                    const float vcq[2] = {clampf(vq[0], 0.f, 1.f), clampf(vq[1], 0.f, 1.f)};
                    const float dq[2] = {xc.x - yc.x, xc.y - yc.y};
                    float ngq[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (LEAN) {
                            if (rowB && row_inner) err_acc = fmaf(vcq[q], alI[q], err_acc);        // the SSIM part; the L1 part was added at phase H
                            ngq[q] = (rowB & (vq[q] == vcq[q])) ? ngSm[q] : 0.f;
                            continue;
                        }
                        const float err = fmaf(alpha, vcq[q], oma * fabsf(dq[q]));
                        if (rowB && row_inner) {
                            err_acc = fmaf(EXTRA ? err * mk[q] : err, T.mI[q], err_acc);
                            if (EXTRA && eout_base && T.colI[q]) eout_base[ch * eout_sc + rb * eout_sh + (T.ix0 + q) * eout_sw] = err;
                        }
                        ngq[q] = (rowB & (vq[q] == vcq[q])) ? (EXTRA ? ngSm[q] * mk[q] : ngSm[q]) : 0.f;      // the clamp passes gradient inside [0,1] only
                    }
                    if (GRAD && LEAN) ssim_sums_bwd_lean_p(ss, tot.x, tot.y, p2(ngq[0], ngq[1]), co.a, co.b, co.c);
                    else if (GRAD) ssim_sums_bwd_p(ss, tot.x, tot.y, p2(ngq[0], ngq[1]), co.a, co.b, co.c);
                    if (GRAD && GIMG) cod = ssim_sums_bwd_x_p(ss, tot.x, tot.y, p2(ngq[0], ngq[1]));
                }
                hs_add(s2[ch], H, s1[ch]);        // s1 <- H(t-1) + H(t)
                s2[ch] = H;
                co3[ch] = co; cod3[ch] = cod;
            };
            auto phase_c1 = [&](const int ch) {
                // ---------------- horizontal sums of the coefficient row t-1 ----------------
                const CS co = co3[ch];
                CS Hc;
                {
                    const float al = p2lo(co.a), ah = p2hi(co.a), bl = p2lo(co.b), bh = p2hi(co.b), cl = p2lo(co.c), chh = p2hi(co.c);
                    const float to_right[3] = {ah, bh, chh}, to_left[3] = {al, bl, cl};
                    float from_left[3], from_right[3];
                    neighbor_exchange<3>(who, to_right, to_left, from_left, from_right);
                    const float ta = al + ah, tb = bl + bh, tc = cl + chh;
                    Hc.a = p2add(p2(from_left[0], from_right[0]), p2s(ta));
                    Hc.b = p2add(p2(from_left[1], from_right[1]), p2s(tb));
                    Hc.c = p2add(p2(from_left[2], from_right[2]), p2s(tc));
                }
                P2 Hd = p2s(0.f);
                if (GIMG) {
                    const float dl = p2lo(cod3[ch]), dh = p2hi(cod3[ch]);
                    const float to_right[1] = {dh}, to_left[1] = {dl};
                    float from_left[1], from_right[1];
                    neighbor_exchange<1>(who, to_right, to_left, from_left, from_right);
                    Hd = p2add(p2(from_left[0], from_right[0]), p2s(dl + dh));
                }
                Hc3[ch] = Hc; Hd3[ch] = Hd;
            };
            auto phase_c2 = [&](const int ch) {
                // ---------------- phase C (row t-2): dL/d(warped) ----------------
                // (computed every step, stored when the row exists: see phase B)
                const CS Hc = Hc3[ch]; const P2 Hd = Hd3[ch];
                if (SNDE_SSIM_STRAIGHT || doC) {
                    CS bs; cs_add(c1[ch], Hc, bs);
                    const F2 xf = sm.X[rc & 7][ch][lane], yf = sm.Y[rc & 3][ch][lane];
                    const P2 xv = p2f(xf), yv = p2f(yf);
                    const float d0 = xf.x - yf.x, d1 = xf.y - yf.y;
                    const float l0 = times_sign(gl[0], d0), l1 = times_sign(gl[1], d1);      // g * sign(x - y), as two selects (no branch)
                    const P2 po = p2sub(p2fma(p2add(yv, yv), bs.b, p2fma(xv, bs.c, bs.a)), p2(l0, l1));
                    // (phase-major: stored after the last channel's loads -- a shared-memory store in between would pin the order
                    // of the next channel's loads behind this channel's whole dependency chain)
                    if (SNDE_SSIM_PHASED) po3[ch] = f2p(po);
                    else if (doC) sm.PO[rc & 1][ch][lane] = f2p(po);
                    if (GIMG && gown_base && doC) {
                        // d loss / dx(p) = box(c_sx) + 2 x box(c_sxx) + y box(c_sxy) + (1 - alpha) g sign(x - y)
                        const F2 pox = f2p(p2add(p2fma(p2add(xv, xv), bs.b, p2fma(yv, bs.c, p2add(xa1[ch].d, Hd))), p2(l0, l1)));
                        float *dst = gown_base + (ch * gown_sc + rc * gown_sh + T.ix0 * gown_sw);
                        if (T.colI[0]) atomic_add_float(dst, pox.x);
                        if (T.colI[1]) atomic_add_float(dst + gown_sw, pox.y);
                    }
                }
                cs_add(c2[ch], Hc, c1[ch]);   // c1 <- Hc(t-2) + Hc(t-1)
                c2[ch] = Hc;
                if (GIMG) { xa1[ch].d = p2add(xa2[ch].d, Hd); xa2[ch].d = Hd; }
            };
#if SNDE_SSIM_PHASED
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) phase_load(ch);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) phase_a(ch);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) phase_b(ch);
            if (GRAD) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) phase_c1(ch);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) phase_c2(ch);
                if ((SNDE_SSIM_STRAIGHT || doC) && doC) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) sm.PO[rc & 1][ch][lane] = po3[ch];
                }
            }
#else
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                phase_load(ch); phase_a(ch); phase_b(ch);
                if (GRAD) { phase_c1(ch); phase_c2(ch); }
            }
#endif
            }
            // ---- owner row t+3 starts its way into the ring (its slot held row t-5); row t+2 must have landed before the
            //      next barrier publishes it ----
            rows.prefetch_row(sm, T, lane, t + 3);
            row_copy_wait<1>();
            if (disp_in) rows.disparity_to_depth(sm, T, lane, t + 2);
        };
        for (int t = T.t_begin; t <= T.t_end + 1; ++t) row_step(t);
        row_copy_wait<0>();
        pair_barrier(who, bar);                  // the geometry warp is through with the ring: the next pass / task may refill it
        // ---- publish this pass: error sum of pair n ----
        {
            float v[1] = {err_acc};
            warp_sum<1>(who, v);
            if (T.n_others > 0 && lane == 0 && v[0] != 0.f) atomic_add_double(sc.loss + (T.clsF ? 0 : 1), (double)v[0]);
        }
    }
}

// ===============================================================================================================
// geometry warp
// ===============================================================================================================
template <bool GRAD, bool GRADK, bool EXTRA, bool FAST, bool DISP, bool GIMG = false>
struct GeoLane {
    const DevLoss &p;
    const DevScale &sc;
    PairSmem &sm;
    const StripTask &T;
    int lane;
    float xt[2];
    P2 xtp;                   // (xt[0], xt[1])
    float mxs[2], mys[2], mxi[2], myi[2];      // 0/1: the horizontal / vertical smoothness pair starting at column q exists (.. and is an inner column)
    // owner
    float *gd_base; int gd_sh, gd_sw;
    // current other image
    const float *oth_base; int oth_sc, oth_sh, oth_sw;
    float *wout_base; int wout_sc, wout_sh, wout_sw;
    float *gim_base; int gim_sc, gim_sh, gim_sw;      // gradient image of the OTHER image (GIMG: the scatter part)
    float P[12], Ki[9];       // projection and inverse intrinsics of the current pass
    // accumulators
    float dP[12], dK[9], sx_acc, sy_acc;
    F2 sy_prev;               // signed smoothness weights of the vertical pairs starting one row above
    unsigned vbits;           // SNDE_LERP_BLEND: validity of the lane's two pixels, two bits per issued row, newest lowest

    SNDE_STRIP_FN GeoLane(const DevLoss &p_, const DevScale &sc_, PairSmem &sm_, const StripTask &T_, int lane_)
        : p(p_), sc(sc_), sm(sm_), T(T_), lane(lane_) {}

    // ---- phase A, first half: reproject the lane's two pixels of a row (in the image) and start the 24 gathers;
    //      the loads land in registers that nothing touches until the next step (software prefetch) ----
    // The two pixels run as packed fp32 pairs (lane-wise the scalar chain of reproject_fast / make_taps_valid).
    // wts = {wx0, wx1, wy0, wy1}, each the pair over the two pixels (all zero for a pixel that samples nothing); border
    // mode also needs the derivative gates, kept in dgate = {dx, dy}.
    SNDE_STRIP_FN void a_issue(float yt, const F2 d, P2 (&wts)[4], P2 (&dgate)[2], float (&taps)[24], bool row_ok = true) {
        // (lanes whose columns lie outside the image go through the same motions with zero weights on pixel 0)
        const float *ob = oth_base;
        SNDE_PIN_PTR(ob);
        const int h = T.h, w = T.w;
        ReprojP rp; reproject_fast_p(Ki, P, xtp, yt, p2f(d), rp);
        P2 xp, yp; flow_coords_p(rp, xtp, yt, xp, yp);
        const float xq[2] = {p2lo(xp), p2hi(xp)}, yq[2] = {p2lo(yp), p2hi(yp)};
        int o00[2], dyo[2], dxo[2];
        bool valid[2];
        float wq[2][4], gq[2][2];
        if (EXTRA && p.padding) {                  // border mode samples everywhere; the clamps are per pixel
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                Taps tp; make_taps(xq[q], yq[q], h, w, p.padding, tp);
                valid[q] = T.in[q] && row_ok;
                o00[q] = valid[q] ? tp.y0 * oth_sh + tp.x0 * (FAST ? 1 : oth_sw) : 0;
                // border mode may have clamped both taps onto the same pixel: step 0
                dyo[q] = (tp.y1 - tp.y0) * oth_sh; dxo[q] = (tp.x1 - tp.x0) * (FAST ? 1 : oth_sw);
                wq[q][0] = tp.wx0; wq[q][1] = tp.wx1; wq[q][2] = tp.wy0; wq[q][3] = tp.wy1;
                gq[q][0] = tp.dx; gq[q][1] = tp.dy;
            }
        } else {
            // Outside [0,w-1) x [0,h-1) the reference's clamped-corner weights cancel to zero (up to rounding
            // residue that grows with the distance from the image); the fused path writes the exact zero.
            // Inside, none of the clamps is active: x0 = floor(x), weights (x0 + 1) - x and x - x0 (make_taps_valid).
            const float x0f[2] = {floorf(xq[0]), floorf(xq[1])}, y0f[2] = {floorf(yq[0]), floorf(yq[1])};
            const P2 x0p = p2(x0f[0], x0f[1]), y0p = p2(y0f[0], y0f[1]);
            const P2 wx0 = p2sub(p2add(x0p, p2s(1.f)), xp), wx1 = p2sub(xp, x0p);
            const P2 wy0 = p2sub(p2add(y0p, p2s(1.f)), yp), wy1 = p2sub(yp, y0p);
            wq[0][0] = p2lo(wx0); wq[1][0] = p2hi(wx0); wq[0][1] = p2lo(wx1); wq[1][1] = p2hi(wx1);
            wq[0][2] = p2lo(wy0); wq[1][2] = p2hi(wy0); wq[0][3] = p2lo(wy1); wq[1][3] = p2hi(wy1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                valid[q] = tap_valid(xq[q], yq[q], h, w, 0) != 0.f && T.in[q] && row_ok;
                // an invalid pixel reads the 2x2 block at the image origin (h, w >= 2 is checked on the host)
                o00[q] = valid[q] ? (int)y0f[q] * oth_sh + (int)x0f[q] * (FAST ? 1 : oth_sw) : 0;
                // row / column steps to the second tap: constants, so that the four taps of a channel are two base
                // addresses with immediate offsets
                dyo[q] = oth_sh; dxo[q] = FAST ? 1 : oth_sw;
                gq[q][0] = gq[q][1] = 1.f;
            }
        }
        if (SNDE_LERP_BLEND && !EXTRA) {
            // wts = {wx1, wy1, valid as 0/1}: the fractional parts as they are, validity as one factor
            wts[0] = p2(wq[0][1], wq[1][1]); wts[1] = p2(wq[0][3], wq[1][3]);
            wts[2] = p2(valid[0] ? 1.f : 0.f, valid[1] ? 1.f : 0.f);
            vbits = (vbits >> 2) | (valid[0] ? 0x100u : 0u) | (valid[1] ? 0x200u : 0u);       // four pushes later the row's bits are bits 0, 1
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) wts[i] = p2(valid[0] ? wq[0][i] : 0.f, valid[1] ? wq[1][i] : 0.f);
            dgate[0] = p2(gq[0][0], gq[1][0]); dgate[1] = p2(gq[0][1], gq[1][1]);
        }
        // Branch-free per pixel: an invalid pixel samples address 0 of the image with four zero weights.  The two tap rows
        // are walked channel by channel as pinned pointers (one 64-bit add per row and channel).
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#if SNDE_GATHER_WALK
            const float *q0 = ob + o00[q];
            const float *q1 = q0 + dyo[q];
#endif
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
#if SNDE_GATHER_WALK
                SNDE_PIN_PTR(q0); SNDE_PIN_PTR(q1);
#else
                const float *q0 = ob + (o00[q] + ch * oth_sc);
                const float *q1 = q0 + dyo[q];
#endif
                taps[12 * q + 4 * ch + 0] = ldg(q0);             // (y0, x0)
                taps[12 * q + 4 * ch + 1] = ldg(q1);             // (y1, x0)
                taps[12 * q + 4 * ch + 2] = ldg(q0 + dxo[q]);    // (y0, x1)
                taps[12 * q + 4 * ch + 3] = ldg(q1 + dxo[q]);    // (y1, x1)
#if SNDE_GATHER_WALK
                q0 += oth_sc; q1 += oth_sc;
#endif
            }
        }
    }

    // ---- phase A, second half (one step later): blend the gathered taps -> warped row (+ tap derivatives) ----
    SNDE_STRIP_FN void a_finish(int iy, const P2 (&wts)[4], const P2 (&dgate)[2], const float (&taps)[24], bool want_g,
                                F2 (&y)[3], F2 (&g)[6]) const {
        if (SNDE_LERP_BLEND && !EXTRA) {
            // top = i00 + wx1 (i10 - i00), bot = i01 + wx1 (i11 - i01): the two rows interpolated along x; their difference is
            // d/dy, the warped value is top + wy1 (bot - top), d/dx is the x differences interpolated along y.  The same
            // numbers as the four-weight form up to rounding (wx0 + wx1 = wy0 + wy1 = 1 exactly for an in-image sample).
            const P2 wx1 = wts[0], wy1 = wts[1], vm = wts[2];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const P2 i00 = p2(taps[4 * ch + 0], taps[12 + 4 * ch + 0]), i01 = p2(taps[4 * ch + 1], taps[12 + 4 * ch + 1]);
                const P2 i10 = p2(taps[4 * ch + 2], taps[12 + 4 * ch + 2]), i11 = p2(taps[4 * ch + 3], taps[12 + 4 * ch + 3]);
                const P2 a = p2sub(i10, i00), b = p2sub(i11, i01);
                const P2 top = p2fma(wx1, a, i00), bot = p2fma(wx1, b, i01);
                const P2 gy = p2sub(bot, top);
                y[ch] = f2p(p2mul(p2fma(wy1, gy, top), vm));
                // (formed on every row -- gy is needed for the value anyway -- and stored by the caller on inner rows only)
                if (GRAD) { g[2 * ch] = f2p(p2fma(wy1, p2sub(b, a), a)); g[2 * ch + 1] = f2p(gy); }
                else g[2 * ch] = g[2 * ch + 1] = F2{0.f, 0.f};
            }
            return;
        }
        const P2 wx0 = wts[0], wx1 = wts[1], wy0 = wts[2], wy1 = wts[3];
        const P2 w00 = p2mul(wx0, wy0), w01 = p2mul(wx0, wy1), w10 = p2mul(wx1, wy0), w11 = p2mul(wx1, wy1);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const P2 i00 = p2(taps[4 * ch + 0], taps[12 + 4 * ch + 0]), i01 = p2(taps[4 * ch + 1], taps[12 + 4 * ch + 1]);
            const P2 i10 = p2(taps[4 * ch + 2], taps[12 + 4 * ch + 2]), i11 = p2(taps[4 * ch + 3], taps[12 + 4 * ch + 3]);
            const P2 yv = p2fma(w11, i11, p2fma(w10, i10, p2fma(w01, i01, p2mul(w00, i00))));      // tap_blend_fma
            y[ch] = f2p(yv);
            g[2 * ch] = g[2 * ch + 1] = F2{0.f, 0.f};         // (dead unless want_g: the caller does not store them then)
            if (GRAD && want_g) {
                P2 gx = p2fma(wy1, p2sub(i11, i01), p2mul(wy0, p2sub(i10, i00)));
                P2 gy = p2fma(wx1, p2sub(i11, i10), p2mul(wx0, p2sub(i01, i00)));
                if (EXTRA) { gx = p2mul(gx, dgate[0]); gy = p2mul(gy, dgate[1]); }
                g[2 * ch] = f2p(gx); g[2 * ch + 1] = f2p(gy);
            }
        }
        if (EXTRA && wout_base && T.do_photo && iy >= T.y0 && iy < T.y1) {
            const float yv[2][3] = {{y[0].x, y[1].x, y[2].x}, {y[0].y, y[1].y, y[2].y}};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (!T.colI[q]) continue;
                float *o = wout_base + iy * wout_sh + (T.ix0 + q) * wout_sw;
                o[0] = yv[q][0]; o[wout_sc] = yv[q][1]; o[2 * wout_sc] = yv[q][2];
            }
        }
    }

    // ---- row r = t-3 is finished: phase G, smoothness, depth-gradient write ----
    // SMOOTH: this pass also does the owner's smoothness term (the first pass over an owner)
    template <bool SMOOTH>
    SNDE_STRIP_FN void finish_row(int r, int k, int who, bool smooth_now = true) {
        const int y0 = T.y0, ix0 = T.ix0;
        const bool doG = r >= y0;                  // an inner row (r <= y1-1 by construction of the loop)
        float gd[2] = {0.f, 0.f};
        // Later passes over the same owner add to what this lane wrote in the earlier ones -- by a no-return add, because a
        // load here would share the gathers' scoreboard and its use would wait for all of them (measured: 7 % of all stall
        // samples).  Only the pass that turns the finished sum into a disparity gradient has to read it back (from L2:
        // the adds of the passes in between do not go through L1); that read is issued first thing in the row step.
        const bool to_disp = (DISP || (EXTRA && (p.flags & FLAG_DISP))) && k == T.n_pass - 1;
        F2 gd_old = F2{0.f, 0.f};
        if (FAST && GRAD && k > 0 && to_disp && doG && T.colI[0]) gd_old = ldcg2(gd_base + (r * gd_sh + ix0));
        const bool g_on = GRAD && T.do_photo && doG;                          // (uniform over the warp)
        // (SNDE_LERP_BLEND: the tap derivatives of an invalid pixel are not zero by themselves; its validity bit switches it off)
        const bool lerp = SNDE_LERP_BLEND && !EXTRA;
        const bool on[2] = {g_on && T.colI[0] && (!lerp || (vbits & 1u)), g_on && (FAST ? T.colI[0] : T.colI[1]) && (!lerp || (vbits & 2u))};
        if (GRAD && (SNDE_GEO_STRAIGHT || on[0] || on[1])) {
            // both pixels as packed pairs (lane-wise reproject_nodiv / reproject_bwd_fast); a FAST lane holds two inner
            // pixels or none, a generic one may hold a single inner pixel.  SNDE_GEO_STRAIGHT: no branch around this -- an
            // idle pixel (halo lane, row outside the segment, whatever the rings hold there) runs with zero depth, zero
            // upstream gradient and a zero reciprocal and adds exact zeros to every sum.
            const float yt2 = g_on ? sm.YT[r & 7] : 0.f;
            const F2 d2f = sm.D[r & 7][lane];
            const P2 d2 = p2(on[0] ? d2f.x : 0.f, on[1] ? d2f.y : 0.f);
            P2 g[6], po[3];
#pragma unroll
            for (int i = 0; i < 6; ++i) g[i] = p2f(sm.G[r & 3][i][lane]);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) po[ch] = p2f(sm.PO[r & 1][ch][lane]);
            const P2 gxa = p2fma(po[2], g[4], p2fma(po[1], g[2], p2mul(po[0], g[0])));
            const P2 gya = p2fma(po[2], g[5], p2fma(po[1], g[3], p2mul(po[0], g[1])));
            const P2 gx = p2(on[0] ? p2lo(gxa) : 0.f, on[1] ? p2hi(gxa) : 0.f);
            const P2 gy = p2(on[0] ? p2lo(gya) : 0.f, on[1] ? p2hi(gya) : 0.f);
            ReprojP rp; reproject_nodiv_p(Ki, P, xtp, yt2, d2, rp);
            P2 gu[3], gcam[3];
            const P2 gdp = reproject_bwd_fast_p(P, rp, gx, gy, gu, gcam, on[0], on[1]);
            const float gdq[2] = {p2lo(gdp), p2hi(gdp)};
            if (GIMG && gim_base) {
                // ---- image-gradient scatter of row r (index_select backward, utils_edited.py:473-476): the sample position
                //      again (with the division), its four taps and weights, dL/d(warped) times the weight to every tap.  Lanes
                //      whose taps land on the same pixel are grouped (match.any): one atomic per distinct pixel and channel. ----
                ReprojP rq; reproject_fast_p(Ki, P, xtp, yt2, d2, rq);
                P2 xsp, ysp; flow_coords_p(rq, xtp, yt2, xsp, ysp);
                const float xs[2] = {p2lo(xsp), p2hi(xsp)}, ys[2] = {p2lo(ysp), p2hi(ysp)};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    Taps tp; make_taps(xs[q], ys[q], T.h, T.w, (EXTRA && p.padding) ? 1 : 0, tp);
                    const bool live = on[q] && ((EXTRA && p.padding) || tap_valid(xs[q], ys[q], T.h, T.w, 0) != 0.f);
                    const float wk[4] = {tp.wx0 * tp.wy0, tp.wx0 * tp.wy1, tp.wx1 * tp.wy0, tp.wx1 * tp.wy1};
                    const int ok[4] = {tp.y0 * gim_sh + tp.x0 * gim_sw, tp.y1 * gim_sh + tp.x0 * gim_sw,
                                       tp.y0 * gim_sh + tp.x1 * gim_sw, tp.y1 * gim_sh + tp.x1 * gim_sw};
                    const float pq[3] = {q ? p2hi(po[0]) : p2lo(po[0]), q ? p2hi(po[1]) : p2lo(po[1]), q ? p2hi(po[2]) : p2lo(po[2])};
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const TapGroup grp = tap_group(lane, live ? (long long)ok[k4] : -(long long)(1 + lane * 8 + q * 4 + k4));
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float sum = group_sum(lane, grp, live ? wk[k4] * pq[ch] : 0.f);
                            if (grp.leader && live) atomic_add_float(gim_base + (ch * gim_sc + ok[k4]), sum);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                gd[q] = gdq[q];               // (exact zero for an idle pixel)
                float guq[3], camq[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) { guq[j] = q ? p2hi(gu[j]) : p2lo(gu[j]); camq[j] = q ? p2hi(rp.cam[j]) : p2lo(rp.cam[j]); }
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    dP[j * 4 + 0] = fmaf(guq[j], camq[0], dP[j * 4 + 0]);
                    dP[j * 4 + 1] = fmaf(guq[j], camq[1], dP[j * 4 + 1]);
                    dP[j * 4 + 2] = fmaf(guq[j], camq[2], dP[j * 4 + 2]);
                    dP[j * 4 + 3] += guq[j];
                }
                if (GRADK) {
                    const float dq = q ? p2hi(d2) : p2lo(d2);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float gray = (q ? p2hi(gcam[j]) : p2lo(gcam[j])) * dq;
                        dK[j * 3 + 0] = fmaf(gray, xt[q], dK[j * 3 + 0]);
                        dK[j * 3 + 1] = fmaf(gray, yt2, dK[j * 3 + 1]);
                        dK[j * 3 + 2] += gray;
                    }
                }
            }
        }
        // smoothness of row r (first pass only).  One step earlier (row y0-1) only the vertical pair weights are
        // formed: they are the "pair above" of the first inner row.
        // Straight-line like the back-projection above: the loads and the arithmetic run on every row, a row that does
        // not exist (outside the image / before the segment) contributes selected zeros.
        if (SMOOTH && (SNDE_SMOOTH_ALWAYS || smooth_now)) {
            const int h = T.h;
            // (SNDE_SMOOTH_ALWAYS: no branch around the term -- on the later passes over an owner it runs with every pair switched
            // off, so that it shares a basic block, and the scheduler's attention, with the back-projection above)
            const bool r_in = (!SNDE_SMOOTH_ALWAYS || smooth_now) && r >= y0 - 1 && r >= 0 && r < h, below = r_in && r + 1 < h;
            F2 xa[3], xb[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { xa[ch] = sm.X[r & 7][ch][lane]; xb[ch] = sm.X[(r + 1) & 7][ch][lane]; }
            const F2 da = sm.D[r & 7][lane], db = sm.D[(r + 1) & 7][lane];
            float nb_x[4];                                  // right neighbour column: image (3) + depth
            {
                const float to_left[4] = {xa[0].x, xa[1].x, xa[2].x, da.x};
                take_from_right<4>(who, to_left, nb_x);
            }
            const float xq[2][3] = {{xa[0].x, xa[1].x, xa[2].x}, {xa[0].y, xa[1].y, xa[2].y}};
            const float xn[2][3] = {{xa[0].y, xa[1].y, xa[2].y}, {nb_x[0], nb_x[1], nb_x[2]}};
            const float dq[2] = {da.x, da.y}, dn[2] = {da.y, nb_x[3]};
            const float xd[2][3] = {{xb[0].x, xb[1].x, xb[2].x}, {xb[0].y, xb[1].y, xb[2].y}};
            const float dd[2] = {db.x, db.y};
            float vxs[2], vys[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                // pairs that do not exist (image border, halo column) are switched off by 0/1 factors fixed per task
                {
                    const float wgt = edge_weight3_fast(xq[q][0], xq[q][1], xq[q][2], xn[q][0], xn[q][1], xn[q][2]);
                    const float v = (dq[q] - dn[q]) * wgt;
                    const float sgn = v > 0.f ? wgt : (v < 0.f ? -wgt : 0.f);
                    vxs[q] = r_in ? sgn * mxs[q] : 0.f;
                    if (doG && r_in) sx_acc = fmaf(fabsf(v), mxi[q], sx_acc);
                }
                {
                    const float wgt = edge_weight3_fast(xq[q][0], xq[q][1], xq[q][2], xd[q][0], xd[q][1], xd[q][2]);
                    const float v = (dq[q] - dd[q]) * wgt;
                    const float sgn = v > 0.f ? wgt : (v < 0.f ? -wgt : 0.f);
                    vys[q] = below ? sgn * mys[q] : 0.f;
                    if (doG && below) sy_acc = fmaf(fabsf(v), myi[q], sy_acc);
                }
            }
            const F2 sxw = F2{vxs[0], vxs[1]}, syw = F2{vys[0], vys[1]};
            // the horizontal pair that ENDS at the lane's left column starts in the left neighbour lane
            float sx_left;
            {
                const float to_right[1] = {sxw.y};
                float from_left[1];
                take_from_left<1>(who, to_right, from_left);
                sx_left = from_left[0];
            }
            if (GRAD) {                                     // (stored for inner rows only)
                gd[0] += sc.gsx * (sxw.x - sx_left) + sc.gsy * (syw.x - sy_prev.x);
                gd[1] += sc.gsx * (sxw.y - sxw.x) + sc.gsy * (syw.y - sy_prev.y);
            }
            sy_prev = syw;
        }
        if (GRAD && doG && T.in[0]) {
            // Every lane writes only its own inner pixels; later passes over the same owner add to what this lane wrote.
            // Disparity input: the last pass turns the finished sum into d/d(disparity) = -(g * depth^2), autograd's
            // reciprocal backward -grad * (result * result).
            // (DISP: the lean disparity instance; the general EXTRA instance looks at the flag)
            F2 dd = F2{0.f, 0.f};
            if (DISP || EXTRA) dd = sm.D[r & 7][lane];
            if (FAST) {
                if (T.colI[0]) {     // inner lanes hold two inner pixels (w even)
                    float *dst = gd_base + (r * gd_sh + ix0);
                    F2 o = F2{gd[0], gd[1]};
                    if (k > 0 && !to_disp) add_pair_noreturn(dst, o);
                    else {
                        if (k > 0) { o.x += gd_old.x; o.y += gd_old.y; }
                        if (to_disp) { o.x = SNDE_XMUL(-o.x, SNDE_XMUL(dd.x, dd.x)); o.y = SNDE_XMUL(-o.y, SNDE_XMUL(dd.y, dd.y)); }
                        *reinterpret_cast<F2 *>(dst) = o;
                    }
                }
            } else {
                const float ddq[2] = {dd.x, dd.y};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (!T.colI[q]) continue;
                    float *dst = gd_base + (r * gd_sh + (ix0 + q) * gd_sw);
                    float o = k > 0 ? *dst + gd[q] : gd[q];
                    if (to_disp) o = SNDE_XMUL(-o, SNDE_XMUL(ddq[q], ddq[q]));
                    *dst = o;
                }
            }
        }
    }
};

// GRAD: also back-propagate.  GRADK: also accumulate d/dKinv.  EXTRA: honour the optional explainability masks,
// warped / error-map outputs and border padding.  FAST: planar tensors with unit x-stride and 8-byte aligned rows
// (64-bit loads / stores of the lane's pixel pair).  DISP: the depth maps hold disparities (lean instance; the EXTRA
// instance handles the flag at run time).
template <bool GRAD, bool GRADK, bool EXTRA, bool FAST, bool DISP, bool GIMG = false>
SNDE_STRIP_FN void geo_role(const DevLoss &p, int task, int who, PairSmem &sm, int bar) {
    const int lane = who & 31;
    StripTask T;
    decode_task(p, task, lane, T);
    const DevScale &sc = p.sc[T.scale];
    const int B = p.B;
    GeoLane<GRAD, GRADK, EXTRA, FAST, DISP, GIMG> L(p, sc, sm, T, lane);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        L.xt[q] = T.in[q] ? ldg(sc.xtab + (T.ix0 + q)) : 0.f;
        L.mxs[q] = (T.colB[q] && T.ix0 + q + 1 < T.w) ? 1.f : 0.f; L.mys[q] = T.colB[q] ? 1.f : 0.f;
        L.mxi[q] = L.mxs[q] * T.mI[q]; L.myi[q] = L.mys[q] * T.mI[q];
    }
    L.xtp = p2(L.xt[0], L.xt[1]);
    const bool clsF = T.clsF;
    const int m = T.m, b = T.b;
    const long long depth_n = clsF ? m : (long long)B + m;
    L.gd_base = GRAD ? sc.gdepth.p + depth_n * sc.gdepth.sn : nullptr; L.gd_sh = sc.gdepth.sh; L.gd_sw = sc.gdepth.sw;
    SNDE_PIN_PTR(L.gd_base);
    L.sx_acc = 0.f; L.sy_acc = 0.f; L.vbits = 0u;
#pragma unroll
    for (int i = 0; i < 9; ++i) L.dK[i] = 0.f;
    const float *Kinv = sc.Kinv + 9 * b;
    const int t_begin = T.t_begin, t_end = T.t_end, h = T.h;

    for (int k = 0; k < T.n_pass; ++k) {
        const long long n = clsF ? (long long)k * B + b : m;      // index into the N-sized tensors
        const DevImage &oth = clsF ? sc.src : sc.tgt;
        L.oth_base = oth.p + (clsF ? n : (long long)b) * oth.sn; L.oth_sc = oth.sc; L.oth_sh = oth.sh; L.oth_sw = oth.sw;
        L.wout_base = nullptr; L.wout_sc = L.wout_sh = L.wout_sw = 0;
        L.gim_base = nullptr; L.gim_sc = L.gim_sh = L.gim_sw = 0;
        if (GIMG) {
            const DevImageOut &go = clsF ? sc.gimg_src : sc.gimg_tgt;
            L.gim_base = go.p ? go.p + (clsF ? n : (long long)b) * go.sn : nullptr; L.gim_sc = go.sc; L.gim_sh = go.sh; L.gim_sw = go.sw;
        }
        if (EXTRA) {
            const DevImageOut &wout = clsF ? sc.warped_fwd : sc.warped_bwd;
            L.wout_base = wout.p ? wout.p + n * wout.sn : nullptr; L.wout_sc = wout.sc; L.wout_sh = wout.sh; L.wout_sw = wout.sw;
        }
        if (T.n_others > 0) {
            const float *Pn = sc.P + (n * 2 + (clsF ? 0 : 1)) * 12;
#pragma unroll
            for (int i = 0; i < 12; ++i) L.P[i] = ldg(Pn + i);
#pragma unroll
            for (int i = 0; i < 9; ++i) L.Ki[i] = ldg(Kinv + i);
            // Touch every entry here: the loads share a scoreboard with the tap gathers, and the compiler waits for that
            // scoreboard at the first use of a loaded value on every path.  Left to the row loop, that first use sits
            // right behind the gathers of the next row and makes the warp wait for all of them (measured: 10 % of all
            // stall samples on one instruction).
            // (an impossible store keeps the touching sum alive through ptxas)
            float touch = L.xt[0] + L.xt[1];
#pragma unroll
            for (int i = 0; i < 12; ++i) touch += L.P[i];
#pragma unroll
            for (int i = 0; i < 9; ++i) touch += L.Ki[i];
            if (touch == 1.2345e38f) sm.pad_[0] = 1;
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) L.P[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 9; ++i) L.Ki[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) L.dP[i] = 0.f;
        L.sy_prev = F2{0.f, 0.f};

        // prologue: owner rows t_begin .. t_begin+2 on their way into the ring; the gathers of row t_begin started
        P2 wts[4], dgate[2]; float taps[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) taps[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) wts[i] = p2s(0.f);
        dgate[0] = dgate[1] = p2s(0.f);
        pair_barrier(who, bar);           // the SSIM warp has brought owner rows t_begin, t_begin+1 into the ring
        if (T.do_photo && t_begin < h) L.a_issue(sm.YT[t_begin & 7], sm.D[t_begin & 7][lane], wts, dgate, taps);

        auto row_loop = [&](auto smooth_tag) {
        constexpr bool SMOOTH = decltype(smooth_tag)::value;
        for (int t = t_begin; t <= t_end + 1; ++t) {
            // ---- finish row t: the warped row goes to the SSIM warp, the tap derivatives stay here ----
#if SNDE_GEO_STRAIGHT
            {
                // (unconditional: a row whose gathers were never started has four zero weights -> a zero row, which is what
                // the pools want below the image; past t_end the row is written and never read)
                const bool want_g = t >= T.y0 && t < T.y1;
                F2 yr[3], gr[6];
                L.a_finish(t, wts, dgate, taps, want_g, yr, gr);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) sm.Y[t & 3][ch][lane] = yr[ch];
                if (GRAD && want_g) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) sm.G[t & 3][i][lane] = gr[i];
                }
            }
#else
            if (T.do_photo && t <= t_end) {
                const bool want_g = t >= T.y0 && t < T.y1;
                if (t < h) {
                    F2 yr[3], gr[6];
                    L.a_finish(t, wts, dgate, taps, want_g, yr, gr);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) sm.Y[t & 3][ch][lane] = yr[ch];
                    if (GRAD && want_g) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) sm.G[t & 3][i][lane] = gr[i];
                    }
                } else {                      // rows below the image are zero rows of the pools (never inner rows)
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) sm.Y[t & 3][ch][lane] = F2{0.f, 0.f};
                }
            }
#endif
            pair_barrier(who, bar);       // owner rows <= t+1 and warped rows <= t are published; dL/d(warped) of row t-3 is ready
#if SNDE_GEO_STRAIGHT
            // ---- start row t+1 (its gathers fly while the rest of this step runs) and take row t-3 back through the
            //      projection, as one basic block ----
            L.a_issue(sm.YT[(t + 1) & 7], sm.D[(t + 1) & 7][lane], wts, dgate, taps, T.do_photo && t + 1 <= t_end && t + 1 < h);
            L.template finish_row<SMOOTH>(t - 3, k, who, SNDE_SMOOTH_SPLIT || (T.do_smooth && k == 0));
#else
            // ---- start row t+1: its gathers fly while the rest of this step runs ----
            if (T.do_photo && t + 1 <= t_end && t + 1 < h) L.a_issue(sm.YT[(t + 1) & 7], sm.D[(t + 1) & 7][lane], wts, dgate, taps);
            // ---- row t-3: back-projection, depth gradient ----
            if (t - 3 >= T.y0 - 1) L.template finish_row<SMOOTH>(t - 3, k, who, SNDE_SMOOTH_SPLIT || (T.do_smooth && k == 0));
#endif
        }
        };
        // (SNDE_SMOOTH_SPLIT: two copies of the row loop -- the smoothness term belongs to the first pass over an owner only)
#if SNDE_SMOOTH_SPLIT
        if (T.do_smooth && k == 0) row_loop(SmoothOn{}); else row_loop(SmoothOff{});
#else
        row_loop(SmoothOn{});
#endif
        pair_barrier(who, bar);           // through with the ring (see the SSIM warp)

        // ---- publish this pass: d/dP of pair n (the reduction runs unconditionally: warp-collective code stays uniform) ----
        {
            const float mine = warp_sum_scatter<12>(who, L.dP);
            const int slot = warp_sum_slot(lane);
            if (GRAD && T.n_others > 0 && !(lane & 1) && slot < 12 && mine != 0.f)
                atomic_add_double(sc.gP + (n * 2 + (clsF ? 0 : 1)) * 12 + slot, (double)mine);
        }
    }
    // ---- publish the per-owner sums: smoothness, d/dKinv ----
    {
        float v[11];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = L.dK[i];
        v[9] = L.sx_acc; v[10] = L.sy_acc;
        const float mine = warp_sum_scatter<11>(who, v);
        const int slot = warp_sum_slot(lane);
        if (!(lane & 1) && mine != 0.f) {
            if (GRADK && slot < 9) atomic_add_double(sc.gKinv + 9 * b + slot, (double)mine);
            if (slot == 9 || slot == 10) atomic_add_double(sc.loss + 2 + (slot - 9), (double)mine);
        }
    }
}

}  // namespace snde
