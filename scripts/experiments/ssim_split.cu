// EXPERIMENT (not shipped): the SSIM + L1 forward/backward of the fused loss as a kernel of its own -- one WARP per
// (pass, strip, segment, colour channel), no shared memory, no barrier, rows streamed straight from global memory.  Question:
// does a light (< 96 registers), barrier-free per-channel warp at 20+ warps per SM issue fast enough that a split pipeline
// (warp forward -> this kernel -> geometry backward, intermediates through HBM/L2) could beat the 247 us fused strip kernel?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I../../self-supervised-normal-depth-estimation_b200/csrc -I../../include \
//        -o ssim_split ssim_split.cu && ./ssim_split [B S H W]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "snde_strip.cuh"

using namespace snde;

struct SplitScale {
    const float *tgt, *src;      // planar (n, 3, h, w), contiguous
    const float *Y;              // warped images, (2N, 3, h, w)
    float *PO;                   // dL/d(warped), (2N, 3, h, w)
    double *loss;                // [2]: class F, class R
    int h, w, strips, segs, seg_h;
    long long task0;             // first warp task of the scale
    float gphoto;
};
struct SplitParams { int B, S, ns; float alpha, oma; long long total; SplitScale s[6]; };

#ifndef SPLIT_WARPS
#define SPLIT_WARPS 6
#endif
#ifndef SPLIT_MIN_CTAS
#define SPLIT_MIN_CTAS 4
#endif

__global__ void __launch_bounds__(SPLIT_WARPS * 32, SPLIT_MIN_CTAS) ssim_channel_kernel(const __grid_constant__ SplitParams p) {
    const int lane = threadIdx.x & 31, who = lane;
    const long long task = (long long)blockIdx.x * SPLIT_WARPS + (threadIdx.x >> 5);
    if (task >= p.total) return;
    int si = 0;
#pragma unroll
    for (int i = 1; i < 6; ++i) if (i < p.ns && task >= p.s[i].task0) si = i;
    const SplitScale &sc = p.s[si];
    const int N = p.B * p.S;
    long long local = task - sc.task0;
    const int ch = (int)(local % 3); local /= 3;
    const int strip = (int)(local % sc.strips); local /= sc.strips;
    const int seg = (int)(local % sc.segs); local /= sc.segs;
    const int pass = (int)local;                                  // [0, N): class F (n = pass), [N, 2N): class R (m = pass - N)
    const bool clsF = pass < N;
    const int h = sc.h, w = sc.w;
    const long long plane = (long long)h * w;
    const float *xbase = (clsF ? sc.tgt + (long long)(pass % p.B) * 3 * plane : sc.src + (long long)(pass - N) * 3 * plane) + ch * plane;
    const float *ybase = sc.Y + ((long long)pass * 3 + ch) * plane;
    float *pobase = sc.PO + ((long long)pass * 3 + ch) * plane;
    const int ix0 = strip * StripCfg::TW - 2 + 2 * lane;
    const bool in0 = ix0 >= 0 && ix0 < w;
    bool colB[2], colI[2]; float mI[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int c = 2 * lane + q;
        colB[q] = in0 && c >= 1 && c <= StripCfg::RW - 2;
        colI[q] = in0 && c >= 2 && c <= StripCfg::RW - 3;
        mI[q] = colI[q] ? 1.f : 0.f;
    }
    const int y0 = seg * sc.seg_h, y1 = y0 + sc.seg_h < h ? y0 + sc.seg_h : h;
    const int t_begin = y0 - 2 > 0 ? y0 - 2 : 0, t_end = y1 + 1;
    const float alpha = p.alpha, oma = p.oma, g1 = p.oma * sc.gphoto;
    float ngSm[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) ngSm[q] = colB[q] ? 0.5f * p.alpha * sc.gphoto : 0.f;

    // rows go global -> shared by cp.async (zero-filled outside the image), three rows ahead, into a ring of 8 rows per warp:
    // no register ever holds a value that has not arrived.  (Versions 1-3 of this experiment loaded into registers: the select
    // behind the load, then the register rotation, then ptxas' own copy of the loaded pair each waited for the load -- 50-60 % of
    // all stall samples on one instruction, 149-178 us.)
    __shared__ F2 ring_all[SPLIT_WARPS][8][2][32];
    F2 (*ring)[2][32] = ring_all[threadIdx.x >> 5];
    auto prefetch = [&](int iy) {
        const bool take = in0 && iy >= 0 && iy < h;
        const int off = take ? iy * w + ix0 : 0;
        row_copy_async<8>(&ring[iy & 7][0][lane], xbase + off, take);
        row_copy_async<8>(&ring[iy & 7][1][lane], ybase + off, take);
        row_copy_commit();
    };

    HS s1, s2; CS c1, c2;
    hs_zero(s1); hs_zero(s2); cs_zero(c1); cs_zero(c2);
    float err_acc = 0.f;
    prefetch(t_begin - 2); prefetch(t_begin - 1);                 // (rows t-1, t-2 of the first steps: zero rows or the rows above the segment)
    prefetch(t_begin); prefetch(t_begin + 1); prefetch(t_begin + 2);
    for (int t = t_begin; t <= t_end; ++t) {
        row_copy_wait<2>();
        const F2 xr = ring[t & 7][0][lane], yr = ring[t & 7][1][lane];
        const F2 x1 = ring[(t - 1) & 7][0][lane], yy1 = ring[(t - 1) & 7][1][lane];
        const F2 x2 = ring[(t - 2) & 7][0][lane], yy2 = ring[(t - 2) & 7][1][lane];
        prefetch(t + 3);
        const int rb = t - 1, rc = t - 2;
        const bool doB = t >= y0, doC = t >= y0 + 2;
        const bool rowB = doB && rb >= 0 && rb < h;
        const bool row_inner = rb >= y0 && rb < y1;
        HS H;
        {
            const float to_right[2] = {xr.y, yr.y}, to_left[2] = {xr.x, yr.x};
            float from_left[2], from_right[2];
            neighbor_exchange<2>(who, to_right, to_left, from_left, from_right);
            hsum_row(xr, yr, from_left[0], from_right[0], from_left[1], from_right[1], H);
        }
        CS co;
        {
            HS tot; hs_add(s1, H, tot);
            SsimP ss; ssim_sums_p(tot.x, tot.y, tot.q, tot.xy, ss);
            const P2 v = p2fma(p2s(-0.5f), ss.S, p2s(0.5f));
            const float vq[2] = {p2lo(v), p2hi(v)};
            const float vcq[2] = {clampf(vq[0], 0.f, 1.f), clampf(vq[1], 0.f, 1.f)};
            const float dq[2] = {x1.x - yy1.x, x1.y - yy1.y};
            float ngq[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float err = fmaf(alpha, vcq[q], oma * fabsf(dq[q]));
                if (rowB && row_inner) err_acc = fmaf(err, mI[q], err_acc);
                ngq[q] = (rowB && vq[q] == vcq[q]) ? ngSm[q] : 0.f;
            }
            ssim_sums_bwd_p(ss, tot.x, tot.y, p2(ngq[0], ngq[1]), co.a, co.b, co.c);
        }
        hs_add(s2, H, s1);
        s2 = H;
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
        {
            CS bs; cs_add(c1, Hc, bs);
            const P2 xv = p2f(x2), yv = p2f(yy2);
            const float d0 = x2.x - yy2.x, d1 = x2.y - yy2.y;
            const float l0 = d0 > 0.f ? g1 : (d0 < 0.f ? -g1 : 0.f), l1 = d1 > 0.f ? g1 : (d1 < 0.f ? -g1 : 0.f);
            const P2 po = p2sub(p2fma(p2add(yv, yv), bs.b, p2fma(xv, bs.c, bs.a)), p2(l0, l1));
            if (doC && colI[0]) *reinterpret_cast<F2 *>(pobase + (rc * w + ix0)) = f2p(po);
        }
        cs_add(c2, Hc, c1);
        c2 = Hc;
    }
    row_copy_wait<0>();
    {
        float v[1] = {err_acc};
        warp_sum<1>(who, v);
        if (lane == 0 && v[0] != 0.f) atomicAdd(sc.loss + (clsF ? 0 : 1), (double)v[0]);
    }
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 32, S = argc > 2 ? atoi(argv[2]) : 2, H = argc > 3 ? atoi(argv[3]) : 128, W = argc > 4 ? atoi(argv[4]) : 416;
    const int ns = 4, N = B * S, seg_h = argc > 5 ? atoi(argv[5]) : 64;
    SplitParams p{};
    p.B = B; p.S = S; p.ns = ns; p.alpha = 0.85f; p.oma = 0.15f;
    long long total = 0, px = 0;
    for (int s = 0; s < ns; ++s) {
        SplitScale &sc = p.s[s];
        sc.h = H >> s; sc.w = W >> s;
        const long long plane = (long long)sc.h * sc.w;
        std::vector<float> host((size_t)(2 * N) * 3 * plane);
        for (size_t i = 0; i < host.size(); ++i) host[i] = 0.5f + 0.4f * sinf(0.013f * (float)(i % 100003)) * cosf(0.7f * (float)(i % 911));
        float *tgt, *src, *Y, *PO; double *loss;
        cudaMalloc(&tgt, (size_t)B * 3 * plane * 4); cudaMalloc(&src, (size_t)N * 3 * plane * 4);
        cudaMalloc(&Y, host.size() * 4); cudaMalloc(&PO, host.size() * 4); cudaMalloc(&loss, 16);
        cudaMemcpy(tgt, host.data(), (size_t)B * 3 * plane * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(src, host.data() + 17, (size_t)N * 3 * plane * 4, cudaMemcpyHostToDevice);
        for (size_t i = 0; i < host.size(); ++i) host[i] = 0.5f + 0.4f * sinf(0.011f * (float)(i % 100019) + 0.3f);
        cudaMemcpy(Y, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(PO, 0, host.size() * 4); cudaMemset(loss, 0, 16);
        sc.tgt = tgt; sc.src = src; sc.Y = Y; sc.PO = PO; sc.loss = loss;
        sc.strips = (sc.w + StripCfg::TW - 1) / StripCfg::TW; sc.seg_h = seg_h < sc.h ? seg_h : sc.h; sc.segs = (sc.h + sc.seg_h - 1) / sc.seg_h;
        sc.task0 = total; sc.gphoto = 1.f / (float)(3 * plane * N);
        total += (long long)2 * N * sc.segs * sc.strips * 3;
        px += 2 * N * plane;
    }
    p.total = total;
    const int blocks = (int)((total + SPLIT_WARPS - 1) / SPLIT_WARPS);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, ssim_channel_kernel);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ssim_channel_kernel, SPLIT_WARPS * 32, 0);
    printf("B %d S %d %dx%d seg %d: %lld warp tasks, %d blocks of %d warps, %d regs, %d CTAs/SM (%d warps/SM), %lld warped px\n", B, S, H, W, seg_h, total, blocks,
           SPLIT_WARPS, fa.numRegs, occ, occ * SPLIT_WARPS, px);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) ssim_channel_kernel<<<blocks, SPLIT_WARPS * 32>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) ssim_channel_kernel<<<blocks, SPLIT_WARPS * 32>>>(p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double l[2]; cudaMemcpy(l, p.s[0].loss, 16, cudaMemcpyDeviceToHost);
    printf("ssim_channel_kernel: %.1f us per launch; scale-0 error sums %g %g (x%d launches)\n", ms * 1e3 / reps, l[0], l[1], reps + 5);
    return 0;
}
