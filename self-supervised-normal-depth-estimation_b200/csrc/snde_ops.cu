// snde_ops.cu -- reference-signature kernels of the C ABI (include/snde.h): pose matrices, compute_rigid_flow,
// bilinear_sampler / flow_warp, image_similarity / DSSIM, smooth_loss.  Each materialises its output like the
// reference op it replaces.  Built with -fmad=false: these follow the reference's arithmetic op by op.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "snde.h"
#include "snde_math.cuh"

using namespace snde;

// views with 64-bit strides for the reference-signature kernels (any layout, any size)
struct LImage { const float *p; long long sn, sc, sh, sw; };
struct LImageOut { float *p; long long sn, sc, sh, sw; };
struct LMap { const float *p; long long sn, sh, sw; };
struct LMapOut { float *p; long long sn, sh, sw; };
static inline LImage ldev(const SndeImage &v) { return LImage{v.ptr, v.sn, v.sc, v.sh, v.sw}; }
static inline LImageOut ldev(const SndeImageOut &v) { return LImageOut{v.ptr, v.sn, v.sc, v.sh, v.sw}; }
static inline LMap ldev(const SndeMap &v) { return LMap{v.ptr, v.sn, v.sh, v.sw}; }
static inline LMapOut ldev(const SndeMapOut &v) { return LMapOut{v.ptr, v.sn, v.sh, v.sw}; }
__device__ __forceinline__ float ld_img(const LImage &im, long long n, int c, int y, int x) {
    return __ldg(im.p + n * im.sn + c * im.sc + (long long)y * im.sh + (long long)x * im.sw);
}
__device__ __forceinline__ float ld_map(const LMap &m, long long n, int y, int x) {
    return __ldg(m.p + n * m.sn + (long long)y * m.sh + (long long)x * m.sw);
}

#define SNDE_CHECK_LAUNCH() do { if (cudaGetLastError() != cudaSuccess) return SNDE_ERR_CUDA; } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// =============================================================================================
// Reference-signature pieces
// =============================================================================================
__global__ void __launch_bounds__(128) pose_matrices_kernel(int B, const float *__restrict__ poses, long long stride,
                                                            const float *__restrict__ K, int reverse, float *__restrict__ P,
                                                            float *__restrict__ Kinv) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float vec[6], Km[9], T[12], Ti[12], Pm[12], Ki[9];
    for (int j = 0; j < 6; ++j) vec[j] = poses[b * stride + j];
    for (int j = 0; j < 9; ++j) Km[j] = K[b * 9 + j];
    pose_to_T(vec, T, Ti);
    k_times_T(Km, reverse ? Ti : T, Pm);
    for (int j = 0; j < 12; ++j) P[b * 12 + j] = Pm[j];
    if (!inverse3(Km, Ki)) for (int j = 0; j < 9; ++j) Ki[j] = nanf("");
    for (int j = 0; j < 9; ++j) Kinv[b * 9 + j] = Ki[j];
}

extern "C" int snde_pose_matrices(int B, const float *poses, int64_t pose_stride, const float *K, int reverse_pose,
                                  float *P, float *Kinv, void *stream) {
    if (!poses || !K || !P || !Kinv) return SNDE_ERR_NULL;
    if (B <= 0) return SNDE_ERR_SHAPE;
    pose_matrices_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(B, poses, pose_stride, K, reverse_pose, P, Kinv);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

__global__ void __launch_bounds__(128) pose_matrices_bwd_kernel(int B, const float *__restrict__ poses, long long stride,
                                                                const float *__restrict__ K, int reverse,
                                                                const double *__restrict__ gP, const double *__restrict__ gKinv,
                                                                float *__restrict__ grad_poses, float *__restrict__ grad_K) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float vec[6], Km[9];
    for (int j = 0; j < 6; ++j) vec[j] = poses[b * stride + j];
    for (int j = 0; j < 9; ++j) Km[j] = K[b * 9 + j];
    double gv[6] = {0, 0, 0, 0, 0, 0}, gK[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    pose_prep_bwd(vec, Km, reverse ? nullptr : gP + b * 12, reverse ? gP + b * 12 : nullptr,
                  (gKinv && grad_K) ? gKinv + b * 9 : nullptr, gv, grad_K ? gK : nullptr);
    for (int j = 0; j < 6; ++j) grad_poses[b * 6 + j] = (float)gv[j];
    if (grad_K) for (int j = 0; j < 9; ++j) grad_K[b * 9 + j] = (float)gK[j];
}

extern "C" int snde_pose_matrices_bwd(int B, const float *poses, int64_t pose_stride, const float *K, int reverse_pose,
                                      const double *gP, const double *gKinv, float *grad_poses, float *grad_K, void *stream) {
    if (!poses || !K || !gP || !grad_poses) return SNDE_ERR_NULL;
    if (B <= 0) return SNDE_ERR_SHAPE;
    pose_matrices_bwd_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(B, poses, pose_stride, K, reverse_pose, gP, gKinv,
                                                                            grad_poses, grad_K);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// ---- compute_rigid_flow, per-pixel part: grid (ceil(h*w/256), B) ----
__global__ void __launch_bounds__(256) rigid_flow_fwd_kernel(int h, int w, LMap depth, const float *__restrict__ Kinv,
                                                             const float *__restrict__ P, const float *__restrict__ xtab,
                                                             const float *__restrict__ ytab, float *__restrict__ flow) {
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= h * w) return;
    const int iy = pix / w, ix = pix - iy * w;
    float Ki[9], Pm[12];
    for (int j = 0; j < 9; ++j) Ki[j] = __ldg(Kinv + b * 9 + j);
    for (int j = 0; j < 12; ++j) Pm[j] = __ldg(P + b * 12 + j);
    const float xt = __ldg(xtab + ix), yt = __ldg(ytab + iy);
    Reproj rp; reproject(Ki, Pm, xt, yt, ld_map(depth, b, iy, ix), rp);
    float2 o; o.x = rp.xn - xt; o.y = rp.yn - yt;
    reinterpret_cast<float2 *>(flow)[(size_t)b * h * w + pix] = o;
}

extern "C" int snde_rigid_flow_fwd(int B, int h, int w, SndeMap depth, const float *Kinv, const float *P, const float *xtab,
                                   const float *ytab, float *flow, void *stream) {
    if (!depth.ptr || !Kinv || !P || !xtab || !ytab || !flow) return SNDE_ERR_NULL;
    if (B <= 0 || h <= 0 || w <= 0) return SNDE_ERR_SHAPE;
    if ((long long)B * h * w >= (1LL << 31) || B > 65535) return SNDE_ERR_TOO_LARGE;
    dim3 grid((h * w + 255) / 256, B);
    rigid_flow_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(h, w, ldev(depth), Kinv, P, xtab, ytab, flow);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

template <int NV>
__device__ __forceinline__ void block_atomic_add(float (&v)[NV], double *dst, float *red /*[8][NV]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp * NV + i] = x;
    }
    __syncthreads();
    if (tid < NV) {
        double s = 0.0;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) s += (double)red[wq * NV + tid];
        if (s != 0.0) atomicAdd(dst + tid, s);
    }
    __syncthreads();
}

template <bool GRADK>
__global__ void __launch_bounds__(256) rigid_flow_bwd_kernel(int h, int w, LMap depth, const float *__restrict__ Kinv,
                                                             const float *__restrict__ P, const float *__restrict__ xtab,
                                                             const float *__restrict__ ytab, const float *__restrict__ gflow,
                                                             LMapOut gdepth, double *__restrict__ gP, double *__restrict__ gKinv) {
    __shared__ float red[8 * 12];
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    float aP[12], aK[9];
#pragma unroll
    for (int j = 0; j < 12; ++j) aP[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) aK[j] = 0.f;
    if (pix < h * w) {
        const int iy = pix / w, ix = pix - iy * w;
        float Ki[9], Pm[12];
        for (int j = 0; j < 9; ++j) Ki[j] = __ldg(Kinv + b * 9 + j);
        for (int j = 0; j < 12; ++j) Pm[j] = __ldg(P + b * 12 + j);
        const float xt = __ldg(xtab + ix), yt = __ldg(ytab + iy), d = ld_map(depth, b, iy, ix);
        Reproj rp; reproject(Ki, Pm, xt, yt, d, rp);
        const float2 g = reinterpret_cast<const float2 *>(gflow)[(size_t)b * h * w + pix];
        float gu[3], gcam[3];
        const float gd = reproject_bwd(Pm, rp, g.x, g.y, gu, gcam);
        gdepth.p[b * gdepth.sn + (long long)iy * gdepth.sh + (long long)ix * gdepth.sw] = gd;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            aP[q * 4 + 0] = gu[q] * rp.cam[0]; aP[q * 4 + 1] = gu[q] * rp.cam[1]; aP[q * 4 + 2] = gu[q] * rp.cam[2]; aP[q * 4 + 3] = gu[q];
            if (GRADK) { const float gray = gcam[q] * d; aK[q * 3 + 0] = gray * xt; aK[q * 3 + 1] = gray * yt; aK[q * 3 + 2] = gray; }
        }
    }
    block_atomic_add<12>(aP, gP + b * 12, red);
    if (GRADK) block_atomic_add<9>(aK, gKinv + b * 9, red);
}

extern "C" int snde_rigid_flow_bwd(int B, int h, int w, SndeMap depth, const float *Kinv, const float *P, const float *xtab,
                                   const float *ytab, const float *gflow, SndeMapOut gdepth, double *gP, double *gKinv,
                                   void *stream) {
    if (!depth.ptr || !Kinv || !P || !xtab || !ytab || !gflow || !gdepth.ptr || !gP) return SNDE_ERR_NULL;
    if (B <= 0 || h <= 0 || w <= 0) return SNDE_ERR_SHAPE;
    if ((long long)B * h * w >= (1LL << 31) || B > 65535) return SNDE_ERR_TOO_LARGE;
    dim3 grid((h * w + 255) / 256, B);
    if (gKinv) rigid_flow_bwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(h, w, ldev(depth), Kinv, P, xtab, ytab, gflow, ldev(gdepth), gP, gKinv);
    else rigid_flow_bwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(h, w, ldev(depth), Kinv, P, xtab, ytab, gflow, ldev(gdepth), gP, gKinv);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// ---- bilinear_sampler / flow_warp ----
__global__ void __launch_bounds__(256) sampler_fwd_kernel(long long total, int C, int hs, int ws, LImage img, int ht, int wt,
                                                          const float *__restrict__ coords, int add_grid,
                                                          const float *__restrict__ xtab, const float *__restrict__ ytab, int border,
                                                          float *__restrict__ out, float *__restrict__ valid) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % wt); const long long t = i / wt; const int iy = (int)(t % ht); const long long n = t / ht;
        const float2 cf = reinterpret_cast<const float2 *>(coords)[i];
        float x = cf.x, y = cf.y;
        if (add_grid) { x = __ldg(xtab + ix) + x; y = __ldg(ytab + iy) + y; }
        Taps tp; make_taps(x, y, hs, ws, border, tp);
        const float *base = img.p + n * img.sn;
        const long long o00 = (long long)tp.y0 * img.sh + (long long)tp.x0 * img.sw, o01 = (long long)tp.y1 * img.sh + (long long)tp.x0 * img.sw;
        const long long o10 = (long long)tp.y0 * img.sh + (long long)tp.x1 * img.sw, o11 = (long long)tp.y1 * img.sh + (long long)tp.x1 * img.sw;
        for (int c = 0; c < C; ++c) {
            const float *bc = base + c * img.sc;
            out[i * C + c] = tap_blend(tp, __ldg(bc + o00), __ldg(bc + o01), __ldg(bc + o10), __ldg(bc + o11));
        }
        if (valid) valid[i] = tap_valid(x, y, hs, ws, border);
    }
}

extern "C" int snde_bilinear_sampler_fwd(int N, int C, int hs, int ws, SndeImage img, int ht, int wt, const float *coords,
                                         int add_grid, const float *xtab, const float *ytab, int padding_mode, float *out,
                                         float *valid, void *stream) {
    if (!img.ptr || !coords || !out || (add_grid && (!xtab || !ytab))) return SNDE_ERR_NULL;
    if (N <= 0 || C <= 0 || hs <= 0 || ws <= 0 || ht <= 0 || wt <= 0) return SNDE_ERR_SHAPE;
    if (padding_mode != SNDE_PAD_ZEROS && padding_mode != SNDE_PAD_BORDER) return SNDE_ERR_ARG;
    const long long total = (long long)N * ht * wt;
    if (total * C >= (1LL << 31) || (long long)N * hs * ws * C >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    sampler_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(total, C, hs, ws, ldev(img), ht, wt, coords, add_grid, xtab, ytab,
                                                             padding_mode, out, valid);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// Lanes of the active mask that share `key`, as seen by each lane: the peers above it, whether it is the lowest of its
// group, and the size of the largest group in the warp (the trip count of the reduction below; 1 = no collision at all).
struct WarpTapGroups { unsigned active, above; int max_group; bool leader; };
__device__ __forceinline__ WarpTapGroups warp_tap_groups(unsigned active, unsigned long long key) {
    const int lane = (int)(threadIdx.x & 31);
    const unsigned peers = __match_any_sync(active, key);
    WarpTapGroups g;
    g.active = active;
    g.above = lane == 31 ? 0u : peers & ~((2u << lane) - 1u);
    g.leader = (peers & ((1u << lane) - 1u)) == 0u;
    g.max_group = (int)__reduce_max_sync(active, (unsigned)__popc(peers));
    return g;
}
// sum of v over the lanes of this lane's group that lie above it, plus its own: the group's total in its leader
__device__ __forceinline__ float warp_group_sum(const WarpTapGroups &g, float v) {
    const int lane = (int)(threadIdx.x & 31);
    float sum = v;
    unsigned rest = g.above;
    for (int it = 1; it < g.max_group; ++it) {
        const int src = rest ? __ffs((int)rest) - 1 : lane;
        const float o = __shfl_sync(g.active, v, src);
        if (rest) { sum += o; rest &= rest - 1u; }
    }
    return sum;
}

__global__ void __launch_bounds__(256) sampler_bwd_kernel(long long total, int C, int hs, int ws, LImage img, int ht, int wt,
                                                          const float *__restrict__ coords, int add_grid,
                                                          const float *__restrict__ xtab, const float *__restrict__ ytab, int border,
                                                          const float *__restrict__ gout, float *__restrict__ gcoords, LImageOut gimg) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % wt); const long long t = i / wt; const int iy = (int)(t % ht); const long long n = t / ht;
        const float2 cf = reinterpret_cast<const float2 *>(coords)[i];
        float x = cf.x, y = cf.y;
        if (add_grid) { x = __ldg(xtab + ix) + x; y = __ldg(ytab + iy) + y; }
        Taps tp; make_taps(x, y, hs, ws, border, tp);
        const float *base = img.p + n * img.sn;
        const long long o00 = (long long)tp.y0 * img.sh + (long long)tp.x0 * img.sw, o01 = (long long)tp.y1 * img.sh + (long long)tp.x0 * img.sw;
        const long long o10 = (long long)tp.y0 * img.sh + (long long)tp.x1 * img.sw, o11 = (long long)tp.y1 * img.sh + (long long)tp.x1 * img.sw;
        const float w00 = tp.wx0 * tp.wy0, w01 = tp.wx0 * tp.wy1, w10 = tp.wx1 * tp.wy0, w11 = tp.wx1 * tp.wy1;
        float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
        // image-gradient scatter (index_select backward, utils_edited.py:473-476), WARP-AGGREGATED: lanes of the warp whose tap
        // lands on the same pixel are found with match.any on the pixel's address, the lowest of them adds the others'
        // contributions (in lane order) and issues ONE atomic per distinct address and channel.
        WarpTapGroups grp[4];
        long long goff[4] = {0, 0, 0, 0};
        if (gimg.p) {
            const unsigned active = __activemask();
            goff[0] = (long long)tp.y0 * gimg.sh + (long long)tp.x0 * gimg.sw; goff[1] = (long long)tp.y1 * gimg.sh + (long long)tp.x0 * gimg.sw;
            goff[2] = (long long)tp.y0 * gimg.sh + (long long)tp.x1 * gimg.sw; goff[3] = (long long)tp.y1 * gimg.sh + (long long)tp.x1 * gimg.sw;
#pragma unroll
            for (int k = 0; k < 4; ++k) grp[k] = warp_tap_groups(active, (unsigned long long)(n * gimg.sn + goff[k]));
        }
        for (int c = 0; c < C; ++c) {
            const float *bc = base + c * img.sc;
            const float g = gout[i * C + c];
            g00 += g * __ldg(bc + o00); g01 += g * __ldg(bc + o01); g10 += g * __ldg(bc + o10); g11 += g * __ldg(bc + o11);
            if (gimg.p) {
                float *gb = gimg.p + n * gimg.sn + c * gimg.sc;
                const float wk[4] = {w00, w01, w10, w11};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float sum = warp_group_sum(grp[k], wk[k] * g);
                    if (grp[k].leader) atomicAdd(gb + goff[k], sum);
                }
            }
        }
        const float gwx0 = g00 * tp.wy0 + g01 * tp.wy1, gwx1 = g10 * tp.wy0 + g11 * tp.wy1;
        const float gwy0 = g00 * tp.wx0 + g10 * tp.wx1, gwy1 = g01 * tp.wx0 + g11 * tp.wx1;
        float2 o; o.x = (gwx1 - gwx0) * tp.dx; o.y = (gwy1 - gwy0) * tp.dy;
        reinterpret_cast<float2 *>(gcoords)[i] = o;
    }
}

extern "C" int snde_bilinear_sampler_bwd(int N, int C, int hs, int ws, SndeImage img, int ht, int wt, const float *coords,
                                         int add_grid, const float *xtab, const float *ytab, int padding_mode, const float *gout,
                                         float *gcoords, SndeImageOut gimg, void *stream) {
    if (!img.ptr || !coords || !gout || !gcoords || (add_grid && (!xtab || !ytab))) return SNDE_ERR_NULL;
    if (N <= 0 || C <= 0 || hs <= 0 || ws <= 0 || ht <= 0 || wt <= 0) return SNDE_ERR_SHAPE;
    if (padding_mode != SNDE_PAD_ZEROS && padding_mode != SNDE_PAD_BORDER) return SNDE_ERR_ARG;
    const long long total = (long long)N * ht * wt;
    if (total * C >= (1LL << 31) || (long long)N * hs * ws * C >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    sampler_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(total, C, hs, ws, ldev(img), ht, wt, coords, add_grid, xtab, ytab,
                                                             padding_mode, gout, gcoords, ldev(gimg));
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// ---- deterministic image-gradient scatter: 64-bit fixed-point accumulation (two passes) ----
__global__ void __launch_bounds__(256) absmax_kernel(long long n, const float *__restrict__ v, unsigned *__restrict__ out_bits) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = fabsf(__ldg(v + i));
        if (a > m && a < __int_as_float(0x7f800000)) m = a;            // ignore inf / NaN: they poison the result either way
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

// exponent e with 2^e >= max|g| (from the bits the first pass left), as the shift of the fixed-point scale 2^(38 - e)
__device__ __forceinline__ int fixed_shift(const unsigned *max_bits) {
    const int e = (int)((*max_bits >> 23) & 0xff) - 127 + 1;
    return 38 - e;
}

__global__ void __launch_bounds__(256) sampler_gimg_fixed_kernel(long long total, int C, int hs, int ws, int ht, int wt,
                                                                 const float *__restrict__ coords, int add_grid,
                                                                 const float *__restrict__ xtab, const float *__restrict__ ytab, int border,
                                                                 const float *__restrict__ gout, long long *__restrict__ acc,
                                                                 const unsigned *__restrict__ max_bits) {
    const int sh = fixed_shift(max_bits);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % wt); const long long t = i / wt; const int iy = (int)(t % ht); const long long n = t / ht;
        const float2 cf = reinterpret_cast<const float2 *>(coords)[i];
        float x = cf.x, y = cf.y;
        if (add_grid) { x = __ldg(xtab + ix) + x; y = __ldg(ytab + iy) + y; }
        Taps tp; make_taps(x, y, hs, ws, border, tp);
        const float w00 = tp.wx0 * tp.wy0, w01 = tp.wx0 * tp.wy1, w10 = tp.wx1 * tp.wy0, w11 = tp.wx1 * tp.wy1;
        for (int c = 0; c < C; ++c) {
            const float g = gout[i * C + c];
            unsigned long long *plane = reinterpret_cast<unsigned long long *>(acc) + ((n * C + c) * hs) * (long long)ws;
            auto add = [&](int yy, int xx, float w) {
                const long long q = __double2ll_rn(ldexp((double)(w * g), sh));     // the product is rounded to fp32 first, like the default mode
                if (q) atomicAdd(plane + (long long)yy * ws + xx, (unsigned long long)q);
            };
            add(tp.y0, tp.x0, w00); add(tp.y1, tp.x0, w01); add(tp.y0, tp.x1, w10); add(tp.y1, tp.x1, w11);
        }
    }
}

__global__ void __launch_bounds__(256) fixed_to_gimg_kernel(long long total, int C, int hs, int ws, const long long *__restrict__ acc,
                                                            const unsigned *__restrict__ max_bits, LImageOut gimg) {
    const int sh = fixed_shift(max_bits);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % ws); long long t = i / ws; const int y = (int)(t % hs); t /= hs; const int c = (int)(t % C); const long long n = t / C;
        const long long q = acc[i];
        if (q) {
            float *dst = gimg.p + n * gimg.sn + c * gimg.sc + (long long)y * gimg.sh + (long long)x * gimg.sw;
            *dst += (float)ldexp((double)q, -sh);
        }
    }
}

extern "C" int snde_bilinear_sampler_bwd_deterministic(int N, int C, int hs, int ws, SndeImage img, int ht, int wt, const float *coords,
                                                       int add_grid, const float *xtab, const float *ytab, int padding_mode,
                                                       const float *gout, float *gcoords, SndeImageOut gimg, long long *scratch,
                                                       void *stream) {
    if (!gimg.ptr || !scratch) return SNDE_ERR_NULL;
    SndeImageOut none; memset(&none, 0, sizeof none);
    int rc = snde_bilinear_sampler_bwd(N, C, hs, ws, img, ht, wt, coords, add_grid, xtab, ytab, padding_mode, gout, gcoords, none, stream);
    if (rc) return rc;
    const long long total = (long long)N * ht * wt, cells = (long long)N * C * hs * ws;
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(scratch, 0, sizeof(long long) * (size_t)(cells + 1), st) != cudaSuccess) return SNDE_ERR_CUDA;
    unsigned *max_bits = reinterpret_cast<unsigned *>(scratch + cells);
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    absmax_kernel<<<blocks, 256, 0, st>>>(total * C, gout, max_bits);
    sampler_gimg_fixed_kernel<<<blocks, 256, 0, st>>>(total, C, hs, ws, ht, wt, coords, add_grid, xtab, ytab, padding_mode, gout, scratch, max_bits);
    const int blocks2 = (int)((cells + 255) / 256 < 148 * 32 ? (cells + 255) / 256 : 148 * 32);
    fixed_to_gimg_kernel<<<blocks2, 256, 0, st>>>(cells, C, hs, ws, scratch, max_bits, ldev(gimg));
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// ---- image_similarity / DSSIM (materialising, reference-signature version) ----
// one thread per (n, y, x); 5x5 neighbourhoods come straight from global memory (L1/L2 resident).
// The fused strip kernel is the fast path; this one exists so the reference's own call sequence runs.
__device__ __forceinline__ float img_or_zero(const LImage &im, long long n, int c, int y, int x, int h, int w) {
    return (y >= 0 && y < h && x >= 0 && x < w) ? ld_img(im, n, c, y, x) : 0.f;
}

__device__ __forceinline__ void ssim_at(const LImage &X, const LImage &Y, long long n, int c, int cy, int cx, int h, int w, Ssim &ss) {
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
            const float xv = img_or_zero(X, n, c, cy + dr, cx + dc, h, w), yv = img_or_zero(Y, n, c, cy + dr, cx + dc, h, w);
            sx = sx + xv; sy = sy + yv; sxx = sxx + xv * xv; syy = syy + yv * yv; sxy = sxy + xv * yv;
        }
    ssim_from_sums(sx, sy, sxx, syy, sxy, ss);
}

__global__ void __launch_bounds__(256) similarity_fwd_kernel(long long total, int h, int w, float alpha, float oma, LImage X, LImage Y,
                                                             float *__restrict__ err) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % w); const long long t = i / w; const int iy = (int)(t % h); const long long n = t / h;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Ssim ss; ssim_at(X, Y, n, c, iy, ix, h, w, ss);
            float v;
            err[i * 3 + c] = similarity_value(alpha, oma, ss.S, ld_img(X, n, c, iy, ix), ld_img(Y, n, c, iy, ix), v);
        }
    }
}

extern "C" int snde_image_similarity_fwd(int N, int h, int w, float alpha, SndeImage x, SndeImage y, float *err, void *stream) {
    if (!x.ptr || !y.ptr || !err) return SNDE_ERR_NULL;
    if (N <= 0 || h <= 0 || w <= 0) return SNDE_ERR_SHAPE;
    const long long total = (long long)N * h * w;
    if (total * 3 >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    similarity_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(total, h, w, alpha, (float)(1.0 - (double)alpha), ldev(x), ldev(y), err);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// Backward, tiled: a CTA owns a 32 x 8 tile of one image.  Per channel it stages the tile + a 2-pixel halo of x and y in
// shared memory (zeros outside the image = the pools' padding), forms the four SSIM backward coefficients once per window
// centre of the tile + 1-pixel halo (each centre is shared by the 9 pixels around it), then every pixel sums its 3 x 3
// neighbourhood of coefficients.  Same operations in the same order as the per-pixel formulation it replaces (81 window taps
// and 9 SSIM evaluations per pixel and channel, straight from global memory), so the results are bit-identical to it.
constexpr int SIM_TX = 32, SIM_TY = 8;

__global__ void __launch_bounds__(SIM_TX * SIM_TY) similarity_bwd_kernel(int h, int w, float alpha, float oma, LImage X, LImage Y,
                                                                         const float *__restrict__ gerr, float *__restrict__ gx,
                                                                         float *__restrict__ gy) {
    __shared__ float sx_[SIM_TY + 4][SIM_TX + 4], sy_[SIM_TY + 4][SIM_TX + 4];
    __shared__ float cf[4][SIM_TY + 2][SIM_TX + 2];
    const long long n = blockIdx.z;
    const int x0 = blockIdx.x * SIM_TX, y0 = blockIdx.y * SIM_TY;
    const int tid = threadIdx.x, tx = tid % SIM_TX, ty = tid / SIM_TX;
    const int ix = x0 + tx, iy = y0 + ty;
    const bool inside = ix < w && iy < h;
    const long long pix = (n * h + iy) * (long long)w + ix;
    for (int c = 0; c < 3; ++c) {
        __syncthreads();                                     // the previous channel's readers are done
        for (int i = tid; i < (SIM_TY + 4) * (SIM_TX + 4); i += SIM_TX * SIM_TY) {
            const int r = i / (SIM_TX + 4), q = i - r * (SIM_TX + 4);
            const int gy_ = y0 + r - 2, gx_ = x0 + q - 2;
            sx_[r][q] = img_or_zero(X, n, c, gy_, gx_, h, w);
            sy_[r][q] = img_or_zero(Y, n, c, gy_, gx_, h, w);
        }
        __syncthreads();
        for (int i = tid; i < (SIM_TY + 2) * (SIM_TX + 2); i += SIM_TX * SIM_TY) {
            const int r = i / (SIM_TX + 2), q = i - r * (SIM_TX + 2);
            const int qy = y0 + r - 1, qx = x0 + q - 1;
            float cmx = 0.f, cmy = 0.f, ce = 0.f, cxy = 0.f;
            if (qy >= 0 && qy < h && qx >= 0 && qx < w) {
                float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
                for (int dr = 0; dr < 3; ++dr)
#pragma unroll
                    for (int dc = 0; dc < 3; ++dc) {
                        const float xv = sx_[r + dr][q + dc], yv = sy_[r + dr][q + dc];
                        sx = sx + xv; sy = sy + yv; sxx = sxx + xv * xv; syy = syy + yv * yv; sxy = sxy + xv * yv;
                    }
                Ssim ss; ssim_from_sums(sx, sy, sxx, syy, sxy, ss);
                const float v = (1.f - ss.S) / 2.f;
                const float g = gerr[((n * h + qy) * (long long)w + qx) * 3 + c];
                const float gS = (v >= 0.f && v <= 1.f) ? -0.5f * alpha * g : 0.f;
                ssim_bwd_coeffs(ss, gS, cmx, cmy, ce, cxy);
            }
            cf[0][r][q] = cmx; cf[1][r][q] = cmy; cf[2][r][q] = ce; cf[3][r][q] = cxy;
        }
        __syncthreads();
        if (inside) {
            float bmx = 0.f, bmy = 0.f, be = 0.f, bxy = 0.f;
#pragma unroll
            for (int dr = 0; dr < 3; ++dr)
#pragma unroll
                for (int dc = 0; dc < 3; ++dc) {
                    const int qy = iy + dr - 1, qx = ix + dc - 1;
                    if (qy < 0 || qy >= h || qx < 0 || qx >= w) continue;
                    bmx += cf[0][ty + dr][tx + dc]; bmy += cf[1][ty + dr][tx + dc]; be += cf[2][ty + dr][tx + dc]; bxy += cf[3][ty + dr][tx + dc];
                }
            const float xv = sx_[ty + 2][tx + 2], yv = sy_[ty + 2][tx + 2];
            const float l1 = oma * gerr[pix * 3 + c] * signf(xv - yv);
            if (gx) gx[pix * 3 + c] = (bmx + 2.f * xv * be + yv * bxy) / 9.f + l1;
            if (gy) gy[pix * 3 + c] = (bmy + 2.f * yv * be + xv * bxy) / 9.f - l1;
        }
    }
}

extern "C" int snde_image_similarity_bwd(int N, int h, int w, float alpha, SndeImage x, SndeImage y, const float *gerr, float *gx,
                                         float *gy, void *stream) {
    if (!x.ptr || !y.ptr || !gerr) return SNDE_ERR_NULL;
    if (N <= 0 || h <= 0 || w <= 0) return SNDE_ERR_SHAPE;
    const long long total = (long long)N * h * w;
    if (total * 3 >= (1LL << 31) || N > 65535 || (h + SIM_TY - 1) / SIM_TY > 65535) return SNDE_ERR_TOO_LARGE;
    if (!gx && !gy) return SNDE_OK;
    const dim3 grid((w + SIM_TX - 1) / SIM_TX, (h + SIM_TY - 1) / SIM_TY, N);
    similarity_bwd_kernel<<<grid, SIM_TX * SIM_TY, 0, as_stream(stream)>>>(h, w, alpha, (float)(1.0 - (double)alpha), ldev(x), ldev(y), gerr, gx, gy);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// ---- smooth_loss ----
__device__ __forceinline__ float pair_weight(const LImage &I, long long m, int y0, int x0, int y1, int x1) {
    return edge_weight3(ld_img(I, m, 0, y0, x0), ld_img(I, m, 1, y0, x0), ld_img(I, m, 2, y0, x0),
                        ld_img(I, m, 0, y1, x1), ld_img(I, m, 1, y1, x1), ld_img(I, m, 2, y1, x1));
}

__global__ void __launch_bounds__(256) smooth_fwd_kernel(long long total, int h, int w, LMap D, LImage I, double *__restrict__ sums) {
    __shared__ float red[8 * 2];
    float acc[2] = {0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % w); const long long t = i / w; const int iy = (int)(t % h); const long long m = t / h;
        const float d = ld_map(D, m, iy, ix);
        if (ix + 1 < w) acc[0] += fabsf((d - ld_map(D, m, iy, ix + 1)) * pair_weight(I, m, iy, ix, iy, ix + 1));
        if (iy + 1 < h) acc[1] += fabsf((d - ld_map(D, m, iy + 1, ix)) * pair_weight(I, m, iy, ix, iy + 1, ix));
    }
    block_atomic_add<2>(acc, sums, red);
}

extern "C" int snde_smooth_loss_fwd(int M, int h, int w, SndeMap depth, SndeImage img, double *sums, void *stream) {
    if (!depth.ptr || !img.ptr || !sums) return SNDE_ERR_NULL;
    if (M <= 0 || h < 2 || w < 2) return SNDE_ERR_SHAPE;
    const long long total = (long long)M * h * w;
    if (total * 3 >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    smooth_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(total, h, w, ldev(depth), ldev(img), sums);
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

__global__ void __launch_bounds__(256) smooth_bwd_kernel(long long total, int h, int w, LMap D, LImage I, const float *__restrict__ gscale,
                                                         double inv_nx, double inv_ny, LMapOut G) {
    const float gs = __ldg(gscale);
    const float cx = (float)((double)gs * inv_nx), cy = (float)((double)gs * inv_ny);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % w); const long long t = i / w; const int iy = (int)(t % h); const long long m = t / h;
        const float d = ld_map(D, m, iy, ix);
        float g = 0.f;
        if (ix + 1 < w) { const float wg = pair_weight(I, m, iy, ix, iy, ix + 1); g += cx * signf((d - ld_map(D, m, iy, ix + 1)) * wg) * wg; }
        if (ix >= 1)    { const float wg = pair_weight(I, m, iy, ix - 1, iy, ix); g -= cx * signf((ld_map(D, m, iy, ix - 1) - d) * wg) * wg; }
        if (iy + 1 < h) { const float wg = pair_weight(I, m, iy, ix, iy + 1, ix); g += cy * signf((d - ld_map(D, m, iy + 1, ix)) * wg) * wg; }
        if (iy >= 1)    { const float wg = pair_weight(I, m, iy - 1, ix, iy, ix); g -= cy * signf((ld_map(D, m, iy - 1, ix) - d) * wg) * wg; }
        G.p[m * G.sn + (long long)iy * G.sh + (long long)ix * G.sw] = g;
    }
}

extern "C" int snde_smooth_loss_bwd(int M, int h, int w, SndeMap depth, SndeImage img, const float *gscale_dev, SndeMapOut gdepth,
                                    void *stream) {
    if (!depth.ptr || !img.ptr || !gscale_dev || !gdepth.ptr) return SNDE_ERR_NULL;
    if (M <= 0 || h < 2 || w < 2) return SNDE_ERR_SHAPE;
    const long long total = (long long)M * h * w;
    if (total * 3 >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    smooth_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(total, h, w, ldev(depth), ldev(img), gscale_dev,
                                                            1.0 / ((double)M * h * (w - 1)), 1.0 / ((double)M * (h - 1) * w), ldev(gdepth));
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

// =============================================================================================
// frames -> normalised fp32 pyramid (caller-side prep)
// reference: models/baseline.py:761-767 (uint8 -> float, *= 1/255, * 2 - 1), :772-775 (sources concatenated source-major) and
// utils/utils_edited.py:89-112 (scale_pyramid: F.interpolate(mode='area')).
// One CTA owns one tile of TR rows x 128 columns of one channel plane (TR = max(16, 2^(num_scales-1))).  The bytes come in
// as 32-bit words along the row, become (u8 * 1/255) * 2 - 1 op by op, go to shared memory and straight out as level 0 with
// 16-byte stores (a warp writes one 512-byte row segment); the coarser levels are then formed from the shared tile: every
// output sums its 2^s x 2^s window in row-major order and multiplies by the (power of two) reciprocal count, like
// adaptive_avg_pool2d, and consecutive threads write consecutive outputs of a row.
// Traffic: 1 byte read + 4 * (1 + 1/4 + 1/16 + ...) bytes written per full-resolution pixel and channel.
// `groups`: the channel axis holds groups * 3 channels (the DataLoader's (B, 3*S, H, W) source block); group g of frame b
// lands in image g * n + b of the outputs -- the source-major order of the loss path -- so one launch does all S sources.
// =============================================================================================
struct PyrLevels { float *p[SNDE_MAX_SCALES]; };

// grid: x = tile of the plane (tiles_x * tiles_y), y = channel + 3 * group, z = frame.  Everything inside a plane is 32-bit
// index arithmetic with shifts (the kernel is instruction-bound, not bandwidth-bound: ncu, profiles/).
template <int NS>
__global__ void __launch_bounds__(256) frames_to_pyramid_kernel(int n, int H, int W, int tiles_x, const uint8_t *__restrict__ frames,
                                                                long long sn, long long sc, long long sh, long long sw, PyrLevels lv) {
    constexpr int R = 1 << (NS - 1);
    constexpr int TR = R > 16 ? R : 16, TC = 128;
    __shared__ __align__(16) float tile[TR][TC];
    const int tx = (int)(blockIdx.x % (unsigned)tiles_x), ty = (int)(blockIdx.x / (unsigned)tiles_x);
    const int cg = (int)blockIdx.y, g = cg / 3, c = cg - 3 * g;
    const int b = (int)blockIdx.z;
    const int x0 = tx * TC, y0 = ty * TR;
    const int cols = min(TC, W - x0), rows = min(TR, H - y0);        // multiples of R (W % R == H % R == 0)
    const uint8_t *src = frames + b * sn + (long long)cg * sc + (long long)y0 * sh + (long long)x0 * sw;
    const long long plane = ((long long)g * n + b) * 3 + c;
    float *dst0 = lv.p[0] + (plane * H + y0) * (long long)W + x0;
    const int tid = (int)threadIdx.x;
    // fast path: bytes contiguous along x and 4-byte aligned rows, level-0 rows 16-byte aligned
    const bool words = sw == 1 && ((sh | sc | sn) & 3) == 0 && ((uintptr_t)frames & 3) == 0 && (W & 3) == 0;
    const bool vst = words && ((uintptr_t)lv.p[0] & 15) == 0;
    // ---- level 0: bytes -> normalised floats -> shared tile + global; thread = (row y, 4 columns x4), 8 rows per sweep ----
    {
        const int x4 = (tid & 31) * 4;
        if (x4 < cols) {
            const int sh32 = (int)sh;
            for (int y = tid >> 5; y < rows; y += 8) {
                float v[4];
                if (words) {
                    const unsigned w4 = __ldg(reinterpret_cast<const unsigned *>(src + y * sh32 + x4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = (float)__byte_perm(w4, 0u, 0x4440u + k);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = x4 + k < cols ? (float)__ldg(src + (long long)y * sh + (long long)(x4 + k) * sw) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = __fsub_rn(__fmul_rn(__fmul_rn(v[k], 1.f / 255.f), 2.f), 1.f);   // (u8 * 1/255) * 2 - 1, op by op
                const float4 q = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(&tile[y][x4]) = q;
                float *d = dst0 + y * W + x4;
                if (vst) *reinterpret_cast<float4 *>(d) = q;                                    // a warp writes one 512-byte row segment
                else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (x4 + k < cols) d[k] = v[k];
                }
            }
        }
    }
    __syncthreads();
    // ---- levels 1 .. NS-1 from the shared tile: every output sums its 2^s x 2^s window in row-major order; the window rows come
    //      in as 8- / 16-byte shared loads, consecutive threads take consecutive outputs of a row ----
#pragma unroll
    for (int s = 1; s < NS; ++s) {
        constexpr int dummy = 0; (void)dummy;
        const int k = 1 << s, ws = W >> s;
        const int ocols = cols >> s, orows = rows >> s;
        const int lcols = 7 - s;                                     // log2 of the outputs per full tile row (TC >> s)
        float *dst = lv.p[s] + (plane * (H >> s) + (y0 >> s)) * (long long)ws + (x0 >> s);
        const float inv = 1.f / (float)(k * k);
        for (int i = tid; i < (orows << lcols); i += 256) {
            const int oy = i >> lcols, ox = i & ((1 << lcols) - 1);
            if (ox >= ocols) continue;
            float acc = 0.f;
            const float *row = &tile[oy << s][ox << s];
#pragma unroll
            for (int dy = 0; dy < k; ++dy, row += TC) {
                if (s == 1) {
                    const float2 q = *reinterpret_cast<const float2 *>(row);
                    acc = __fadd_rn(__fadd_rn(acc, q.x), q.y);
                } else {
#pragma unroll
                    for (int dx = 0; dx < k; dx += 4) {
                        const float4 q = *reinterpret_cast<const float4 *>(row + dx);
                        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, q.x), q.y), q.z), q.w);
                    }
                }
            }
            dst[oy * ws + ox] = __fmul_rn(acc, inv);
        }
    }
}

extern "C" int snde_frames_to_pyramid_grouped(int n, int groups, int H, int W, int num_scales, const uint8_t *frames, int64_t sn,
                                              int64_t sc, int64_t sh, int64_t sw, float *const *levels, void *stream) {
    if (!frames || !levels) return SNDE_ERR_NULL;
    if (n <= 0 || groups <= 0 || H <= 0 || W <= 0 || num_scales <= 0 || num_scales > SNDE_MAX_SCALES) return SNDE_ERR_SHAPE;
    const int R = 1 << (num_scales - 1);
    if (H % R || W % R) return SNDE_ERR_SHAPE;
    if ((long long)n * groups * 3 * H * W >= (1LL << 40)) return SNDE_ERR_TOO_LARGE;
    PyrLevels lv;
    for (int s = 0; s < SNDE_MAX_SCALES; ++s) lv.p[s] = nullptr;
    for (int s = 0; s < num_scales; ++s) { if (!levels[s]) return SNDE_ERR_NULL; lv.p[s] = levels[s]; }
    const int TR = R > 16 ? R : 16;
    const int tiles_x = (W + 127) / 128, tiles_y = (H + TR - 1) / TR;
    if (n > 65535 || groups * 3 > 65535 || (long long)sh * TR + W >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    const dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)(3 * groups), (unsigned)n);
    cudaStream_t st = as_stream(stream);
    switch (num_scales) {
        case 1: frames_to_pyramid_kernel<1><<<grid, 256, 0, st>>>(n, H, W, tiles_x, frames, sn, sc, sh, sw, lv); break;
        case 2: frames_to_pyramid_kernel<2><<<grid, 256, 0, st>>>(n, H, W, tiles_x, frames, sn, sc, sh, sw, lv); break;
        case 3: frames_to_pyramid_kernel<3><<<grid, 256, 0, st>>>(n, H, W, tiles_x, frames, sn, sc, sh, sw, lv); break;
        case 4: frames_to_pyramid_kernel<4><<<grid, 256, 0, st>>>(n, H, W, tiles_x, frames, sn, sc, sh, sw, lv); break;
        case 5: frames_to_pyramid_kernel<5><<<grid, 256, 0, st>>>(n, H, W, tiles_x, frames, sn, sc, sh, sw, lv); break;
        default: frames_to_pyramid_kernel<6><<<grid, 256, 0, st>>>(n, H, W, tiles_x, frames, sn, sc, sh, sw, lv); break;
    }
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}

extern "C" int snde_frames_to_pyramid(int n, int H, int W, int num_scales, const uint8_t *frames, int64_t sn, int64_t sc, int64_t sh,
                                      int64_t sw, float *const *levels, void *stream) {
    return snde_frames_to_pyramid_grouped(n, 1, H, W, num_scales, frames, sn, sc, sh, sw, levels, stream);
}
