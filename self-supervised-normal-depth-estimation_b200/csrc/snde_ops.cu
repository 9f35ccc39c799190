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
        for (int c = 0; c < C; ++c) {
            const float *bc = base + c * img.sc;
            const float g = gout[i * C + c];
            g00 += g * __ldg(bc + o00); g01 += g * __ldg(bc + o01); g10 += g * __ldg(bc + o10); g11 += g * __ldg(bc + o11);
            if (gimg.p) {
                float *gb = gimg.p + n * gimg.sn + c * gimg.sc;
                atomicAdd(gb + (long long)tp.y0 * gimg.sh + (long long)tp.x0 * gimg.sw, w00 * g);
                atomicAdd(gb + (long long)tp.y1 * gimg.sh + (long long)tp.x0 * gimg.sw, w01 * g);
                atomicAdd(gb + (long long)tp.y0 * gimg.sh + (long long)tp.x1 * gimg.sw, w10 * g);
                atomicAdd(gb + (long long)tp.y1 * gimg.sh + (long long)tp.x1 * gimg.sw, w11 * g);
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

__global__ void __launch_bounds__(256) similarity_bwd_kernel(long long total, int h, int w, float alpha, float oma, LImage X, LImage Y,
                                                             const float *__restrict__ gerr, float *__restrict__ gx, float *__restrict__ gy) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % w); const long long t = i / w; const int iy = (int)(t % h); const long long n = t / h;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float bmx = 0.f, bmy = 0.f, be = 0.f, bxy = 0.f;
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    const int qy = iy + dr, qx = ix + dc;
                    if (qy < 0 || qy >= h || qx < 0 || qx >= w) continue;
                    Ssim ss; ssim_at(X, Y, n, c, qy, qx, h, w, ss);
                    const float v = (1.f - ss.S) / 2.f;
                    const float g = gerr[((n * h + qy) * (long long)w + qx) * 3 + c];
                    const float gS = (v >= 0.f && v <= 1.f) ? -0.5f * alpha * g : 0.f;
                    float cmx, cmy, ce, cxy; ssim_bwd_coeffs(ss, gS, cmx, cmy, ce, cxy);
                    bmx += cmx; bmy += cmy; be += ce; bxy += cxy;
                }
            const float xv = ld_img(X, n, c, iy, ix), yv = ld_img(Y, n, c, iy, ix);
            const float l1 = oma * gerr[i * 3 + c] * signf(xv - yv);
            if (gx) gx[i * 3 + c] = (bmx + 2.f * xv * be + yv * bxy) / 9.f + l1;
            if (gy) gy[i * 3 + c] = (bmy + 2.f * yv * be + xv * bxy) / 9.f - l1;
        }
    }
}

extern "C" int snde_image_similarity_bwd(int N, int h, int w, float alpha, SndeImage x, SndeImage y, const float *gerr, float *gx,
                                         float *gy, void *stream) {
    if (!x.ptr || !y.ptr || !gerr) return SNDE_ERR_NULL;
    if (N <= 0 || h <= 0 || w <= 0) return SNDE_ERR_SHAPE;
    const long long total = (long long)N * h * w;
    if (total * 3 >= (1LL << 31)) return SNDE_ERR_TOO_LARGE;
    if (!gx && !gy) return SNDE_OK;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    similarity_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(total, h, w, alpha, (float)(1.0 - (double)alpha), ldev(x), ldev(y), gerr, gx, gy);
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
// reference: models/baseline.py:761-767 (uint8 -> float, *= 1/255, * 2 - 1) and
// utils/utils_edited.py:89-112 (scale_pyramid: F.interpolate(mode='area')).
// One thread owns one R x R block of one channel plane (R = 2^(num_scales-1)): it reads R rows of R bytes and writes
// the block of every level; neighbouring threads own neighbouring blocks of a row, so every access is coalesced.
// Sums run over each window in row-major order and are divided by the (power of two) count, like adaptive_avg_pool2d.
// =============================================================================================
struct PyrLevels { float *p[SNDE_MAX_SCALES]; };

template <int NS>
__global__ void __launch_bounds__(128) frames_to_pyramid_kernel(long long total, int H, int W, const uint8_t *__restrict__ frames,
                                                                long long sn, long long sc, long long sh, long long sw, PyrLevels lv) {
    constexpr int R = 1 << (NS - 1);
    const int bw = W / R, bh = H / R;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int bx = (int)(i % bw);
        long long r = i / bw;
        const int by = (int)(r % bh);
        r /= bh;
        const int c = (int)(r % 3);
        const long long n = r / 3;
        const uint8_t *src = frames + n * sn + c * sc + (long long)(by * R) * sh + (long long)(bx * R) * sw;
        float v[R][R];
#pragma unroll
        for (int y = 0; y < R; ++y) {
            if (sw == 1 && R == 8 && (((uintptr_t)(src + y * sh)) & 7) == 0) {
                const uint2 w8 = __ldg(reinterpret_cast<const uint2 *>(src + y * sh));
#pragma unroll
                for (int x = 0; x < 8; ++x) v[y][x] = (float)(((x < 4 ? w8.x : w8.y) >> (8 * (x & 3))) & 0xffu);
            } else {
#pragma unroll
                for (int x = 0; x < R; ++x) v[y][x] = (float)__ldg(src + y * sh + x * sw);
            }
#pragma unroll
            for (int x = 0; x < R; ++x) v[y][x] = __fsub_rn(__fmul_rn(__fmul_rn(v[y][x], 1.f / 255.f), 2.f), 1.f);   // (u8 * 1/255) * 2 - 1, op by op
        }
        const long long plane = n * 3 + c;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            constexpr int dummy = 0; (void)dummy;
            const int k = 1 << s, hs = H >> s, ws = W >> s, m = R >> s;      // window size, level size, block size at this level
            float *dst = lv.p[s] + (plane * hs + (long long)by * m) * ws + (long long)bx * m;
#pragma unroll
            for (int y = 0; y < m; ++y)
#pragma unroll
                for (int x = 0; x < m; ++x) {
                    float acc = 0.f;
#pragma unroll
                    for (int dy = 0; dy < k; ++dy)
#pragma unroll
                        for (int dx = 0; dx < k; ++dx) acc = __fadd_rn(acc, v[y * k + dy][x * k + dx]);
                    dst[(long long)y * ws + x] = s == 0 ? v[y][x] : __fmul_rn(acc, 1.f / (float)(k * k));
                }
        }
    }
}

extern "C" int snde_frames_to_pyramid(int n, int H, int W, int num_scales, const uint8_t *frames, int64_t sn, int64_t sc, int64_t sh,
                                      int64_t sw, float *const *levels, void *stream) {
    if (!frames || !levels) return SNDE_ERR_NULL;
    if (n <= 0 || H <= 0 || W <= 0 || num_scales <= 0 || num_scales > 5) return SNDE_ERR_SHAPE;     // R x R floats live in registers: R <= 16
    const int R = 1 << (num_scales - 1);
    if (H % R || W % R) return SNDE_ERR_SHAPE;
    if ((long long)n * 3 * H * W >= (1LL << 40)) return SNDE_ERR_TOO_LARGE;
    PyrLevels lv;
    for (int s = 0; s < SNDE_MAX_SCALES; ++s) lv.p[s] = nullptr;
    for (int s = 0; s < num_scales; ++s) { if (!levels[s]) return SNDE_ERR_NULL; lv.p[s] = levels[s]; }
    const long long total = (long long)n * 3 * (H / R) * (W / R);
    const long long want = (total + 127) / 128;
    const int blocks = (int)(want < 148LL * 32 ? want : 148LL * 32);
    cudaStream_t st = as_stream(stream);
    switch (num_scales) {
        case 1: frames_to_pyramid_kernel<1><<<blocks, 128, 0, st>>>(total, H, W, frames, sn, sc, sh, sw, lv); break;
        case 2: frames_to_pyramid_kernel<2><<<blocks, 128, 0, st>>>(total, H, W, frames, sn, sc, sh, sw, lv); break;
        case 3: frames_to_pyramid_kernel<3><<<blocks, 128, 0, st>>>(total, H, W, frames, sn, sc, sh, sw, lv); break;
        case 4: frames_to_pyramid_kernel<4><<<blocks, 128, 0, st>>>(total, H, W, frames, sn, sc, sh, sw, lv); break;
        default: frames_to_pyramid_kernel<5><<<blocks, 128, 0, st>>>(total, H, W, frames, sn, sc, sh, sw, lv); break;
    }
    SNDE_CHECK_LAUNCH();
    return SNDE_OK;
}
