/*
 * snde.h -- C ABI of libsnde.so: B200 (sm_100a) kernels for the self-supervised
 * rigid-warp + photometric (L1 + 3x3 SSIM) + edge-aware smoothness loss path.
 *
 * The reference (Eisoc/Self-supervised-Normal-Depth-Estimation) is pure Python: its
 * "plugin boundary" for this path is a set of Python functions in two modules.  Each
 * entry point below names the reference function(s) it replaces (file:line under the
 * reference root).  The Python side (snde_b200/loss_functions.py, utils_edited.py)
 * re-exports the reference names on top of these calls via ctypes.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.  All tensors are fp32 DEVICE
 *    pointers owned by the caller; the library never allocates, frees or keeps them.
 *  - Strides are in ELEMENTS.  Images are addressed as logical (n, c, y, x) through
 *    (sn, sc, sh, sw), so planar NCHW memory, NHWC memory and permuted views all work
 *    (the reference hands NHWC views of NCHW memory around).
 *  - Every call launches asynchronously on the given stream (a cudaStream_t passed as
 *    void*) and returns 0 or a negative SndeStatus; no internal synchronisation.
 *  - Batch order of every N = S*B sized tensor is source-major: n = src*B + b
 *    (models/baseline.py:772-775, 915-922).
 */
#ifndef SNDE_H_
#define SNDE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SNDE_API __attribute__((visibility("default")))
#else
#define SNDE_API
#endif

#define SNDE_VERSION 200      /* 0.2.0: grouped pyramid prep, image-gradient outputs of the fused loss */
#define SNDE_MAX_SCALES 6
#define SNDE_MAX_SOURCES 8     /* source views per target in the fused loss */

typedef enum SndeStatus {
    SNDE_OK = 0,
    SNDE_ERR_NULL = -1,        /* a required pointer is NULL */
    SNDE_ERR_SHAPE = -2,       /* non-positive / unsupported dimension (C must be 3 where stated; h,w >= 2 for smoothness) */
    SNDE_ERR_TOO_LARGE = -3,   /* an index would not fit 31 bits (N*h*w >= 2^31) */
    SNDE_ERR_WORKSPACE = -4,   /* workspace missing, too small or misaligned (needs 16-byte alignment) */
    SNDE_ERR_ARG = -5,         /* bad enum / flag / scalar */
    SNDE_ERR_CUDA = -6         /* cudaGetLastError() after launch != cudaSuccess */
} SndeStatus;

typedef enum SndePadding {
    SNDE_PAD_ZEROS = 0,        /* the reference's sampler: value 0 outside [0,w-1) x [0,h-1)  (utils_edited.py:435-484) */
    SNDE_PAD_BORDER = 1        /* extension (unpinned by the reference): clamp coordinates to the image, == grid_sample(border, align_corners=True) */
} SndePadding;

/* 3-channel image batch, logical (n, c, y, x) */
typedef struct SndeImage { const float *ptr; int64_t sn, sc, sh, sw; } SndeImage;
typedef struct SndeImageOut { float *ptr; int64_t sn, sc, sh, sw; } SndeImageOut;
/* single-channel map batch, logical (n, y, x) */
typedef struct SndeMap { const float *ptr; int64_t sn, sh, sw; } SndeMap;
typedef struct SndeMapOut { float *ptr; int64_t sn, sh, sw; } SndeMapOut;

SNDE_API int snde_version(void);
SNDE_API const char *snde_strerror(int status);

/* meshgrid 1-D table  t[i] = (linspace(-1,1,n)[i] + 1) * 0.5 * (n-1), bit-identical to the
 * reference's (utils/utils_edited.py:307-317; NOT exact integers).  HOST function, host buffer. */
SNDE_API int snde_fill_grid_table(int n, float *host_out);

/* ------------------------------------------------------------------------------------------
 * Fused loss: everything GeoNetModel.build_rigid_warp_flow (models/baseline.py:867-970) and
 * build_losses (:1115-1161) compute between (pyramids, depth, poses, intrinsics) and
 * loss_total, forward AND backward in one pass:
 *   pose_vec2mat/euler2mat/inverse/K4@T  (utils_edited.py:171-214, 268-289, 343-357)
 *   meshgrid, pixel2cam, cam2pixel, compute_rigid_flow, flow_warp, bilinear_sampler (:216-486)
 *   DSSIM, image_similarity (:121-141, loss_functions.py:4-6), torch.mean, smooth_loss (:8-24)
 * and their autograd backward to depth, poses and (extension) intrinsics.
 *
 *   loss_rigid_warp  = sum_s photo_weight[s] * (mean(fwd_err_s) + mean(bwd_err_s))
 *   loss_disp_smooth = sum_s smooth_weight[s] * smooth_loss(depth_s, cat(tgt_s, src_s))
 * (the reference uses photo_weight = w_rigid*S/2 and smooth_weight = w_smooth/2^s).
 * Gradients are d(loss_rigid_warp + loss_disp_smooth)/d(input), i.e. for upstream gradient 1.
 * ---------------------------------------------------------------------------------------- */
typedef struct SndeLossScale {
    int32_t h, w;
    SndeImage tgt;              /* (B,3,h,w) */
    SndeImage src;              /* (S*B,3,h,w) source-major */
    SndeMap depth;              /* (B*(1+S),h,w): B target depths then S*B source depths (models/baseline.py:905-912) */
    SndeMapOut grad_depth;      /* same shape; written completely when SNDE_LOSS_GRAD is set */
    const float *K;             /* (B,3,3) contiguous: this scale's intrinsics (compute_multi_scale_intrinsics, utils_edited.py:149) */
    const float *xtab, *ytab;   /* device copies of snde_fill_grid_table(w), (h) */
    float photo_weight, smooth_weight;
    SndeMap mask_fwd, mask_bwd;         /* optional (ptr may be NULL): explainability weight per pixel, (S*B,h,w); extension */
    SndeImageOut warped_fwd, warped_bwd;/* optional outputs (S*B,3,h,w): flow_warp results (logged by the reference every 10k it) */
    SndeImageOut err_fwd, err_bwd;      /* optional outputs (S*B,3,h,w): image_similarity maps */
    /* optional (SNDE_LOSS_GRAD; extension, section 8(f3)): d loss_rigid_warp / d image, ACCUMULATED into buffers the caller has
     * zero-filled -- (B,3,h,w) and (S*B,3,h,w), any strides.  Two parts land in the same buffer: the dense part (the image as
     * the x argument of image_similarity) and the scatter part (the image as sampled by bilinear_sampler: index_select
     * backward, utils_edited.py:473-476), the latter with warp-aggregated atomics (lanes hitting the same pixel are found with
     * match.any; one atomic per distinct address).  The smoothness term gives the images no gradient (as in the reference,
     * where they are data).  Summation order varies from run to run in the last bits; the deterministic alternative is the
     * stand-alone snde_bilinear_sampler_bwd_deterministic. */
    SndeImageOut grad_tgt_img, grad_src_img;
} SndeLossScale;

enum {
    SNDE_LOSS_GRAD = 1,         /* also produce grad_depth / grad_poses (/ grad_K) */
    SNDE_LOSS_GRAD_K = 2,       /* also produce grad_K (needs SNDE_LOSS_GRAD) */
    SNDE_LOSS_NO_PHOTO = 4,     /* skip the warp/photometric term */
    SNDE_LOSS_NO_SMOOTH = 8,    /* skip the smoothness term */
    SNDE_LOSS_DISPARITY = 16    /* the `depth` maps hold DISPARITIES (DispNetS heads, models/DispNetS.py:113): the kernel uses
                                   depth = 1/disp (models/baseline.py:835) and grad_depth receives d loss / d disparity
                                   = -(g * depth^2), what autograd's reciprocal backward computes (section 8(f2)) */
};

typedef struct SndeLossArgs {
    int32_t B, S, num_scales;
    int32_t flags;              /* SNDE_LOSS_* */
    int32_t padding_mode;       /* SndePadding */
    float alpha;                /* SSIM weight (reference: simi_alpha = 0.85) */
    const float *poses;         /* (B,S,6) contiguous [tx,ty,tz,rx,ry,rz] */
    SndeLossScale scale[SNDE_MAX_SCALES];
    float *loss_out;            /* 3 + 3*num_scales floats: total, rigid_warp, disp_smooth, then per scale (mean fwd_err, mean bwd_err, smooth_loss) */
    float *grad_poses;          /* (B,S,6) contiguous, written when SNDE_LOSS_GRAD */
    float *grad_K;              /* (B,num_scales,3,3) contiguous, written when SNDE_LOSS_GRAD_K */
    void *workspace;            /* device scratch, >= snde_loss_workspace_bytes(), 16-byte aligned; contents undefined on entry */
    size_t workspace_bytes;
} SndeLossArgs;

/* sizeof(SndeLossArgs) as compiled: lets an FFI binding verify its struct layout */
SNDE_API size_t snde_sizeof_loss_args(void);
SNDE_API size_t snde_loss_workspace_bytes(int B, int S, int num_scales);
/* launches: 3 kernels (matrix prep + zeroing of the accumulators, persistent fused strip kernel over every scale/direction/strip, finalize),
 * chained by programmatic dependent launch: each waits for its predecessor on the stream with griddepcontrol.wait */
SNDE_API int snde_loss_fwdbwd(const SndeLossArgs *args, void *stream);
/* number of strip tasks (owner x 60-column strip x row segment) the fused kernel works through for these shapes on a
 * B200 (host only; used by tests/bench) */
SNDE_API int64_t snde_loss_num_tiles(const SndeLossArgs *args);

/* ------------------------------------------------------------------------------------------
 * Reference-signature pieces (each materialises its output exactly like the reference).
 * ---------------------------------------------------------------------------------------- */

/* pose_vec2mat + optional inverse + K4 @ pose, and inverse(K): the per-sample matrices of
 * compute_rigid_flow (utils_edited.py:343-357, :231).  poses (B,6) with element stride
 * pose_stride between samples, K (B,3,3) contiguous -> P (B,3,4) rows 0..2 of the 4x4
 * projection, Kinv (B,3,3). */
SNDE_API int snde_pose_matrices(int B, const float *poses, int64_t pose_stride, const float *K, int reverse_pose,
                       float *P, float *Kinv, void *stream);
/* backward of the above: gP (B,3,4), gKinv (B,3,3) doubles (accumulators of snde_rigid_flow_bwd) ->
 * grad_poses (B,6) contiguous [written], grad_K (B,3,3) [written, may be NULL] */
SNDE_API int snde_pose_matrices_bwd(int B, const float *poses, int64_t pose_stride, const float *K, int reverse_pose,
                           const double *gP, const double *gKinv, float *grad_poses, float *grad_K, void *stream);

/* compute_rigid_flow's per-pixel part (utils_edited.py:331-362): depth (B,h,w) -> flow (B,h,w,2) contiguous */
SNDE_API int snde_rigid_flow_fwd(int B, int h, int w, SndeMap depth, const float *Kinv, const float *P,
                        const float *xtab, const float *ytab, float *flow, void *stream);
/* gflow (B,h,w,2) contiguous -> gdepth (written); gP (B,3,4), gKinv (B,3,3) double accumulators (+=; zero them first; gKinv may be NULL) */
SNDE_API int snde_rigid_flow_bwd(int B, int h, int w, SndeMap depth, const float *Kinv, const float *P,
                        const float *xtab, const float *ytab, const float *gflow,
                        SndeMapOut gdepth, double *gP, double *gKinv, void *stream);

/* bilinear_sampler (utils_edited.py:402-486): img (N,C,hs,ws) strided, coords (N,ht,wt,2) contiguous
 * -> out (N,ht,wt,C) contiguous (NHWC, like the reference).  add_grid != 0: coords are a FLOW and the
 * meshgrid tables (of wt / ht entries) are added first == flow_warp (utils_edited.py:386-399). */
SNDE_API int snde_bilinear_sampler_fwd(int N, int C, int hs, int ws, SndeImage img, int ht, int wt, const float *coords,
                              int add_grid, const float *xtab, const float *ytab, int padding_mode,
                              float *out, float *valid /* optional (N,ht,wt) */, void *stream);
/* gout (N,ht,wt,C) -> gcoords (N,ht,wt,2) written; gimg optional (N,C,hs,ws strided; += with atomics, zero it first) */
SNDE_API int snde_bilinear_sampler_bwd(int N, int C, int hs, int ws, SndeImage img, int ht, int wt, const float *coords,
                              int add_grid, const float *xtab, const float *ytab, int padding_mode,
                              const float *gout, float *gcoords, SndeImageOut gimg, void *stream);

/* Deterministic variant of the image-gradient scatter of snde_bilinear_sampler_bwd (the float atomics of the default
 * mode land in an arbitrary order, so its last bits vary from run to run).  Two passes: every contribution w*g is
 * added as a 64-bit fixed-point integer -- integer addition is associative, the sum does not depend on the order --
 * with a scale derived on the device from max|gout| (quantum = 2^-38 of it, room for 2^24 contributions per pixel);
 * the second pass converts the sums and adds them to gimg.  scratch: device, (N*C*hs*ws + 1) int64, overwritten.
 * gcoords is written exactly as by the default mode. */
SNDE_API int snde_bilinear_sampler_bwd_deterministic(int N, int C, int hs, int ws, SndeImage img, int ht, int wt, const float *coords,
                                            int add_grid, const float *xtab, const float *ytab, int padding_mode,
                                            const float *gout, float *gcoords, SndeImageOut gimg, long long *scratch, void *stream);

/* image_similarity(alpha,x,y) = alpha*DSSIM(x,y) + (1-alpha)*|x-y| (loss_functions.py:4-6, utils_edited.py:121-141).
 * alpha = 1 gives DSSIM.  x,y (N,3,h,w) strided -> err (N,h,w,3) contiguous NHWC */
SNDE_API int snde_image_similarity_fwd(int N, int h, int w, float alpha, SndeImage x, SndeImage y, float *err, void *stream);
/* gerr (N,h,w,3) contiguous -> gx, gy (N,h,w,3) contiguous, written (either may be NULL) */
SNDE_API int snde_image_similarity_bwd(int N, int h, int w, float alpha, SndeImage x, SndeImage y, const float *gerr,
                              float *gx, float *gy, void *stream);

/* smooth_loss(depth, image) (loss_functions.py:8-24): depth (M,h,w) strided, image (M,3,h,w) strided.
 * sums: 2 doubles (sum|dx*wx|, sum|dy*wy|), += (zero first); loss = sums[0]/(M*h*(w-1)) + sums[1]/(M*(h-1)*w) */
SNDE_API int snde_smooth_loss_fwd(int M, int h, int w, SndeMap depth, SndeImage img, double *sums, void *stream);
/* gdepth (M,h,w) written = gscale * dloss/ddepth, gscale read from a DEVICE float (the upstream gradient) */
SNDE_API int snde_smooth_loss_bwd(int M, int h, int w, SndeMap depth, SndeImage img, const float *gscale_dev,
                         SndeMapOut gdepth, void *stream);

/* ------------------------------------------------------------------------------------------
 * Caller-side prep just before the path: the image part of
 * GeoNetModel.iter_data_preparation (models/baseline.py:761-767: uint8 -> float, * 1/255, * 2 - 1)
 * fused with scale_pyramid (utils/utils_edited.py:89-112: F.interpolate(mode='area') to
 * int(H/2^s) x int(W/2^s)).  One pass: every byte is read once, every level written once.
 *   frames : uint8 (n,3,H,W) addressed through element (= byte) strides (sn, sc, sh, sw), so
 *            NHWC camera buffers work too
 *   levels : HOST array of num_scales DEVICE pointers; level s is planar contiguous fp32
 *            (n,3,H>>s,W>>s) = the mean over 2^s x 2^s blocks of the normalised pixels
 * H and W must be multiples of 2^(num_scales-1) (the sizes at which the reference's
 * int(H/2^s) pyramid agrees with its networks); else SNDE_ERR_SHAPE. */
SNDE_API int snde_frames_to_pyramid(int n, int H, int W, int num_scales, const uint8_t *frames, int64_t sn, int64_t sc,
                           int64_t sh, int64_t sw, float *const *levels, void *stream);
/* Same, for a frame block whose channel axis holds groups * 3 channels -- the DataLoader's source block (B, 3*S, H, W),
 * models/baseline.py:772-775: group g of frame b becomes image g * n + b of every level (the source-major order
 * n = src * B + b of the loss path), so ONE launch prepares all S sources.  levels[s] is (groups*n, 3, H>>s, W>>s).
 * num_scales <= SNDE_MAX_SCALES (6). */
SNDE_API int snde_frames_to_pyramid_grouped(int n, int groups, int H, int W, int num_scales, const uint8_t *frames, int64_t sn,
                                            int64_t sc, int64_t sh, int64_t sw, float *const *levels, void *stream);

/* ------------------------------------------------------------------------------------------
 * Depth <-> normal consistency (BASELINE config 3).  The reference holds the inference-only GeoNet
 * (Qi et al. 2018) refinement inside NNET.forward -- depth-to-normal, models/baseline.py:350-446, and normal-to-depth,
 * :495-547, over a k x k window (k = 9) with dilation `rate` (4) -- with a scrambled patch extraction (F.unfold on NHWC
 * tensors), a hard-coded batch of 4, no loss and no backward.  These entry points implement the INTENDED computation
 * (spatial windows, zero padding) and define a consistency loss on it; parity is pinned by oracle/normal_depth.py.
 *   grid     g_i = (xgrid[x_i], ygrid[y_i], 1)   (the reference: linspace(-0.6, 0.6, w), linspace(-0.4, 0.4, h), :308-317)
 *   D2N      n'_i = normalize((det M > 1e-5 ? M : I)^-1 b),  M = sum_j P_j P_j^T, b = sum_j P_j over the window's j with
 *            n_j . n_i > thresh,  P_j = d_j g_j
 *   N2D      d'_i = clamp( sum_j [n_j.g_i > 1e-5] (n_j.g_j)/(n_j.g_i) d_j (n_j.n_i) / (sum_j [..] (n_j.n_i) + 1e-5), 0, clamp_max)
 *   loss     w_depth * mean_i |d_i - d'_i| + w_normal * mean_i (1 - n_i . n'_i)
 * snde_dn_consistency_fwd writes loss_out = {total, depth term, normal term} and, when given, n', d' and the per-pixel
 * scratch `sq` (2 floats per pixel) that snde_dn_consistency_bwd needs; the backward writes d loss / d depth * (*gscale_dev)
 * (through d_i and through the d_j inside d'_i; normals are data and n' is a constant: the normal term has no gradient).
 * ---------------------------------------------------------------------------------------- */
typedef struct SndeDnArgs {
    int32_t B, h, w;
    int32_t k, rate;                    /* window: k x k taps (odd, <= 15), `rate` pixels apart (<= 16) */
    float thresh, clamp_max;            /* reference: 0.95, 10.0 */
    float w_depth, w_normal;
    SndeMap depth;                      /* (B,h,w) linear depth */
    SndeImage normal;                   /* (B,3,h,w) unit normals, n . ray > 0 */
    const float *xgrid, *ygrid;         /* device arrays (w), (h) */
    SndeImageOut normal_from_depth;     /* optional output (B,3,h,w) */
    SndeMapOut depth_from_normal;       /* optional output (B,h,w) */
    float *loss_out;                    /* device, 3 floats */
    float *sq;                          /* optional device scratch, 2*B*h*w floats (required by the backward) */
    double *sums;                       /* device scratch, 2 doubles (zeroed by the call) */
} SndeDnArgs;
SNDE_API int snde_dn_consistency_fwd(const SndeDnArgs *args, void *stream);
SNDE_API int snde_dn_consistency_bwd(const SndeDnArgs *args, const float *gscale_dev, SndeMapOut grad_depth, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SNDE_H_ */
