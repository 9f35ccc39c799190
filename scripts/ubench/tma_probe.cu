// TMA row-load probe: which of (descriptor nested deep in a large __grid_constant__ struct, 4-D boxes, negative start
// columns, boxes wider than the tensor) the hardware accepts.  usage: tma_probe PAD_KIND RANK W X0
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct alignas(64) TM { unsigned long long o[16]; };
template <int PAD> struct Args { unsigned char pad[PAD]; TM tm; float *out; int rank, x0, y, n; };
__device__ unsigned su32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int PAD>
__global__ void probe(const __grid_constant__ Args<PAD> a) {
    extern __shared__ __align__(128) unsigned char raw[];
    float *dst = reinterpret_cast<float *>(raw);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(raw + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const int bytes = a.rank == 4 ? 768 : 256;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(bytes) : "memory");
        if (a.rank == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(su32(dst)), "l"(&a.tm), "r"(a.x0), "r"(a.y), "r"(0), "r"(a.n), "r"(su32(bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(su32(dst)), "l"(&a.tm), "r"(a.x0), "r"(a.y), "r"(a.n), "r"(su32(bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(su32(bar)) : "memory");
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) a.out[i] = dst[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int PAD> int run(int rank, int w, int x0) {
    const int h = 6, C = 3, n = 2;
    std::vector<float> host((size_t)w * h * C * n);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (float)i + 1.f;
    float *img, *out; cudaMalloc(&img, host.size() * 4); cudaMalloc(&out, 768 * 4);
    cudaMemcpy(img, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Args<PAD> a; a.out = out; a.rank = rank; a.x0 = x0; a.y = 1; a.n = 1;
    cuuint64_t dims[4], str[3]; cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    if (rank == 4) { dims[0] = w; dims[1] = h; dims[2] = C; dims[3] = n; str[0] = w * 4; str[1] = (cuuint64_t)w * h * 4; str[2] = (cuuint64_t)w * h * C * 4; box[0] = 64; box[1] = 1; box[2] = 3; box[3] = 1; }
    else { dims[0] = w; dims[1] = h; dims[2] = C * n; str[0] = w * 4; str[1] = (cuuint64_t)w * h * 4; box[0] = 64; box[1] = 1; box[2] = 1; }
    CUresult r = ((Enc)fp)((CUtensorMap *)&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, img, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("pad %d rank %d w %d x0 %d: encode rc %d sizeof(args) %zu\n", PAD, rank, w, x0, (int)r, sizeof(a));
    if (r) return 1;
    probe<PAD><<<1, 64, 2048>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e) return 2;
    std::vector<float> got(768); cudaMemcpy(got.data(), out, 768 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    const int nch = rank == 4 ? 3 : 1;
    for (int c = 0; c < nch; ++c)
        for (int i = 0; i < 64; ++i) {
            const int x = x0 + i;
            const size_t plane = rank == 4 ? (size_t)(1 * C + c) : (size_t)1;
            const float want = (x >= 0 && x < w) ? host[plane * w * h + (size_t)1 * w + x] : 0.f;
            if (got[c * 64 + i] != want) { if (bad < 4) printf("  mismatch c %d i %d got %g want %g\n", c, i, got[c * 64 + i], want); ++bad; }
        }
    printf("  %s (%d mismatches)\n", bad ? "WRONG" : "ok", bad);
    return bad ? 3 : 0;
}
int main(int argc, char **argv) {
    const int kind = atoi(argv[1]), rank = atoi(argv[2]), w = atoi(argv[3]), x0 = atoi(argv[4]);
    return kind == 0 ? run<64>(rank, w, x0) : kind == 1 ? run<2048>(rank, w, x0) : run<5056>(rank, w, x0);
}
