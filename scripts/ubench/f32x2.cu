// Micro-benchmark: issue rate of FFMA vs FFMA2 / FADD2 / FMUL2 (packed fp32 pairs, sm_100a) and of SHFL / LDS mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#define ILP 8
template <int MODE> __global__ void k(float *out, int iters, float s) {
    float a[ILP]; u64 p[ILP]; unsigned q[ILP], q2[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = pack(a[i], a[i] + 1.f); q[i] = threadIdx.x + i; q2[i] = q[i] * 3u; }
    const unsigned qs = (unsigned)iters | 1u;
    const u64 ps = pack(s, s * 0.5f), pt = pack(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], s, 1e-3f);
            if (MODE == 1) p[i] = ffma2(p[i], ps, pt);
            if (MODE == 2) p[i] = fadd2(p[i], pt);
            if (MODE == 3) p[i] = fmul2(p[i], ps);
            if (MODE == 4) { a[i] = fmaf(a[i], s, 1e-3f); p[i] = ffma2(p[i], ps, pt); }   // mix 1:1
            if (MODE == 5) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1) + 1e-3f;        // SHFL + FADD
            if (MODE == 6) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            // does a packed op hold the ISSUE port for two cycles, or only the FMA pipe?  ALU-pipe companions (LOP3) tell:
            if (MODE == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(qs), "r"(q2[i]));                        // ALU only
            if (MODE == 8) { p[i] = ffma2(p[i], ps, pt); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(qs), "r"(q2[i])); }   // FFMA2 + 1 ALU
            if (MODE == 9) { p[i] = ffma2(p[i], ps, pt); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(qs), "r"(q2[i]));
                             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q2[i]) : "r"(qs), "r"(q[i])); }                        // FFMA2 + 2 ALU
            if (MODE == 10) { a[i] = fmaf(a[i], s, 1e-3f); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(qs), "r"(q2[i])); } // FFMA + 1 ALU
        }
    }
    float acc = 0.f;
    for (int i = 0; i < ILP; ++i) { float x, y; unpack(p[i], x, y); acc += a[i] + x + y + (float)(q[i] ^ q2[i]); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char *name, int per_iter_inst) {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 2, threads = 512;
    k<MODE><<<blocks, threads>>>(out, 100, 0.999f);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * threads / 32 * iters * ILP * per_iter_inst;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %7.1f G warp-inst/s  = %.2f warp-inst/clk/SM at %d MHz (nominal)\n", name, ms, winst / ms / 1e6,
           winst / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
// dependent-issue latency: ONE warp per SM runs a single dependent chain; cycles per instruction = latency
template <int MODE> __global__ void lat(float *out, int iters, float s) {
    float a = threadIdx.x * 1e-3f; u64 p = pack(a, a + 1.f);
    const u64 ps = pack(s, s * 0.5f), pt = pack(1e-3f, 2e-3f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a = fmaf(a, s, 1e-3f);
            if (MODE == 1) p = ffma2(p, ps, pt);
            if (MODE == 2) p = fadd2(p, pt);
            if (MODE == 3) { a = fmaf(a, s, 1e-3f); float x, y; unpack(p, x, y); p = ffma2(pack(a, y), ps, pt); unpack(p, x, y); a = x; }   // FFMA -> FFMA2 -> FFMA
            if (MODE == 4) a = __shfl_xor_sync(0xffffffffu, a, 1);
            if (MODE == 5) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); a = r; }
        }
    }
    long long t1 = clock64();
    float x, y; unpack(p, x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + x + y;
    if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long *>(out + 8192)[0] = t1 - t0;
}
template <int MODE> void runlat(const char *name, int inst_per_iter) {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int iters = 4000;
    lat<MODE><<<148, 32>>>(out, iters, 0.999f); cudaDeviceSynchronize();
    lat<MODE><<<148, 32>>>(out, iters, 0.999f); cudaDeviceSynchronize();
    long long clk; cudaMemcpy(&clk, out + 8192, 8, cudaMemcpyDeviceToHost);
    printf("latency %-24s %6.2f cycles per dependent instruction\n", name, (double)clk / ((double)iters * 16 * inst_per_iter));
    cudaFree(out);
}
int main() {
    runlat<0>("FFMA", 1); runlat<1>("FFMA2", 1); runlat<2>("FADD2", 1); runlat<3>("FFMA+FFMA2 alternating", 2); runlat<4>("SHFL", 1); runlat<5>("MUFU.RCP", 1);
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FADD2", 1); run<3>("FMUL2", 1); run<4>("FFMA+FFMA2", 2); run<5>("SHFL+FADD", 2); run<6>("MUFU.RCP", 1);
    run<7>("LOP3 (ALU pipe)", 1); run<8>("FFMA2 + 1 LOP3", 2); run<9>("FFMA2 + 2 LOP3", 3); run<10>("FFMA + 1 LOP3", 2);
    return 0;
}
