// snde_team.cu -- the TEAM kernel: the fused loss's strip program run by five specialised warps per strip (snde_team.cuh),
// 20 warps per SM, register budgets per role (setmaxnreg).  Its own translation unit: the rings and the per-row barrier have
// other sizes here than in the pair kernel (snde_fused.cu), so everything of snde_strip.cuh lives in its own namespace.
#define snde snde_team
#define SNDE_G_SLOTS 8
#define SNDE_TEAM_THREADS 160
#include <cuda_runtime.h>

#include "snde.h"
#include "snde_team.cuh"

using namespace snde_team;

#ifndef SNDE_TEAM_REGS_GEO
#define SNDE_TEAM_REGS_GEO 128
#endif
#ifndef SNDE_TEAM_REGS_SSIM
#define SNDE_TEAM_REGS_SSIM 64
#endif
#ifndef SNDE_TEAM_LB
#define SNDE_TEAM_LB TeamCfg::NT
#endif
#define SNDE_STR2(x) #x
#define SNDE_STR(x) SNDE_STR2(x)

// Warp w of role group g (= threads 128 g .. 128 g + 127) serves team w: the five warps of a team have the same index modulo
// four, i.e. they share a scheduler.  Role groups are warpgroups, which is what setmaxnreg needs.
template <bool DISP>
__global__ void __launch_bounds__(SNDE_TEAM_LB, 1) loss_team_kernel(const __grid_constant__ DevLoss p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
    const int role = warp / TeamCfg::TEAMS, team = warp % TeamCfg::TEAMS;
    PairSmem &sm = reinterpret_cast<PairSmem *>(smem_raw)[team];
    const int bar = 1 + team;
#if !defined(SNDE_TEAM_NO_SETMAXNREG)
    if (role >= 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 " SNDE_STR(SNDE_TEAM_REGS_SSIM) ";");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 " SNDE_STR(SNDE_TEAM_REGS_GEO) ";");
#endif
    if (role == 0) {
        for (;;) {
            if (lane == 0) sm.task = atomicAdd(p.counter, 1);
            pair_barrier(lane, bar);
            const int task = __shfl_sync(0xffffffffu, sm.task, 0);
            if (task >= p.total_tasks) break;
#if !defined(SNDE_TEAM_ONLY_ROLE) || SNDE_TEAM_ONLY_ROLE == 0
            team_geo_fwd_role<true, DISP>(p, task, lane, sm, bar);
#endif
        }
    } else if (role == 1) {
        for (;;) {
            pair_barrier(lane, bar);
            const int task = __shfl_sync(0xffffffffu, sm.task, 0);
            if (task >= p.total_tasks) break;
#if !defined(SNDE_TEAM_ONLY_ROLE) || SNDE_TEAM_ONLY_ROLE == 1
            team_geo_bwd_role<true, DISP>(p, task, lane, sm, bar);
#endif
        }
    } else {
        const int ch = role - 2;
        for (;;) {
            pair_barrier(lane, bar);
            const int task = __shfl_sync(0xffffffffu, sm.task, 0);
            if (task >= p.total_tasks) break;
#if !defined(SNDE_TEAM_ONLY_ROLE) || SNDE_TEAM_ONLY_ROLE == 2
            team_ssim_role<true, DISP>(p, task, lane, sm, bar, ch);
#endif
        }
    }
}

// launched by snde_loss_fwdbwd (snde_fused.cu) for the lean fast-path instances; `d` was built by build_dev_loss there (the
// DevLoss layout is the same in both translation units)
extern "C" int snde_launch_team_kernel(const void *dev_loss, int disparity, int sm_count, long long tiles, void *stream) {
    const DevLoss &d = *reinterpret_cast<const DevLoss *>(dev_loss);
    const size_t smem = sizeof(PairSmem) * TeamCfg::TEAMS;
    long long ctas = sm_count;
    const long long need = (tiles + TeamCfg::TEAMS - 1) / TeamCfg::TEAMS;
    if (ctas > need) ctas = need;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto kern = disparity ? loss_team_kernel<true> : loss_team_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SNDE_ERR_CUDA;
    kern<<<(unsigned)ctas, TeamCfg::NT, smem, st>>>(d);
    return cudaGetLastError() == cudaSuccess ? SNDE_OK : SNDE_ERR_CUDA;
}
