// One translation unit per row-block count: nvcc -DCAFE_RB=<1..8> prune_inst.cu (see __graft_entry__.build()).
// Geometries compiled (gw, ng, cps, pw):
//   every RB      (4, 2, 2, 2)   two consumer groups                                  matrix size <= 256
//   RB <= 5       (4, 3, 4, 1)   three consumer groups, 40 KB ring stages (the default where it fits)   matrix size <= 160
//   RB <= 5       (4, 3, 2, 2)   three consumer groups, 20 KB ring stages
//   RB >= 5       (8, 1, 1, 1)   one group of eight warps, 64*RB rows                 matrix size 257 .. 512
//   RB == 5       a few more, selected by the CAFE_B200_GEOM environment variable (measurements in profiles/)
#include <cstdio>
#include <cstdlib>

#include "prune_launch.h"
#include "prune.cuh"

#ifndef CAFE_RB
#error "compile with -DCAFE_RB=<1..8>"
#endif

namespace cafe {

namespace {

template <int RB, int GW, int NG, int CPS, int PW>
cudaError_t go(const PruneParams& p, int grid, int smem, cudaStream_t s, bool attr_only)
{
    if (attr_only) return cudaFuncSetAttribute(prune_kernel<RB, GW, NG, CPS, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    prune_kernel<RB, GW, NG, CPS, PW><<<grid, pg_threads(NG, GW, PW), smem, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && getenv("CAFE_B200_DEBUG")) {
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, prune_kernel<RB, GW, NG, CPS, PW>);
        fprintf(stderr, "prune_kernel<%d,%d,%d,%d,%d>: %s; threads %d regs %d maxThreads %d dyn smem requested %d allowed %d static %zu local %zu\n", RB, GW, NG,
                CPS, PW, cudaGetErrorString(e), pg_threads(NG, GW, PW), a.numRegs, a.maxThreadsPerBlock, smem, a.maxDynamicSharedSizeBytes,
                a.sharedSizeBytes, a.localSizeBytes);
    }
    return e;
}

}  // namespace

#define CAFE_CAT2(a, b) a##b
#define CAFE_CAT(a, b) CAFE_CAT2(a, b)
#define GEOM(GW, NG, CPS, PW) \
    if (g.gw == GW && g.ng == NG && g.cps == CPS && g.pw == PW) return go<CAFE_RB, GW, NG, CPS, PW>(p, grid, smem, s, attr_only);

cudaError_t CAFE_CAT(prune_launch_rb, CAFE_RB)(const PruneGeom& g, const PruneParams& p, int grid, int smem, cudaStream_t s, bool attr_only)
{
    GEOM(4, 2, 2, 2)
#if CAFE_RB <= 5
    GEOM(4, 3, 4, 1)
    GEOM(4, 3, 2, 2)
#endif
#if CAFE_RB >= 5
    GEOM(8, 1, 1, 1)
#endif
#if CAFE_RB == 5
    GEOM(4, 3, 2, 1)
    GEOM(4, 3, 3, 2)
    GEOM(4, 3, 4, 2)
    GEOM(4, 2, 4, 2)
#endif
    return cudaErrorInvalidConfiguration;
}

#if CAFE_RB == 1
bool prune_geometry_compiled(const PruneGeom& g)
{
    if (g.rb < 1 || g.rb > 8) return false;
    if (g.gw == 4 && g.ng == 2 && g.cps == 2 && g.pw == 2) return true;
    if (g.gw == 4 && g.ng == 3 && g.cps == 2 && g.pw == 2) return g.rb <= 5;
    if (g.gw == 4 && g.ng == 3 && g.cps == 4 && g.pw == 1) return g.rb <= 5;
    if (g.gw == 8 && g.ng == 1 && g.cps == 1 && g.pw == 1) return g.rb >= 5;
    if (g.rb == 5 && g.gw == 4)
        return (g.ng == 3 && g.cps == 2 && g.pw == 1) || (g.ng == 3 && g.cps == 3 && g.pw == 2) || (g.ng == 3 && g.cps == 4 && g.pw == 2) ||
               (g.ng == 2 && g.cps == 4 && g.pw == 2);
    return false;
}
#endif

}  // namespace cafe
