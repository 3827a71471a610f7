// Launchers of the pruning kernel's instantiations.  The kernel is compiled once per row-block count RB (prune_inst.cu
// with -DCAFE_RB=1..8, in parallel), each translation unit exporting one function that selects among the geometries
// compiled for that RB.  attr_only = only raise the kernel's dynamic shared-memory limit (done once at create).
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace cafe {

struct PruneGeom;

#define CAFE_DECLARE_PRUNE_RB(RB) \
    cudaError_t prune_launch_rb##RB(const PruneGeom& g, const PruneParams& p, int grid, int smem, cudaStream_t s, bool attr_only);
CAFE_DECLARE_PRUNE_RB(1)
CAFE_DECLARE_PRUNE_RB(2)
CAFE_DECLARE_PRUNE_RB(3)
CAFE_DECLARE_PRUNE_RB(4)
CAFE_DECLARE_PRUNE_RB(5)
CAFE_DECLARE_PRUNE_RB(6)
CAFE_DECLARE_PRUNE_RB(7)
CAFE_DECLARE_PRUNE_RB(8)
#undef CAFE_DECLARE_PRUNE_RB

// true when (rb, gw, ng, cps, pw) is one of the compiled geometries
bool prune_geometry_compiled(const PruneGeom& g);

}  // namespace cafe
