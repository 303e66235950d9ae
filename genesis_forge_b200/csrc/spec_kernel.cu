// One specialised instance of the fused post-physics kernel.
//
// Compiled by genesis_forge_b200/spec.py with -DGFB_SPEC_HEADER="<generated header>".  The header
// defines, as compile-time constants, the STRUCTURE of one term table and its slab plan:
//     namespace gfb_spec { TILE, PHASES, kP (gfb_program_head, live values zeroed), kPlan (gfb::Plan) }
// post_kernel.cuh then reads structure from those constants (term loops unroll, switches fold,
// shared-memory offsets become immediates) and live values from the kernel parameters as usual.
#include <cuda_runtime.h>

#include "../../include/gfb200.h"
#include "plan.h"

#include GFB_SPEC_HEADER

#define GFB_SPEC 1
#include "post_kernel.cuh"

namespace {
// host copies of the structure for the library's match test
const gfb_program_head h_canon = gfb_spec::kP_host;
const gfb::Plan h_plan = gfb_spec::kPlan_host;
int smem_attr = 0;
}  // namespace

extern "C" int gfb_spec_info(int* tile, unsigned* phases, const void** canon, const void** plan, int* head_bytes,
                             int* plan_bytes) {
  *tile = gfb_spec::TILE;
  *phases = gfb_spec::PHASES;
  *canon = &h_canon;
  *plan = &h_plan;
  *head_bytes = (int)sizeof(gfb_program_head);
  *plan_bytes = (int)sizeof(gfb::Plan);
  return 0;
}

extern "C" int gfb_spec_blocks_per_sm(unsigned smem) {
  if ((int)smem > 44 * 1024 && (int)smem > smem_attr) {
    if (cudaFuncSetAttribute(gfb::post_kernel<gfb_spec::TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return 0;
    smem_attr = (int)smem;
  }
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gfb::post_kernel<gfb_spec::TILE>, gfb_spec::TILE, smem) != cudaSuccess)
    n = 0;
  return n;
}

extern "C" int gfb_spec_launch(const gfb::KParams* kp, int grid, unsigned smem, void* stream) {
  if ((int)smem > 44 * 1024 && (int)smem > smem_attr) {
    if (cudaFuncSetAttribute(gfb::post_kernel<gfb_spec::TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return 1;
    smem_attr = (int)smem;
  }
  gfb::post_kernel<gfb_spec::TILE><<<grid, gfb_spec::TILE, smem, static_cast<cudaStream_t>(stream)>>>(*kp);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
