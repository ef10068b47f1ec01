// tile.cu -- fused tile-resident multi-gate pass (placeholder until the kernel lands).
#include "common.cuh"

namespace b200 {

int launch_tile_pass(b200_state*, const uint64_t*, const uint64_t*, int, const double*) {
  set_error("B200_OP_TILE: fused tile pass is not built into this library");
  return 1;
}

}  // namespace b200
