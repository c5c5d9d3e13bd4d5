// Stand-in for <cooperative_groups.h> (tests only): grid-wide sync of the host emulation, see cuda_runtime.h here.
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group {
  void sync() const { emu::grid_sync(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
