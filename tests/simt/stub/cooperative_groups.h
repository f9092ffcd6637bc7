// Stub for the host SIMT interpreter build (tests/simt): grid-wide barriers need co-resident
// blocks, the interpreter runs blocks one after the other -- the cooperative kernel is never launched.
#pragma once
#include <stdlib.h>
namespace cooperative_groups {
struct grid_group {
  void sync() const { abort(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
