// Host-side launch helpers shared by the .cu translation units.
#pragma once
#include "common.cuh"

static inline int px_clamp_blocks(size_t work_items, int per_block, int max_blocks) {
  size_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (max_blocks > 0 && b > (size_t)max_blocks) b = max_blocks;
  if (b > PX_MAX_BLOCKS) b = PX_MAX_BLOCKS;
  return (int)b;
}

// entry p = pointer of rank (rank + p) % world  (entry 0 is local)
static inline PeerPtrs px_rotate(const void* const* ptrs, int rank, int world) {
  PeerPtrs R{};
  for (int p = 0; p < world; ++p) R.p[p] = const_cast<void*>(ptrs[(rank + p) % world]);
  return R;
}

#define PX_DISPATCH_WORLD(world, M, T)                      \
  switch (world) {                                          \
    case 1: M(T, 1); break;                                 \
    case 2: M(T, 2); break;                                 \
    case 3: M(T, 3); break;                                 \
    case 4: M(T, 4); break;                                 \
    case 5: M(T, 5); break;                                 \
    case 6: M(T, 6); break;                                 \
    case 7: M(T, 7); break;                                 \
    case 8: M(T, 8); break;                                 \
    default: break;                                         \
  }
