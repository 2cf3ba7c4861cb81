// Sparse path: remote-gather lookup, local aggregation (dedup), P2P push to
// the owning rank, owner-side accumulate + sparse optimizer, async (Hogwild)
// remote apply.
//
// What this replaces in the reference: worker GPU →
// local-chief CPU SparseConditionalAccumulator → gRPC → PS CPU accumulator
// (sorted two-pointer merge, whole value tensor re-allocated per apply,
// tensorflow/core/kernels/sparse_conditional_accumulator.h:192-319) → chief
// take_grad → serial CPU SparseApplyAdagrad row loop
// (tensorflow/core/kernels/training_ops.cc:1338-1351) → token queues; and for
// lookups dynamic_partition → per-shard PS CPU gather → gRPC → dynamic_stitch
// (tensorflow/python/ops/embedding_ops.py:151-209,
//  gather_functor_gpu.cu.h:32-70, dynamic_partition_op_gpu.cu.cc:60-110).
//
// Step protocol (sync mode), all flags are monotonically increasing step
// numbers living in each rank's signal area:
//   lookup(t)  waits  applied[o] >= t-1  for every owner o      (rows fresh)
//   push(t)    writes rows+ids+count into owner's ring[src=me], then
//              st.release.sys pushed[me] = t at the owner
//   claim(t)   (owner) waits pushed[s] >= t for every source s, merges
//              duplicate rows across sources
//   apply(t)   (owner) optimizer on every touched row once, then publishes
//              applied[me] = t to every rank and bumps the local step counter
#include "common.cuh"
#include "launch.h"

struct TableGeom {
  int V, P, W, rows_per_part, D4;   // D4 = padded row length in float4 units
  int strategy;                     // 0 mod, 1 div
  int replicated;                   // 1: every rank holds the full table (AR mode)
  int extras, base;                 // div strategy
};

// per-rank sparse control block (local memory, one per table)
struct SparseCtl {
  uint32_t step;        // completed steps
  uint32_t push_done;   // CTA ticket counter (push kernel)
  uint32_t apply_done;  // CTA ticket counter (apply kernel)
  int32_t n_uniq;
  int32_t owner_cnt[PX_MAX_RANKS];
};

// flags inside the per-table symmetric header: [pushed[W] | applied[W] | cnt[W]]
#define PX_TBL_HDR_WORDS (3 * PX_MAX_RANKS)

__device__ __forceinline__ void geom_map(const TableGeom& g, int id, int& owner, int& local) {
  if (g.replicated) { owner = 0; local = id; return; }
  int p, idx;
  if (g.strategy == 0) { p = id % g.P; idx = id / g.P; }
  else {
    const int thr = g.extras * (g.base + 1);
    if (id < thr) { p = id / (g.base + 1); idx = id - p * (g.base + 1); }
    else { p = (id - g.extras) / max(g.base, 1); idx = id - (p * g.base + g.extras); }
  }
  owner = p % g.W;
  local = (p / g.W) * g.rows_per_part + idx;
}

__device__ __forceinline__ uint32_t hash_id(int id) {
  uint32_t x = (uint32_t)id * 2654435761u;
  return x ^ (x >> 15);
}

// ------------------------------------------------------------------ lookup
// out[i,:] = table_owner(ids[i])[local(ids[i]), :]; LPR lanes cooperate on a row.
template <typename IdT, typename OutT>
__global__ void __launch_bounds__(256)
px_sparse_lookup_kernel(const IdT* __restrict__ ids, int n, float* const* __restrict__ tables,
                        OutT* __restrict__ out, int32_t* __restrict__ pend_ids, TableGeom g,
                        const uint32_t* applied, const SparseCtl* ctl, int lpr, int wait) {
  if (wait) {
    if (threadIdx.x < g.W) {
      const uint32_t need = ctl->step;
      while ((int32_t)(ld_acquire_sys(applied + threadIdx.x) - need) < 0) { }
    }
    __syncthreads();
  }
  const int rows_per_block = blockDim.x / lpr;
  const int sub = threadIdx.x % lpr;
  for (int i = blockIdx.x * rows_per_block + threadIdx.x / lpr; i < n;
       i += gridDim.x * rows_per_block) {
    const long long idl = (long long)ids[i];
    const bool valid = idl >= 0 && idl < g.V;
    const int id = valid ? (int)idl : 0;
    if (pend_ids != nullptr && sub == 0) pend_ids[i] = valid ? id : -1;
    int owner, local;
    geom_map(g, id, owner, local);
    const float4* src = reinterpret_cast<const float4*>(tables[g.replicated ? 0 : owner]) +
                        (size_t)local * g.D4;
    OutT* dst = out + (size_t)i * g.D4 * 4;
    for (int c = sub; c < g.D4; c += lpr) {
      uint4 v = valid ? ld_v4(src + c) : make_uint4(0, 0, 0, 0);   // OOB -> zeros
      if (sizeof(OutT) == 4) {
        st_v4(reinterpret_cast<float4*>(dst) + c, v);
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(__uint_as_float(v.x), __uint_as_float(v.y));
        __nv_bfloat162 hi = __floats2bfloat162_rn(__uint_as_float(v.z), __uint_as_float(v.w));
        uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        *reinterpret_cast<uint2*>(reinterpret_cast<char*>(dst) + (size_t)c * 8) = o;
      }
    }
  }
}

// ------------------------------------------------------------------- dedup
// Outputs (compact, per unique id u): uniq_id[u], uniq_k[u] (index inside its
// owner bucket), uniq_cnt[u] (how many positions carry that id) and, per
// position i, pos2u[i] (its unique slot, -1 for padding).  ctl->n_uniq,
// ctl->owner_cnt[o].  (Arrays are named uniq_head/next at the ABI for
// historical reasons: uniq_head == uniq_cnt, next == pos2u.)

// Single-CTA variant: the hash table lives in shared memory ("local
// aggregation dedups indices in SMEM before shipping").  n <= smem capacity.
__global__ void __launch_bounds__(1024)
px_sparse_dedup_smem_kernel(const int32_t* __restrict__ pend_ids, int n, int hbits,
                            int32_t* __restrict__ uniq_id, int32_t* __restrict__ uniq_k,
                            int32_t* __restrict__ uniq_head, int32_t* __restrict__ next,
                            SparseCtl* ctl, TableGeom g, int dedup) {
  extern __shared__ int32_t smem[];
  const int H = 1 << hbits;
  int32_t* keys = smem;          // [H]
  int32_t* slot_u = smem + H;    // [H]
  __shared__ int s_nuniq;
  __shared__ int s_owner_cnt[PX_MAX_RANKS];
  for (int h = threadIdx.x; h < H; h += blockDim.x) keys[h] = -1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) uniq_head[i] = 0;
  if (threadIdx.x < PX_MAX_RANKS) s_owner_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_nuniq = 0;
  __syncthreads();
  // pass 1: insert
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int id = pend_ids[i];
    if (id < 0) continue;
    if (!dedup) {
      int owner, local; geom_map(g, id, owner, local);
      const int u = atomicAdd(&s_nuniq, 1);
      uniq_id[u] = id; uniq_k[u] = atomicAdd(&s_owner_cnt[owner], 1);
      uniq_head[u] = 1; next[i] = u;
      continue;
    }
    uint32_t h = hash_id(id) & (H - 1);
    while (true) {
      const int old = atomicCAS(&keys[h], -1, id);
      if (old == -1) {
        int owner, local; geom_map(g, id, owner, local);
        const int u = atomicAdd(&s_nuniq, 1);
        slot_u[h] = u; uniq_id[u] = id; uniq_k[u] = atomicAdd(&s_owner_cnt[owner], 1);
        break;
      }
      if (old == id) break;
      h = (h + 1) & (H - 1);
    }
  }
  __syncthreads();
  // pass 2: map positions to their unique entry and count them
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int id = pend_ids[i];
    if (id < 0) { next[i] = -1; continue; }
    if (!dedup) continue;
    uint32_t h = hash_id(id) & (H - 1);
    while (keys[h] != id) h = (h + 1) & (H - 1);
    const int u = slot_u[h];
    next[i] = u;
    atomicAdd(&uniq_head[u], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) ctl->n_uniq = s_nuniq;
  if (threadIdx.x < PX_MAX_RANKS) ctl->owner_cnt[threadIdx.x] = s_owner_cnt[threadIdx.x];
}

// Multi-CTA variant, hash table in global memory (L2-resident): pass A inserts,
// pass B links and clears the slots it visits.
__global__ void __launch_bounds__(256)
px_sparse_dedup_insert_kernel(const int32_t* __restrict__ pend_ids, int n, int hbits,
                              int32_t* keys, int32_t* slot_u, int32_t* __restrict__ uniq_id,
                              int32_t* __restrict__ uniq_k, int32_t* __restrict__ uniq_head,
                              int32_t* __restrict__ next, SparseCtl* ctl, TableGeom g, int dedup) {
  const int H = 1 << hbits;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int id = pend_ids[i];
    if (id < 0) { next[i] = -1; continue; }
    if (!dedup) {
      int owner, local; geom_map(g, id, owner, local);
      const int u = atomicAdd(&ctl->n_uniq, 1);
      uniq_id[u] = id; uniq_k[u] = atomicAdd(&ctl->owner_cnt[owner], 1);
      uniq_head[u] = 1; next[i] = u;
      continue;
    }
    uint32_t h = hash_id(id) & (H - 1);
    while (true) {
      const int old = atomicCAS(&keys[h], -1, id);
      if (old == -1) {
        int owner, local; geom_map(g, id, owner, local);
        const int u = atomicAdd(&ctl->n_uniq, 1);
        slot_u[h] = u; uniq_id[u] = id; uniq_k[u] = atomicAdd(&ctl->owner_cnt[owner], 1);
        uniq_head[u] = 0;
        break;
      }
      if (old == id) break;
      h = (h + 1) & (H - 1);
    }
  }
}
__global__ void __launch_bounds__(256)
px_sparse_dedup_link_kernel(const int32_t* __restrict__ pend_ids, int n, int hbits,
                            const int32_t* keys, const int32_t* slot_u, int32_t* uniq_head,
                            int32_t* __restrict__ next) {
  const int H = 1 << hbits;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int id = pend_ids[i];
    if (id < 0) continue;
    uint32_t h = hash_id(id) & (H - 1);
    while (ld_volatile_u32(reinterpret_cast<const uint32_t*>(keys) + h) != (uint32_t)id)
      h = (h + 1) & (H - 1);
    const int u = slot_u[h];
    next[i] = u;
    atomicAdd(&uniq_head[u], 1);
  }
}

// -------------------------------------------------------------------- push
// One warp per unique id: sum the rows of all positions carrying it (fp32),
// scale, and store the row + its local index into the owner's receive ring
// (or every rank's ring in replicated/AR mode) over NVLink.  The last CTA
// publishes counts and the `pushed` flag.  CH = row chunks of 32 float4 held
// in registers so the duplicate list is walked once (rows up to 512 floats);
// longer rows re-walk per group of CH chunks.
template <typename GradT>
__device__ __forceinline__ float4 ld_grad4(const GradT* base, size_t f4_index) {
  if (sizeof(GradT) == 4) {
    const uint4 v = ld_v4_stream(reinterpret_cast<const float4*>(base) + f4_index);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z),
                       __uint_as_float(v.w));
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(base) +
                                                    f4_index * 8);
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                       __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
  }
}

__device__ __forceinline__ void sparse_update4(int kind, float lr, float a, float b, float eps,
                                               float nesterov, const float4& g, float4& w,
                                               float4& s0, float4& s1) {
  float gg[4] = {g.x, g.y, g.z, g.w};
  float ww[4] = {w.x, w.y, w.z, w.w};
  float a0[4] = {s0.x, s0.y, s0.z, s0.w};
  float a1[4] = {s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float gi = gg[i];
    switch (kind) {
      case 0: ww[i] = fmaf(-lr, gi, ww[i]); break;
      case 1:
        a0[i] = fmaf(a, a0[i], gi);
        ww[i] = nesterov != 0.f ? fmaf(-lr, fmaf(a, a0[i], gi), ww[i]) : fmaf(-lr, a0[i], ww[i]);
        break;
      case 2:
        a0[i] = fmaf(gi, gi, a0[i]);
        ww[i] = fmaf(-lr * gi, rsqrtf(a0[i]), ww[i]);
        break;
      case 3:
        a0[i] = fmaf(a, a0[i], (1.f - a) * gi);
        a1[i] = fmaf(b, a1[i], (1.f - b) * gi * gi);
        ww[i] -= lr * a0[i] / (sqrtf(a1[i]) + eps);
        break;
      case 4:
        a0[i] = fmaf(a, a0[i], (1.f - a) * gi * gi);
        a1[i] = fmaf(b, a1[i], lr * gi * rsqrtf(a0[i] + eps));
        ww[i] -= a1[i];
        break;
    }
  }
  w = make_float4(ww[0], ww[1], ww[2], ww[3]);
  s0 = make_float4(a0[0], a0[1], a0[2], a0[3]);
  s1 = make_float4(a1[0], a1[1], a1[2], a1[3]);
}

// Kernel A — position-parallel (LPR lanes per position): a position whose id
// is unique in this batch goes straight to its destination (owner's ring, or a
// remote optimizer application in async mode); positions sharing an id are
// summed with vector atomics into a local fp32 staging row.  Duplicate-heavy
// (Zipfian) batches therefore cost O(1) depth instead of a serial list walk.
struct PushDst {
  char* const* rings; size_t ring_ids_off; int cap; int rank;           // sync: rings
  float* const* tables; float* const* slot0s; float* const* slot1s;      // async: remote apply
  const float* hp; int kind;
};

__device__ __forceinline__ void emit_row4(const PushDst& d, const TableGeom& g, int owner,
                                          int local, int k, int c, float4 v, bool async) {
  if (!async) {
    const uint4 o = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z),
                               __float_as_uint(v.w));
    if (g.replicated) {
      for (int p = 0; p < g.W; ++p) {
        const int q = (d.rank + p) % g.W;
        st_v4_stream(reinterpret_cast<float4*>(d.rings[q]) + ((size_t)d.rank * d.cap + k) * g.D4 + c, o);
      }
    } else {
      st_v4_stream(reinterpret_cast<float4*>(d.rings[owner]) + ((size_t)d.rank * d.cap + k) * g.D4 + c, o);
    }
  } else {
    float4* pw = reinterpret_cast<float4*>(d.tables[owner]) + (size_t)local * g.D4 + c;
    float4* p0 = d.slot0s ? reinterpret_cast<float4*>(d.slot0s[owner]) + (size_t)local * g.D4 + c : nullptr;
    float4* p1 = d.slot1s ? reinterpret_cast<float4*>(d.slot1s[owner]) + (size_t)local * g.D4 + c : nullptr;
    float4 w = *pw, s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
    if (p0) s0 = *p0;
    if (p1) s1 = *p1;
    sparse_update4(d.kind, d.hp[0], d.hp[1], d.hp[2], d.hp[3], d.hp[7], v, w, s0, s1);
    *pw = w;
    if (p0) *p0 = s0;
    if (p1) *p1 = s1;
  }
}

__device__ __forceinline__ void emit_id(const PushDst& d, const TableGeom& g, int owner, int local,
                                        int k) {
  if (g.replicated) {
    for (int p = 0; p < g.W; ++p)
      reinterpret_cast<int32_t*>(d.rings[p] + d.ring_ids_off)[(size_t)d.rank * d.cap + k] = local;
  } else {
    reinterpret_cast<int32_t*>(d.rings[owner] + d.ring_ids_off)[(size_t)d.rank * d.cap + k] = local;
  }
}

template <typename GradT, bool ASYNC>
__global__ void __launch_bounds__(256)
px_sparse_scatter_kernel(const GradT* __restrict__ pend_grads, int n,
                         const int32_t* __restrict__ pos2u, const int32_t* __restrict__ uniq_id,
                         const int32_t* __restrict__ uniq_k, const int32_t* __restrict__ uniq_cnt,
                         float* __restrict__ staging, const SparseCtl* ctl, PushDst d,
                         uint32_t* const* __restrict__ hdrs, TableGeom g, float scale, int lpr) {
  if (!ASYNC) {   // do not overwrite a ring the owner may still be draining
    if (threadIdx.x < g.W) {
      const uint32_t need = ctl->step;
      const uint32_t* applied = hdrs[d.rank] + PX_MAX_RANKS;
      while ((int32_t)(ld_acquire_sys(applied + threadIdx.x) - need) < 0) { }
    }
    __syncthreads();
  }
  const int rows_per_block = blockDim.x / lpr;
  const int sub = threadIdx.x % lpr;
  const float mul = ASYNC ? scale * d.hp[6] : scale;
  for (int i = blockIdx.x * rows_per_block + threadIdx.x / lpr; i < n;
       i += gridDim.x * rows_per_block) {
    const int u = pos2u[i];
    if (u < 0) continue;
    const int cnt = uniq_cnt[u];
    if (cnt == 1) {
      const int id = uniq_id[u], k = uniq_k[u];
      int owner, local;
      geom_map(g, id, owner, local);
      for (int c = sub; c < g.D4; c += lpr) {
        float4 v = ld_grad4<GradT>(pend_grads, (size_t)i * g.D4 + c);
        v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
        emit_row4(d, g, owner, local, k, c, v, ASYNC);
      }
      if (!ASYNC && sub == 0) emit_id(d, g, owner, local, k);
    } else {
      float4* dst = reinterpret_cast<float4*>(staging) + (size_t)u * g.D4;
      for (int c = sub; c < g.D4; c += lpr)
        atomicAdd(dst + c, ld_grad4<GradT>(pend_grads, (size_t)i * g.D4 + c));
    }
  }
}

// Kernel B — unique-parallel: flush the staged (duplicate) rows, re-zero the
// staging rows, then the last CTA publishes counts + `pushed` flag (sync) or
// bumps the step (async) and re-arms the local counters.
template <bool ASYNC>
__global__ void __launch_bounds__(256)
px_sparse_flush_kernel(const int32_t* __restrict__ uniq_id, const int32_t* __restrict__ uniq_k,
                       const int32_t* __restrict__ uniq_cnt, float* __restrict__ staging,
                       SparseCtl* ctl, PushDst d, uint32_t* const* __restrict__ hdrs, TableGeom g,
                       float scale, int lpr) {
  const int rows_per_block = blockDim.x / lpr;
  const int sub = threadIdx.x % lpr;
  const int n_uniq = ctl->n_uniq;
  const float mul = ASYNC ? scale * d.hp[6] : scale;
  for (int u = blockIdx.x * rows_per_block + threadIdx.x / lpr; u < n_uniq;
       u += gridDim.x * rows_per_block) {
    if (uniq_cnt[u] <= 1) continue;
    const int id = uniq_id[u], k = uniq_k[u];
    int owner, local;
    geom_map(g, id, owner, local);
    float4* src = reinterpret_cast<float4*>(staging) + (size_t)u * g.D4;
    for (int c = sub; c < g.D4; c += lpr) {
      float4 v = src[c];
      src[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
      emit_row4(d, g, owner, local, k, c, v, ASYNC);
    }
    if (!ASYNC && sub == 0) emit_id(d, g, owner, local, k);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctl->push_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  const uint32_t step = ctl->step + 1;
  if (!ASYNC) {
    if (threadIdx.x < g.W) {
      const int o = threadIdx.x;
      const int cnt = g.replicated ? ctl->owner_cnt[0] : ctl->owner_cnt[o];
      uint32_t* hdr = hdrs[o];
      reinterpret_cast<volatile int32_t*>(hdr + 2 * PX_MAX_RANKS)[d.rank] = cnt;   // cnt[src]
      __threadfence_system();
      st_release_sys(hdr + d.rank, step);                                          // pushed[src]
    }
    __syncthreads();
  }
  if (threadIdx.x < PX_MAX_RANKS) ctl->owner_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    ctl->n_uniq = 0; ctl->push_done = 0;
    if (ASYNC) ctl->step = step;
  }
}

// ------------------------------------------------------------------ owner
__device__ __forceinline__ void wait_pushed(const uint32_t* hdr, const SparseCtl* ctl, int W) {
  if (threadIdx.x < W) {
    const uint32_t need = ctl->step + 1;
    while ((int32_t)(ld_acquire_sys(hdr + threadIdx.x) - need) < 0) { }
  }
  __syncthreads();
}

// claim: merge duplicate rows arriving from different sources.  The first
// entry to claim a row becomes its accumulator; later ones add into it.
__global__ void __launch_bounds__(256)
px_sparse_claim_kernel(char* ring, uint32_t* hdr, size_t ring_ids_off, int cap, int32_t* slotmap,
                       const SparseCtl* ctl, TableGeom g) {
  wait_pushed(hdr, ctl, g.W);
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int32_t* cnt = reinterpret_cast<const int32_t*>(hdr + 2 * PX_MAX_RANKS);
  const int32_t* ids = reinterpret_cast<const int32_t*>(ring + ring_ids_off);
  float4* rows = reinterpret_cast<float4*>(ring);
  for (int s = 0; s < g.W; ++s) {
    const int c = ld_volatile_u32(reinterpret_cast<const uint32_t*>(cnt) + s);
    for (int j = blockIdx.x * warps + (threadIdx.x >> 5); j < c; j += gridDim.x * warps) {
      const int e = s * cap + j;
      const int r = ids[e];
      int old = 0;
      if (lane == 0) old = atomicCAS(&slotmap[r], -1, e);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old != -1) {
        for (int cidx = lane; cidx < g.D4; cidx += 32) {
          const float4 v = rows[(size_t)e * g.D4 + cidx];
          atomicAdd(&rows[(size_t)old * g.D4 + cidx], v);
        }
      }
    }
  }
}

// apply: every claimed row gets exactly one optimizer application with the
// summed gradient (× avg).  `use_slotmap`=0 when entries are known unique
// (world 1 with local aggregation): claim is skipped entirely.
__global__ void __launch_bounds__(256)
px_sparse_apply_kernel(char* ring, uint32_t* hdr, size_t ring_ids_off, int cap, int32_t* slotmap,
                       float* table, float* slot0, float* slot1, const float* hp, float avg,
                       int kind, SparseCtl* ctl, uint32_t* const* __restrict__ hdrs, TableGeom g,
                       int rank, int use_slotmap) {
  wait_pushed(hdr, ctl, g.W);
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int32_t* cnt = reinterpret_cast<const int32_t*>(hdr + 2 * PX_MAX_RANKS);
  const int32_t* ids = reinterpret_cast<const int32_t*>(ring + ring_ids_off);
  const float4* rows = reinterpret_cast<const float4*>(ring);
  const float lr = hp[0], ha = hp[1], hb = hp[2], eps = hp[3], nesterov = hp[7];
  const float gmul = avg * hp[6];
  for (int s = 0; s < g.W; ++s) {
    const int c = ld_volatile_u32(reinterpret_cast<const uint32_t*>(cnt) + s);
    for (int j = blockIdx.x * warps + (threadIdx.x >> 5); j < c; j += gridDim.x * warps) {
      const int e = s * cap + j;
      const int r = ids[e];
      if (use_slotmap) {
        int owner_e = 0;
        if (lane == 0) owner_e = slotmap[r];
        owner_e = __shfl_sync(0xffffffffu, owner_e, 0);
        if (owner_e != e) continue;
      }
      for (int cidx = lane; cidx < g.D4; cidx += 32) {
        float4 gv = rows[(size_t)e * g.D4 + cidx];
        gv.x *= gmul; gv.y *= gmul; gv.z *= gmul; gv.w *= gmul;
        float4* pw = reinterpret_cast<float4*>(table) + (size_t)r * g.D4 + cidx;
        float4 w = *pw, s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
        float4* p0 = slot0 ? reinterpret_cast<float4*>(slot0) + (size_t)r * g.D4 + cidx : nullptr;
        float4* p1 = slot1 ? reinterpret_cast<float4*>(slot1) + (size_t)r * g.D4 + cidx : nullptr;
        if (p0) s0 = *p0;
        if (p1) s1 = *p1;
        sparse_update4(kind, lr, ha, hb, eps, nesterov, gv, w, s0, s1);
        *pw = w;
        if (p0) *p0 = s0;
        if (p1) *p1 = s1;
      }
      __syncwarp();
      if (use_slotmap && lane == 0) slotmap[r] = -1;
    }
  }
  // ---- completion: publish applied[me] = step to every rank
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctl->apply_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  const uint32_t step = ctl->step + 1;
  if (threadIdx.x < g.W) st_release_sys(hdrs[threadIdx.x] + PX_MAX_RANKS + rank, step);
  __syncthreads();
  if (threadIdx.x == 0) { ctl->step = step; ctl->apply_done = 0; }
}

// ---------------------------------------------------------------------------
extern "C" {

struct PxTableGeom { int V, P, W, rows_per_part, D4, strategy, replicated, extras, base; };
static inline TableGeom to_geom(const PxTableGeom* g) {
  TableGeom t; t.V = g->V; t.P = g->P; t.W = g->W; t.rows_per_part = g->rows_per_part;
  t.D4 = g->D4; t.strategy = g->strategy; t.replicated = g->replicated; t.extras = g->extras;
  t.base = g->base; return t;
}

size_t px_sparse_ctl_bytes() { return sizeof(SparseCtl); }
int px_sparse_hdr_words() { return PX_TBL_HDR_WORDS; }

static inline int pick_lpr(int D4) { int l = 1; while (l < D4 && l < 32) l <<= 1; return l; }

// ids_is64: 1 = int64 ids, 0 = int32.  out_dtype 0 fp32 / 1 bf16.
int px_sparse_lookup(const void* ids, int ids_is64, int n, void* tables_dev, void* out,
                     int out_dtype, int32_t* pend_ids, const PxTableGeom* g, const void* hdr_mine,
                     const void* ctl, int wait, cudaStream_t stream) {
  if (n <= 0) return 0;
  const TableGeom G = to_geom(g);
  const int lpr = pick_lpr(G.D4);
  const int threads = 256, rpb = threads / lpr;
  int blocks = (n + rpb - 1) / rpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const uint32_t* applied = reinterpret_cast<const uint32_t*>(hdr_mine) + PX_MAX_RANKS;
#define LK(IdT, OutT)                                                                          \
  px_sparse_lookup_kernel<IdT, OutT><<<blocks, threads, 0, stream>>>(                          \
      (const IdT*)ids, n, (float* const*)tables_dev, (OutT*)out, pend_ids, G, applied,         \
      (const SparseCtl*)ctl, lpr, wait)
  if (ids_is64) { if (out_dtype == 0) LK(long long, float); else LK(long long, __nv_bfloat16); }
  else { if (out_dtype == 0) LK(int, float); else LK(int, __nv_bfloat16); }
#undef LK
  return (int)cudaGetLastError();
}

// Local aggregation.  hbits: log2 of hash size (>= 2n).  If smem_ok the
// single-CTA shared-memory variant is used, else the global two-pass variant
// (keys/slot_u: device scratch of 2^hbits int32 each; keys are reset here).
int px_sparse_dedup(const int32_t* pend_ids, int n, int hbits, int32_t* keys, int32_t* slot_u,
                    int32_t* uniq_id, int32_t* uniq_k, int32_t* uniq_head, int32_t* next,
                    void* ctl, const PxTableGeom* g, int dedup, int use_smem,
                    cudaStream_t stream) {
  const TableGeom G = to_geom(g);
  if (use_smem) {
    const size_t smem = (size_t)2 * sizeof(int32_t) << hbits;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(px_sparse_dedup_smem_kernel,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    if (smem > 200 * 1024) return -2;
    px_sparse_dedup_smem_kernel<<<1, 1024, smem, stream>>>(pend_ids, n, hbits, uniq_id, uniq_k,
                                                           uniq_head, next, (SparseCtl*)ctl, G,
                                                           dedup);
  } else {
    int blocks = (n + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    if (dedup) cudaMemsetAsync(keys, 0xff, sizeof(int32_t) << hbits, stream);
    px_sparse_dedup_insert_kernel<<<blocks, 256, 0, stream>>>(
        pend_ids, n, hbits, keys, slot_u, uniq_id, uniq_k, uniq_head, next, (SparseCtl*)ctl, G,
        dedup);
    if (dedup)
      px_sparse_dedup_link_kernel<<<blocks, 256, 0, stream>>>(pend_ids, n, hbits, keys, slot_u,
                                                              uniq_head, next);
  }
  return (int)cudaGetLastError();
}

// grad_dtype 0 fp32 / 1 bf16.  rings_dev/hdrs_dev: device arrays of `world` pointers.
// sync push: scatter (position-parallel) + flush (duplicates, flags).
int px_sparse_push(const void* pend_grads, int grad_dtype, int n, const int32_t* pos2u,
                   const int32_t* uniq_id, const int32_t* uniq_k, const int32_t* uniq_cnt,
                   float* staging, void* ctl, void* rings_dev, void* hdrs_dev,
                   size_t ring_ids_off, int cap, const PxTableGeom* g, float scale, int rank,
                   int max_blocks, cudaStream_t stream) {
  const TableGeom G = to_geom(g);
  PushDst d{};
  d.rings = (char* const*)rings_dev; d.ring_ids_off = ring_ids_off; d.cap = cap; d.rank = rank;
  const int lpr = pick_lpr(G.D4), rpb = 256 / lpr;
  int blocks = (n + rpb - 1) / rpb;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  if (grad_dtype == 0)
    px_sparse_scatter_kernel<float, false><<<blocks, 256, 0, stream>>>(
        (const float*)pend_grads, n, pos2u, uniq_id, uniq_k, uniq_cnt, staging,
        (const SparseCtl*)ctl, d, (uint32_t* const*)hdrs_dev, G, scale, lpr);
  else
    px_sparse_scatter_kernel<__nv_bfloat16, false><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)pend_grads, n, pos2u, uniq_id, uniq_k, uniq_cnt, staging,
        (const SparseCtl*)ctl, d, (uint32_t* const*)hdrs_dev, G, scale, lpr);
  // only ids carried by several positions are staged: a small grid suffices
  const int fblocks = blocks > 96 ? 96 : blocks;
  px_sparse_flush_kernel<false><<<fblocks, 256, 0, stream>>>(
      uniq_id, uniq_k, uniq_cnt, staging, (SparseCtl*)ctl, d, (uint32_t* const*)hdrs_dev, G,
      scale, lpr);
  return (int)cudaGetLastError();
}

int px_sparse_claim(void* ring, void* hdr, size_t ring_ids_off, int cap, int32_t* slotmap,
                    const void* ctl, const PxTableGeom* g, int blocks, cudaStream_t stream) {
  const TableGeom G = to_geom(g);
  if (blocks < 1) blocks = 1;
  px_sparse_claim_kernel<<<blocks, 256, 0, stream>>>((char*)ring, (uint32_t*)hdr, ring_ids_off,
                                                     cap, slotmap, (const SparseCtl*)ctl, G);
  return (int)cudaGetLastError();
}

int px_sparse_apply(void* ring, void* hdr, size_t ring_ids_off, int cap, int32_t* slotmap,
                    float* table, float* slot0, float* slot1, const float* hp, float avg, int kind,
                    void* ctl, void* hdrs_dev, const PxTableGeom* g, int rank, int use_slotmap,
                    int blocks, cudaStream_t stream) {
  const TableGeom G = to_geom(g);
  if (blocks < 1) blocks = 1;
  px_sparse_apply_kernel<<<blocks, 256, 0, stream>>>(
      (char*)ring, (uint32_t*)hdr, ring_ids_off, cap, slotmap, table, slot0, slot1, hp, avg, kind,
      (SparseCtl*)ctl, (uint32_t* const*)hdrs_dev, G, rank, use_slotmap);
  return (int)cudaGetLastError();
}

int px_sparse_async_apply(const void* pend_grads, int grad_dtype, int n, const int32_t* pos2u,
                          const int32_t* uniq_id, const int32_t* uniq_k,
                          const int32_t* uniq_cnt, float* staging, void* ctl, void* tables_dev,
                          void* slot0s_dev, void* slot1s_dev, const float* hp, float scale,
                          int kind, const PxTableGeom* g, int max_blocks, cudaStream_t stream) {
  const TableGeom G = to_geom(g);
  PushDst d{};
  d.tables = (float* const*)tables_dev; d.slot0s = (float* const*)slot0s_dev;
  d.slot1s = (float* const*)slot1s_dev; d.hp = hp; d.kind = kind;
  const int lpr = pick_lpr(G.D4), rpb = 256 / lpr;
  int blocks = (n + rpb - 1) / rpb;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  if (grad_dtype == 0)
    px_sparse_scatter_kernel<float, true><<<blocks, 256, 0, stream>>>(
        (const float*)pend_grads, n, pos2u, uniq_id, uniq_k, uniq_cnt, staging,
        (const SparseCtl*)ctl, d, nullptr, G, scale, lpr);
  else
    px_sparse_scatter_kernel<__nv_bfloat16, true><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)pend_grads, n, pos2u, uniq_id, uniq_k, uniq_cnt, staging,
        (const SparseCtl*)ctl, d, nullptr, G, scale, lpr);
  const int fblocks = blocks > 96 ? 96 : blocks;
  px_sparse_flush_kernel<true><<<fblocks, 256, 0, stream>>>(
      uniq_id, uniq_k, uniq_cnt, staging, (SparseCtl*)ctl, d, nullptr, G, scale, lpr);
  return (int)cudaGetLastError();
}

}  // extern "C"
