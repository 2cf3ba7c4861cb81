// Sparse path, one *group* of co-indexed tables per launch:
//   px_sparse_lookup_kernel  remote-gather lookup (NVLink peer loads), all member tables
//   px_sparse_push_kernel    local aggregation (SMEM dedup) + P2P push + `pushed` flag:
//                            ONE launch per group and step
//   px_sparse_owner_kernel   owner side: cross-source merge + sparse optimizer + `applied`
//                            flag: ONE (cooperative) launch per group and step
// (async / Hogwild mode: the push kernel applies the optimizer remotely, no owner kernel.)
//
// A group is a set of row-partitioned tables that share (V, P, strategy, owner map) and are
// looked up with the SAME ids in one call (LM1B: softmax_w + softmax_b; any single table is a
// group of one).  The ids are deduplicated once and every member table's rows travel together.
//
// What this replaces in the reference: worker GPU → local-chief CPU
// SparseConditionalAccumulator → gRPC → PS CPU accumulator (sorted two-pointer merge, whole
// value tensor re-allocated per apply, tensorflow/core/kernels/sparse_conditional_accumulator.h
// :192-319) → chief take_grad → serial CPU SparseApplyAdagrad row loop (tensorflow/core/kernels/
// training_ops.cc:1338-1351) → token queues; and for lookups dynamic_partition → per-shard PS CPU
// gather → gRPC → dynamic_stitch (tensorflow/python/ops/embedding_ops.py:151-209,
// gather_functor_gpu.cu.h:32-70, dynamic_partition_op_gpu.cu.cc:60-110).
//
// Step protocol (sync mode); flags are monotonically increasing step numbers in the group's
// symmetric header, written with st.release.sys and polled with ld.acquire.sys:
//   lookup(t)  waits applied[o] >= t-1 for every owner o   (rows fresh, receive rings drained)
//   push(t)    writes rows + local row ids + counts into the owner's ring[src = me], then
//              pushed[me] = t at the owner
//   owner(t)   waits pushed[s] >= t for every source s, links the entries of every touched row
//              into a list (one atomicExch per entry), grid barrier, then the list head sums the
//              (bf16 or fp32) wire rows in fp32 and applies the optimizer once per row; publishes
//              applied[me] = t to every rank.
//
// Wire format ("boundary between workers and servers", graph_transform_lib.py:1315-1370): rows
// cross NVLink in the gradient's own dtype (bf16 gradients stay bf16; the widening cast runs on the
// owner, after the wire) unless the boundary optimisation is switched off, in which case the
// sender widens to fp32 first.
#include "common.cuh"
#include "launch.h"
#include "optim_rules.cuh"

#define PX_GRP_MAX 4            // member tables per group

struct GroupGeom {
  int V, P, W, rows_per_part;
  int strategy;                 // 0 mod, 1 div
  int replicated;               // 1: every rank holds the full table (AR mode)
  int extras, base;             // div strategy
  const int* part_owner;        // [P] owner rank of partition p (byte-greedy placement)
  const int* part_slot;         // [P] index of partition p among its owner's partitions
};

// per-rank control block of a group (local memory)
struct SparseCtl {
  uint32_t step;                // completed steps
  uint32_t push_done;           // CTA ticket counter (push kernel)
  uint32_t apply_done;          // CTA ticket counter (owner kernel)
  uint32_t bar;                 // grid barrier arrivals (owner kernel)
  int32_t n_dup;                // staging rows handed out this step
  int32_t overflow;             // #positions that fell back to un-deduplicated entries (stat)
  int32_t owner_cnt[PX_MAX_RANKS];
  unsigned long long t_push[2]; // %globaltimer at push start / flag publication
  unsigned long long t_own[3];  // owner kernel: start / all sources arrived / applied published
  unsigned long long t_dbg[8];  // push kernel, CTA 0: end of each internal phase (profiling aid)
};

// group header (symmetric): [pushed[R] | applied[R] | cnt[R]]
#define PX_GRP_HDR_WORDS (3 * PX_MAX_RANKS)

__device__ __forceinline__ unsigned long long px_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

#define PX_SMEM_PARTS 1024

__device__ __forceinline__ void geom_part(const GroupGeom& g, int id, int& p, int& idx) {
  if (g.strategy == 0) { p = id % g.P; idx = id / g.P; }
  else {
    const int thr = g.extras * (g.base + 1);
    if (id < thr) { p = id / (g.base + 1); idx = id - p * (g.base + 1); }
    else { p = (id - g.extras) / max(g.base, 1); idx = id - (p * g.base + g.extras); }
  }
}

__device__ __forceinline__ void geom_map(const GroupGeom& g, int id, int& owner, int& local) {
  if (g.replicated) { owner = 0; local = id; return; }
  int p, idx;
  geom_part(g, id, p, idx);
  owner = __ldg(g.part_owner + p);
  local = __ldg(g.part_slot + p) * g.rows_per_part + idx;
}

__device__ __forceinline__ uint32_t hash_cta(int id) {
  uint32_t x = (uint32_t)id * 2654435761u;
  return x ^ (x >> 15);
}
__device__ __forceinline__ uint32_t hash_slot(int id) {
  uint32_t x = (uint32_t)id * 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  return x ^ (x >> 16);
}

// ------------------------------------------------------------------ lookup
struct LookupTable {
  const void* const* srcs;      // device array[W]: fp32 tables or bf16 shadows of every rank
  void* out;                    // [n, D4*4] rows
  int D4;
  int src_bf16;                 // 1: srcs are bf16 shadow copies (out is bf16 too)
  int out_bf16;
};
struct LookupArgs {
  LookupTable t[PX_GRP_MAX];
  int nt;
};

// out_t[i,:] = table_t@owner(ids[i])[local(ids[i]), :] for every member table; LPR lanes per row.
template <typename IdT>
__global__ void __launch_bounds__(256)
px_sparse_lookup_kernel(const IdT* __restrict__ ids, int n, LookupArgs a,
                        int32_t* __restrict__ pend_ids, GroupGeom g, const uint32_t* applied,
                        const SparseCtl* ctl, int lpr, int wait) {
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
#pragma unroll 1
    for (int t = 0; t < a.nt; ++t) {
      const LookupTable& T = a.t[t];
      const char* src = reinterpret_cast<const char*>(T.srcs[owner]);
      if (T.src_bf16) {
        // bf16 shadow → bf16 rows: 8 elements per 16-byte load, half the NVLink/HBM bytes
        const int nv = (T.D4 + 1) / 2;             // 16-byte vectors per row (row padded to 8)
        const uint4* s = reinterpret_cast<const uint4*>(src) + (size_t)local * nv;
        uint4* d = reinterpret_cast<uint4*>(T.out) + (size_t)i * nv;
        for (int c = sub; c < nv; c += lpr)
          st_v4(d + c, valid ? ld_v4(s + c) : make_uint4(0, 0, 0, 0));
      } else {
        const float4* s = reinterpret_cast<const float4*>(src) + (size_t)local * T.D4;
        for (int c = sub; c < T.D4; c += lpr) {
          const uint4 v = valid ? ld_v4(s + c) : make_uint4(0, 0, 0, 0);   // OOB -> zeros
          if (!T.out_bf16) {
            st_v4(reinterpret_cast<float4*>(T.out) + (size_t)i * T.D4 + c, v);
          } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(__uint_as_float(v.x), __uint_as_float(v.y));
            __nv_bfloat162 hi = __floats2bfloat162_rn(__uint_as_float(v.z), __uint_as_float(v.w));
            *reinterpret_cast<uint2*>(reinterpret_cast<char*>(T.out) +
                                      ((size_t)i * T.D4 + c) * 8) =
                make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          }
        }
      }
    }
  }
}

// -------------------------------------------------------------------- push
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

template <typename WireT>
__device__ __forceinline__ void st_wire4(char* row_base, int c, const float4& v) {
  if (sizeof(WireT) == 4) {
    st_v4_stream(reinterpret_cast<float4*>(row_base) + c,
                 make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z),
                            __float_as_uint(v.w)));
  } else {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    const uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
                 ::"l"(row_base + (size_t)c * 8), "r"(o.x), "r"(o.y) : "memory");
  }
}
template <typename WireT>
__device__ __forceinline__ float4 ld_wire4(const char* row_base, int c) {
  if (sizeof(WireT) == 4) {
    const uint4 v = ld_v4_stream(reinterpret_cast<const float4*>(row_base) + c);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z),
                       __uint_as_float(v.w));
  } else {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(v.x), "=r"(v.y) : "l"(row_base + (size_t)c * 8) : "memory");
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                       __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
  }
}


// optimizer on 4 elements of one table row (+ bf16 shadow refresh); used by the owner kernel on
// local rows and by the async push on remote rows
template <int FAM>
__device__ __forceinline__ void px_row_apply4(int kind, const PxHP& h, const float4& g,
                                              float* table, float* slot0, float* slot1,
                                              float* slot2, __nv_bfloat16* shadow, size_t row,
                                              int D4, int c) {
  const size_t off = row * D4 + c;
  float4* pw = reinterpret_cast<float4*>(table) + off;
  float4* p0 = slot0 ? reinterpret_cast<float4*>(slot0) + off : nullptr;
  float4* p1 = slot1 ? reinterpret_cast<float4*>(slot1) + off : nullptr;
  float4* p2 = (FAM == 1 && slot2) ? reinterpret_cast<float4*>(slot2) + off : nullptr;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w = *pw, s0 = z, s1 = z, s2 = z;
  if (p0) s0 = *p0;
  if (p1) s1 = *p1;
  if (p2) s2 = *p2;
  px_rule4<FAM>(kind, h, g, w, s0, s1, s2);
  *pw = w;
  if (p0) *p0 = s0;
  if (p1) *p1 = s1;
  if (p2) *p2 = s2;
  if (shadow) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<char*>(shadow) +
                              (row * ((D4 + 1) / 2 * 2) + c) * 8) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// same for 8 consecutive elements (two float4 groups 2*c2, 2*c2+1): every load is issued
// before the first store (table / slot pointers may alias as far as the compiler knows, so
// two back-to-back px_row_apply4 calls would serialise load -> store -> load)
template <int FAM>
__device__ __forceinline__ void px_row_apply8(int kind, const PxHP& h, const float* g,
                                              float* table, float* slot0, float* slot1,
                                              float* slot2, __nv_bfloat16* shadow, size_t row,
                                              int D4, int c2) {
  const size_t off = row * D4 + 2 * c2;
  float4* pw = reinterpret_cast<float4*>(table) + off;
  float4* p0 = slot0 ? reinterpret_cast<float4*>(slot0) + off : nullptr;
  float4* p1 = slot1 ? reinterpret_cast<float4*>(slot1) + off : nullptr;
  float4* p2 = (FAM == 1 && slot2) ? reinterpret_cast<float4*>(slot2) + off : nullptr;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w[2] = {pw[0], pw[1]}, s0[2] = {z, z}, s1[2] = {z, z}, s2[2] = {z, z};
  if (p0) { s0[0] = p0[0]; s0[1] = p0[1]; }
  if (p1) { s1[0] = p1[0]; s1[1] = p1[1]; }
  if (p2) { s2[0] = p2[0]; s2[1] = p2[1]; }
  px_rule4<FAM>(kind, h, make_float4(g[0], g[1], g[2], g[3]), w[0], s0[0], s1[0], s2[0]);
  px_rule4<FAM>(kind, h, make_float4(g[4], g[5], g[6], g[7]), w[1], s0[1], s1[1], s2[1]);
  pw[0] = w[0]; pw[1] = w[1];
  if (p0) { p0[0] = s0[0]; p0[1] = s0[1]; }
  if (p1) { p1[0] = s1[0]; p1[1] = s1[1]; }
  if (p2) { p2[0] = s2[0]; p2[1] = s2[1]; }
  if (shadow) {
    const float f[8] = {w[0].x, w[0].y, w[0].z, w[0].w, w[1].x, w[1].y, w[1].z, w[1].w};
    st_v4(reinterpret_cast<uint4*>(shadow) + row * ((D4 + 1) / 2) + c2,
          Vec16<__nv_bfloat16>::pack(f));
  }
}

struct PushTable {
  const void* grads;            // [n, D4*4] gradient rows (GradT)
  float* staging;               // [n/2+1, D4*4] fp32 rows for ids carried by several positions
  char* const* rings;           // sync: device array[W] of receive-ring bases
  float* const* tables;         // async: device arrays[W] for the remote optimizer application
  float* const* slot0s;
  float* const* slot1s;
  float* const* slot2s;
  __nv_bfloat16* const* shadows;
  const float* hp;
  int D4, kind;
  float scale;                  // ScaleGradients factor when it runs on the sender
};
struct PushArgs {
  PushTable t[PX_GRP_MAX];
  int nt;
  int32_t* const* ring_ids;     // device array[W]: id rings ([W_src][cap] local rows)
  uint32_t* const* hdrs;        // device array[W]: group headers
  int cap, rank;
};

// destination of one row (all member tables): the owner's ring slot (sync) or the owner's
// table row itself (async)
template <typename WireT, bool ASYNC, int FAM>
__device__ __forceinline__ void emit_row(const PushArgs& a, const GroupGeom& g, int t, int owner,
                                         int local, int k, int c, float4 v) {
  const PushTable& T = a.t[t];
  if (!ASYNC) {
    const size_t row_bytes = (size_t)T.D4 * 4 * sizeof(WireT);
    if (g.replicated) {
      for (int p = 0; p < g.W; ++p) {
        const int q = (a.rank + p) % g.W;
        st_wire4<WireT>(T.rings[q] + ((size_t)a.rank * a.cap + k) * row_bytes, c, v);
      }
    } else {
      st_wire4<WireT>(T.rings[owner] + ((size_t)a.rank * a.cap + k) * row_bytes, c, v);
    }
  } else {
    px_row_apply4<FAM>(T.kind, px_load_hp(T.hp), v, T.tables[owner],
                       T.slot0s ? T.slot0s[owner] : nullptr, T.slot1s ? T.slot1s[owner] : nullptr,
                       T.slot2s ? T.slot2s[owner] : nullptr,
                       T.shadows ? T.shadows[owner] : nullptr, (size_t)local, T.D4, c);
  }
}

__device__ __forceinline__ void emit_id(const PushArgs& a, const GroupGeom& g, int owner,
                                        int local, int k) {
  if (g.replicated) {
    for (int p = 0; p < g.W; ++p) a.ring_ids[p][(size_t)a.rank * a.cap + k] = local;
  } else {
    a.ring_ids[owner][(size_t)a.rank * a.cap + k] = local;
  }
}

// ONE launch: local aggregation + push + flag.
// The id space is partitioned over the CTAs by a hash, so every CTA owns all positions of "its"
// ids: it deduplicates them in a shared-memory hash table ("local aggregation dedups indices in
// SMEM before shipping"), reserves ring slots with one global atomic per (CTA, owner), ships rows
// whose id is unique in the batch straight from the gradient buffer, sums rows sharing an id with
// vector atomics into a local fp32 staging row (O(1) depth for Zipfian batches) and flushes those
// once.  No grid-wide phase is needed; the last CTA publishes the counts and the `pushed` flag.
// Every CTA scans all ids twice; the scans read them from a shared-memory staging chunk filled
// with 16-byte loads issued back to back (a one-id-per-trip global loop is bound by L2 latency,
// and unrolling it instead makes the kernel instruction-fetch bound).
// SMEM layout: keys[H] | cnt[H] | kk[H] (k inside the owner bucket) | dup[H] | ids[PX_ID_CHUNK]
//              | work[PX_ID_CHUNK] (positions of the current chunk that belong to this CTA)
#define PX_ID_CHUNK 4096

// stage ids[base, base+m) into shared memory: 16 ids (4 x int4) per thread in flight
__device__ __forceinline__ void stage_ids(const int32_t* __restrict__ ids, int base, int m,
                                          int32_t* ids_s) {
  int4 q[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = (u * blockDim.x + threadIdx.x) * 4;
    q[u] = make_int4(-1, -1, -1, -1);
    if (j + 3 < m) q[u] = __ldg(reinterpret_cast<const int4*>(ids + base + j));
    else {
      if (j < m) q[u].x = __ldg(ids + base + j);
      if (j + 1 < m) q[u].y = __ldg(ids + base + j + 1);
      if (j + 2 < m) q[u].z = __ldg(ids + base + j + 2);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = (u * blockDim.x + threadIdx.x) * 4;
    if (j < PX_ID_CHUNK) *reinterpret_cast<int4*>(ids_s + j) = q[u];
  }
}

template <typename GradT, typename WireT, bool ASYNC, int FAM>
__global__ void __launch_bounds__(256)
px_sparse_push_kernel(const int32_t* __restrict__ pend_ids, int n, PushArgs a, GroupGeom g,
                      SparseCtl* ctl, int hbits, int dedup) {
  extern __shared__ int32_t smem[];
  const int H = 1 << hbits;
  int32_t* keys = smem;
  int32_t* cnt = smem + H;
  int32_t* kk = smem + 2 * H;
  int32_t* dup = smem + 3 * H;
  int32_t* ids_s = smem + 4 * H;
  __shared__ int s_owner_cnt[PX_MAX_RANKS], s_base_k[PX_MAX_RANKS];
  __shared__ int s_ndup, s_base_dup, s_overflow;
  __shared__ bool s_last;
  // placement maps and ring bases are read for every row: keep them in shared memory (a
  // global load each would put two more memory latencies on every row's critical path)
  __shared__ short s_part_owner[PX_SMEM_PARTS], s_part_slot[PX_SMEM_PARTS];
  __shared__ char* s_ring[PX_GRP_MAX][PX_MAX_RANKS];
  const bool parts_in_smem = !g.replicated && g.P <= PX_SMEM_PARTS;
  if (parts_in_smem)
    for (int p = threadIdx.x; p < g.P; p += blockDim.x) {
      s_part_owner[p] = (short)__ldg(g.part_owner + p);
      s_part_slot[p] = (short)__ldg(g.part_slot + p);
    }
  if (!ASYNC && threadIdx.x < a.nt * PX_MAX_RANKS) {
    const int t = threadIdx.x / PX_MAX_RANKS, r = threadIdx.x % PX_MAX_RANKS;
    s_ring[t][r] = r < g.W ? a.t[t].rings[r] : nullptr;
  }
#define GEOM_MAP(id, owner, local)                                                     \
  do {                                                                                 \
    if (parts_in_smem) {                                                               \
      int p_, idx_;                                                                    \
      geom_part(g, (id), p_, idx_);                                                    \
      owner = s_part_owner[p_];                                                        \
      local = s_part_slot[p_] * g.rows_per_part + idx_;                                \
    } else geom_map(g, (id), owner, local);                                            \
  } while (0)
  const int G = gridDim.x, c_me = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool raw_all = !dedup;
#define PX_DBG(i) do { if (c_me == 0 && threadIdx.x == 0) ctl->t_dbg[i] = px_globaltimer(); } while (0)
  if (c_me == 0 && threadIdx.x == 0) ctl->t_push[0] = px_globaltimer();
  for (int h = threadIdx.x; h < H; h += blockDim.x) { keys[h] = -1; cnt[h] = 0; }
  if (threadIdx.x < PX_MAX_RANKS) s_owner_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_ndup = 0; s_overflow = 0; }
  __syncthreads();
  // ---- pass 1: insert my ids, count positions per id (raw mode: count my positions per owner)
  for (int base = 0; base < n; base += PX_ID_CHUNK) {
    const int m = min(PX_ID_CHUNK, n - base);
    stage_ids(pend_ids, base, m, ids_s);
    __syncthreads();
#pragma unroll 1
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const int id = ids_s[j];
      if (id < 0) continue;
      if (raw_all) {
        if ((base + j) % G == c_me) {
          int owner, local;
          GEOM_MAP(id, owner, local);
          atomicAdd(&s_owner_cnt[owner], 1);
        }
        continue;
      }
      if ((int)(((unsigned long long)hash_cta(id) * (unsigned)G) >> 32) != c_me) continue;
      uint32_t h = hash_slot(id) & (H - 1);
      int probes = 0;
      while (true) {
        const int old = atomicCAS(&keys[h], -1, id);
        if (old == -1 || old == id) { atomicAdd(&cnt[h], 1); break; }
        h = (h + 1) & (H - 1);
        if (++probes >= H) { atomicAdd(&s_overflow, 1); break; }   // table full: raw entry later
      }
    }
    __syncthreads();
  }
  PX_DBG(0);
  if (dedup) {
    // ---- pass 2: one ring slot per unique id, one staging row per duplicated id
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
      const int id = keys[h];
      if (id < 0) continue;
      int owner, local;
      GEOM_MAP(id, owner, local);
      kk[h] = atomicAdd(&s_owner_cnt[owner], 1);
      dup[h] = cnt[h] > 1 ? atomicAdd(&s_ndup, 1) : -1;
    }
    __syncthreads();
    if (s_overflow > 0) {
      // the (statistically never hit) SMEM overflow: positions whose id did not fit travel
      // as raw entries after the deduplicated ones; count them per owner
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int id = pend_ids[i];
        if (id < 0) continue;
        if ((int)(((unsigned long long)hash_cta(id) * (unsigned)G) >> 32) != c_me) continue;
        uint32_t h = hash_slot(id) & (H - 1);
        int probes = 0;
        while (keys[h] != id && keys[h] != -1 && ++probes <= H) h = (h + 1) & (H - 1);
        if (keys[h] != id) {
          int owner, local;
          GEOM_MAP(id, owner, local);
          atomicAdd(&s_owner_cnt[owner], 1);
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < PX_MAX_RANKS) {
    const int m = s_owner_cnt[threadIdx.x];
    s_base_k[threadIdx.x] = m > 0 ? atomicAdd(&ctl->owner_cnt[threadIdx.x], m) : 0;
    s_owner_cnt[threadIdx.x] = 0;              // re-used as the raw-entry cursor below
  }
  if (threadIdx.x == 0) {
    s_base_dup = s_ndup > 0 ? atomicAdd(&ctl->n_dup, s_ndup) : 0;
    if (s_overflow > 0) atomicAdd(&ctl->overflow, s_overflow);
  }
  __syncthreads();
  // raw entries take the slots after the deduplicated ones of this CTA
  if (dedup && s_overflow > 0) {
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
      if (keys[h] < 0) continue;
      int owner, local;
      GEOM_MAP(keys[h], owner, local);
      atomicMax(&s_owner_cnt[owner], kk[h] + 1);
    }
    __syncthreads();
  }
  PX_DBG(1);
  // ---- pass 3: ship unique rows, stage duplicated ones.  Per chunk: (a) every thread scans
  // the staged ids and appends its CTA's positions to a work list in shared memory, (b) the
  // list is processed round-robin by half-warps (16 lanes per row, four 16-byte accesses in
  // flight per lane) — balanced whatever the id distribution is.
  int32_t* work = ids_s + PX_ID_CHUNK;
  __shared__ int s_nwork;
  const int sub = lane & 15, half = lane >> 4;
  const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
  for (int base = 0; base < n; base += PX_ID_CHUNK) {
    const int m = min(PX_ID_CHUNK, n - base);
    if (threadIdx.x == 0) s_nwork = 0;
    stage_ids(pend_ids, base, m, ids_s);
    __syncthreads();
#pragma unroll 1
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const int id = ids_s[j];
      if (id < 0) continue;
      const bool mine = raw_all ? ((base + j) % G == c_me)
          : ((int)(((unsigned long long)hash_cta(id) * (unsigned)G) >> 32) == c_me);
      if (mine) work[atomicAdd(&s_nwork, 1)] = j;
    }
    __syncthreads();
    if (base == 0) PX_DBG(2);
    const int nwork = s_nwork;
#pragma unroll 1
    for (int wi = wid * 2 + half; wi < nwork; wi += nwarps * 2) {
      const int j = work[wi];
      const int pi = base + j;
      const int pid = ids_s[j];
      int ph = -1;
      if (!raw_all) {
        uint32_t h = hash_slot(pid) & (H - 1);
        int probes = 0;
        while (keys[h] != pid && keys[h] != -1 && ++probes <= H) h = (h + 1) & (H - 1);
        if (keys[h] == pid) ph = (int)h;
      }
      int owner, local;
      GEOM_MAP(pid, owner, local);
      int k = 0, pcnt = 1;
      if (ph < 0) {                       // raw entry: its own ring slot
        if (sub == 0) k = s_base_k[owner] + atomicAdd(&s_owner_cnt[owner], 1);
        k = __shfl_sync(hmask, k, half * 16);
      } else {
        k = s_base_k[owner] + kk[ph];
        pcnt = cnt[ph];
      }
      if (pcnt == 1) {
#pragma unroll 1
        for (int t = 0; t < a.nt; ++t) {
          const PushTable& T = a.t[t];
          const float mul = ASYNC ? T.scale * T.hp[HP_GSCALE] : T.scale;
          if (!ASYNC && sizeof(GradT) == 2 && sizeof(WireT) == 2 && (T.D4 & 1) == 0 &&
              !g.replicated) {
            // bf16 gradient -> bf16 wire: 16-byte copies, four per lane in flight
            const int nv = T.D4 / 2;
            const uint4* src = reinterpret_cast<const uint4*>(T.grads) + (size_t)pi * nv;
            uint4* dst = reinterpret_cast<uint4*>(
                s_ring[t][owner] + ((size_t)a.rank * a.cap + k) * ((size_t)T.D4 * 8));
#pragma unroll 1
            for (int c = sub; c < nv; c += 64) {
              uint4 v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                v[u] = c + 16 * u < nv ? ld_v4_stream(src + c + 16 * u) : make_uint4(0, 0, 0, 0);
              if (mul != 1.f) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float f[8];
                  Vec16<__nv_bfloat16>::unpack(v[u], f);
#pragma unroll
                  for (int q = 0; q < 8; ++q) f[q] *= mul;
                  v[u] = Vec16<__nv_bfloat16>::pack(f);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (c + 16 * u < nv) st_v4_stream(dst + c + 16 * u, v[u]);
            }
            continue;
          }
#pragma unroll 1
          for (int c = sub; c < T.D4; c += 16) {
            float4 v = ld_grad4<GradT>(reinterpret_cast<const GradT*>(T.grads),
                                       (size_t)pi * T.D4 + c);
            v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
            emit_row<WireT, ASYNC, FAM>(a, g, t, owner, local, k, c, v);
          }
        }
        if (!ASYNC && sub == 0) emit_id(a, g, owner, local, k);
      } else {
        const int d = s_base_dup + dup[ph];
#pragma unroll 1
        for (int t = 0; t < a.nt; ++t) {
          const PushTable& T = a.t[t];
          float4* dst = reinterpret_cast<float4*>(T.staging) + (size_t)d * T.D4;
          for (int c = sub; c < T.D4; c += 16)
            atomicAdd(dst + c, ld_grad4<GradT>(reinterpret_cast<const GradT*>(T.grads),
                                               (size_t)pi * T.D4 + c));
        }
      }
    }
    __syncthreads();
  }
  PX_DBG(3);
  // every position of my duplicated ids is staged now (they are all mine)
  // ---- pass 4: flush duplicated ids (warp per id), re-zero the staging rows
  if (dedup && s_ndup > 0) {
    for (int h0 = wid * 32; h0 < H; h0 += nwarps * 32) {
      unsigned dm = __ballot_sync(0xffffffffu, keys[h0 + lane] >= 0 && cnt[h0 + lane] > 1);
      while (dm) {
        const int h = h0 + __ffs(dm) - 1;
        dm &= dm - 1;
        int owner, local;
        GEOM_MAP(keys[h], owner, local);
        const int k = s_base_k[owner] + kk[h];
        const int d = s_base_dup + dup[h];
#pragma unroll 1
        for (int t = 0; t < a.nt; ++t) {
          const PushTable& T = a.t[t];
          const float mul = ASYNC ? T.scale * T.hp[HP_GSCALE] : T.scale;
          float4* src = reinterpret_cast<float4*>(T.staging) + (size_t)d * T.D4;
          for (int c = lane; c < T.D4; c += 32) {
            float4 v = __ldcg(src + c);
            __stcg(src + c, make_float4(0.f, 0.f, 0.f, 0.f));
            v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
            emit_row<WireT, ASYNC, FAM>(a, g, t, owner, local, k, c, v);
          }
        }
        if (!ASYNC && lane == 0) emit_id(a, g, owner, local, k);
      }
    }
  }
  PX_DBG(4);
  // ---- completion: last CTA publishes counts + `pushed` (sync) / bumps the step (async).
  // One fence per CTA: the barrier orders every thread's stores before thread 0's
  // system-scope fence (cumulativity), which orders them before the ticket and the flag.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(&ctl->push_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  PX_DBG(5);
  if (!s_last) return;
  __threadfence_system();
  const uint32_t step = ctl->step + 1;
  if (!ASYNC) {
    if (threadIdx.x < g.W) {
      const int o = threadIdx.x;
      const int cn = g.replicated ? ctl->owner_cnt[0] : ctl->owner_cnt[o];
      uint32_t* hdr = a.hdrs[o];
      reinterpret_cast<volatile int32_t*>(hdr + 2 * PX_MAX_RANKS)[a.rank] = cn;   // cnt[src]
      __threadfence_system();
      st_release_sys(hdr + a.rank, step);                                         // pushed[src]
    }
    __syncthreads();
  }
  if (threadIdx.x < PX_MAX_RANKS) ctl->owner_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    ctl->n_dup = 0; ctl->push_done = 0;
    ctl->t_push[1] = px_globaltimer();
    if (ASYNC) ctl->step = step;
  }
}

// ------------------------------------------------------------------ owner
struct OwnerTable {
  char* ring;                   // my receive ring: [W_src][cap][D4*4] WireT
  float* table; float* slot0; float* slot1; float* slot2;
  __nv_bfloat16* shadow;        // bf16 copy read by lookups (or null)
  const float* hp;
  int D4, kind;
  float avg;                    // 1/num_workers (average_sparse) × owner-side gradient scale
};
struct OwnerArgs {
  OwnerTable t[PX_GRP_MAX];
  int nt;
  const int32_t* ring_ids;      // [W_src][cap]
  uint32_t* hdr;                // my group header
  uint32_t* const* hdrs;        // every rank's header (to publish `applied`)
  int32_t* slotmap;             // [rows_local], -1 when idle
  int32_t* next;                // [W_src * cap]
  int cap, rank, use_merge;
  int fixed_cnt;                // >= 0: library-collective arm — every source delivered exactly this
                                // many entries (negative row id = not mine), no flags involved
};

// One warp that waits for every source's `pushed` flag.  Launched right before the owner kernel
// on the comm stream: the owner's (cooperative, whole-GPU) grid then starts with its inputs
// complete instead of spinning with every register file of the device allocated while the
// backward pass on the main stream is starved — measured at N=2: 31 us of owner spinning cost
// 65 us of forward/backward time.
__global__ void px_sparse_wait_kernel(const uint32_t* hdr, const SparseCtl* ctl, int W) {
  if (threadIdx.x < W) {
    const uint32_t need = ctl->step + 1;
    while ((int32_t)(ld_acquire_sys(hdr + threadIdx.x) - need) < 0) { __nanosleep(200); }
  }
}

// ONE launch: wait for every source, merge rows that several sources touched, apply the sparse
// optimizer once per touched row, publish `applied`.  Launched cooperatively when use_merge (one
// grid barrier between linking and applying).
template <typename WireT, int FAM>
__global__ void __launch_bounds__(256, FAM == 0 ? 4 : 2)
px_sparse_owner_kernel(OwnerArgs a, GroupGeom g, SparseCtl* ctl) {
  __shared__ bool s_last;
  const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) ctl->t_own[0] = px_globaltimer();
  if (a.fixed_cnt < 0) {
    if (threadIdx.x < g.W) {
      const uint32_t need = ctl->step + 1;
      while ((int32_t)(ld_acquire_sys(a.hdr + threadIdx.x) - need) < 0) { }
    }
    __syncthreads();
  }
  if (stamp) ctl->t_own[1] = px_globaltimer();
  // apply phase: 16 lanes per entry (two entries per warp in flight)
  const int lane = threadIdx.x & 15, warps = blockDim.x >> 4;
  const unsigned hmask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
  const uint32_t* cnt = a.hdr + 2 * PX_MAX_RANKS;
  // entries of all sources form ONE index space [0, total): a (source, j) double loop would
  // hand every half-warp one entry per source — W entries in sequence for the first few
  // half-warps and nothing for the rest
  __shared__ int s_pre[PX_MAX_RANKS + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < g.W; ++s) {
      s_pre[s] = acc;
      acc += a.fixed_cnt >= 0 ? a.fixed_cnt : (int)ld_volatile_u32(cnt + s);
    }
    s_pre[g.W] = acc;
  }
  __syncthreads();
  const int total = s_pre[g.W];
  if (a.use_merge) {
    // link: every entry pushes itself on the list of its row (at most one entry per source when
    // the senders aggregate locally, so lists are <= W long)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      int s = 0;
      while (i >= s_pre[s + 1]) ++s;
      const int e = s * a.cap + (i - s_pre[s]);
      const int r = a.ring_ids[e];
      if (r >= 0) a.next[e] = atomicExch(&a.slotmap[r], e);
    }
    // grid barrier (all CTAs are co-resident: cooperative launch)
    if (stamp) ctl->t_dbg[6] = px_globaltimer();
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(&ctl->bar, 1u);
      while (ld_volatile_u32(&ctl->bar) < gridDim.x) { }
      __threadfence();
    }
    __syncthreads();
    if (stamp) ctl->t_dbg[7] = px_globaltimer();
  }
  {
    for (int i = blockIdx.x * warps + (threadIdx.x >> 4); i < total; i += gridDim.x * warps) {
      int s = 0;
      while (i >= s_pre[s + 1]) ++s;
      const int e = s * a.cap + (i - s_pre[s]);
      const int r = a.ring_ids[e];
      {
        // the rows of a batch are scattered over a multi-GB table: every touch is a DRAM
        // (and usually a TLB) miss.  Prefetch this half-warp's NEXT entry's table / slot rows
        // into L2 now, so that miss overlaps the work on the current entry.
        const int in = i + gridDim.x * warps;
        if (in < total) {
          int sn = s;
          while (in >= s_pre[sn + 1]) ++sn;
          const int rn = a.ring_ids[sn * a.cap + (in - s_pre[sn])];
          if (rn >= 0) {
            for (int t = 0; t < a.nt; ++t) {
              const OwnerTable& T = a.t[t];
              const size_t off = (size_t)rn * T.D4 * 16;             // row offset in bytes
              const int lines = (T.D4 * 16 + 127) / 128;
              for (int l = lane; l < lines; l += 16) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(T.table) + off + l * 128));
                if (T.slot0)
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(T.slot0) + off + l * 128));
                if (T.slot1)
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(T.slot1) + off + l * 128));
              }
            }
          }
        }
      }
      if (r < 0) continue;
      if (a.use_merge) {
        if (__ldcg(a.slotmap + r) != e) continue;          // not the list head
      }
#pragma unroll 1
      for (int t = 0; t < a.nt; ++t) {
        const OwnerTable& T = a.t[t];
        const size_t row_bytes = (size_t)T.D4 * 4 * sizeof(WireT);
        const float gmul = T.avg * T.hp[HP_GSCALE];
        const PxHP hp = px_load_hp(T.hp);
        if (sizeof(WireT) == 2 && (T.D4 & 1) == 0) {
          // bf16 wire rows: one 16-byte load carries 8 elements = two float4 groups
          for (int c2 = lane; c2 < T.D4 / 2; c2 += 16) {
            float f[8];
            Vec16<__nv_bfloat16>::unpack(
                ld_v4_stream(reinterpret_cast<const uint4*>(T.ring + (size_t)e * row_bytes) + c2), f);
            if (a.use_merge) {
              for (int x = __ldcg(a.next + e); x != -1; x = __ldcg(a.next + x)) {
                float o[8];
                Vec16<__nv_bfloat16>::unpack(
                    ld_v4_stream(reinterpret_cast<const uint4*>(T.ring + (size_t)x * row_bytes) + c2), o);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] += o[q];
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] *= gmul;
            px_row_apply8<FAM>(T.kind, hp, f, T.table, T.slot0, T.slot1, T.slot2, T.shadow,
                               (size_t)r, T.D4, c2);
          }
          continue;
        }
        for (int cidx = lane; cidx < T.D4; cidx += 16) {
          float4 gv = ld_wire4<WireT>(T.ring + (size_t)e * row_bytes, cidx);
          if (a.use_merge) {
            for (int x = __ldcg(a.next + e); x != -1; x = __ldcg(a.next + x)) {
              const float4 o = ld_wire4<WireT>(T.ring + (size_t)x * row_bytes, cidx);
              gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
            }
          }
          gv.x *= gmul; gv.y *= gmul; gv.z *= gmul; gv.w *= gmul;
          px_row_apply4<FAM>(T.kind, hp, gv, T.table, T.slot0, T.slot1, T.slot2, T.shadow,
                             (size_t)r, T.D4, cidx);
        }
      }
      __syncwarp(hmask);
      if (a.use_merge && lane == 0) a.slotmap[r] = -1;
    }
  }
  // ---- completion: publish applied[me] = step to every rank (one fence per CTA, see push)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(&ctl->apply_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  const uint32_t step = ctl->step + 1;
  if (a.fixed_cnt < 0 && threadIdx.x < g.W)
    st_release_sys(a.hdrs[threadIdx.x] + PX_MAX_RANKS + a.rank, step);
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl->step = step; ctl->apply_done = 0; ctl->bar = 0;
    ctl->t_own[2] = px_globaltimer();
  }
}

// ---------------------------------------------------------------------------
template <typename GT, typename WT, bool AS, int FAM>
static void launch_push(int blocks, size_t smem, cudaStream_t stream, const int32_t* pend_ids,
                        int n, const PushArgs& a, const GroupGeom& G, SparseCtl* ctl, int hbits,
                        int dedup) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(px_sparse_push_kernel<GT, WT, AS, FAM>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                         4 * 4 * 8192 + 2 * PX_ID_CHUNK * 4);
    attr = true;
  }
  px_sparse_push_kernel<GT, WT, AS, FAM><<<blocks, 256, smem, stream>>>(pend_ids, n, a, G, ctl,
                                                                         hbits, dedup);
}

extern "C" {

struct PxGroupGeom {
  int V, P, W, rows_per_part, strategy, replicated, extras, base;
  const int* part_owner; const int* part_slot;
};
static inline GroupGeom to_geom(const PxGroupGeom* g) {
  GroupGeom t; t.V = g->V; t.P = g->P; t.W = g->W; t.rows_per_part = g->rows_per_part;
  t.strategy = g->strategy; t.replicated = g->replicated; t.extras = g->extras; t.base = g->base;
  t.part_owner = g->part_owner; t.part_slot = g->part_slot; return t;
}

// C mirrors of the per-table descriptors (plain pointers / ints, filled from Python)
struct PxLookupTable { const void* srcs; void* out; int D4, src_bf16, out_bf16, pad; };
struct PxPushTable {
  const void* grads; float* staging; void* rings; void* tables; void* slot0s; void* slot1s;
  void* slot2s; void* shadows; const float* hp; int D4, kind; float scale; int pad;
};
struct PxOwnerTable {
  void* ring; float* table; float* slot0; float* slot1; float* slot2; void* shadow;
  const float* hp; int D4, kind; float avg; int pad;
};

size_t px_sparse_ctl_bytes() { return sizeof(SparseCtl); }
int px_sparse_hdr_words() { return PX_GRP_HDR_WORDS; }
int px_sparse_group_max() { return PX_GRP_MAX; }
// byte offset of the device timestamps inside SparseCtl (t_push[2], t_own[3]: 5 x u64)
int px_sparse_ctl_time_offset() { return (int)offsetof(SparseCtl, t_push); }
int px_sparse_ctl_overflow_offset() { return (int)offsetof(SparseCtl, overflow); }

static inline int pick_lpr(int D4) { int l = 1; while (l < D4 && l < 32) l <<= 1; return l; }

// ids_is64: 1 = int64 ids, 0 = int32.
int px_sparse_lookup(const void* ids, int ids_is64, int n, const PxLookupTable* tabs, int nt,
                     int32_t* pend_ids, const PxGroupGeom* g, const void* hdr_mine,
                     const void* ctl, int wait, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (nt < 1 || nt > PX_GRP_MAX) return -4;
  const GroupGeom G = to_geom(g);
  LookupArgs a{};
  a.nt = nt;
  int maxv = 1;
  for (int t = 0; t < nt; ++t) {
    a.t[t].srcs = (const void* const*)tabs[t].srcs; a.t[t].out = tabs[t].out;
    a.t[t].D4 = tabs[t].D4; a.t[t].src_bf16 = tabs[t].src_bf16; a.t[t].out_bf16 = tabs[t].out_bf16;
    const int v = tabs[t].src_bf16 ? (tabs[t].D4 + 1) / 2 : tabs[t].D4;
    if (v > maxv) maxv = v;
  }
  const int lpr = pick_lpr(maxv);
  const int threads = 256, rpb = threads / lpr;
  int blocks = (n + rpb - 1) / rpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const uint32_t* applied = reinterpret_cast<const uint32_t*>(hdr_mine) + PX_MAX_RANKS;
  if (ids_is64)
    px_sparse_lookup_kernel<long long><<<blocks, threads, 0, stream>>>(
        (const long long*)ids, n, a, pend_ids, G, applied, (const SparseCtl*)ctl, lpr, wait);
  else
    px_sparse_lookup_kernel<int><<<blocks, threads, 0, stream>>>(
        (const int*)ids, n, a, pend_ids, G, applied, (const SparseCtl*)ctl, lpr, wait);
  return (int)cudaGetLastError();
}

static inline int push_hbits(int n, int blocks) {
  // >= 4x the expected ids per CTA, 1024..8192 slots (16 B of SMEM per slot)
  long long want = 4LL * ((n + blocks - 1) / blocks);
  int hb = 10;
  while ((1LL << hb) < want && hb < 13) ++hb;
  return hb;
}

// grad_dtype / wire_dtype: 0 fp32, 1 bf16.  async: 1 = remote optimizer application (Hogwild).
// All member tables of a group use the same optimizer kind family.
int px_sparse_push(const int32_t* pend_ids, int n, const PxPushTable* tabs, int nt,
                   int grad_dtype, int wire_dtype, int async, void* ring_ids_dev, void* hdrs_dev,
                   int cap, const PxGroupGeom* g, void* ctl, int rank, int dedup, int max_blocks,
                   cudaStream_t stream) {
  if (nt < 1 || nt > PX_GRP_MAX) return -4;
  const GroupGeom G = to_geom(g);
  PushArgs a{};
  a.nt = nt; a.ring_ids = (int32_t* const*)ring_ids_dev; a.hdrs = (uint32_t* const*)hdrs_dev;
  a.cap = cap; a.rank = rank;
  int fam = 0;
  for (int t = 0; t < nt; ++t) {
    PushTable& T = a.t[t];
    T.grads = tabs[t].grads; T.staging = tabs[t].staging; T.rings = (char* const*)tabs[t].rings;
    T.tables = (float* const*)tabs[t].tables; T.slot0s = (float* const*)tabs[t].slot0s;
    T.slot1s = (float* const*)tabs[t].slot1s; T.slot2s = (float* const*)tabs[t].slot2s;
    T.shadows = (__nv_bfloat16* const*)tabs[t].shadows; T.hp = tabs[t].hp;
    T.D4 = tabs[t].D4; T.kind = tabs[t].kind; T.scale = tabs[t].scale;
    if (t == 0) fam = PX_KIND_FAMILY(T.kind);
    else if (fam != PX_KIND_FAMILY(T.kind)) return -6;
  }
  int blocks = (n + 15) / 16;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  const int hbits = push_hbits(n, blocks);
  const size_t smem = ((size_t)4 * sizeof(int32_t) << hbits) + 2 * PX_ID_CHUNK * sizeof(int32_t);
  SparseCtl* C = (SparseCtl*)ctl;
#define PUSH(GT, WT, AS, FAM) launch_push<GT, WT, AS, FAM>(blocks, smem, stream, pend_ids, n, a, G, C, hbits, dedup)
  if (async) {
    if (grad_dtype == 0) { if (fam == 0) PUSH(float, float, true, 0); else PUSH(float, float, true, 1); }
    else { if (fam == 0) PUSH(__nv_bfloat16, float, true, 0); else PUSH(__nv_bfloat16, float, true, 1); }
  } else if (grad_dtype == 0) {
    if (wire_dtype != 0) return -5;             // never narrow fp32 gradients
    PUSH(float, float, false, 0);
  } else {
    if (wire_dtype == 0) PUSH(__nv_bfloat16, float, false, 0);
    else PUSH(__nv_bfloat16, __nv_bfloat16, false, 0);
  }
#undef PUSH
  return (int)cudaGetLastError();
}

int px_sparse_owner(const PxOwnerTable* tabs, int nt, int wire_dtype, const int32_t* ring_ids,
                    void* hdr, void* hdrs_dev, int32_t* slotmap, int32_t* next, int cap,
                    const PxGroupGeom* g, void* ctl, int rank, int use_merge, int blocks,
                    int fixed_cnt, cudaStream_t stream) {
  if (nt < 1 || nt > PX_GRP_MAX) return -4;
  GroupGeom G = to_geom(g);
  OwnerArgs a{};
  a.fixed_cnt = fixed_cnt;
  a.nt = nt; a.ring_ids = ring_ids; a.hdr = (uint32_t*)hdr; a.hdrs = (uint32_t* const*)hdrs_dev;
  a.slotmap = slotmap; a.next = next; a.cap = cap; a.rank = rank; a.use_merge = use_merge;
  int fam = 0;
  for (int t = 0; t < nt; ++t) {
    OwnerTable& T = a.t[t];
    T.ring = (char*)tabs[t].ring; T.table = tabs[t].table; T.slot0 = tabs[t].slot0;
    T.slot1 = tabs[t].slot1; T.slot2 = tabs[t].slot2; T.shadow = (__nv_bfloat16*)tabs[t].shadow;
    T.hp = tabs[t].hp; T.D4 = tabs[t].D4; T.kind = tabs[t].kind; T.avg = tabs[t].avg;
    if (t == 0) fam = PX_KIND_FAMILY(T.kind);
    else if (fam != PX_KIND_FAMILY(T.kind)) return -6;
  }
  if (blocks < 1) blocks = 1;
  if (fixed_cnt < 0 && G.W > 1)
    px_sparse_wait_kernel<<<1, 32, 0, stream>>>((const uint32_t*)hdr, (const SparseCtl*)ctl, G.W);
  static int max_coop = 0;
  if (max_coop == 0) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, px_sparse_owner_kernel<float, 1>, 256,
                                                  0);
    max_coop = per_sm * sms;
    if (max_coop < 1) max_coop = 1;
  }
  SparseCtl* C = (SparseCtl*)ctl;
  const void* fn;
  if (wire_dtype == 0) fn = fam == 0 ? (const void*)px_sparse_owner_kernel<float, 0>
                                     : (const void*)px_sparse_owner_kernel<float, 1>;
  else fn = fam == 0 ? (const void*)px_sparse_owner_kernel<__nv_bfloat16, 0>
                     : (const void*)px_sparse_owner_kernel<__nv_bfloat16, 1>;
  void* args[] = {&a, &G, &C};
  cudaError_t e;
  if (use_merge) {
    // one grid barrier inside: every CTA must be resident
    if (blocks > max_coop) blocks = max_coop;
    if (blocks > 148 * 4) blocks = 148 * 4;
    e = cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, 0, stream);
  } else {
    e = cudaLaunchKernel(fn, dim3(blocks), dim3(256), args, 0, stream);
  }
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

}  // extern "C"
