// Fused dense step: reduce-scatter (pull) → fp32 optimizer on the owned slice
// → all-gather of the *updated parameters* (push), in ONE kernel.
//
// What it replaces in the reference:
//   AR / HYBRID dense : ncclAllReduce on the fusion buffer + `tf.div` +
//     one Apply<Optimizer> Eigen kernel per variable on every replica
//     (horovod/common/ops/nccl_operations.cc:60-109,
//      horovod/tensorflow/__init__.py:76-81,
//      tensorflow/core/kernels/training_ops_gpu.cu.cc:28-283)
//   PS dense (sync)   : ConditionalAccumulator.take_grad(num_workers) on the PS
//     CPU, chief applies, token queues, mirror-variable refresh
//     (graph_transform_lib.py:330-582, :584-704)
// Both are the same data movement on an NVSwitch box: the rank that owns
// slice r is that slice's "parameter server".  Owning the slice also means
// only 1/W of the fp32 master weights and optimizer slots live on each GPU.
//
// MODE 0 FUSED        : barrier, reduce, update, push params, barrier
// MODE 1 REDUCE_ONLY  : barrier, reduce → fp32 scratch + Σg² (for global-norm
//                       clipping the norm must be known before any update)
// MODE 2 UPDATE_PUSH  : scratch·clip → update, push params, barrier
// World 1 degenerates to a fused multi-tensor optimizer (also used as the
// local update after a plain all-reduce in "replicated" AR mode).
#include "common.cuh"
#include "launch.h"
#include "optim_rules.cuh"

struct DenseStepArgs {
  PeerPtrs grads;    // rotated, element type T
  PeerPtrs params;   // rotated, element type T
  float* master;     // [slice] fp32 master weights of the owned slice
  float* slot0;      // [slice] or null
  float* slot1;      // [slice] or null
  float* slot2;      // [slice] or null (centered RMSProp)
  float* ema;        // [slice] or null
  float* red;        // [slice] fp32 scratch (modes 1/2) or null
  const float* hp;   // device hyper-parameters (8 floats)
  const float* clip; // device scalar multiplier or null
  float* sumsq;      // device scalar accumulator or null
  size_t n;          // bucket elements, multiple of W * VN
  float avg;         // 1/num_workers (or 1)
  float ema_decay;
  int rank, ch_start, ch_end, kind, mode;
  int use_mc;        // 1: grads.p[0] / params.p[0] are NVSwitch multicast addresses
};

// NVLS: the switch reduces (fp32 accumulate) / broadcasts; see collectives.cu
template <typename T>
__device__ __forceinline__ uint4 ds_mm_ld_reduce(const T* p) {
  uint4 r;
  if (sizeof(T) == 2)
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  else
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void ds_mm_st(void* p, const uint4& r) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

template <int VN>
__device__ __forceinline__ void ld_f32(const float* p, float* f) {
#pragma unroll
  for (int i = 0; i < VN / 4; ++i) {
    const uint4 v = ld_v4(p + 4 * i);
    f[4 * i] = __uint_as_float(v.x); f[4 * i + 1] = __uint_as_float(v.y);
    f[4 * i + 2] = __uint_as_float(v.z); f[4 * i + 3] = __uint_as_float(v.w);
  }
}
template <int VN>
__device__ __forceinline__ void st_f32(float* p, const float* f) {
#pragma unroll
  for (int i = 0; i < VN / 4; ++i)
    st_v4(p + 4 * i, make_uint4(__float_as_uint(f[4 * i]), __float_as_uint(f[4 * i + 1]),
                                __float_as_uint(f[4 * i + 2]), __float_as_uint(f[4 * i + 3])));
}

// Rank-level (not CTA-paired) synchronisation, so the grid is not limited to the
// PX_MAX_BLOCKS barrier slots of `px_block_barrier` and the optimizer phase can use the whole
// GPU's HBM bandwidth:
//  * start: CTA 0 announces "this rank reached the kernel" (all earlier work on the stream —
//    the gradients — is complete) in slot 0 of `ch_start`; EVERY CTA waits until every peer has.
//  * end: the last CTA to finish (ticket) fences, announces "all my parameter stores are out"
//    in slot 0 of `ch_end` and waits for the same from every peer; the kernel — and with it the
//    stream — completes only then.  Slot 1 of `ch_end` holds the ticket counter.
__device__ __forceinline__ unsigned long long px_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void px_rank_signal(uint32_t* const* pads, int slot, int rank, int world,
                                               uint32_t e) {
  if (threadIdx.x < world)
    st_release_sys(pads[threadIdx.x] + (size_t)slot * PX_MAX_RANKS + rank, e);
}
__device__ __forceinline__ void px_rank_wait(uint32_t* const* pads, int slot, int rank, int world,
                                             uint32_t e) {
  if (threadIdx.x < world) {
    const uint32_t* mine = pads[rank] + (size_t)slot * PX_MAX_RANKS + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) { }
  }
  __syncthreads();
}

template <typename T, int W, int FAM>
__global__ void __launch_bounds__(512, (W == 1 && FAM == 0) ? 2 : 1)
px_dense_step_kernel(DenseStepArgs a, uint32_t* const* pads, uint32_t* epoch_ctr) {
  constexpr int VN = Vec16<T>::N;
  const int mode = a.mode, kind = a.kind;
  const int slot_s = a.ch_start * PX_MAX_BLOCKS, slot_e = a.ch_end * PX_MAX_BLOCKS;
  // profiling aid: %globaltimer stamps of the last launch in the (otherwise unused) epoch slots
  // of the last channel: [0] CTA 0 start, [1] CTA 0 past the start wait, [2] CTA 0 loop done,
  // [3] last CTA fenced, [4] last CTA past the end wait
  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(
      epoch_ctr + (PX_NUM_CHANNELS - 1) * PX_MAX_BLOCKS);
  const bool stamp0 = blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp0) dbg[0] = px_timer();
  uint32_t e_start = 0;
  if (W > 1 && mode != 2) {
    e_start = ld_volatile_u32(epoch_ctr + slot_s) + 1;
    if (blockIdx.x == 0) px_rank_signal(pads, slot_s, a.rank, W, e_start);
    px_rank_wait(pads, slot_s, a.rank, W, e_start);
  }
  if (stamp0) dbg[1] = px_timer();
  const size_t slice = a.n / W;
  const size_t nvec = slice / VN;
  const size_t base = (size_t)a.rank * slice;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const PxHP h = px_load_hp(a.hp);
  float gmul = a.avg * a.hp[HP_GSCALE];
  if (mode == 2) gmul = 1.f;                       // already applied in REDUCE
  if (mode != 1 && a.clip != nullptr) gmul *= *a.clip;
  float ss = 0.f;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const size_t e = v * VN;                        // element offset inside the slice
    // ---- issue EVERY load of this vector before touching any loaded value: the helpers are
    // volatile asm (program order), so a dependent FADD between the peer loads and the
    // master/slot loads would serialise two memory latencies (one of them an NVLink round
    // trip) per vector
    float g[VN];
    uint4 in[W];
    uint4 mcv = make_uint4(0, 0, 0, 0);
    if (mode == 2) {
      ld_f32<VN>(a.red + e, g);
    } else if (a.use_mc) {
      // one switch-side reduction instead of W peer loads
      mcv = ds_mm_ld_reduce<T>(reinterpret_cast<const T*>(a.grads.p[0]) + base + e);
    } else {
#pragma unroll
      for (int p = 0; p < W; ++p)
        in[p] = ld_v4_stream(reinterpret_cast<const T*>(a.grads.p[p]) + base + e);
    }
    float w[VN], s0[VN], s1[VN], s2[FAM == 1 ? VN : 1], m[VN];
    if (mode != 1) {
      ld_f32<VN>(a.master + e, w);
      if (a.slot0) ld_f32<VN>(a.slot0 + e, s0);
      if (a.slot1) ld_f32<VN>(a.slot1 + e, s1);
      if (FAM == 1 && a.slot2) ld_f32<VN>(a.slot2 + e, s2);
      if (a.ema) ld_f32<VN>(a.ema + e, m);
    }
    // ---- math
    if (mode != 2) {
      if (a.use_mc) {
        Vec16<T>::unpack(mcv, g);
      } else {
#pragma unroll
        for (int i = 0; i < VN; ++i) g[i] = 0.f;
#pragma unroll
        for (int p = 0; p < W; ++p) {
          float f[VN];
          Vec16<T>::unpack(in[p], f);
#pragma unroll
          for (int i = 0; i < VN; ++i) g[i] += f[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) g[i] *= gmul;
    if (mode == 1) {
#pragma unroll
      for (int i = 0; i < VN; ++i) ss += g[i] * g[i];
      st_f32<VN>(a.red + e, g);
      continue;
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float gi = h.wd != 0.f ? fmaf(h.wd, w[i], g[i]) : g[i];
      px_rule<FAM>(kind, h, gi, w[i], s0[i], s1[i], s2[FAM == 1 ? i : 0]);
    }
    st_f32<VN>(a.master + e, w);
    if (a.slot0) st_f32<VN>(a.slot0 + e, s0);
    if (a.slot1) st_f32<VN>(a.slot1 + e, s1);
    if (FAM == 1 && a.slot2) st_f32<VN>(a.slot2 + e, s2);
    if (a.ema) {
#pragma unroll
      for (int i = 0; i < VN; ++i) m[i] -= (1.f - a.ema_decay) * (m[i] - w[i]);
      st_f32<VN>(a.ema + e, m);
    }
    const uint4 out = Vec16<T>::pack(w);
    if (a.use_mc) {
      ds_mm_st(reinterpret_cast<T*>(a.params.p[0]) + base + e, out);   // switch broadcast
    } else {
#pragma unroll
      for (int p = 0; p < W; ++p)
        st_v4_stream(reinterpret_cast<T*>(a.params.p[p]) + base + e, out);
    }
  }
  if (mode == 1 && a.sumsq != nullptr) block_atomic_sum(ss, a.sumsq);
  if (stamp0) dbg[2] = px_timer();
  if (W > 1) {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();          // one fence per CTA (cumulative over the barrier)
      s_last = atomicAdd(epoch_ctr + slot_e + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      if (threadIdx.x == 0) dbg[3] = px_timer();
      if (mode != 1) {
        const uint32_t e_end = ld_volatile_u32(epoch_ctr + slot_e) + 1;
        px_rank_signal(pads, slot_e, a.rank, W, e_end);
        px_rank_wait(pads, slot_e, a.rank, W, e_end);
        if (threadIdx.x == 0) epoch_ctr[slot_e] = e_end;
      }
      if (threadIdx.x == 0) dbg[4] = px_timer();
      if (threadIdx.x == 0) {
        if (mode != 2) epoch_ctr[slot_s] = e_start;
        epoch_ctr[slot_e + 1] = 0;
      }
    }
  }
}

// device timestamp (ns) — a graph-capturable probe for "exposed communication" measurements
__global__ void px_stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

// scale = max_norm / max(sqrt(total), max_norm)  (tf.clip_by_global_norm);
// also exports the norm and zeroes the accumulator for the next step.
__global__ void px_clip_scale_kernel(const float* sumsq, float max_norm, float* scale_out,
                                     float* norm_out, float* zero_after) {
  const float norm = sqrtf(*sumsq);
  *scale_out = max_norm / fmaxf(norm, max_norm);
  if (norm_out) *norm_out = norm;
  if (zero_after) *zero_after = 0.f;
}

// Asynchronous PS dense apply (Hogwild): this rank's gradient is applied,
// un-averaged, straight onto every owner's master slice over NVLink, and the
// refreshed values are pulled back into the local parameter mirror.
// Reference: sync=False ⇒ no accumulators, update ops race on the PS
// variables (ps/between_graph_parallel.py:137-146).
template <typename T, int W, int FAM>
__global__ void __launch_bounds__(512)
px_dense_async_kernel(const T* __restrict__ my_grads, T* __restrict__ my_params,
                      PeerPtrs master, PeerPtrs slot0, PeerPtrs slot1, PeerPtrs slot2,
                      const float* hp,
                      const float* clip, size_t n, int kind, int rank) {
  constexpr int VN = Vec16<T>::N;
  const size_t slice = n / W;
  const size_t nvec = n / VN;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const PxHP h = px_load_hp(hp);
  float gmul = hp[HP_GSCALE];
  if (clip != nullptr) gmul *= *clip;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const size_t e = v * VN;
    const int owner = (int)(e / slice);
    const size_t le = e - (size_t)owner * slice;
    float *pm = nullptr, *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
#pragma unroll
    for (int p = 0; p < W; ++p)     // master/slots arrive in NATURAL rank order
      if (p == owner) {
        pm = reinterpret_cast<float*>(master.p[p]) + le;
        p0 = slot0.p[p] ? reinterpret_cast<float*>(slot0.p[p]) + le : nullptr;
        p1 = slot1.p[p] ? reinterpret_cast<float*>(slot1.p[p]) + le : nullptr;
        p2 = (FAM == 1 && slot2.p[p]) ? reinterpret_cast<float*>(slot2.p[p]) + le : nullptr;
      }
    float g[VN], w[VN], s0[VN], s1[VN], s2[FAM == 1 ? VN : 1];
    Vec16<T>::unpack(ld_v4(my_grads + e), g);
    ld_f32<VN>(pm, w);
    if (p0) ld_f32<VN>(p0, s0);
    if (p1) ld_f32<VN>(p1, s1);
    if (FAM == 1 && p2) ld_f32<VN>(p2, s2);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      float gi = g[i] * gmul;
      if (h.wd != 0.f) gi = fmaf(h.wd, w[i], gi);
      px_rule<FAM>(kind, h, gi, w[i], s0[i], s1[i], s2[FAM == 1 ? i : 0]);
    }
    st_f32<VN>(pm, w);
    if (p0) st_f32<VN>(p0, s0);
    if (p1) st_f32<VN>(p1, s1);
    if (FAM == 1 && p2) st_f32<VN>(p2, s2);
    st_v4(my_params + e, Vec16<T>::pack(w));
  }
}

// Σx² of a local buffer (used for clipping in the async path)
template <typename T>
__global__ void __launch_bounds__(512)
px_sumsq_kernel(const T* __restrict__ x, size_t n, float mul, float* out) {
  constexpr int VN = Vec16<T>::N;
  const size_t nvec = n / VN;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float ss = 0.f;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float f[VN];
    Vec16<T>::unpack(ld_v4(x + v * VN), f);
#pragma unroll
    for (int i = 0; i < VN; ++i) ss += f[i] * f[i] * mul * mul;
  }
  block_atomic_sum(ss, out);
}

extern "C" {

// dtype 0 fp32 / 1 bf16.  grads/params: `world` pointers in natural order.
int px_dense_step(const void* const* grads, const void* const* params, float* master,
                  float* slot0, float* slot1, float* slot2, float* ema, float* red, const float* hp,
                  const float* clip, float* sumsq, size_t n, float avg, float ema_decay,
                  int kind, int mode, int dtype, void* pads_dev, void* epoch_ctr, int ch_start,
                  int ch_end, int rank, int world, int max_blocks, int use_mc,
                  cudaStream_t stream) {
  if (world < 1 || world > 8) return -3;
  const int vn = dtype == 0 ? 4 : 8;
  if (n % ((size_t)world * vn) != 0) return -1;
  DenseStepArgs a;
  a.use_mc = use_mc;
  if (use_mc) {            // entry 0 = multicast address; no peer pointers needed
    a.grads = PeerPtrs{}; a.params = PeerPtrs{};
    a.grads.p[0] = const_cast<void*>(grads[0]);
    a.params.p[0] = const_cast<void*>(params[0]);
  } else {
  a.grads = px_rotate(grads, rank, world);
  a.params = px_rotate(params, rank, world);
  }
  a.master = master; a.slot0 = slot0; a.slot1 = slot1; a.slot2 = slot2; a.ema = ema; a.red = red;
  a.hp = hp; a.clip = clip; a.sumsq = sumsq; a.n = n; a.avg = avg; a.ema_decay = ema_decay;
  a.rank = rank; a.ch_start = ch_start; a.ch_end = ch_end; a.kind = kind; a.mode = mode;
  const int threads = 512;
  // rank-level barriers: the grid is sized for HBM bandwidth, not by barrier slots
  size_t b = (n / world / vn + threads - 1) / threads;
  const size_t cap = world == 1 ? 148 * 4 : (size_t)(max_blocks > 0 ? max_blocks : 148);
  int blocks = (int)(b < 1 ? 1 : (b > cap ? cap : b));
#define LAUNCH(T, W)                                                                       \
  do {                                                                                     \
    if (PX_KIND_FAMILY(kind) == 0)                                                         \
      px_dense_step_kernel<T, W, 0><<<blocks, threads, 0, stream>>>(                       \
          a, (uint32_t* const*)pads_dev, (uint32_t*)epoch_ctr);                            \
    else                                                                                   \
      px_dense_step_kernel<T, W, 1><<<blocks, threads, 0, stream>>>(                       \
          a, (uint32_t* const*)pads_dev, (uint32_t*)epoch_ctr);                            \
  } while (0)
  if (dtype == 0) { PX_DISPATCH_WORLD(world, LAUNCH, float); }
  else { PX_DISPATCH_WORLD(world, LAUNCH, __nv_bfloat16); }
#undef LAUNCH
  return (int)cudaGetLastError();
}

int px_stamp(void* slot, cudaStream_t stream) {
  px_stamp_kernel<<<1, 1, 0, stream>>>((unsigned long long*)slot);
  return (int)cudaGetLastError();
}

int px_clip_scale(const float* sumsq, float max_norm, float* scale_out, float* norm_out,
                  float* zero_after, cudaStream_t stream) {
  px_clip_scale_kernel<<<1, 1, 0, stream>>>(sumsq, max_norm, scale_out, norm_out, zero_after);
  return (int)cudaGetLastError();
}

int px_dense_async(const void* my_grads, void* my_params, const void* const* master,
                   const void* const* slot0, const void* const* slot1,
                   const void* const* slot2, const float* hp,
                   const float* clip, size_t n, int kind, int dtype, int rank, int world,
                   int max_blocks, cudaStream_t stream) {
  if (world < 1 || world > 8) return -3;
  const int vn = dtype == 0 ? 4 : 8;
  if (n % ((size_t)world * vn) != 0) return -1;
  PeerPtrs M{}, S0{}, S1{}, S2{};
  for (int i = 0; i < world; ++i) {
    M.p[i] = const_cast<void*>(master[i]);
    S0.p[i] = slot0 ? const_cast<void*>(slot0[i]) : nullptr;
    S1.p[i] = slot1 ? const_cast<void*>(slot1[i]) : nullptr;
    S2.p[i] = slot2 ? const_cast<void*>(slot2[i]) : nullptr;
  }
  const int blocks = px_clamp_blocks(n / vn, 512, max_blocks);
#define LAUNCH(T, W)                                                                   \
  do {                                                                                 \
    if (PX_KIND_FAMILY(kind) == 0)                                                     \
      px_dense_async_kernel<T, W, 0><<<blocks, 512, 0, stream>>>(                      \
          (const T*)my_grads, (T*)my_params, M, S0, S1, S2, hp, clip, n, kind, rank);  \
    else                                                                               \
      px_dense_async_kernel<T, W, 1><<<blocks, 512, 0, stream>>>(                      \
          (const T*)my_grads, (T*)my_params, M, S0, S1, S2, hp, clip, n, kind, rank);  \
  } while (0)
  if (dtype == 0) { PX_DISPATCH_WORLD(world, LAUNCH, float); }
  else { PX_DISPATCH_WORLD(world, LAUNCH, __nv_bfloat16); }
#undef LAUNCH
  return (int)cudaGetLastError();
}

int px_sumsq(const void* x, size_t n, int dtype, float mul, float* out, cudaStream_t stream) {
  const int vn = dtype == 0 ? 4 : 8;
  const int blocks = px_clamp_blocks(n / vn, 512 * 4, 148 * 2);
  if (dtype == 0) px_sumsq_kernel<float><<<blocks, 512, 0, stream>>>((const float*)x, n, mul, out);
  else px_sumsq_kernel<__nv_bfloat16><<<blocks, 512, 0, stream>>>((const __nv_bfloat16*)x, n, mul, out);
  return (int)cudaGetLastError();
}

}  // extern "C"
