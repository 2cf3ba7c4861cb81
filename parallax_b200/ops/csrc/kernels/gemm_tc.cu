// tcgen05 / TMEM / TMA GEMM for the skinny recurrent products of the LSTM
// (M = batch = 128·k rows, huge K or huge N) — hand-written for sm_100a.
//
//   C[M,N] (bf16) = A[M,K] (bf16, K-contiguous) · B[N,K]^T (bf16, K-contiguous)
//                   (+ addend[M,N] bf16)                        "TN" GEMM
//
// Roles (one CTA = 8 warps): warp 0 TMA producer (cp.async.bulk.tensor, 128B
// swizzle), warp 1 single-thread tcgen05.mma issuer (UMMA 128×BN×16, fp32
// accumulator in TMEM), warp 2 TMEM allocator, warps 4-7 epilogue
// (tcgen05.ld 32x32b → registers → global).  STAGES-deep smem ring with
// full/empty mbarriers; MMA completion is signalled with tcgen05.commit.
//
// Split-K: grid.z CTAs each reduce a K-slice and red.add their fp32 tile into
// an L2-resident workspace; the last CTA to arrive for a tile (atomic ticket)
// converts (+addend) to bf16, stores C and re-zeroes the workspace — so a
// [128, 8192]·[8192, 512] product runs on 4·32 CTAs instead of 4, without a
// second launch.  The reference reaches these products through cuBLAS
// (tensorflow/core/kernels/matmul_op.cc:252-369 → cuda_blas.cc:2229); cuBLAS'
// heuristic picks a 64x8-tile kernel for this shape that takes ~11 µs.
#include <cuda.h>
#include "common.cuh"

namespace tc {

constexpr int BM = 128;       // UMMA M
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// K-major operand tile, 128B swizzle: rows at 128 B pitch, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address
  d |= (uint64_t)0 << 16;                             // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // SBO
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
               "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmArgs {
  __nv_bfloat16* C;            // [M, N]
  const __nv_bfloat16* addend; // [M, N] or null
  float* ws;                   // [M, N] fp32, zero between calls (split-K only)
  unsigned int* tickets;       // [(M/128) * (N/BN)] zero between calls (split-K only)
  int M, N, K;                 // K = full reduction length
  int k_per_split;             // multiple of BK
};

// CLUSTER = true: the K-splits of one output tile form a thread-block cluster (1,1,splits) and
// reduce their fp32 partial tiles through distributed shared memory — each CTA parks its tile in
// its own SMEM (the drained pipeline stages), cluster barrier, then every CTA sums 1/splits of
// the rows straight out of its peers' SMEM (`ld.shared::cluster`), adds the addend and stores
// bf16.  No L2 `red.add`, no ticket, no read-back pass, no workspace.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BN, int STAGES, bool CLUSTER>
__global__ void __launch_bounds__(256, 1)
px_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [STAGES][A 16 KB][B BN*128 B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  __shared__ unsigned int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int k0 = split * g.k_per_split;
  const int num_kb = g.k_per_split / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::
                 "r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------ TMA producer ------------------------------
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      uint8_t* sb = sa + A_BYTES;
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      tma_load_2d(sa, &tmap_a, &full_bar[s], k0 + kb * BK, m_tile * BM);
      tma_load_2d(sb, &tmap_b, &full_bar[s], k0 + kb * BK, n_tile * BN);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------- MMA issuer -------------------------------
    // instr desc: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major both,
    // N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t sb = sa + A_BYTES;
      const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        // advance 32 B (= UMMA_K bf16) inside the swizzle atom: +2 in the
        // 16-byte-granular start-address field
        umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                  (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(&empty_bar[s]);          // frees the smem slot when the MMAs retire
    }
    umma_commit(tmem_full);                // accumulator complete
  } else if (warp >= 4) {
    // -------------------------------- epilogue --------------------------------
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int wq = warp - 4;                         // TMEM lane quarter
    const int row = m_tile * BM + wq * 32 + lane;    // output row of this thread
    const bool splitk = gridDim.z > 1;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
      const size_t off = (size_t)row * g.N + (size_t)n_tile * BN + c0;
      if (CLUSTER) {
        // park the fp32 partial in shared memory (row stride BN+4 floats)
        float4* dst = reinterpret_cast<float4*>(smem) + ((size_t)(wq * 32 + lane) * (BN + 4) + c0) / 4;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          dst[i / 4] = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                   __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      } else if (row < g.M) {
        if (splitk) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            atomicAdd(reinterpret_cast<float4*>(g.ws + off + i),
                      make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                  __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
        } else {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
          if (g.addend) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              float a[8];
              Vec16<__nv_bfloat16>::unpack(ld_v4(g.addend + off + i), a);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[i + j] += a[j];
            }
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) st_v4(g.C + off + i, Vec16<__nv_bfloat16>::pack(f + i));
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)BN) : "memory");
  }
  if (CLUSTER) {
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    cluster_sync_all();                               // every split's partial is in its SMEM
    const int rows_per = BM / (int)csize;             // csize divides 128
    const int nvec = rows_per * BN / 4;               // float4 groups this CTA reduces
    const uint32_t my_base = smem_u32(smem);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const int rr = (int)crank * rows_per + (v * 4) / BN, cc = (v * 4) % BN;
      const uint32_t off_b = (uint32_t)(((size_t)rr * (BN + 4) + cc) * 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t p = 0; p < csize; ++p) {
        uint32_t raddr;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(my_base + off_b), "r"(p));
        float4 x;
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(raddr) : "memory");
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
      const int row = m_tile * BM + rr;
      if (row < g.M) {
        const size_t off = (size_t)row * g.N + (size_t)n_tile * BN + cc;
        if (g.addend) {
          const uint2 a2 = *reinterpret_cast<const uint2*>(g.addend + off);
          acc.x += __uint_as_float(a2.x << 16); acc.y += __uint_as_float(a2.x & 0xffff0000u);
          acc.z += __uint_as_float(a2.y << 16); acc.w += __uint_as_float(a2.y & 0xffff0000u);
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
        *reinterpret_cast<uint2*>(g.C + off) =
            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
    cluster_sync_all();                               // nobody exits while a peer reads its SMEM
    return;
  }
  if (gridDim.z > 1) {
    // last-arriving split for this tile finalises: ws (+addend) -> bf16 C, ws := 0
    __threadfence();
    __syncthreads();
    const unsigned int tile_id = m_tile * gridDim.x + n_tile;
    if (threadIdx.x == 0) s_last = (atomicAdd(&g.tickets[tile_id], 1u) == gridDim.z - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int idx = threadIdx.x; idx < BM * BN / 8; idx += blockDim.x) {
        const int rr = idx / (BN / 8), cc = (idx % (BN / 8)) * 8;
        const int row = m_tile * BM + rr;
        if (row >= g.M) continue;
        const size_t off = (size_t)row * g.N + (size_t)n_tile * BN + cc;
        float f[8];
        const uint4 lo = __ldcg(reinterpret_cast<const uint4*>(g.ws + off));
        const uint4 hi = __ldcg(reinterpret_cast<const uint4*>(g.ws + off + 4));
        f[0] = __uint_as_float(lo.x); f[1] = __uint_as_float(lo.y); f[2] = __uint_as_float(lo.z);
        f[3] = __uint_as_float(lo.w); f[4] = __uint_as_float(hi.x); f[5] = __uint_as_float(hi.y);
        f[6] = __uint_as_float(hi.z); f[7] = __uint_as_float(hi.w);
        if (g.addend) {
          float a[8];
          Vec16<__nv_bfloat16>::unpack(ld_v4(g.addend + off), a);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += a[j];
        }
        st_v4(g.C + off, Vec16<__nv_bfloat16>::pack(f));
        __stcg(reinterpret_cast<uint4*>(g.ws + off), make_uint4(0, 0, 0, 0));
        __stcg(reinterpret_cast<uint4*>(g.ws + off + 4), make_uint4(0, 0, 0, 0));
      }
      if (threadIdx.x == 0) g.tickets[tile_id] = 0;
    }
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                 "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) {
  const float e = __expf(-2.f * fabsf(x));
  return copysignf((1.f - e) / (1.f + e), x);
}

// ---------------------------------------------------------------------------
// Recurrent LSTM step with the whole cell fused into the GEMM epilogue:
//   gates = xw_t + h_{t-1} · Wh          (tcgen05, accumulator in TMEM)
//   c_t = σ(f+fb)·c_{t-1} + σ(i)·tanh(j);  m_t = σ(o)·tanh(c_t)
// The 4S gate columns are stored GATE-INTERLEAVED: tile n (128 columns) holds
// [i | j | f | o] × 32 hidden units (n·32 … n·32+31), so one CTA owns every
// gate of its 32 units and the cell update never leaves the SM.  Each epilogue
// thread owns one batch row of the TMEM tile.  Replaces cuBLAS addmm + the
// stand-alone cell kernel (and the [B,4S] pre-activation round trip).
struct LstmArgs {
  const __nv_bfloat16* xw;   // [M, 4S] permuted layout (input half + bias)
  const float* c_prev;       // [M, S]
  float* c_new;              // [M, S]
  __nv_bfloat16* m_out;      // [M, S]
  __nv_bfloat16* act;        // [M, 4S] permuted layout: σ(i) | tanh(j) | σ(f+fb) | σ(o)
  int M, S, K;
  float forget_bias;
};

template <int STAGES>
__global__ void __launch_bounds__(256, 1)
px_lstm_gates_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                        const __grid_constant__ CUtensorMap tmap_b, LstmArgs g) {
  constexpr int BN = 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int num_kb = g.K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::
                 "r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      tma_load_2d(sa, &tmap_a, &full_bar[s], kb * BK, m_tile * BM);
      tma_load_2d(sa + A_BYTES, &tmap_b, &full_bar[s], kb * BK, n_tile * BN);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
      const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k)
        umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                  (kb | k) != 0 ? 1u : 0u);
      umma_commit(&empty_bar[s]);
    }
    umma_commit(tmem_full);
  } else if (warp >= 4) {
    const int wq = warp - 4;
    const int row = m_tile * BM + wq * 32 + lane;
    const size_t gcol0 = (size_t)row * 4 * g.S + (size_t)n_tile * BN;   // permuted column base
    const size_t ccol0 = (size_t)row * g.S + (size_t)n_tile * 32;       // hidden-unit base
    // The addend (xw) and c_{t-1} do not depend on the MMA: fetch this row's
    // 128 + 32 values while TMA/UMMA are still filling the accumulator.
    uint4 xv[4][4];      // [q][gate]
    uint4 cv[4][2];
    if (row < g.M) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) xv[q][gi] = ld_v4(g.xw + gcol0 + gi * 32 + q * 8);
        cv[q][0] = ld_v4(g.c_prev + ccol0 + q * 8);
        cv[q][1] = ld_v4(g.c_prev + ccol0 + q * 8 + 4);
      }
    }
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {                                       // 8 hidden units per pass
      uint32_t ri[8], rj[8], rf[8], ro[8];
      tmem_ld8(trow + 0 * 32 + q * 8, ri);
      tmem_ld8(trow + 1 * 32 + q * 8, rj);
      tmem_ld8(trow + 2 * 32 + q * 8, rf);
      tmem_ld8(trow + 3 * 32 + q * 8, ro);
      tmem_ld_wait();
      if (row < g.M) {
        float xi[8], xj[8], xf[8], xo[8], cp[8];
        Vec16<__nv_bfloat16>::unpack(xv[q][0], xi);
        Vec16<__nv_bfloat16>::unpack(xv[q][1], xj);
        Vec16<__nv_bfloat16>::unpack(xv[q][2], xf);
        Vec16<__nv_bfloat16>::unpack(xv[q][3], xo);
        Vec16<float>::unpack(cv[q][0], cp);
        Vec16<float>::unpack(cv[q][1], cp + 4);
        float si[8], tj[8], sf[8], so[8], cn[8], mo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          // pre-activations stay fp32 (TMEM accumulator + bf16 addend)
          si[u] = sigm(__uint_as_float(ri[u]) + xi[u]);
          tj[u] = tanh_(__uint_as_float(rj[u]) + xj[u]);
          sf[u] = sigm(__uint_as_float(rf[u]) + xf[u] + g.forget_bias);
          so[u] = sigm(__uint_as_float(ro[u]) + xo[u]);
          cn[u] = sf[u] * cp[u] + si[u] * tj[u];
          mo[u] = so[u] * tanh_(cn[u]);
        }
        st_v4(g.c_new + ccol0 + q * 8, Vec16<float>::pack(cn));
        st_v4(g.c_new + ccol0 + q * 8 + 4, Vec16<float>::pack(cn + 4));
        st_v4(g.m_out + ccol0 + q * 8, Vec16<__nv_bfloat16>::pack(mo));
        st_v4(g.act + gcol0 + 0 * 32 + q * 8, Vec16<__nv_bfloat16>::pack(si));
        st_v4(g.act + gcol0 + 1 * 32 + q * 8, Vec16<__nv_bfloat16>::pack(tj));
        st_v4(g.act + gcol0 + 2 * 32 + q * 8, Vec16<__nv_bfloat16>::pack(sf));
        st_v4(g.act + gcol0 + 3 * 32 + q * 8, Vec16<__nv_bfloat16>::pack(so));
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)BN) : "memory");
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128B swizzle
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols,
                     uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -10;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

}  // namespace tc

extern "C" {

// C[M,N] = A[M,K] · B[N,K]^T (+ addend).  splits > 1 needs ws (fp32 [M,N], zero)
// and tickets (uint32 [tiles], zero).  Returns 0 or a negative error.
// cluster != 0: the splits of a tile reduce through DSMEM in a (1,1,splits) cluster (splits must
// divide 128 and be <= 16; 16 needs the non-portable cluster size); ws / tickets unused.
int px_gemm_tc(const void* A, const void* B, void* C, const void* addend, float* ws,
               unsigned int* tickets, int M, int N, int K, int splits, int bn, int cluster,
               cudaStream_t stream) {
  using namespace tc;
  if (M % BM || K % BK || (bn != 64 && bn != 128) || N % bn) return -1;
  if (splits < 1 || K % (splits * BK)) return -2;
  if (cluster && (splits < 2 || splits > 16 || (BM % splits) != 0)) return -4;
  if (splits > 1 && !cluster && (!ws || !tickets)) return -3;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, M, K, BM);
  if (rc) return rc;
  rc = make_tmap(&tb, B, N, K, bn);
  if (rc) return rc;
  GemmArgs g;
  g.C = (__nv_bfloat16*)C; g.addend = (const __nv_bfloat16*)addend; g.ws = ws;
  g.tickets = tickets; g.M = M; g.N = N; g.K = K; g.k_per_split = K / splits;
  dim3 grid(N / bn, M / BM, splits);
  constexpr int STAGES = 4;
  if (cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = splits;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e;
    if (bn == 128) {
      constexpr int SMEM = STAGES * (BM * BK * 2 + 128 * BK * 2) + 1024 + 256;
      static bool setc128 = false;
      if (!setc128) {
        cudaFuncSetAttribute(px_gemm_tc_kernel<128, STAGES, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        cudaFuncSetAttribute(px_gemm_tc_kernel<128, STAGES, true>,
                             cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        setc128 = true;
      }
      cfg.dynamicSmemBytes = SMEM;
      e = cudaLaunchKernelEx(&cfg, px_gemm_tc_kernel<128, STAGES, true>, ta, tb, g);
    } else {
      constexpr int SMEM = STAGES * (BM * BK * 2 + 64 * BK * 2) + 1024 + 256;
      static bool setc64 = false;
      if (!setc64) {
        cudaFuncSetAttribute(px_gemm_tc_kernel<64, STAGES, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        cudaFuncSetAttribute(px_gemm_tc_kernel<64, STAGES, true>,
                             cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        setc64 = true;
      }
      cfg.dynamicSmemBytes = SMEM;
      e = cudaLaunchKernelEx(&cfg, px_gemm_tc_kernel<64, STAGES, true>, ta, tb, g);
    }
    return e == cudaSuccess ? (int)cudaGetLastError() : (int)e;
  }
  if (bn == 128) {
    constexpr int SMEM = STAGES * (BM * BK * 2 + 128 * BK * 2) + 1024 + 256;
    static bool set128 = false;
    if (!set128) {
      cudaFuncSetAttribute(px_gemm_tc_kernel<128, STAGES, false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      set128 = true;
    }
    px_gemm_tc_kernel<128, STAGES, false><<<grid, 256, SMEM, stream>>>(ta, tb, g);
  } else {
    constexpr int SMEM = STAGES * (BM * BK * 2 + 64 * BK * 2) + 1024 + 256;
    static bool set64 = false;
    if (!set64) {
      cudaFuncSetAttribute(px_gemm_tc_kernel<64, STAGES, false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      set64 = true;
    }
    px_gemm_tc_kernel<64, STAGES, false><<<grid, 256, SMEM, stream>>>(ta, tb, g);
  }
  return (int)cudaGetLastError();
}

// Fused recurrent LSTM step (see px_lstm_gates_tc_kernel).  h: [M,K] bf16;
// WhP: [4S, K] bf16 gate-interleaved rows; xw/act: [M,4S] gate-interleaved.
int px_lstm_gates_tc(const void* h, const void* WhP, const void* xw, const float* c_prev,
                     float* c_new, void* m_out, void* act, int M, int S, int K,
                     float forget_bias, cudaStream_t stream) {
  using namespace tc;
  if (M % BM || K % BK || S % 32) return -1;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, h, M, K, BM);
  if (rc) return rc;
  rc = make_tmap(&tb, WhP, 4 * (uint64_t)S, K, 128);
  if (rc) return rc;
  LstmArgs g;
  g.xw = (const __nv_bfloat16*)xw; g.c_prev = c_prev; g.c_new = c_new;
  g.m_out = (__nv_bfloat16*)m_out; g.act = (__nv_bfloat16*)act;
  g.M = M; g.S = S; g.K = K; g.forget_bias = forget_bias;
  constexpr int STAGES = 4;
  constexpr int SMEM = STAGES * (BM * BK * 2 + 128 * BK * 2) + 1024 + 256;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(px_lstm_gates_tc_kernel<STAGES>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    set = true;
  }
  dim3 grid(4 * S / 128, M / BM, 1);
  px_lstm_gates_tc_kernel<STAGES><<<grid, 256, SMEM, stream>>>(ta, tb, g);
  return (int)cudaGetLastError();
}

}  // extern "C"
