// 2-CTA (cta_group::2) persistent tcgen05 GEMM for sm_100a: a CTA pair (cluster of 2, same TPC) owns one 256 x 256 output
// tile.  Each CTA stages its own 128 rows of A and only HALF of B (128 of the 256 N rows); the leader CTA's single MMA
// thread issues tcgen05.mma.cta_group::2 (M = 256), which reads A from both CTAs' shared memory and the two halves of B,
// and leaves rows [0,128) of the accumulator in the leader's TMEM and rows [128,256) in the peer's.
//
// Why: an SS MMA is shared-memory-bandwidth bound: a 1-CTA 128x256x16 step reads 12 KB per 64 math
// clocks, over the ~128 B/clk port; the pair reads 4 KB (A) + 4 KB (half B) per CTA for the same math per SM.
//
// Pipeline (per CTA unless noted):
//   warp0 lane0  TMA producer     waits local empty[s]; leader: arrive.expect_tx(2 x stage bytes) on its full[s];
//                                 peer: remote arrive on the leader's full[s]; both: cp.async.bulk.tensor...cta_group::2
//                                 whose complete_tx lands on the LEADER's full[s] (peer bit of the barrier address cleared)
//   warp1 lane0  MMA issuer       leader only: waits full[s], 4 x tcgen05.mma.cta_group::2, tcgen05.commit multicast to
//                                 empty[s] of both CTAs (and to tmem_full[a] of both after the last k-block)
//   warps 2..5   epilogue         wait local tmem_full[a]; TMEM -> registers -> global (bf16 / fp32, optional +=);
//                                 arrive on the LEADER's tmem_empty[a] (count 8 = 4 warps x 2 CTAs)
//
// All four operand layouts (TN / NN / NT / TT) as in gemm_tcgen05.cu; bf16 inputs.
#include "sm100_common.cuh"

namespace smx {

constexpr int k2BlockM = 256;       // per pair (128 per CTA)
constexpr int k2BlockN = 256;       // per pair (each CTA stages 128 rows of B)
constexpr int k2RowBytes = 128;
constexpr int k2Threads = 192;
constexpr int k2ABytes = 128 * k2RowBytes;           // 16 KB
constexpr int k2BBytes = 128 * k2RowBytes;           // 16 KB (half of B)
constexpr int k2StageBytes = k2ABytes + k2BBytes;    // 32 KB
constexpr int k2Stages = 7;          // plain GEMM
constexpr int k2StagesRS = 6;        // fused reduce-scatter epilogue: one stage is traded for the store staging buffers
constexpr int k2StgWarpBytes = 32 * 128;                   // staging tile per epilogue warp: 32 rows x 64 bf16, 128B-swizzled
constexpr int k2SmemBytesFor(int stages, bool rs) {
  return rs ? stages * k2StageBytes + 1024 + 1024 + 4 * k2StgWarpBytes : stages * k2StageBytes + 1024 + 256;
}
struct RsMaps {
  CUtensorMap m[8];          // [owner]: my slot in that rank's staging buffer as a [rs_rows, N] bf16 tensor, box 64 x 32
};
struct NoMaps {
  int unused;
};
template <bool RS> struct RsArg { using type = NoMaps; };
template <> struct RsArg<true> { using type = RsMaps; };
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // clears the CTA-rank bit of a shared::cluster address (-> CTA 0 of the pair)

struct Gemm2Params {
  int G, M, N, K;
  int m_tiles, n_tiles;      // 256 x 256 tiles per group
  long long ldc, c_batch_stride;
  float alpha;
  int accumulate;
  // fused all-gather prologue (see gemm_tcgen05.cu): flags[row / ag_rows_per_flag] >= ag_epoch <=> those rows of A landed
  const uint32_t* ag_flags;
  int ag_rows_per_flag, ag_first_m_blk;
  uint32_t ag_epoch;
  // fused reduce-scatter epilogue (RS kernels only): row block [o * rs_rows, (o+1) * rs_rows) of the product belongs to
  // rank o; my partial tile goes into slot `rs_rank` of o's [world][rs_rows, N] bf16 staging buffer (peer-mapped memory)
  // through 128B-swizzled shared memory and TMA tensor stores.  The last CTA to finish adds 1 to rs_counter[r] on every rank.
  int rs_world, rs_rank, rs_rows;
  uint32_t* rs_counter[8];
  uint32_t* rs_ticket;       // local, zero-initialised CTA ticket (last CTA signals)
  // in-kernel all-gather push (AG kernels): warps 6-7 of every CTA copy row chunks of this rank's shard of A (which already
  // sits at its place in the local gathered buffer) into every peer's gathered buffer -- multimem.st through the NVLS
  // multicast alias when there is one, plain peer stores otherwise -- and raise flag[rank * chunks + c] on every rank.
  const uint4* ag_src;           // my shard inside my gathered buffer
  uint4* ag_dst[8];              // every rank's mapping of those same rows in ITS gathered buffer
  uint4* ag_mc_dst;              // multicast alias (nullptr: P2P stores)
  uint32_t* ag_peer_flags[8];    // every rank's flags + rank * chunks
  long long ag_chunk_vec;        // 16-byte vectors per chunk
  int ag_chunks, ag_world, ag_rank, ag_shard_rows;
};

SMX_DEVICE void ag_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool A_MN, bool B_MN, bool OUT_F32, bool RS = false, bool AG = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads + 64, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 void* __restrict__ c_ptr, const Gemm2Params p,
                 const __grid_constant__ typename RsArg<RS>::type rs_maps) {
  constexpr int BLOCK_K = 64, UMMA_K = 16, kKSteps = BLOCK_K / UMMA_K;
  constexpr uint32_t kMnChunkBytes = BLOCK_K * k2RowBytes;      // 8 KB: one 64-wide MN chunk of an MN-major tile
  constexpr int k2Stages = RS ? k2StagesRS : smx::k2Stages;     // shadows the namespace constant inside the kernel
  static_assert(!(RS && OUT_F32), "the reduce-scatter epilogue stores bf16 partial tiles");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + k2Stages * k2StageBytes);
  uint64_t* empty_bar = full_bar + k2Stages;
  uint64_t* tmem_full_bar = empty_bar + k2Stages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < k2Stages; ++i) {
      mbar_init(&full_bar[i], 2);      // leader's own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[i], 1);     // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);  // 4 epilogue warps in each of the 2 CTAs (only the leader's copy is waited on)
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<2>(tmem_base_ptr, 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  const int tiles_per_group = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_per_group * p.G;
  const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  // grouped rasterisation: 4 (256-row) M-tiles down before the next N-tile
  auto tile_coords = [&](int t, int& g, int& m_blk, int& n_blk) {
    g = t / tiles_per_group;
    int r = t - g * tiles_per_group;
    constexpr int kGroupM = 4;
    int width = kGroupM * p.n_tiles;
    int group_id = r / width;
    int first_m = group_id * kGroupM;
    int gsz = min(p.m_tiles - first_m, kGroupM);
    int in_group = r - group_id * width;
    m_blk = first_m + in_group % gsz;
    n_blk = in_group / gsz;
    if (p.ag_flags || RS) {       // rotated start: own rows first (all-gather) / the next rank's rows first (reduce-scatter)
      m_blk += p.ag_first_m_blk;
      if (m_blk >= p.m_tiles) m_blk -= p.m_tiles;
    }
  };

  // Work list: full 256 x 256 tiles for the rounds that occupy every pair, then the tiles of the last partial round cut
  // into two 128-column slices when that lets the tail fit in (about) half a round (same idea as gemm_tcgen05.cu).
  const int full_tiles = (num_tiles / num_pairs) * num_pairs;
  const int tail_tiles = num_tiles - full_tiles;
  const int split_log2 = (tail_tiles > 0 && tail_tiles * 2 <= num_pairs) ? 1 : 0;
  const int num_work = full_tiles + (tail_tiles << split_log2);
  auto work_coords = [&](int w, int& g, int& m_blk, int& n_base, int& n_eff) {
    int t = w, part = 0;
    n_eff = k2BlockN;
    if (w >= full_tiles) {
      const int j = w - full_tiles;
      t = full_tiles + (j >> split_log2);
      part = j & ((1 << split_log2) - 1);
      n_eff = k2BlockN >> split_log2;
    }
    int n_blk;
    tile_coords(t, g, m_blk, n_blk);
    n_base = n_blk * k2BlockN + part * n_eff;
  };

  if (warp_idx == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair; w < num_work; w += num_pairs) {
        int g, m_blk, n_base, n_eff;
        work_coords(w, g, m_blk, n_base, n_eff);
        const int half = n_eff >> 1;                               // B rows staged by each CTA (128 or 64)
        const int m0 = m_blk * k2BlockM + (int)cta_rank * 128;     // my rows of A / of the output
        const int n0 = n_base + (int)cta_rank * half;              // my half of B
        if (p.ag_flags && !(AG && m0 / p.ag_shard_rows == p.ag_rank)) {
          // gathered A: wait until my 128 rows of this tile have been delivered by their owner's gather warps / kernel
          const int last_row = min(m0 + 128, p.M) - 1;
          const uint64_t t0 = global_timer_ns();
          for (int f = m0 / p.ag_rows_per_flag; f <= last_row / p.ag_rows_per_flag; ++f) {
            uint32_t v;
            do {
              asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ag_flags + f) : "memory");
              if ((int32_t)(v - p.ag_epoch) < 0 && global_timer_ns() - t0 > 20000000000ull) {
                printf("simumax_b200: all-gather GEMM (2-CTA) timed out on flag %d (have %u want %u)\n", f, v, p.ag_epoch);
                __trap();
              }
            } while ((int32_t)(v - p.ag_epoch) < 0);
          }
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 41);
          uint8_t* sa = smem + stage * k2StageBytes;
          uint8_t* sb = sa + k2ABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (k2ABytes + half * k2RowBytes));
          else mbar_arrive_cluster(&full_bar[stage], 0);
          const uint32_t bar = smem_u32(&full_bar[stage]) & kPeerBitMask;     // bytes are counted on the leader's barrier
          const int k0 = kb * BLOCK_K;
          if constexpr (!A_MN) {
            tma_load_3d_2sm(&tmap_a, bar, sa, k0, m0, g);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_3d_2sm(&tmap_a, bar, sa + c * kMnChunkBytes, m0 + c * 64, k0, g);
          }
          if constexpr (!B_MN) {
#pragma unroll
            for (int c = 0; c < 2; ++c)      // 64-row boxes
              if (c * 64 < half) tma_load_3d_2sm(&tmap_b, bar, sb + c * (64 * k2RowBytes), k0, n0 + c * 64, g);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if (c * 64 < half) tma_load_3d_2sm(&tmap_b, bar, sb + c * kMnChunkBytes, n0 + c * 64, k0, g);
          }
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (leader) {
      constexpr uint32_t a_kstep = A_MN ? (UMMA_K * k2RowBytes) >> 4 : (UMMA_K * 2) >> 4;
      constexpr uint32_t b_kstep = B_MN ? (UMMA_K * k2RowBytes) >> 4 : (UMMA_K * 2) >> 4;
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem), A_MN ? kMnChunkBytes : 16, 1024);
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem + k2ABytes), B_MN ? kMnChunkBytes : 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int w = pair; w < num_work; w += num_pairs) {
        const int n_eff = (w >= full_tiles) ? (k2BlockN >> split_log2) : k2BlockN;
        const uint32_t idesc = make_idesc(1u, 1u, A_MN ? 1u : 0u, B_MN ? 1u : 0u, k2BlockM, n_eff);
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1, 42);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * k2BlockN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase, 43);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = a_desc0 + (uint64_t)stage * (k2StageBytes >> 4);
            const uint64_t b_desc = b_desc0 + (uint64_t)stage * (k2StageBytes >> 4);
#pragma unroll
            for (int k = 0; k < kKSteps; ++k)
              umma_f16<2>(tmem_d, a_desc + k * a_kstep, b_desc + k * b_kstep, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit<2>(&empty_bar[stage]);                       // frees the stage in both CTAs
            if (kb == num_k_blocks - 1) umma_commit<2>(&tmem_full_bar[as]);   // accumulator ready in both CTAs
          }
          __syncwarp();
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (AG && warp_idx >= 6) {
    // ============================ all-gather push (both CTAs, 64 threads) ============================
    const int gt = threadIdx.x - k2Threads;
    for (int c = blockIdx.x; c < p.ag_chunks; c += gridDim.x) {
      const uint4* src = p.ag_src + (long long)c * p.ag_chunk_vec;
      for (long long i0 = 0; i0 < p.ag_chunk_vec; i0 += 64 * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * 64 + gt;
          if (i < p.ag_chunk_vec)
            asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i) : "memory");
        }
        if (p.ag_mc_dst != nullptr) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * 64 + gt;
            if (i < p.ag_chunk_vec)
              asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.ag_mc_dst + (long long)c * p.ag_chunk_vec + i),
                           "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
          }
        } else {
#pragma unroll 1
          for (int s = 1; s < p.ag_world; ++s) {
            int peer = p.ag_rank - s;             // the rank that needs my rows first is my predecessor
            if (peer < 0) peer += p.ag_world;
            uint4* dst = p.ag_dst[peer] + (long long)c * p.ag_chunk_vec;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const long long i = i0 + u * 64 + gt;
              if (i < p.ag_chunk_vec)
                asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
            }
          }
        }
      }
      __threadfence_system();
      asm volatile("bar.sync 2, 64;" ::: "memory");       // the two gather warps
      if (gt < p.ag_world) ag_st_release_sys(p.ag_peer_flags[gt] + c, p.ag_epoch);
    }
  } else if (warp_idx < 6) {
    // ============================ epilogue (both CTAs, own 128 rows) ============================
    const int q = warp_idx % 4;
    using OutT = typename std::conditional<OUT_F32, float, __nv_bfloat16>::type;
    OutT* c_base = reinterpret_cast<OutT*>(c_ptr);
    const bool vec_ok = OUT_F32 ? (p.ldc % 4 == 0) : (p.ldc % 8 == 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int w = pair; w < num_work; w += num_pairs) {
      int g, m_blk, n0, n_eff;
      work_coords(w, g, m_blk, n0, n_eff);
      const int row = m_blk * k2BlockM + (int)cta_rank * 128 + q * 32 + lane;
      mbar_wait(&tmem_full_bar[as], aphase, 44);
      tcgen05_fence_after();
      OutT* c_row = c_base + (long long)g * p.c_batch_stride + (long long)row * p.ldc;
      if constexpr (RS) {
        // ---- fused reduce-scatter: TMEM -> bf16 -> 128B-swizzled smem tile (32 rows x 64 columns per warp) -> one TMA
        // tensor store per chunk into slot `rs_rank` of the owner's staging buffer (the tensor map clips the M / N edges).
        const uint32_t stg_warp = smem_u32(smem) + k2Stages * k2StageBytes + 1024 + q * k2StgWarpBytes;
        const uint32_t stg_row = stg_warp + lane * 128;
        const int row0 = m_blk * k2BlockM + (int)cta_rank * 128 + q * 32;      // first row of this warp (one owner per warp)
        const int owner = min(row0 / p.rs_rows, p.rs_world - 1);
        const int slot_row0 = row0 - owner * p.rs_rows;
#pragma unroll 1
        for (int c = 0; c < n_eff / 64; ++c) {
          uint32_t r0[32], r1[32];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * k2BlockN + c * 64;
          tmem_ld_32x32b_x32(taddr, r0);
          tmem_ld_32x32b_x32(taddr + 32, r1);
          tmem_ld_wait();
          if (lane == 0) tma_store_wait_read<0>();      // the previous store has finished reading the staging tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v4(stg_row + ((j ^ (lane & 7)) << 4),
                         pack_bf16x2(__uint_as_float(r0[8 * j]) * p.alpha, __uint_as_float(r0[8 * j + 1]) * p.alpha),
                         pack_bf16x2(__uint_as_float(r0[8 * j + 2]) * p.alpha, __uint_as_float(r0[8 * j + 3]) * p.alpha),
                         pack_bf16x2(__uint_as_float(r0[8 * j + 4]) * p.alpha, __uint_as_float(r0[8 * j + 5]) * p.alpha),
                         pack_bf16x2(__uint_as_float(r0[8 * j + 6]) * p.alpha, __uint_as_float(r0[8 * j + 7]) * p.alpha));
            st_shared_v4(stg_row + (((4 + j) ^ (lane & 7)) << 4),
                         pack_bf16x2(__uint_as_float(r1[8 * j]) * p.alpha, __uint_as_float(r1[8 * j + 1]) * p.alpha),
                         pack_bf16x2(__uint_as_float(r1[8 * j + 2]) * p.alpha, __uint_as_float(r1[8 * j + 3]) * p.alpha),
                         pack_bf16x2(__uint_as_float(r1[8 * j + 4]) * p.alpha, __uint_as_float(r1[8 * j + 5]) * p.alpha),
                         pack_bf16x2(__uint_as_float(r1[8 * j + 6]) * p.alpha, __uint_as_float(r1[8 * j + 7]) * p.alpha));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            tma_store_3d_addr(&rs_maps.m[owner], stg_warp, n0 + c * 64, slot_row0, 0);
            tma_store_commit();
          }
        }
      } else
#pragma unroll 1
      for (int c = 0; c < n_eff / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * k2BlockN + c * 32;
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < p.M && col0 < p.N) {
          if (vec_ok && col0 + 32 <= p.N) {
            if constexpr (OUT_F32) {
              float4* dst = reinterpret_cast<float4*>(c_row + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 v = make_float4(__uint_as_float(r[4 * j]) * p.alpha, __uint_as_float(r[4 * j + 1]) * p.alpha,
                                       __uint_as_float(r[4 * j + 2]) * p.alpha, __uint_as_float(r[4 * j + 3]) * p.alpha);
                if (p.accumulate == 2) {
                  // fp32 main-grad accumulation as a fire-and-forget vector reduction: the read-modify-write happens at the L2
                  // slice, the epilogue neither waits for the old value nor moves it through the SM
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                  continue;
                }
                if (p.accumulate) {
                  float4 o = dst[j];
                  v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                dst[j] = v;
              }
            } else {
              uint4* dst = reinterpret_cast<uint4*>(c_row + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[8 * j + e]) * p.alpha;
                if (p.accumulate) {
                  uint4 o = dst[j];
                  const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float2 of = __bfloat1622float2(ob[e]);
                    f[2 * e] += of.x;
                    f[2 * e + 1] += of.y;
                  }
                }
                uint4 v;
                v.x = pack_bf16x2(f[0], f[1]);
                v.y = pack_bf16x2(f[2], f[3]);
                v.z = pack_bf16x2(f[4], f[5]);
                v.w = pack_bf16x2(f[6], f[7]);
                dst[j] = v;
              }
            }
          } else {
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < p.N) {
                float v = __uint_as_float(r[j]) * p.alpha;
                if constexpr (OUT_F32) {
                  if (p.accumulate) v += c_row[col0 + j];
                  c_row[col0 + j] = v;
                } else {
                  if (p.accumulate) v += __bfloat162float(c_row[col0 + j]);
                  c_row[col0 + j] = __float2bfloat16(v);
                }
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty_bar[as]);
        else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if constexpr (RS) {
      // every store of this CTA has landed at its owner -> one release-increment on every rank's completion counter
      if (lane == 0) {
        tma_store_wait_all<0>();
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps
      if (warp_idx == 2) {
        // last CTA of the grid (ticket) announces this rank's completion: +1 on every rank's counter, release at system scope
        uint32_t last = 0;
        if (lane == 0) {
          __threadfence();
          uint32_t t;
          asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(p.rs_ticket) : "memory");
          last = (t == gridDim.x - 1) ? 1u : 0u;
          if (last) *p.rs_ticket = 0;                       // self-resetting: the next launch starts from zero
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && lane < p.rs_world) {
          __threadfence_system();
          asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p.rs_counter[lane]), "r"(1u) : "memory");
        }
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();          // nobody may still be reading the peer's smem / signalling its barriers
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

inline void clear_ag_push(Gemm2Params& p) {
  p.ag_src = nullptr; p.ag_mc_dst = nullptr; p.ag_chunk_vec = 0; p.ag_chunks = 0; p.ag_world = 1; p.ag_rank = 0; p.ag_shard_rows = 1 << 30;
  for (int i = 0; i < 8; ++i) { p.ag_dst[i] = nullptr; p.ag_peer_flags[i] = nullptr; }
}

}  // namespace smx

namespace {

template <bool A_MN, bool B_MN, bool OUT_F32, bool RS = false, bool AG = false>
int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, void* C, const smx::Gemm2Params& p, int num_sms,
                cudaStream_t stream, int* grid_out = nullptr,
                const typename smx::RsArg<RS>::type& maps = typename smx::RsArg<RS>::type()) {
  auto kern = smx::gemm_2cta_kernel<A_MN, B_MN, OUT_F32, RS, AG>;
  constexpr int kThreads = smx::k2Threads + (AG ? 64 : 0);
  constexpr int kSmem = smx::k2SmemBytesFor(RS ? smx::k2StagesRS : smx::k2Stages, RS);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // A persistent grid must be fully co-resident: ask the driver how many CTA pairs fit at once (GPCs with an odd
  // number of usable TPC slots make this smaller than num_sms / 2; oversubscribing would run a second serial wave).
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms & ~1, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    max_clusters = (e == cudaSuccess && n > 0) ? n : num_sms / 2;
    if (e != cudaSuccess) (void)cudaGetLastError();
    if (getenv("SMX_GEMM_VERBOSE")) fprintf(stderr, "simumax_b200: gemm_2cta max active clusters = %d (rc %d)\n", n, (int)e);
  }
  const int tiles = p.G * p.m_tiles * p.n_tiles;
  int max_pairs = num_sms / 2 < max_clusters ? num_sms / 2 : max_clusters;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  if (grid_out) *grid_out = 2 * pairs;
  kern<<<2 * pairs, kThreads, kSmem, stream>>>(ta, tb, C, p, maps);
  return (int)cudaGetLastError();
}

}  // namespace

namespace {
const uint32_t* g2_ag_flags = nullptr;
int g2_ag_rows = 1, g2_ag_first = 0;
uint32_t g2_ag_epoch = 0;
}  // namespace

extern "C" {

// Arms the fused all-gather prologue for the NEXT smx_gemm_2cta call; first_m_blk counts 256-row pair tiles.
int smx_gemm_2cta_arm_all_gather(const void* flags, int rows_per_flag, int first_m_blk, unsigned epoch) {
  if (rows_per_flag <= 0) return -15;
  g2_ag_flags = reinterpret_cast<const uint32_t*>(flags);
  g2_ag_rows = rows_per_flag; g2_ag_first = first_m_blk; g2_ag_epoch = epoch;
  return 0;
}

// Same contract as smx_gemm (bf16 inputs only).  Returns -12 when the problem is not worth a 256 x 256 pair tile.
int smx_gemm_2cta(const void* A, const void* B, void* C, int G, int M, int N, int K, int a_mn, int b_mn, int out_f32,
                  int accumulate, float alpha, long long lda, long long ldb, long long ldc, long long a_bs,
                  long long b_bs, long long c_bs, int num_sms, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  // the armed all-gather state is one-shot: capture and clear it before any early return, so a refused call can never leak
  // its flags / epoch into the next, unrelated GEMM
  const uint32_t* armed_flags = g2_ag_flags;
  const int armed_rows = g2_ag_rows, armed_first = g2_ag_first;
  const uint32_t armed_epoch = g2_ag_epoch;
  g2_ag_flags = nullptr;
  if (G <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
  if (M < 256 || N < 256) return -12;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = smx::make_tmap_3d(&ta, dt, A, K, M, G, lda * 2, a_bs * 2, 64, 128);
  else       rc = smx::make_tmap_3d(&ta, dt, A, M, K, G, lda * 2, a_bs * 2, 64, 64);
  if (rc) return rc;
  if (!b_mn) rc = smx::make_tmap_3d(&tb, dt, B, K, N, G, ldb * 2, b_bs * 2, 64, 64);
  else       rc = smx::make_tmap_3d(&tb, dt, B, N, K, G, ldb * 2, b_bs * 2, 64, 64);
  if (rc) return rc;
  smx::Gemm2Params p;
  p.G = G; p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + smx::k2BlockM - 1) / smx::k2BlockM;
  p.n_tiles = (N + smx::k2BlockN - 1) / smx::k2BlockN;
  // fp32 accumulate-into-C epilogue: read-modify-write (p.accumulate = 1) or red.global.add.v4.f32 (p.accumulate = 2).
  // Whole-step A/B on one box (Llama-3-8B, 8 x 4096 tokens, profiles/final_bench_n1_red{0,1}.json): 1608.6 ms with the read-modify-write
  // epilogue, 1522.1 ms with RED (+5.7 %), so RED is the default.  Op by op at its own sustained, power-capped clock the picture is
  // mixed (profiles/calib_ops_v8_red_epilogue.json vs v7: RED +7 ... +28 % for outputs <= 16 MiB, -7 % above 128 MiB): the RED epilogue
  // draws more power, which an isolated sweep charges to the GEMM itself and a real step spreads over the lighter kernels around it.
  // SMX_GEMM_RED=0 forces read-modify-write, SMX_GEMM_RED=auto picks by output size (<= SMX_GEMM_RED_MAX_MIB, default 32).
  static const int red_mode = [] { const char* e = getenv("SMX_GEMM_RED"); return !e ? 1 : (e[0] == '0' ? 0 : (e[0] == 'a' ? 2 : 1)); }();
  static const long long red_max_bytes = [] { const char* e = getenv("SMX_GEMM_RED_MAX_MIB"); return (long long)(e ? atoi(e) : 32) << 20; }();
  const bool use_red = red_mode == 1 || (red_mode == 2 && (long long)G * M * N * 4 <= red_max_bytes);
  p.ldc = ldc; p.c_batch_stride = c_bs; p.alpha = alpha; p.accumulate = (accumulate && out_f32 && use_red) ? 2 : accumulate;
  p.ag_flags = armed_flags; p.ag_rows_per_flag = armed_rows; p.ag_first_m_blk = armed_first; p.ag_epoch = armed_epoch;
  p.rs_world = 1; p.rs_rank = 0; p.rs_rows = M;
  for (int i = 0; i < 8; ++i) p.rs_counter[i] = nullptr;
  p.rs_ticket = nullptr;
  smx::clear_ag_push(p);
  if (!a_mn && !b_mn) return out_f32 ? launch_2cta<false, false, true>(ta, tb, C, p, num_sms, stream)
                                     : launch_2cta<false, false, false>(ta, tb, C, p, num_sms, stream);
  if (!a_mn && b_mn)  return out_f32 ? launch_2cta<false, true, true>(ta, tb, C, p, num_sms, stream)
                                     : launch_2cta<false, true, false>(ta, tb, C, p, num_sms, stream);
  if (a_mn && b_mn)   return out_f32 ? launch_2cta<true, true, true>(ta, tb, C, p, num_sms, stream)
                                     : launch_2cta<true, true, false>(ta, tb, C, p, num_sms, stream);
  return out_f32 ? launch_2cta<true, false, true>(ta, tb, C, p, num_sms, stream)
                 : launch_2cta<true, false, false>(ta, tb, C, p, num_sms, stream);
}


// Fused GEMM -> reduce-scatter (tensor-parallel row-linear forward / column-linear dgrad with sequence parallelism):
// C_partial = alpha * op(A) op(B) is never materialised locally; row block o of it is stored into slot `rank` of rank o's
// staging buffer slot_ptrs[o] ([world][M / world, N] bf16, peer-mapped) from the epilogue, and every CTA finally adds 1 to
// counter_ptrs[r] on every rank r is incremented once per rank and call (by the last CTA to finish, `ticket` is a local
// zero-initialised word), so a consumer may read its slots once its own counter has advanced by `world`.
// Layouts TN (b_mn = 0) and NN (b_mn = 1); returns -12 when the problem is too small for pair tiles, -13 on bad shapes.
int smx_gemm_2cta_reduce_scatter(const void* A, const void* B, void* const* slot_ptrs, void* const* counter_ptrs, void* ticket,
                                 int world, int rank, int M, int N, int K, int b_mn, float alpha, long long lda,
                                 long long ldb, int num_sms, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (M < 256 || N < 256) return -12;
  if (world < 1 || world > 8 || M % world || N % 8 || (M / world) % 128) return -13;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta, tb;
  int rc = smx::make_tmap_3d(&ta, dt, A, K, M, 1, lda * 2, 0, 64, 128);
  if (rc) return rc;
  if (!b_mn) rc = smx::make_tmap_3d(&tb, dt, B, K, N, 1, ldb * 2, 0, 64, 64);
  else       rc = smx::make_tmap_3d(&tb, dt, B, N, K, 1, ldb * 2, 0, 64, 64);
  if (rc) return rc;
  smx::Gemm2Params p;
  p.G = 1; p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + smx::k2BlockM - 1) / smx::k2BlockM;
  p.n_tiles = (N + smx::k2BlockN - 1) / smx::k2BlockN;
  p.ldc = N; p.c_batch_stride = 0; p.alpha = alpha; p.accumulate = 0;
  p.ag_flags = nullptr; p.ag_rows_per_flag = 1; p.ag_epoch = 0;
  smx::clear_ag_push(p);
  p.rs_world = world; p.rs_rank = rank; p.rs_rows = M / world;
  p.rs_ticket = reinterpret_cast<uint32_t*>(ticket);
  // start with the rows of the next rank and finish with my own, so every owner's NVLink ingress sees one source at a time
  p.ag_first_m_blk = (int)(((long long)((rank + 1) % world) * p.rs_rows) / smx::k2BlockM);
  smx::RsMaps maps;
  for (int i = 0; i < 8; ++i) {
    p.rs_counter[i] = i < world ? reinterpret_cast<uint32_t*>(counter_ptrs[i]) : nullptr;
    const int o = i < world ? i : 0;       // unused entries replicate owner 0 (never dereferenced)
    char* slot = reinterpret_cast<char*>(slot_ptrs[o]) + (long long)rank * p.rs_rows * N * 2;
    rc = smx::make_tmap_3d(&maps.m[i], dt, slot, N, p.rs_rows, 1, (uint64_t)N * 2, (uint64_t)p.rs_rows * N * 2, 64, 32);
    if (rc) return rc;
  }
  return b_mn ? launch_2cta<false, true, false, true>(ta, tb, nullptr, p, num_sms, stream, nullptr, maps)
              : launch_2cta<false, false, false, true>(ta, tb, nullptr, p, num_sms, stream, nullptr, maps);
}

// Fused all-gather -> GEMM in ONE kernel (column-parallel linear forward / row-parallel dgrad with sequence parallelism):
// C[M, N] = alpha * A[M, K] op(B), A = concatenation over ranks of [M / world, K] shards.  a_ptrs[r] is rank r's mapping of
// the gathered buffer (this rank's own rows are already in place in a_ptrs[rank]); mc_ptr its NVLS multicast alias or null;
// flag_ptrs[r] rank r's flag array (world * chunks words, epochs only ever grow).  Warps 6-7 of every CTA push this rank's
// rows to all peers chunk by chunk while the MMAs start on the local rows; the TMA producer waits on the flags of a remote
// row block only when it reaches it.  Layouts TN (b_mn = 0) and NN (b_mn = 1); bf16 or fp32 output.
int smx_gemm_2cta_all_gather(void* const* a_ptrs, void* mc_ptr, void* const* flag_ptrs, const void* B, void* C, int world,
                             int rank, int M, int N, int K, int b_mn, int out_f32, float alpha, long long ldb, long long ldc,
                             int chunks, unsigned epoch, int num_sms, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (M < 256 || N < 256) return -12;
  if (world < 2 || world > 8 || M % world || (M / world) % 128 || chunks <= 0 || (M / world) % chunks) return -13;
  const int shard_rows = M / world;
  const long long chunk_bytes = (long long)(shard_rows / chunks) * K * 2;
  if (chunk_bytes % 16) return -13;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta, tb;
  int rc = smx::make_tmap_3d(&ta, dt, a_ptrs[rank], K, M, 1, (uint64_t)K * 2, 0, 64, 128);
  if (rc) return rc;
  if (!b_mn) rc = smx::make_tmap_3d(&tb, dt, B, K, N, 1, ldb * 2, 0, 64, 64);
  else       rc = smx::make_tmap_3d(&tb, dt, B, N, K, 1, ldb * 2, 0, 64, 64);
  if (rc) return rc;
  smx::Gemm2Params p;
  p.G = 1; p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + smx::k2BlockM - 1) / smx::k2BlockM;
  p.n_tiles = (N + smx::k2BlockN - 1) / smx::k2BlockN;
  p.ldc = ldc; p.c_batch_stride = 0; p.alpha = alpha; p.accumulate = 0;
  p.rs_world = 1; p.rs_rank = 0; p.rs_rows = M; p.rs_ticket = nullptr;
  for (int i = 0; i < 8; ++i) p.rs_counter[i] = nullptr;
  p.ag_flags = reinterpret_cast<const uint32_t*>(flag_ptrs[rank]);
  p.ag_rows_per_flag = shard_rows / chunks;
  p.ag_first_m_blk = (int)(((long long)rank * shard_rows) / smx::k2BlockM);     // my own rows first
  p.ag_epoch = epoch;
  smx::clear_ag_push(p);
  const long long shard_bytes = (long long)shard_rows * K * 2;
  p.ag_src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a_ptrs[rank]) + rank * shard_bytes);
  p.ag_mc_dst = mc_ptr ? reinterpret_cast<uint4*>(reinterpret_cast<char*>(mc_ptr) + rank * shard_bytes) : nullptr;
  p.ag_chunk_vec = chunk_bytes / 16; p.ag_chunks = chunks; p.ag_world = world; p.ag_rank = rank; p.ag_shard_rows = shard_rows;
  for (int i = 0; i < world; ++i) {
    p.ag_dst[i] = reinterpret_cast<uint4*>(reinterpret_cast<char*>(a_ptrs[i]) + rank * shard_bytes);
    p.ag_peer_flags[i] = reinterpret_cast<uint32_t*>(flag_ptrs[i]) + (long long)rank * chunks;
  }
  if (!b_mn) return out_f32 ? launch_2cta<false, false, true, false, true>(ta, tb, C, p, num_sms, stream)
                            : launch_2cta<false, false, false, false, true>(ta, tb, C, p, num_sms, stream);
  return out_f32 ? launch_2cta<false, true, true, false, true>(ta, tb, C, p, num_sms, stream)
                 : launch_2cta<false, true, false, false, true>(ta, tb, C, p, num_sms, stream);
}

}  // extern "C"
