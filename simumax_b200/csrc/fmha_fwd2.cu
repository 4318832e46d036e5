// Flash-attention forward, second generation (head dim 128): two 128-query tiles per CTA in ping-pong, 128-key blocks, probabilities
// kept in TMEM.
//
// Why: per 128 x 128 score tile the softmax needs 16384 exp2 = 1024 clk of the SM's 16 MUFU lanes, the two MMAs (Q K^T and P V at
// d = 128) need 1024 clk of the tensor pipe.  The first-generation kernel (fmha_fwd.cu: one query tile per CTA, P through shared
// memory, rescale on the softmax threads) serialises most of that -- ncu showed the tensor pipe 31 % active, the XU pipe 40 %.  Here
//   * a CTA owns 256 queries as two independent tiles; while the softmax warpgroup of tile 0 works on S0_j the tensor core runs
//     P1_{j-1} V and Q1 K_j^T of tile 1, and vice versa -- the MUFU and tensor pipes are both kept busy;
//   * P never touches shared memory: the softmax threads overwrite the first 64 columns of their S tile in TMEM with packed bf16
//     (tcgen05.st) and P V takes its A operand from TMEM (tcgen05.mma with a TMEM A descriptor) -- no st.shared, no proxy fence,
//     and half the shared-memory operand traffic of the P V product;
//   * K / V blocks are 128 keys and shared by both query tiles: half the L2 -> SM traffic per FLOP of the old kernel, and, with
//     GQA, a 2-CTA cluster over two query heads of one KV group loads every K / V block once and multicasts it to both CTAs
//     (cp.async.bulk.tensor ... .multicast::cluster), halving it again.
//
//   TMEM (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,384) | O1 [384,512); P_t aliases S_t columns [0,64).
//   warps: 0 TMA producer, 1 MMA issuer, then the softmax warps: 8 per query tile (two warps share 32 rows and take 64 key columns
//   each) with SMX_FMHA_SPLIT=2, or 4 per tile with one row per thread (default: measured 6 % faster).
//   MMA order per key block j:  P0_j V_j -> Q0 K_{j+1}^T -> P1_j V_j -> Q1 K_{j+1}^T   (tcgen05.mma executes in issue order, so S_t
//   may be overwritten by the next Q K^T right behind the P V that read P_t out of it).
//
// Tried and removed (round 2): double-buffering S per tile so that Q K_{j+1}^T no longer waits for softmax_j.  TMEM has no room for
// four 128-column S buffers next to two O accumulators, so the variant produced scores in 64-key half blocks; N = 64 Q K^T MMAs
// are shared-memory-bandwidth bound and the MMA issuer's period per 128 keys went from 3572 to 4594 clk (638 vs 816 TF at S 4096).
// The in-kernel %clock trace (tools/fmha_fwd_trace.py) of this kernel: softmax 1886 + hand-offs 550 + MMAs 1040 per tile and block.
#include "sm100_common.cuh"

namespace smx {

SMX_DEVICE void tma_load_4d_f2(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// multicast flavour: the box lands at the same smem offset in every CTA of `mask`, each CTA's barrier (same offset) gets the bytes
SMX_DEVICE void tma_load_4d_mc(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}

constexpr int kF2BlockQ = 128;      // rows per query tile
constexpr int kF2Tiles = 2;         // query tiles per CTA
constexpr int kF2BlockKV = 128;
constexpr int kF2D = 128;
constexpr int kF2FirstSoftmaxWarp = 2;       // warp 0 = TMA producer, warp 1 = MMA issuer
template <int SPLIT> constexpr int f2_threads() { return (kF2FirstSoftmaxWarp + 8 * SPLIT) * 32; }
constexpr int kF2Stages = 2;
constexpr int kF2TileBytes = 128 * kF2D * 2;      // 32 KB: one Q tile, one K block, one V block
constexpr int kF2SmemBytes = (kF2Tiles + 2 * kF2Stages) * kF2TileBytes + 256 + 2 * 2 * 128 * 4;   // + barriers + max / sum exchange
SMX_DEVICE void named_bar_sync_f2(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

struct Fmha2Params {
  int B, S_q, S_kv, H, H_kv;
  float scale;
  int causal;
  long long o_stride_b, o_stride_s, o_stride_h;
  long long lse_stride_b, lse_stride_h;
  unsigned int* trace;      // optional: [3 roles][64 blocks][16 slots] SM-clock stamps of CTA (0, 0, 0) (tools/fmha_fwd_trace.py)
};
SMX_DEVICE unsigned int clock32() {
  unsigned int c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  return c;
}
#define F2_TRACE(role, j, slot)                                                                        \
  do {                                                                                                 \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64 && lane == 0)     \
      p.trace[((role) * 64 + (j)) * 16 + (slot)] = clock32();                                          \
  } while (0)

// DIAG (timing experiments only, SMX_FMHA_DIAG): 1 = no MUFU (exp2 replaced by an FMA: wrong numbers, shows the non-MUFU floor),
// 2 = bf16 pack by truncation with PRMT instead of F2FP, 3 = both
template <int CLUSTER, int SPLIT, int DIAG = 0>
__global__ void __launch_bounds__(f2_threads<SPLIT>(), 1)
fmha_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ o_ptr, float* __restrict__ lse_ptr,
                 const Fmha2Params p) {
  constexpr int BQ = kF2BlockQ, BKV = kF2BlockKV, D = kF2D;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) {
    printf("simumax_b200: fmha_fwd2 dynamic smem base %u is not 1024-byte aligned\n", smem_u32(smem_raw));
    __trap();
  }
  uint8_t* sQ = smem;                                   // [2 tiles][2 d-chunks][128 rows x 128 B]
  uint8_t* sK = sQ + kF2Tiles * kF2TileBytes;           // [stages][2 d-chunks][128 keys x 128 B]
  uint8_t* sV = sK + kF2Stages * kF2TileBytes;          // [stages][2 dv-chunks][128 keys x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kF2Stages * kF2TileBytes);
  uint64_t* q_full = bars;                    // 1
  uint64_t* k_full = bars + 1;                // [stages]
  uint64_t* k_empty = k_full + kF2Stages;
  uint64_t* v_full = k_empty + kF2Stages;
  uint64_t* v_empty = v_full + kF2Stages;
  uint64_t* s_full = v_empty + kF2Stages;     // [tiles]
  uint64_t* p_full = s_full + kF2Tiles;       // [tiles]
  uint64_t* o_done = p_full + kF2Tiles;       // [tiles]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_done + kF2Tiles);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const uint32_t cta_rank = CLUSTER > 1 ? cluster_ctarank() : 0;
  // heavy (late) query tiles first: with causal masking tile i has i + 1 units of work
  const int q_tile = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int head = blockIdx.y;                 // CLUSTER == 2: cluster dims (1, 2, 1) -> heads 2c, 2c + 1 share their KV head
  const int batch = blockIdx.z;
  const int kv_head = head / (p.H / p.H_kv);
  const int q0 = q_tile * (BQ * kF2Tiles);
  const int causal_shift = p.S_kv - p.S_q;
  int n_t[kF2Tiles];
#pragma unroll
  for (int t = 0; t < kF2Tiles; ++t) {
    const int qt0 = q0 + t * BQ;
    int kv_end = p.S_kv;
    if (p.causal) kv_end = min(p.S_kv, qt0 + BQ + causal_shift);
    n_t[t] = (qt0 < p.S_q && kv_end > 0) ? (kv_end + BKV - 1) / BKV : 0;
  }
  const int n_max = max(n_t[0], n_t[1]);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kF2Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], CLUSTER);        // every CTA of the cluster must have consumed the block before it is overwritten
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], CLUSTER);
    }
    for (int t = 0; t < kF2Tiles; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4 * SPLIT);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<1>(tmem_base_ptr, 512);
  tcgen05_fence_before();
  if constexpr (CLUSTER > 1) cluster_sync_all();
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kF2Tiles * kF2TileBytes);
#pragma unroll
      for (int t = 0; t < kF2Tiles; ++t)
#pragma unroll
        for (int c = 0; c < D / 64; ++c)
          tma_load_4d_f2(&tmap_q, q_full, sQ + t * kF2TileBytes + c * (BQ * 128), c * 64, q0 + t * BQ, head, batch);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_max; ++j) {
        const int k0 = j * BKV;
        // K_j
        mbar_wait(&k_empty[stage], phase ^ 1, 1);
        mbar_arrive_expect_tx(&k_full[stage], kF2TileBytes);
        if constexpr (CLUSTER > 1) {
          // each CTA fetches one 64-wide d-chunk of the block and multicasts it to both
          const int c = cta_rank;
          tma_load_4d_mc(&tmap_k, &k_full[stage], sK + stage * kF2TileBytes + c * (BKV * 128), c * 64, k0, kv_head, batch, 3);
        } else {
#pragma unroll
          for (int c = 0; c < D / 64; ++c)
            tma_load_4d_f2(&tmap_k, &k_full[stage], sK + stage * kF2TileBytes + c * (BKV * 128), c * 64, k0, kv_head, batch);
        }
        // V_j
        mbar_wait(&v_empty[stage], phase ^ 1, 2);
        mbar_arrive_expect_tx(&v_full[stage], kF2TileBytes);
        if constexpr (CLUSTER > 1) {
          const int c = cta_rank;
          tma_load_4d_mc(&tmap_v, &v_full[stage], sV + stage * kF2TileBytes + c * (BKV * 128), c * 64, k0, kv_head, batch, 3);
        } else {
#pragma unroll
          for (int c = 0; c < D / 64; ++c)
            tma_load_4d_f2(&tmap_v, &v_full[stage], sV + stage * kF2TileBytes + c * (BKV * 128), c * 64, k0, kv_head, batch);
        }
        if (++stage == kF2Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_s = make_idesc(1, 1, 0, 0, BQ, BKV);   // Q (K-major smem) x K (K-major smem) -> S[128, 128]
    constexpr uint32_t idesc_o = make_idesc(1, 1, 0, 1, BQ, D);     // P (TMEM) x V (MN-major smem) -> O[128, 128]
    const uint64_t kQ_km = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t kK_km = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t kV_mn = make_smem_desc_sw128(smem_u32(sV), BKV * 128, 1024);
    constexpr uint64_t kTileStep = kF2TileBytes >> 4;
    auto release = [&](uint64_t* bar) {        // this CTA is done with a K / V block: tell the producer of every CTA of the cluster
      if constexpr (CLUSTER > 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                     "h"(static_cast<uint16_t>(3))
                     : "memory");
      else
        umma_commit<1>(bar);
    };
    auto issue_s = [&](int t, int st) {        // S_t = Q_t K^T
      if (elect_one()) {
        const uint64_t a0 = kQ_km + t * kTileStep;
        const uint64_t b0 = kK_km + st * kTileStep;
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16<1>(tmem_base + t * 128, a0 + ((k / 4) * ((BQ * 128) >> 4) + (k % 4) * 2),
                      b0 + ((k / 4) * ((BKV * 128) >> 4) + (k % 4) * 2), idesc_s, k > 0 ? 1u : 0u);
        umma_commit<1>(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int st, bool first) {   // O_t (+)= P_t V
      if (elect_one()) {
        const uint64_t b0 = kV_mn + st * kTileStep;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)       // 16 keys = 8 packed TMEM columns of P per K step
          umma_f16_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8, b0 + k * ((16 * 128) >> 4), idesc_o,
                      (!first || k > 0) ? 1u : 0u);
        umma_commit<1>(&o_done[t]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 3);
    if (n_max > 0) {
      mbar_wait(&k_full[0], 0, 4);
      tcgen05_fence_after();
      for (int t = 0; t < kF2Tiles; ++t)
        if (n_t[t] > 0) issue_s(t, 0);
      if (elect_one()) release(&k_empty[0]);
      __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_max; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kF2Stages) { nstage = 0; nphase ^= 1; }
      bool v_ready = false, k_ready = false;
      for (int t = 0; t < kF2Tiles; ++t) {
        if (j >= n_t[t]) continue;
        mbar_wait(&p_full[t], j & 1, 5);
        F2_TRACE(0, j, 0 + t * 4);
        if (!v_ready) { mbar_wait(&v_full[stage], phase, 6); v_ready = true; }
        F2_TRACE(0, j, 1 + t * 4);
        tcgen05_fence_after();
        issue_pv(t, stage, j == 0);
        F2_TRACE(0, j, 2 + t * 4);
        if (j + 1 < n_t[t]) {
          if (!k_ready) { mbar_wait(&k_full[nstage], nphase, 7); k_ready = true; }
          tcgen05_fence_after();
          issue_s(t, nstage);
        }
        F2_TRACE(0, j, 3 + t * 4);
      }
      if (elect_one()) {
        release(&v_empty[stage]);                // behind the last P V of this key block
        if (k_ready) release(&k_empty[nstage]);  // behind the last Q K^T of the next one
      }
      __syncwarp();
      stage = nstage;
      phase = nphase;
    }
  } else if (warp_idx >= kF2FirstSoftmaxWarp) {
    // ============================ softmax + epilogue ============================
    // SPLIT == 1: one query row per thread (128 scores in registers).  SPLIT == 2: two warps share 32 rows and take 64 key
    // columns each -- twice the warps per scheduler to cover MUFU / dependency latency and half the live registers; the row
    // maximum is exchanged through shared memory, the row sum stays split until the epilogue.
    constexpr int NC = BKV / SPLIT;                   // score columns per thread
    const int sw = warp_idx - kF2FirstSoftmaxWarp;    // 4 consecutive warps cover the four TMEM lane quadrants (warp_idx % 4)
    const int grp = sw / 4;                           // (tile, column half)
    const int t = grp / SPLIT;
    const int hh = grp % SPLIT;
    const int q = warp_idx % 4;
    const int row_in_tile = q * 32 + lane;
    const int row = q0 + t * BQ + row_in_tile;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t tmem_s = tmem_base + t * 128 + lane_addr;
    const uint32_t tmem_o = tmem_base + 256 + t * 128 + lane_addr;
    const float c_log2 = p.scale * 1.4426950408889634f;
    float* xch = reinterpret_cast<float*>(bars + 32);          // [tile][half][128 rows] floats behind the barrier block
    float* my_x = xch + (t * SPLIT + hh) * BQ + row_in_tile;
    float* peer_x = xch + (t * SPLIT + (hh ^ 1)) * BQ + row_in_tile;
    const int pair_bar = 1 + t * 4 + q;               // named barrier of the two warps that share these rows
    float m_used = -INFINITY;   // running reference max (raw score units)
    float l = 0.f;              // SPLIT == 2: sum over this thread's columns only
    const int n_mine = n_t[t];
    for (int j = 0; j < n_mine; ++j) {
      mbar_wait(&s_full[t], j & 1, 8);
      tcgen05_fence_after();
      if (q == 0 && hh == 0) F2_TRACE(1 + t, j, 0);
      uint32_t r[NC];
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) tmem_ld_32x32b_x32(tmem_s + hh * NC + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      tmem_ld_wait();
      if (q == 0 && hh == 0) F2_TRACE(1 + t, j, 1);
      const int k0 = j * BKV + hh * NC;
      const int qt0 = q0 + t * BQ;
      const bool need_mask = (p.causal && (j * BKV + BKV - 1 > qt0 + causal_shift)) || (j * BKV + BKV > p.S_kv);
      if (need_mask) {
        const int limit = p.causal ? min(row + causal_shift, p.S_kv - 1) : p.S_kv - 1;
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (k0 + i > limit) r[i] = 0xff800000u;   // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < NC; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(r[i + u]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if constexpr (SPLIT == 2) {
        *my_x = mx;
        named_bar_sync_f2(pair_bar, 64);
        mx = fmaxf(mx, *peer_x);
        named_bar_sync_f2(pair_bar, 64);              // the slot is rewritten next block
      }
      // lazy rescale: only move the reference max when it grew by more than 2^8 in the exp2 domain
      float m_new = m_used;
      const bool grow = (mx > m_used) && ((mx - m_used) * c_log2 > 8.f || m_used == -INFINITY);
      if (grow) m_new = mx;
      const bool any_grow = __any_sync(0xffffffffu, grow);       // both warps of a pair see the same rows: same decision
      if (any_grow && j > 0) {
        // O_t may only be touched once P_{j-1} V has retired (the next P V is not issued before this thread signals p_full)
        mbar_wait(&o_done[t], (j - 1) & 1, 9);
        tcgen05_fence_after();
        const float alpha = (m_used == -INFINITY) ? 0.f : exp2f((m_used - m_new) * c_log2);
#pragma unroll 1
        for (int c = 0; c < D / SPLIT / 16; ++c) {    // SPLIT == 2: each warp of the pair rescales its half of the O columns
          uint32_t o[16];
          tmem_ld_32x32b_x16(tmem_o + hh * (D / SPLIT) + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x16(tmem_o + hh * (D / SPLIT) + c * 16, o);
        }
        tmem_st_wait();
        l *= alpha;
      }
      m_used = m_new;
      if (q == 0 && hh == 0) F2_TRACE(1 + t, j, 2);
      const float mc = (m_used == -INFINITY) ? 0.f : m_used * c_log2;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      // P row -> packed bf16 pairs, in place over the registers (keys 2c, 2c + 1 -> 32-bit column c), then NC / 2 TMEM columns.
      // Software-pipelined by hand in chunks of 16: the MUFU results of chunk ch are consumed (row sum, bf16 pack) only while the
      // exponentials of chunk ch + 1 are in flight -- with one or two warps per scheduler nothing else hides the MUFU latency, and
      // ptxas otherwise places every FADD / F2FP two instructions behind the MUFU it depends on.
      float ebuf[32];
#pragma unroll
      for (int ch = 0; ch <= NC / 16; ++ch) {
        if (ch < NC / 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = fmaf(__uint_as_float(r[ch * 16 + i]), c_log2, -mc);
            ebuf[(ch & 1) * 16 + i] = (DIAG & 1) ? fmaf(x, 0.001f, 0.5f) : exp2f(x);
          }
        }
        if (ch > 0) {
          const int pc = ch - 1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float e0 = ebuf[(pc & 1) * 16 + 2 * i], e1 = ebuf[(pc & 1) * 16 + 2 * i + 1];
            sum4[i & 3] += e0 + e1;
            if constexpr (DIAG & 2) r[pc * 8 + i] = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);
            else r[pc * 8 + i] = pack_bf16x2(e0, e1);
          }
        }
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (q == 0 && hh == 0) F2_TRACE(1 + t, j, 3);
#pragma unroll
      for (int c = 0; c < NC / 2 / 16; ++c)
        tmem_st_32x32b_x16(tmem_s + hh * (NC / 2) + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&r[c * 16]));
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (q == 0 && hh == 0) F2_TRACE(1 + t, j, 4);
    }
    // ---- epilogue
    if constexpr (SPLIT == 2) {
      *my_x = l;
      named_bar_sync_f2(pair_bar, 64);
      l += *peer_x;
    }
    if (n_mine > 0) {
      mbar_wait(&o_done[t], (n_mine - 1) & 1, 10);
      tcgen05_fence_after();
    }
    const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
    const bool row_ok = row < p.S_q;
    __nv_bfloat16* orow = o_ptr + batch * p.o_stride_b + (long long)(row_ok ? row : 0) * p.o_stride_s + head * p.o_stride_h +
                          hh * (D / SPLIT);
#pragma unroll 1
    for (int c = 0; c < D / SPLIT / 32; ++c) {
      uint32_t o[32];
      if (n_mine > 0) {
        tmem_ld_32x32b_x32(tmem_o + hh * (D / SPLIT) + c * 32, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          dst[i] = v;
        }
      }
    }
    if (row_ok && lse_ptr && hh == 0)
      lse_ptr[batch * p.lse_stride_b + head * p.lse_stride_h + row] = (l > 0.f) ? m_used * p.scale + __logf(l) : -INFINITY;
  }

  tcgen05_fence_before();
  if constexpr (CLUSTER > 1) cluster_sync_all();      // a peer may still multicast into / commit onto this CTA's shared memory
  else __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

inline int make_tmap_4d_f2(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                           uint64_t s1_bytes, uint64_t s2_bytes, uint64_t s3_bytes, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1_bytes, s2_bytes, s3_bytes};
  cuuint32_t box[4] = {64, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

template <int CLUSTER, int SPLIT, int DIAG = 0>
int launch_fmha_fwd2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, void* o, float* lse,
                     const Fmha2Params& p, cudaStream_t stream) {
  auto kern = fmha_fwd2_kernel<CLUSTER, SPLIT, DIAG>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2SmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((p.S_q + kF2BlockQ * kF2Tiles - 1) / (kF2BlockQ * kF2Tiles), p.H, p.B);
  cfg.blockDim = dim3(f2_threads<SPLIT>());
  cfg.dynamicSmemBytes = kF2SmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CLUSTER;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(o), lse, p);
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

}  // namespace smx

extern "C" {

// Same contract as smx_fmha_fwd (head dim 128 / 128 only).  cluster = 2 pairs adjacent query heads of one KV group (needs an even
// GQA group size); cluster = 0 picks it automatically.
int smx_fmha_fwd2(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S_q, int S_kv, int H, int H_kv,
                  long long q_sb, long long q_ss, long long q_sh, long long k_sb, long long k_ss, long long k_sh, long long v_sb,
                  long long v_ss, long long v_sh, long long o_sb, long long o_ss, long long o_sh, float scale, int causal,
                  int cluster, void* trace, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (H % H_kv) return -30;
  const int group = H / H_kv;
  if (cluster == 0) cluster = (group % 2 == 0) ? 2 : 1;
  if (cluster == 2 && group % 2) return -32;
  CUtensorMap tq, tk, tv;
  int rc = smx::make_tmap_4d_f2(&tq, q, 128, S_q, H, B, q_ss * 2, q_sh * 2, q_sb * 2, smx::kF2BlockQ);
  if (rc) return rc;
  rc = smx::make_tmap_4d_f2(&tk, k, 128, S_kv, H_kv, B, k_ss * 2, k_sh * 2, k_sb * 2, smx::kF2BlockKV);
  if (rc) return rc;
  rc = smx::make_tmap_4d_f2(&tv, v, 128, S_kv, H_kv, B, v_ss * 2, v_sh * 2, v_sb * 2, smx::kF2BlockKV);
  if (rc) return rc;
  smx::Fmha2Params p;
  p.B = B; p.S_q = S_q; p.S_kv = S_kv; p.H = H; p.H_kv = H_kv; p.scale = scale; p.causal = causal;
  p.o_stride_b = o_sb; p.o_stride_s = o_ss; p.o_stride_h = o_sh;
  p.lse_stride_b = (long long)H * S_q; p.lse_stride_h = S_q;
  p.trace = reinterpret_cast<unsigned int*>(trace);
  static const int split = [] { const char* e = getenv("SMX_FMHA_SPLIT"); return (e && e[0] == '2') ? 2 : 1; }();
  static const int diag = [] { const char* e = getenv("SMX_FMHA_DIAG"); return e ? atoi(e) : 0; }();
  if (diag == 1) return smx::launch_fmha_fwd2<1, 1, 1>(tq, tk, tv, o, lse, p, stream);
  if (diag == 2) return smx::launch_fmha_fwd2<1, 1, 2>(tq, tk, tv, o, lse, p, stream);
  if (diag == 3) return smx::launch_fmha_fwd2<1, 1, 3>(tq, tk, tv, o, lse, p, stream);
  if (split == 1)
    return cluster == 2 ? smx::launch_fmha_fwd2<2, 1>(tq, tk, tv, o, lse, p, stream) : smx::launch_fmha_fwd2<1, 1>(tq, tk, tv, o, lse, p, stream);
  return cluster == 2 ? smx::launch_fmha_fwd2<2, 2>(tq, tk, tv, o, lse, p, stream) : smx::launch_fmha_fwd2<1, 2>(tq, tk, tv, o, lse, p, stream);
}

}  // extern "C"
