// Flash-attention backward for sm_100a (head dim 128): five tcgen05 GEMMs per (128-key tile, 64-query tile) with every
// accumulator in TMEM, dK/dV resident in TMEM for the whole CTA (summed over the GQA group's query heads, so no atomics
// for dK/dV) and dQ pushed to an fp32 accumulator with TMA reduce-add.
//
//   per CTA: one (batch, kv head, 128-key tile);  loop over (query head in group) x (64-query tiles at/after the diagonal)
//     S^T  = K Q^T                 M=128(kv) N=64(q)  K=128(d)    A=K  (K-major)   B=Q  (K-major)
//     dP^T = V dO^T                M=128     N=64     K=128       A=V  (K-major)   B=dO (K-major)
//     P^T  = exp2(S^T*c - lse)     softmax warps, one key row per thread -> bf16 -> swizzled smem
//     dS^T = P^T o (dP^T - D) * scale                            -> bf16 -> swizzled smem
//     dV  += P^T dO                M=128(kv) N=128(d) K=64(q)     A=P^T (K-major)  B=dO (MN-major, same smem tile)
//     dK  += dS^T Q                M=128     N=128    K=64        A=dS^T(K-major)  B=Q  (MN-major, same smem tile)
//     dQ^T = K^T dS^T              M=128(d)  N=64(q)  K=128(kv)   A=K  (MN-major, same smem tile)  B=dS^T (MN-major)
//   dQ^T leaves through TMEM -> registers -> smem [64 q][128 d] fp32 -> cp.reduce.async.bulk.tensor (add) into dq_accum.
//
// TMEM columns: S^T [0,64)  dP^T [64,128)  dQ^T [128,192)  dV [192,320)  dK [320,448).
// Warp roles (448 threads): warp0 TMA producer, warp1 MMA issuer (+TMEM alloc), warps 2..9 softmax / dK epilogue (two
// warps share a key row: 32 query columns each, so every SM sub-partition has two softmax warps to hide latency),
// warps 10..13 dQ drain / dV epilogue.  With causal masking CTA x processes key tiles x and n-1-x back to back, which
// gives every CTA the same number of (key tile, query tile) pairs.
//
// Pre-pass: D[b,h,s] = rowsum(dO o O); post-pass: dq = bf16(dq_accum).
// Replaces the fused-attention backward the reference times through TE / cuDNN
// (/root/reference/simu_tools/efficency_test/test_fa_efficiency.py:250-497).
#include "sm100_common.cuh"

namespace smx {

SMX_DEVICE void bwd_tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SMX_DEVICE void bwd_tma_reduce_add_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SMX_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
SMX_DEVICE void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

constexpr int kBwdBKV = 128;
constexpr int kBwdBQ = 64;
constexpr int kBwdD = 128;
constexpr int kBwdSoftmaxWarps = 8;
constexpr int kBwdThreads = (2 + kBwdSoftmaxWarps + 4) * 32;
constexpr int kBwdKVBytes = kBwdBKV * kBwdD * 2;      // 32 KB each for K and V
constexpr int kBwdQBytes = kBwdBQ * kBwdD * 2;        // 16 KB each for Q_i and dO_i
constexpr int kBwdStageBytesUnused = 2 * kBwdQBytes;
constexpr int kBwdStagesUnused = 3;                      // Q/dO ring: a load issued when iteration g-1 retires lands before g+2 needs it
constexpr int kBwdPBytes = kBwdBKV * kBwdBQ * 2;      // 16 KB each for P^T and dS^T
constexpr int kBwdDqBytes = kBwdBQ * kBwdD * 4;       // 32 KB fp32 staging for the dQ reduce
// Per head-dim configuration.  DQK = 128 is the software-pipelined fast path.  DQK = 192 (MLA: qk 192 / v 128) needs a
// 192-column dK accumulator, so dQ^T (two M=128 blocks, the second one half empty) has to alias the S^T / dP^T columns:
// S^T(g+1) / dP^T(g+1) can then only be issued after dQ^T(g) has been drained (no look-ahead), and shared memory only has
// room for a 2-deep Q/dO ring.
template <int DQK>
struct BwdCfg {
  static constexpr int kKBytes = kBwdBKV * DQK * 2;
  static constexpr int kVBytes = kBwdBKV * kBwdD * 2;
  static constexpr int kQBytes = kBwdBQ * DQK * 2;
  static constexpr int kdOBytes = kBwdBQ * kBwdD * 2;
  static constexpr int kStageBytes = kQBytes + kdOBytes;
  static constexpr bool kLookahead = (DQK == 128);
  static constexpr int kStages = kLookahead ? 3 : 2;
  static constexpr int kDqBlocks = (DQK + 127) / 128;
  static constexpr int kTmS = 0, kTmDp = 64;
  static constexpr int kTmDq0 = kLookahead ? 128 : 0;     // block 1 (DQK = 192) lives at kTmDp
  static constexpr int kTmDv = kLookahead ? 192 : 128;
  static constexpr int kTmDk = kLookahead ? 320 : 256;
  // no alignment slack: a kernel without static smem gets a 1024-byte aligned dynamic window (checked at run time)
  static constexpr int kSmemBytes = kKBytes + kVBytes + kStages * kStageBytes + 2 * kBwdPBytes + kBwdDqBytes + 2 * kStages * 64 * 4 + 256;
  static_assert(kSmemBytes <= 227 * 1024, "fmha_bwd shared memory budget");
  static_assert(kTmDk + DQK <= 512, "fmha_bwd TMEM budget");
};

struct FmhaBwdParams {
  int B, S_q, S_kv, H, H_kv;
  float scale;
  int causal;
  int pair;           // causal load balancing: CTA x handles key tiles x and n-1-x
  const float* lse2;  // [B, H, S_pad]  lse * log2(e)   (+inf in the padding and where the row had no keys)
  const float* dsum;  // [B, H, S_pad]  rowsum(dO o O) * scale   (0 in the padding);  S_pad = ceil(S_q / 64) * 64
  long long s_pad;
  long long* trace;   // optional [3 roles][32 iterations][16 events] clock64 stamps of CTA (0,0,0); nullptr = off
  __nv_bfloat16* dk;  // strides in elements
  __nv_bfloat16* dv;
  long long dk_sb, dk_ss, dk_sh, dv_sb, dv_ss, dv_sh;
};

#define SMX_TRACE(role, g, ev)                                                                             \
  do {                                                                                                  \
    if (p.trace && (g) < 32 && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)     \
      p.trace[((role) * 32 + (g)) * 16 + (ev)] = clock64();                                             \
  } while (0)

template <int DQK>
__global__ void __launch_bounds__(kBwdThreads, 1)
fmha_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, const FmhaBwdParams p) {
  using Cfg = BwdCfg<DQK>;
  constexpr int BKV = kBwdBKV, BQ = kBwdBQ, D = kBwdD;      // D = head dim of V / dO / dV
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) {
    printf("simumax_b200: fmha_bwd dynamic smem base %u is not 1024-byte aligned\n", smem_u32(smem_raw));
    __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::kKBytes;
  uint8_t* sQdO = sV + Cfg::kVBytes;                    // [2 stages][Q 16 KB | dO 16 KB]
  uint8_t* sP = sQdO + Cfg::kStages * Cfg::kStageBytes;
  uint8_t* sdS = sP + kBwdPBytes;
  float* sDq = reinterpret_cast<float*>(sdS + kBwdPBytes);
  float* sLse = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sDq) + kBwdDqBytes);   // [2][64]
  float* sDsum = sLse + Cfg::kStages * 64;                                                            // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDsum + Cfg::kStages * 64);
  uint64_t* kv_full = bars;            // 1
  uint64_t* qdo_full = bars + 17;      // [Cfg::kStages]
  uint64_t* qdo_empty = bars + 17 + Cfg::kStages;
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* dp_full = bars + 7;
  uint64_t* dp_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;
  uint64_t* ds_full = bars + 11;
  uint64_t* dkq_done = bars + 12;
  uint64_t* dq_full = bars + 13;
  uint64_t* dq_empty = bars + 14;
  uint64_t* all_done = bars + 15;
  uint64_t* acc_empty = bars + 16;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 17 + 2 * Cfg::kStages);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int kv_head = blockIdx.y;
  const int batch = blockIdx.z;
  const int group = p.H / p.H_kv;
  const int shift = p.S_kv - p.S_q;
  const int n_q_tiles = (p.S_q + BQ - 1) / BQ;
  const int n_kv_tiles = (p.S_kv + BKV - 1) / BKV;
  // causal: CTA x owns key tiles x and (n-1-x) so every CTA sees the same number of (q tile, kv tile) pairs
  const int n_cta_tiles = (p.pair && (int)blockIdx.x != n_kv_tiles - 1 - (int)blockIdx.x) ? 2 : 1;
  auto tile_index = [&](int t) { return t == 0 ? (int)blockIdx.x : n_kv_tiles - 1 - (int)blockIdx.x; };
  auto tile_begin = [&](int k0) { return p.causal ? max(0, k0 - shift) / BQ : 0; };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, kBwdSoftmaxWarps);
    mbar_init(dp_full, 1);
    mbar_init(dp_empty, kBwdSoftmaxWarps);
    mbar_init(p_full, kBwdSoftmaxWarps);
    mbar_init(pv_done, 1);
    mbar_init(ds_full, kBwdSoftmaxWarps);
    mbar_init(dkq_done, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(all_done, 1);
    mbar_init(acc_empty, kBwdSoftmaxWarps + 4);
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<1>(tmem_base_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  const uint32_t tmem_s = tmem_base + Cfg::kTmS;
  const uint32_t tmem_dp = tmem_base + Cfg::kTmDp;
  const uint32_t tmem_dq = tmem_base + Cfg::kTmDq0;      // dQ^T block b sits at tmem_dq + 64 * b
  const uint32_t tmem_dv = tmem_base + Cfg::kTmDv;
  const uint32_t tmem_dk = tmem_base + Cfg::kTmDk;

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int g_base = 0, tw = 0;
      for (int t = 0; t < n_cta_tiles; ++t) {
        const int k0 = tile_index(t) * BKV;
        const int i_begin = tile_begin(k0);
        const int nq_local = max(0, n_q_tiles - i_begin);
        const int n_iter = nq_local * group;
        if (n_iter == 0) continue;
        if (tw > 0) mbar_wait(all_done, (tw - 1) & 1, 20);      // previous tile's MMAs are done with K / V
        mbar_arrive_expect_tx(kv_full, Cfg::kKBytes + Cfg::kVBytes);
#pragma unroll
        for (int c = 0; c < DQK / 64; ++c) bwd_tma_load_4d(&tmap_k, kv_full, sK + c * (BKV * 128), c * 64, k0, kv_head, batch);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) bwd_tma_load_4d(&tmap_v, kv_full, sV + c * (BKV * 128), c * 64, k0, kv_head, batch);
        for (int it = 0; it < n_iter; ++it) {
          const int g = g_base + it;
          const int st = g % Cfg::kStages;
          mbar_wait(&qdo_empty[st], ((g / Cfg::kStages) & 1) ^ 1, 21);
          uint8_t* sQ = sQdO + st * Cfg::kStageBytes;
          uint8_t* sdO = sQ + Cfg::kQBytes;
          mbar_arrive_expect_tx(&qdo_full[st], Cfg::kStageBytes + 2 * 64 * 4);
          const int head = kv_head * group + it / nq_local, q0 = (i_begin + it % nq_local) * BQ;
          const long long stat_off = ((long long)batch * p.H + head) * p.s_pad + q0;
          bulk_load_1d(sLse + st * 64, p.lse2 + stat_off, 64 * 4, &qdo_full[st]);
          bulk_load_1d(sDsum + st * 64, p.dsum + stat_off, 64 * 4, &qdo_full[st]);
#pragma unroll
          for (int c = 0; c < DQK / 64; ++c) bwd_tma_load_4d(&tmap_q, &qdo_full[st], sQ + c * (BQ * 128), c * 64, q0, head, batch);
#pragma unroll
          for (int c = 0; c < D / 64; ++c) bwd_tma_load_4d(&tmap_do, &qdo_full[st], sdO + c * (BQ * 128), c * 64, q0, head, batch);
        }
        g_base += n_iter;
        ++tw;
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_s = make_idesc(1, 1, 0, 0, BKV, BQ);    // K (K-major) x Q (K-major)      -> S^T[128, 64]
    constexpr uint32_t idesc_dv = make_idesc(1, 1, 0, 1, BKV, D);    // P^T (K-major) x dO (MN-major)  -> dV[128, 128]
    constexpr uint32_t idesc_dk = make_idesc(1, 1, 0, 1, BKV, DQK);  // dS^T (K-major) x Q (MN-major)  -> dK[128, DQK]
    constexpr uint32_t idesc_dq = make_idesc(1, 1, 1, 1, D, BQ);     // K (MN-major) x dS^T (MN-major) -> dQ^T[128, 64]
    // every UMMA smem descriptor is a constant plus (stage * stage_stride): build them once, the issue loop is then
    // two 64-bit adds per MMA (issue latency of this single thread is on the critical path: 32 MMAs per iteration)
    const uint64_t kK_km = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t kV_km = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
    const uint64_t kK_mn = make_smem_desc_sw128(smem_u32(sK), BKV * 128, 1024);
    const uint64_t kP_km = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
    const uint64_t kdS_km = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
    const uint64_t kdS_mn = make_smem_desc_sw128(smem_u32(sdS), BKV * 128, 1024);
    const uint64_t kQ_km = make_smem_desc_sw128(smem_u32(sQdO), 16, 1024);
    const uint64_t kQ_mn = make_smem_desc_sw128(smem_u32(sQdO), BQ * 128, 1024);
    const uint64_t kdO_km = make_smem_desc_sw128(smem_u32(sQdO + Cfg::kQBytes), 16, 1024);
    const uint64_t kdO_mn = make_smem_desc_sw128(smem_u32(sQdO + Cfg::kQBytes), BQ * 128, 1024);
    constexpr uint64_t kStageStep = Cfg::kStageBytes >> 4;
    auto issue_s = [&](int g) {       // S^T(g) = K Q^T
      const int st = g % Cfg::kStages;
      mbar_wait(&qdo_full[st], (g / Cfg::kStages) & 1, 22);
      if (g > 0) mbar_wait(s_empty, (g - 1) & 1, 23);
      if (!Cfg::kLookahead && g > 0) mbar_wait(dq_empty, (g - 1) & 1, 38);     // dQ^T(g-1) aliases the S^T / dP^T columns
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t b0 = kQ_km + st * kStageStep;
#pragma unroll
        for (int k = 0; k < DQK / 16; ++k)
          umma_f16<1>(tmem_s, kK_km + ((k / 4) * ((BKV * 128) >> 4) + (k % 4) * 2), b0 + ((k / 4) * ((BQ * 128) >> 4) + (k % 4) * 2),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit<1>(s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int g) {      // dP^T(g) = V dO^T
      const int st = g % Cfg::kStages;
      if (g > 0) mbar_wait(dp_empty, (g - 1) & 1, 24);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t b0 = kdO_km + st * kStageStep;
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16<1>(tmem_dp, kV_km + ((k / 4) * ((BKV * 128) >> 4) + (k % 4) * 2), b0 + ((k / 4) * ((BQ * 128) >> 4) + (k % 4) * 2),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit<1>(dp_full);
      }
      __syncwarp();
    };
    int g_base = 0, tw = 0;
    for (int t = 0; t < n_cta_tiles; ++t) {
      const int k0 = tile_index(t) * BKV;
      const int n_iter = max(0, n_q_tiles - tile_begin(k0)) * group;
      if (n_iter == 0) continue;
      mbar_wait(kv_full, tw & 1, 25);
      if constexpr (Cfg::kLookahead) {
        issue_s(g_base);
        issue_dp(g_base);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int g = g_base + it;
        const int st = g % Cfg::kStages;
        uint8_t* sQ = sQdO + st * Cfg::kStageBytes;
        uint8_t* sdO = sQ + Cfg::kQBytes;
        SMX_TRACE(0, g, 0);
        if constexpr (Cfg::kLookahead) {
          if (it + 1 < n_iter) issue_s(g + 1);
        } else {
          issue_s(g);          // waits for dQ^T(g-1) to be drained (aliased columns)
          issue_dp(g);
        }
        SMX_TRACE(0, g, 1);
        // dV += P^T dO
        mbar_wait(p_full, g & 1, 26);
        SMX_TRACE(0, g, 2);
        if (it == 0 && tw > 0) mbar_wait(acc_empty, (tw - 1) & 1, 36);   // previous tile's dK / dV have been drained
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t b0 = kdO_mn + st * kStageStep;
#pragma unroll
          for (int k = 0; k < BQ / 16; ++k)
            umma_f16<1>(tmem_dv, kP_km + k * 2, b0 + k * ((16 * 128) >> 4), idesc_dv, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit<1>(pv_done);
        }
        __syncwarp();
        SMX_TRACE(0, g, 3);
        if constexpr (Cfg::kLookahead) {
          if (it + 1 < n_iter) issue_dp(g + 1);
        }
        SMX_TRACE(0, g, 4);
        // dK += dS^T Q ;  dQ^T = K^T dS^T
        mbar_wait(ds_full, g & 1, 27);
        SMX_TRACE(0, g, 5);
        if (Cfg::kLookahead && g > 0) mbar_wait(dq_empty, (g - 1) & 1, 28);
        SMX_TRACE(0, g, 6);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t b0 = kQ_mn + st * kStageStep;
#pragma unroll
          for (int k = 0; k < BQ / 16; ++k)
            umma_f16<1>(tmem_dk, kdS_km + k * 2, b0 + k * ((16 * 128) >> 4), idesc_dk, (it > 0 || k > 0) ? 1u : 0u);
          // dQ^T block b = rows [128 b, 128 b + 128) of K^T (for DQK = 192 the upper half of block 1 reads past the K
          // tile into the V tile: finite garbage in TMEM lanes 64..127 that the drain drops)
#pragma unroll
          for (int b = 0; b < Cfg::kDqBlocks; ++b) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              umma_f16<1>(tmem_dq + 64 * b, kK_mn + b * ((2 * BKV * 128) >> 4) + k * ((16 * 128) >> 4),
                          kdS_mn + k * ((16 * 128) >> 4), idesc_dq, k > 0 ? 1u : 0u);
          }
          umma_commit<1>(dq_full);
          umma_commit<1>(dkq_done);
          umma_commit<1>(&qdo_empty[st]);
          if (it + 1 == n_iter) umma_commit<1>(all_done);
        }
        __syncwarp();
        SMX_TRACE(0, g, 7);
      }
      g_base += n_iter;
      ++tw;
    }
  } else if (warp_idx < 2 + kBwdSoftmaxWarps) {
    // ============== softmax / dS: one key row per thread, the 64 query columns split over two warps ==============
    const int quad = warp_idx % 4;
    const int half = (warp_idx - 2) / 4;
    const int col0 = half * 32;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const int tid = threadIdx.x - 64;     // 0..255
    const float c_log2 = p.scale * 1.4426950408889634f;
    int g_base = 0, tw = 0;
    for (int t = 0; t < n_cta_tiles; ++t) {
      const int k0 = tile_index(t) * BKV;
      const int kv_row = k0 + row_in_tile;
      const int i_begin = tile_begin(k0);
      const int nq_local = max(0, n_q_tiles - i_begin);
      const int n_iter = nq_local * group;
      for (int it = 0; it < n_iter; ++it) {
        const int g = g_base + it;
        const int q0 = (i_begin + it % nq_local) * BQ;
        // per-column statistics ride on the same transaction barrier as Q / dO (bulk copies issued by the producer)
        const int st = g % Cfg::kStages;
        const float4* lse4 = reinterpret_cast<const float4*>(sLse + st * 64 + col0);
        const float4* dsum4 = reinterpret_cast<const float4*>(sDsum + st * 64 + col0);
        if (warp_idx == 2) SMX_TRACE(1, g, 0);
        mbar_wait(&qdo_full[st], (g / Cfg::kStages) & 1, 37);
        if (warp_idx == 2) SMX_TRACE(1, g, 1);
        mbar_wait(s_full, g & 1, 29);
        if (warp_idx == 2) SMX_TRACE(1, g, 2);
        tcgen05_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_s + lane_addr + col0, r);
        tmem_ld_wait();
        if (warp_idx == 2) SMX_TRACE(1, g, 3);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        const bool need_mask = (p.causal && (k0 + BKV - 1 > q0 + shift)) || (k0 + BKV > p.S_kv);
        float pr[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 l = lse4[i4];
          pr[4 * i4 + 0] = exp2f(__uint_as_float(r[4 * i4 + 0]) * c_log2 - l.x);
          pr[4 * i4 + 1] = exp2f(__uint_as_float(r[4 * i4 + 1]) * c_log2 - l.y);
          pr[4 * i4 + 2] = exp2f(__uint_as_float(r[4 * i4 + 2]) * c_log2 - l.z);
          pr[4 * i4 + 3] = exp2f(__uint_as_float(r[4 * i4 + 3]) * c_log2 - l.w);
        }
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = (kv_row < p.S_kv) && (!p.causal || kv_row <= q0 + col0 + i + shift);
            pr[i] = ok ? pr[i] : 0.f;
          }
        }
        if (warp_idx == 2) SMX_TRACE(1, g, 4);
        if (g > 0) mbar_wait(pv_done, (g - 1) & 1, 30);     // dV(g-1) has finished reading the P^T tile
        if (warp_idx == 2) SMX_TRACE(1, g, 5);
        {
          uint8_t* prow = sP + row_in_tile * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 pk;
            pk.x = pack_bf16x2(pr[c * 8 + 0], pr[c * 8 + 1]);
            pk.y = pack_bf16x2(pr[c * 8 + 2], pr[c * 8 + 3]);
            pk.z = pack_bf16x2(pr[c * 8 + 4], pr[c * 8 + 5]);
            pk.w = pack_bf16x2(pr[c * 8 + 6], pr[c * 8 + 7]);
            *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (row_in_tile & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (warp_idx == 2) SMX_TRACE(1, g, 6);
        // dS^T = P^T o (dP^T - D) * scale
        mbar_wait(dp_full, g & 1, 31);
        if (warp_idx == 2) SMX_TRACE(1, g, 7);
        tcgen05_fence_after();
        tmem_ld_32x32b_x32(tmem_dp + lane_addr + col0, r);
        tmem_ld_wait();
        if (warp_idx == 2) SMX_TRACE(1, g, 8);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dp_empty);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 dsm = dsum4[i4];
          pr[4 * i4 + 0] *= fmaf(__uint_as_float(r[4 * i4 + 0]), p.scale, -dsm.x);      // (dP - D) * scale, D pre-scaled
          pr[4 * i4 + 1] *= fmaf(__uint_as_float(r[4 * i4 + 1]), p.scale, -dsm.y);
          pr[4 * i4 + 2] *= fmaf(__uint_as_float(r[4 * i4 + 2]), p.scale, -dsm.z);
          pr[4 * i4 + 3] *= fmaf(__uint_as_float(r[4 * i4 + 3]), p.scale, -dsm.w);
        }
        if (warp_idx == 2) SMX_TRACE(1, g, 9);
        if (g > 0) mbar_wait(dkq_done, (g - 1) & 1, 32);    // dK/dQ(g-1) have finished reading the dS^T tile
        if (warp_idx == 2) SMX_TRACE(1, g, 10);
        {
          uint8_t* drow = sdS + row_in_tile * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 pk;
            pk.x = pack_bf16x2(pr[c * 8 + 0], pr[c * 8 + 1]);
            pk.y = pack_bf16x2(pr[c * 8 + 2], pr[c * 8 + 3]);
            pk.z = pack_bf16x2(pr[c * 8 + 4], pr[c * 8 + 5]);
            pk.w = pack_bf16x2(pr[c * 8 + 6], pr[c * 8 + 7]);
            *reinterpret_cast<uint4*>(drow + (((half * 4 + c) ^ (row_in_tile & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
        if (warp_idx == 2) SMX_TRACE(1, g, 11);
      }
      // ---- dK epilogue (rows = keys; this warp writes head-dim columns [half*64, half*64+64))
      if (n_iter > 0) {
        mbar_wait(all_done, tw & 1, 33);
        tcgen05_fence_after();
      }
      const bool row_ok = kv_row < p.S_kv;
      __nv_bfloat16* out = p.dk + batch * p.dk_sb + (long long)(row_ok ? kv_row : 0) * p.dk_ss + kv_head * p.dk_sh + half * (DQK / 2);
#pragma unroll 1
      for (int c = 0; c < DQK / 64; ++c) {
        uint32_t o[32];
        if (n_iter > 0) {
          tmem_ld_32x32b_x32(tmem_dk + lane_addr + half * (DQK / 2) + c * 32, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * i]), __uint_as_float(o[8 * i + 1]));
            v.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]), __uint_as_float(o[8 * i + 3]));
            v.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]), __uint_as_float(o[8 * i + 5]));
            v.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]), __uint_as_float(o[8 * i + 7]));
            dst[i] = v;
          }
        }
      }
      if (n_iter > 0) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        g_base += n_iter;
        ++tw;
      }
    }
  } else {
    // ============================ dQ drain (one head-dim row of dQ^T per thread) ============================
    const int quad = warp_idx % 4;
    const int d_idx = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const int tid = threadIdx.x - (2 + kBwdSoftmaxWarps) * 32;    // 0..127
    int g_base = 0, tw = 0;
    for (int t = 0; t < n_cta_tiles; ++t) {
      const int k0 = tile_index(t) * BKV;
      const int i_begin = tile_begin(k0);
      const int nq_local = max(0, n_q_tiles - i_begin);
      const int n_iter = nq_local * group;
      for (int it = 0; it < n_iter; ++it) {
        const int g = g_base + it;
        const int head = kv_head * group + it / nq_local, q0 = (i_begin + it % nq_local) * BQ;
        if (quad == 2) SMX_TRACE(2, g, 0);
        mbar_wait(dq_full, g & 1, 34);
        if (quad == 2) SMX_TRACE(2, g, 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int b = 0; b < Cfg::kDqBlocks; ++b) {
          uint32_t r[64];
          {
            uint32_t (&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
            uint32_t (&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
            tmem_ld_32x32b_x32(tmem_dq + 64 * b + lane_addr, r0);
            tmem_ld_32x32b_x32(tmem_dq + 64 * b + lane_addr + 32, r1);
            tmem_ld_wait();
          }
          if (b + 1 == Cfg::kDqBlocks) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);
          }
          // the previous bulk reduce must have finished reading the staging buffer
          if (quad == 2) SMX_TRACE(2, g, 2);
          if (tid == 0) tma_store_wait_read<0>();
          named_bar_sync(2, 128);
          if (quad == 2) SMX_TRACE(2, g, 3);
#pragma unroll
          for (int i = 0; i < 64; ++i) sDq[i * D + d_idx] = __uint_as_float(r[i]);
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
          if (tid == 0) {
            // box = 128 head-dim columns starting at 128 b; columns past DQK are out of bounds and dropped by TMA
            bwd_tma_reduce_add_4d(&tmap_dq, sDq, 128 * b, q0, head, batch);
            tma_store_commit();
          }
        }
        if (quad == 2) SMX_TRACE(2, g, 4);
      }
      // ---- dV epilogue (rows = keys)
      if (n_iter > 0) {
        mbar_wait(all_done, tw & 1, 35);
        tcgen05_fence_after();
      }
      const int kv_row = k0 + d_idx;
      const bool row_ok = kv_row < p.S_kv;
      __nv_bfloat16* out = p.dv + batch * p.dv_sb + (long long)(row_ok ? kv_row : 0) * p.dv_ss + kv_head * p.dv_sh;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[32];
        if (n_iter > 0) {
          tmem_ld_32x32b_x32(tmem_dv + lane_addr + c * 32, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * i]), __uint_as_float(o[8 * i + 1]));
            v.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]), __uint_as_float(o[8 * i + 3]));
            v.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]), __uint_as_float(o[8 * i + 5]));
            v.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]), __uint_as_float(o[8 * i + 7]));
            dst[i] = v;
          }
        }
      }
      if (n_iter > 0) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        g_base += n_iter;
        ++tw;
      }
    }
    if (tid == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// Per-row statistics for the main kernel, one warp per (b, h, s) row incl. the padding rows s in [S, S_pad):
//   dsum[b,h,s] = scale * sum_d dO[b,s,h,d] * O[b,s,h,d]      (0 in the padding)
//   lse2[b,h,s] = lse[b,h,s] * log2(e)                         (+inf in the padding and for rows without keys)
__global__ void fmha_bwd_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                     const float* __restrict__ lse, float* __restrict__ dsum, float* __restrict__ lse2,
                                     int B, int S, int S_pad, int H, int D, float scale, long long o_sb, long long o_ss,
                                     long long o_sh, long long do_sb, long long do_ss, long long do_sh) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  const long long total = (long long)B * S_pad * H;
  if (warp >= total) return;
  const int h = warp % H;
  const int s = (warp / H) % S_pad;
  const int b = warp / ((long long)H * S_pad);
  const long long out = ((long long)b * H + h) * S_pad + s;
  if (s >= S) {
    if (lane == 0) { dsum[out] = 0.f; lse2[out] = INFINITY; }
    return;
  }
  const __nv_bfloat16* po = o + b * o_sb + s * o_ss + h * o_sh;
  const __nv_bfloat16* pd = d_o + b * do_sb + s * do_ss + h * do_sh;
  float acc = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const uint2 a = *reinterpret_cast<const uint2*>(po + i);
    const uint2 c = *reinterpret_cast<const uint2*>(pd + i);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 x = __bfloat1622float2(a2[j]), y = __bfloat1622float2(c2[j]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    dsum[out] = acc * scale;
    const float l = lse[((long long)b * H + h) * S + s];
    lse2[out] = (l > -INFINITY) ? l * 1.4426950408889634f : INFINITY;
  }
}

// dq[b,s,h,:] = bf16(dq_accum[b,s,h,:])   (dq strided, accum contiguous [B,S,H,D])
__global__ void fmha_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, long long rows,
                                           int S, int H, int D, long long sb, long long ss, long long sh) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // one thread per 8 elements
  const int per_row = D / 8;
  if (idx >= rows * per_row) return;
  const long long row = idx / per_row;
  const int c = (idx % per_row) * 8;
  const int h = row % H;
  const int s = (row / H) % S;
  const long long b = row / ((long long)H * S);
  const float4 x = *reinterpret_cast<const float4*>(acc + row * D + c);
  const float4 y = *reinterpret_cast<const float4*>(acc + row * D + c + 4);
  uint4 v;
  v.x = pack_bf16x2(x.x, x.y); v.y = pack_bf16x2(x.z, x.w); v.z = pack_bf16x2(y.x, y.y); v.w = pack_bf16x2(y.z, y.w);
  *reinterpret_cast<uint4*>(dq + b * sb + s * ss + h * sh + c) = v;
}

inline int bwd_make_tmap_4d(CUtensorMap* out, CUtensorMapDataType dt, CUtensorMapSwizzle sw, const void* ptr, uint64_t d0,
                            uint64_t d1, uint64_t d2, uint64_t d3, uint64_t s1_bytes, uint64_t s2_bytes,
                            uint64_t s3_bytes, uint32_t box0, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1_bytes, s2_bytes, s3_bytes};
  cuuint32_t box[4] = {box0, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "simumax_b200: fmha_bwd cuTensorMapEncodeTiled failed (%d)\n", (int)r);
    return -2;
  }
  return 0;
}

}  // namespace smx

extern "C" {

// Workspace: stats [2, B, H, S_pad] fp32 (S_pad = ceil(S_q/64)*64: scaled row sums, then lse in log2 units) and
// dq_accum [B,S_q,H,128] fp32, both provided by the caller; dq_accum is zeroed here.
// All of q,k,v,o,do,dq,dk,dv are "bshd" views with element strides (last dim contiguous, 16B aligned).
int smx_fmha_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse, void* dq,
                 void* dk, void* dv, float* dsum, float* dq_accum, int B, int S_q, int S_kv, int H, int H_kv, int d,
                 const long long* strides /* 8 x (sb, ss, sh): q k v o do dq dk dv */, float scale, int causal,
                 cudaStream_t stream, long long* trace) {
  using namespace smx;
  if (d != 128 && d != 192) return -31;      // d = qk head dim; v head dim is 128
  if (H % H_kv) return -30;
  const long long* sq = strides; const long long* sk = strides + 3; const long long* sv = strides + 6;
  const long long* so = strides + 9; const long long* sdo = strides + 12; const long long* sdq = strides + 15;
  const long long* sdk = strides + 18; const long long* sdv = strides + 21;
  CUtensorMap tq, tk, tv, tdo, tdq;
  const long long s_pad = (S_q + kBwdBQ - 1) / kBwdBQ * kBwdBQ;
  float* lse2 = dsum + (long long)B * H * s_pad;
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const auto SW = CU_TENSOR_MAP_SWIZZLE_128B;
  int rc = bwd_make_tmap_4d(&tq, BF, SW, q, d, S_q, H, B, sq[1] * 2, sq[2] * 2, sq[0] * 2, 64, kBwdBQ);
  if (rc) return rc;
  rc = bwd_make_tmap_4d(&tk, BF, SW, k, d, S_kv, H_kv, B, sk[1] * 2, sk[2] * 2, sk[0] * 2, 64, kBwdBKV);
  if (rc) return rc;
  rc = bwd_make_tmap_4d(&tv, BF, SW, v, kBwdD, S_kv, H_kv, B, sv[1] * 2, sv[2] * 2, sv[0] * 2, 64, kBwdBKV);
  if (rc) return rc;
  rc = bwd_make_tmap_4d(&tdo, BF, SW, d_o, kBwdD, S_q, H, B, sdo[1] * 2, sdo[2] * 2, sdo[0] * 2, 64, kBwdBQ);
  if (rc) return rc;
  rc = bwd_make_tmap_4d(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE, dq_accum, d, S_q, H, B,
                        (uint64_t)H * d * 4, (uint64_t)d * 4, (uint64_t)S_q * H * d * 4, kBwdD, kBwdBQ);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(dq_accum, 0, (size_t)B * S_q * H * d * sizeof(float), stream);
  if (e != cudaSuccess) return (int)e;
  {
    const long long rows = (long long)B * s_pad * H;
    const int threads = 256;
    const long long blocks = (rows * 32 + threads - 1) / threads;
    fmha_bwd_dsum_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
        (const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, lse, dsum, lse2, B, S_q, (int)s_pad, H, kBwdD, scale, so[0], so[1],
        so[2], sdo[0], sdo[1], sdo[2]);
  }
  FmhaBwdParams p;
  p.B = B; p.S_q = S_q; p.S_kv = S_kv; p.H = H; p.H_kv = H_kv; p.scale = scale; p.causal = causal;
  p.lse2 = lse2; p.dsum = dsum; p.s_pad = s_pad; p.trace = trace; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.dk_sb = sdk[0]; p.dk_ss = sdk[1]; p.dk_sh = sdk[2]; p.dv_sb = sdv[0]; p.dv_ss = sdv[1]; p.dv_sh = sdv[2];
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(fmha_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg<128>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fmha_bwd_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg<192>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int n_kv_tiles = (S_kv + kBwdBKV - 1) / kBwdBKV;
  p.pair = (causal && n_kv_tiles > 1) ? 1 : 0;
  dim3 grid(p.pair ? (n_kv_tiles + 1) / 2 : n_kv_tiles, H_kv, B);
  if (d == 128) fmha_bwd_kernel<128><<<grid, kBwdThreads, BwdCfg<128>::kSmemBytes, stream>>>(tq, tk, tv, tdo, tdq, p);
  else fmha_bwd_kernel<192><<<grid, kBwdThreads, BwdCfg<192>::kSmemBytes, stream>>>(tq, tk, tv, tdo, tdq, p);
  {
    const long long rows = (long long)B * S_q * H;
    const long long n = rows * (d / 8);
    fmha_bwd_dq_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dq_accum, (__nv_bfloat16*)dq, rows, S_q, H, d,
                                                                                  sdq[0], sdq[1], sdq[2]);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
