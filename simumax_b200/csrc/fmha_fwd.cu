// Flash-attention forward for sm_100a: tcgen05 MMAs with S and O accumulators in TMEM, TMA-fed K/V ring,
// online softmax with lazy (thresholded) rescaling so that O never leaves TMEM until the epilogue.
//
//   per CTA: one (batch, head, 128-query tile).  Loop over 64-key tiles (causal: up to the diagonal):
//     S   = Q K^T            tcgen05.mma  M=128 N=64  K=D_QK      (Q, K K-major in 128B-swizzled smem)
//     P   = exp2(S*c - m*c)  128 softmax threads, one query row each: tcgen05.ld -> registers -> bf16 -> smem
//     O  += P V              tcgen05.mma  M=128 N=D_V K=64        (P K-major smem, V MN-major smem)
//   epilogue: O / l -> bf16 global, lse = m*scale + ln(l).
//
// GQA (kv head = q head / group) and asymmetric head dims (MLA: D_QK=192, D_V=128) are template/arg options.
// Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//
// Warp roles (192 threads): warp0 TMA producer, warp1 MMA issuer (+TMEM alloc), warps 2..5 softmax/epilogue.
//
// Replaces the reference's TE / cuDNN fused-attention measurement back-end
// (/root/reference/simu_tools/efficency_test/test_fa_efficiency.py:250-497).
#include "sm100_common.cuh"

namespace smx {

SMX_DEVICE void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int kFmhaBlockQ = 128;
constexpr int kFmhaBlockKV = 64;
constexpr int kFmhaThreads = 192;
constexpr int kFmhaStages = 2;

template <int D_QK, int D_V>
struct FmhaCfg {
  static constexpr int kQBytes = kFmhaBlockQ * D_QK * 2;
  static constexpr int kKBytes = kFmhaBlockKV * D_QK * 2;
  static constexpr int kVBytes = kFmhaBlockKV * D_V * 2;
  static constexpr int kPBytes = kFmhaBlockQ * kFmhaBlockKV * 2;
  static constexpr int kStageBytes = kKBytes + kVBytes;
  // no alignment slack: the dynamic window of a kernel without static smem starts 1024-byte aligned (checked at run
  // time), and the slack would push two co-resident CTAs (2 x 113.25 KB + 2 x 1 KB reserved) over the 228 KB of an SM
  static constexpr int kSmemBytes = kQBytes + kPBytes + kFmhaStages * kStageBytes + 256;
  static constexpr int kTmemCols = 256;  // S: 64 columns, O: D_V (<=128) columns
};

struct FmhaParams {
  int B, S_q, S_kv, H, H_kv;
  float scale;           // softmax scale (1/sqrt(d))
  int causal;
  long long o_stride_b, o_stride_s, o_stride_h;   // elements
  long long lse_stride_b, lse_stride_h;           // elements (lse is [B, H, S_q] fp32)
};

// P_TMEM (experimental, SMX_FMHA_P_TMEM=1): the probabilities never touch shared memory -- the softmax threads write them as
// packed bf16 into 32 TMEM columns with tcgen05.st and the P V product takes its A operand from TMEM (tcgen05.mma [d], [a], b-desc).
// Saves the 16 KB smem store and 16 KB of operand reads per key tile of a kernel that is shared-memory-bandwidth bound.
template <int D_QK, int D_V, bool P_TMEM = false>
__global__ void __launch_bounds__(kFmhaThreads, (D_QK <= 128 ? 2 : 1))
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ o_ptr,
                float* __restrict__ lse_ptr, const FmhaParams p) {
  using Cfg = FmhaCfg<D_QK, D_V>;
  constexpr int BQ = kFmhaBlockQ, BKV = kFmhaBlockKV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) {
    printf("simumax_b200: fmha_fwd dynamic smem base %u is not 1024-byte aligned\n", smem_u32(smem_raw));
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + Cfg::kQBytes;
  uint8_t* sKV = sP + Cfg::kPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kFmhaStages * Cfg::kStageBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // [stages]
  uint64_t* kv_empty = kv_full + kFmhaStages;
  uint64_t* s_full = kv_empty + kFmhaStages;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  // heavy (late) query tiles first: with causal masking tile i has i+1 units of work
  const int q_tile = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int kv_head = head / (p.H / p.H_kv);
  const int q0 = q_tile * BQ;
  int kv_end = p.S_kv;
  if (p.causal) kv_end = min(p.S_kv, q0 + BQ + (p.S_kv - p.S_q));
  const int n_kv_tiles = (kv_end + BKV - 1) / BKV;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kFmhaStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<1>(tmem_base_ptr, Cfg::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  const uint32_t tmem_s = tmem_base;          // columns [0, 64)
  const uint32_t tmem_o = tmem_base + 64;     // columns [64, 64 + D_V)
  const uint32_t tmem_p = tmem_base + 64 + D_V;   // P_TMEM only: 32 columns of packed bf16 pairs (64 keys)
  static_assert(!P_TMEM || 64 + D_V + 32 <= Cfg::kTmemCols, "no TMEM columns left for P");

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
      for (int c = 0; c < D_QK / 64; ++c) tma_load_4d(&tmap_q, q_full, sQ + c * (BQ * 128), c * 64, q0, head, batch);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1, 1);
        uint8_t* sK = sKV + stage * Cfg::kStageBytes;
        uint8_t* sV = sK + Cfg::kKBytes;
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::kStageBytes);
        const int k0 = j * BKV;
#pragma unroll
        for (int c = 0; c < D_QK / 64; ++c) tma_load_4d(&tmap_k, &kv_full[stage], sK + c * (BKV * 128), c * 64, k0, kv_head, batch);
#pragma unroll
        for (int c = 0; c < D_V / 64; ++c) tma_load_4d(&tmap_v, &kv_full[stage], sV + c * (BKV * 128), c * 64, k0, kv_head, batch);
        if (++stage == kFmhaStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_s = make_idesc(1, 1, 0, 0, BQ, BKV);   // Q (K-major) x K (K-major) -> S[128, 64]
    constexpr uint32_t idesc_o = make_idesc(1, 1, 0, 1, BQ, D_V);   // P (K-major) x V (MN-major) -> O[128, D_V]
    // constant UMMA descriptors (+ stage * stride): the single issuing thread's latency is on the critical path
    const uint64_t kQ_km = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t kP_km = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
    const uint64_t kK_km = make_smem_desc_sw128(smem_u32(sKV), 16, 1024);
    const uint64_t kV_mn = make_smem_desc_sw128(smem_u32(sKV + Cfg::kKBytes), BKV * 128, 1024);
    constexpr uint64_t kStageStep = Cfg::kStageBytes >> 4;
    mbar_wait(q_full, 0, 2);
    int stage = 0;
    uint32_t phase = 0;
    auto issue_qk = [&](int j, int st) {
      // S_j = Q K_j^T  (needs: K_j landed, softmax has drained S_{j-1})
      if (j > 0) mbar_wait(s_empty, (j - 1) & 1, 3);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t b0 = kK_km + st * kStageStep;
#pragma unroll
        for (int k = 0; k < D_QK / 16; ++k)
          umma_f16<1>(tmem_s, kQ_km + ((k / 4) * ((BQ * 128) >> 4) + (k % 4) * 2), b0 + ((k / 4) * ((BKV * 128) >> 4) + (k % 4) * 2),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit<1>(s_full);
      }
      __syncwarp();
    };
    if (n_kv_tiles > 0) {
      mbar_wait(&kv_full[0], 0, 4);
      issue_qk(0, 0);
    }
    for (int j = 0; j < n_kv_tiles; ++j) {
      // look-ahead: S_{j+1} can start as soon as softmax has pulled S_j into registers
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kFmhaStages) { nstage = 0; nphase ^= 1; }
      if (j + 1 < n_kv_tiles) {
        mbar_wait(&kv_full[nstage], nphase, 5);
        issue_qk(j + 1, nstage);
      }
      // O += P_j V_j
      mbar_wait(p_full, j & 1, 6);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t b0 = kV_mn + stage * kStageStep;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          if constexpr (P_TMEM)      // 16 keys = 8 packed columns per K step
            umma_f16_ts(tmem_o, tmem_p + k * 8, b0 + k * ((16 * 128) >> 4), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          else
            umma_f16<1>(tmem_o, kP_km + k * 2, b0 + k * ((16 * 128) >> 4), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit<1>(pv_done);
        umma_commit<1>(&kv_empty[stage]);
      }
      __syncwarp();
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ============================ softmax + epilogue ============================
    const int q = warp_idx % 4;
    const int row_in_tile = q * 32 + lane;
    const int row = q0 + row_in_tile;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const float c_log2 = p.scale * 1.4426950408889634f;
    float m_used = -INFINITY;   // running reference max (raw score units)
    float l = 0.f;
    const int causal_shift = p.S_kv - p.S_q;
    for (int j = 0; j < n_kv_tiles; ++j) {
      mbar_wait(s_full, j & 1, 7);
      tcgen05_fence_after();
      uint32_t r[64];
      {
        uint32_t (&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
        uint32_t (&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
        tmem_ld_32x32b_x32(tmem_s + lane_addr, r0);
        tmem_ld_32x32b_x32(tmem_s + lane_addr + 32, r1);
        tmem_ld_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      const int k0 = j * BKV;
      const bool need_mask = (p.causal && (k0 + BKV - 1 > q0 + causal_shift)) || (k0 + BKV > p.S_kv);
      if (need_mask) {
        const int limit = p.causal ? min(row + causal_shift, p.S_kv - 1) : p.S_kv - 1;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (k0 + i > limit) r[i] = 0xff800000u;   // -inf
      }
      // four independent max chains (a single 64-long dependent chain is pure latency with one warp per scheduler)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mx4[j] = fmaxf(mx4[j], __uint_as_float(r[i + j]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // lazy rescale: only move the reference max when it grew by more than 2^8 in the exp2 domain
      float m_new = m_used;
      bool grow = (mx > m_used) && ((mx - m_used) * c_log2 > 8.f || m_used == -INFINITY);
      if (grow) m_new = mx;
      const bool any_grow = __any_sync(0xffffffffu, grow);
      // P smem and the O accumulator are only safe to touch once PV_{j-1} has retired
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1, 8);
        tcgen05_fence_after();
      }
      if (any_grow && j > 0) {
        const float alpha = (m_used == -INFINITY) ? 0.f : exp2f((m_used - m_new) * c_log2);
#pragma unroll 1
        for (int c = 0; c < D_V / 16; ++c) {
          uint32_t o[16];
          tmem_ld_32x32b_x16(tmem_o + lane_addr + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x16(tmem_o + lane_addr + c * 16, o);
        }
        tmem_st_wait();
        l *= alpha;
      }
      m_used = m_new;
      const float mc = (m_used == -INFINITY) ? 0.f : m_used * c_log2;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (P_TMEM) {
        // P row -> packed bf16 pairs -> 32 TMEM columns of this thread's lane (keys 2c, 2c+1 in column c)
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float e0 = exp2f(__uint_as_float(r[2 * c]) * c_log2 - mc);
          const float e1 = exp2f(__uint_as_float(r[2 * c + 1]) * c_log2 - mc);
          sum4[c & 3] += e0 + e1;
          pk[c] = pack_bf16x2(e0, e1);
        }
        l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        tmem_st_32x32b_x16(tmem_p + lane_addr, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        tmem_st_32x32b_x16(tmem_p + lane_addr + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
      } else {
      // P row -> bf16 -> 128B-swizzled K-major smem tile [128 rows][64 kv]
      uint8_t* prow = sP + row_in_tile * 128;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          e[i] = exp2f(__uint_as_float(r[c16 * 8 + i]) * c_log2 - mc);
          sum4[i & 3] += e[i];
        }
        uint4 pk;
        pk.x = pack_bf16x2(e[0], e[1]);
        pk.y = pack_bf16x2(e[2], e[3]);
        pk.z = pack_bf16x2(e[4], e[5]);
        pk.w = pack_bf16x2(e[6], e[7]);
        *reinterpret_cast<uint4*>(prow + ((c16 ^ (row_in_tile & 7)) << 4)) = pk;
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      __syncwarp();
      }
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue
    if (n_kv_tiles > 0) {
      mbar_wait(pv_done, (n_kv_tiles - 1) & 1, 9);
      tcgen05_fence_after();
    }
    const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
    // tcgen05.ld is warp-collective (.sync.aligned): every lane issues the loads, only the stores are predicated
    // (a partially valid q tile has valid and invalid rows inside one warp).
    const bool row_ok = row < p.S_q;
    __nv_bfloat16* orow = o_ptr + batch * p.o_stride_b + (long long)(row_ok ? row : 0) * p.o_stride_s + head * p.o_stride_h;
#pragma unroll 1
    for (int c = 0; c < D_V / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr + c * 32, o);
      tmem_ld_wait();
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
    if (row_ok && lse_ptr)
      lse_ptr[batch * p.lse_stride_b + head * p.lse_stride_h + row] =
          (l > 0.f) ? m_used * p.scale + __logf(l) : -INFINITY;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, Cfg::kTmemCols);
  }
}

// rank-4 tensor map [d3][d2][d1][d0], 128B swizzle, box = (64, rows, 1, 1)
inline int make_tmap_4d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
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
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "simumax_b200: fmha cuTensorMapEncodeTiled failed (%d) dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu]\n",
            (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3,
            (unsigned long long)s1_bytes, (unsigned long long)s2_bytes, (unsigned long long)s3_bytes);
    return -2;
  }
  return 0;
}

template <int D_QK, int D_V, bool P_TMEM = false>
int launch_fmha_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, void* o, float* lse,
                    const FmhaParams& p, cudaStream_t stream) {
  using Cfg = FmhaCfg<D_QK, D_V>;
  auto kern = fmha_fwd_kernel<D_QK, D_V, P_TMEM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((p.S_q + kFmhaBlockQ - 1) / kFmhaBlockQ, p.H, p.B);
  kern<<<grid, kFmhaThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(o), lse, p);
  return (int)cudaGetLastError();
}

}  // namespace smx

extern "C" {

// q: [B, S_q, H, D_QK], k: [B, S_kv, H_kv, D_QK], v: [B, S_kv, H_kv, D_V] with arbitrary (16B-aligned) batch / seq /
// head strides given in elements (innermost dim contiguous) -- packed qkv rows work without a copy.
// o: [B, S_q, H, D_V] (strides in elements), lse: [B, H, S_q] fp32 contiguous (may be null).
int smx_fmha_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S_q, int S_kv, int H,
                 int H_kv, int d_qk, int d_v, long long q_sb, long long q_ss, long long q_sh, long long k_sb,
                 long long k_ss, long long k_sh, long long v_sb, long long v_ss, long long v_sh, long long o_sb,
                 long long o_ss, long long o_sh, float scale, int causal, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (H % H_kv) return -30;
  CUtensorMap tq, tk, tv;
  int rc = smx::make_tmap_4d(&tq, q, d_qk, S_q, H, B, q_ss * 2, q_sh * 2, q_sb * 2, smx::kFmhaBlockQ);
  if (rc) return rc;
  rc = smx::make_tmap_4d(&tk, k, d_qk, S_kv, H_kv, B, k_ss * 2, k_sh * 2, k_sb * 2, smx::kFmhaBlockKV);
  if (rc) return rc;
  rc = smx::make_tmap_4d(&tv, v, d_v, S_kv, H_kv, B, v_ss * 2, v_sh * 2, v_sb * 2, smx::kFmhaBlockKV);
  if (rc) return rc;
  smx::FmhaParams p;
  p.B = B; p.S_q = S_q; p.S_kv = S_kv; p.H = H; p.H_kv = H_kv; p.scale = scale; p.causal = causal;
  p.o_stride_b = o_sb; p.o_stride_s = o_ss; p.o_stride_h = o_sh;
  p.lse_stride_b = (long long)H * S_q; p.lse_stride_h = S_q;
  static const bool p_tmem = [] { const char* e = getenv("SMX_FMHA_P_TMEM"); return e && e[0] == '1'; }();
  if (d_qk == 128 && d_v == 128 && p_tmem) return smx::launch_fmha_fwd<128, 128, true>(tq, tk, tv, o, lse, p, stream);
  if (d_qk == 128 && d_v == 128) return smx::launch_fmha_fwd<128, 128>(tq, tk, tv, o, lse, p, stream);
  if (d_qk == 64 && d_v == 64) return smx::launch_fmha_fwd<64, 64>(tq, tk, tv, o, lse, p, stream);
  if (d_qk == 192 && d_v == 128) return smx::launch_fmha_fwd<192, 128>(tq, tk, tv, o, lse, p, stream);
  return -31;
}

}  // extern "C"
