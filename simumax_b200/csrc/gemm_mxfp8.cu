// Block-scaled FP8 GEMM (MXFP8: e4m3 payload, one UE8M0 scale per 32 consecutive K elements of every row of A and B)
// on the 5th-gen tensor cores:  C[M,N] (bf16 / fp32) = (A o SFA)[M,K] x (B o SFB)[N,K]^T   ("TN", both K-major).
//
// tcgen05.mma.kind::mxf8f6f4.block_scale applies the scales inside the tensor core: besides the two smem descriptors it
// takes the TMEM addresses of the scale-factor tiles of A and B.  Scale factors of a 128-row x 128-K block (4 scale
// columns, sf_id 0..3 in the instruction descriptor) occupy 4 TMEM columns laid out as  lane = row % 32 (replicated in the
// four 32-lane quadrants), column = row / 32, byte = k-block  -- the layout tcgen05.cp.32x128b.warpx4 produces.  Here four
// "scale" warps (one per quadrant) fetch the scale bytes from global memory with plain loads and write them with
// tcgen05.st, which needs no staging in shared memory.
//
// Warp roles (320 threads): warp0 TMA producer (A, B tiles, 128 fp8 = 128 B per row, 128B swizzle), warp1 MMA issuer,
// warps 2..5 epilogue, warps 6..9 scale-factor loaders.  TMEM: accumulator 256 columns + per-stage scale slots.
//
// Fills the gap the reference leaves open: its B200 config has no fp8 entries at all
// (/root/reference/configs/system/b200_bf16_ceperm.json fp8_matmul: {}), its fp8 probe is per-tensor TE Float8Quantizer
// (/root/reference/simu_tools/efficency_test/test_gemm_efficiency.py:106-116).
#include "sm100_common.cuh"

namespace smx {

constexpr int kMxBlockM = 128, kMxBlockN = 256, kMxBlockK = 128;   // K in fp8 elements (= bytes)
constexpr int kMxThreads = 320;
constexpr int kMxABytes = kMxBlockM * 128, kMxBBytes = kMxBlockN * 128;
constexpr int kMxStageBytes = kMxABytes + kMxBBytes;               // 48 KB
constexpr int kMxStages = 4;
constexpr int kMxSmemBytes = kMxStages * kMxStageBytes + 1024 + 256;
constexpr int kMxSfColsPerStage = 4 + 8;                           // SFA: 128 rows -> 4 columns; SFB: 256 rows -> 8 columns
constexpr int kMxTmemAcc = 0, kMxTmemSf = 256;

struct MxParams {
  int M, N, K;
  long long ldc;
  const uint8_t* sfa;   // [M, K / 32] UE8M0
  const uint8_t* sfb;   // [N, K / 32]
  long long ldsfa, ldsfb;
  int m_tiles, n_tiles;
};

// instruction descriptor of kind::mxf8f6f4.block_scale (cute::UMMA::InstrDescriptorBlockScaled)
__host__ __device__ constexpr uint32_t make_idesc_mx(uint32_t m, uint32_t n, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (0u << 7) /*a e4m3*/ | (0u << 10) /*b e4m3*/ | ((n >> 3) << 17) | (1u << 23) /*scales are E8M0*/ |
         ((m >> 4) << 24) | (a_sf_id << 29);
}

SMX_DEVICE void umma_mxf8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate, uint32_t tsfa,
                          uint32_t tsfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}
SMX_DEVICE void tmem_st_32x32b_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

template <bool OUT_F32>
__global__ void __launch_bounds__(kMxThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ c_ptr,
                  const MxParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kMxStages * kMxStageBytes);
  uint64_t* empty_bar = full_bar + kMxStages;
  uint64_t* sf_full = empty_bar + kMxStages;
  uint64_t* acc_full = sf_full + kMxStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kMxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&sf_full[i], 4);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<1>(tmem_base_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int num_k_blocks = (p.K + kMxBlockK - 1) / kMxBlockK;

  if (warp_idx == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % p.m_tiles) * kMxBlockM, n0 = (t / p.m_tiles) * kMxBlockN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 51);
          uint8_t* sa = smem + stage * kMxStageBytes;
          uint8_t* sb = sa + kMxABytes;
          mbar_arrive_expect_tx(&full_bar[stage], kMxStageBytes);
          tma_load_3d(&tmap_a, &full_bar[stage], sa, kb * kMxBlockK, m0, 0);
#pragma unroll
          for (int c = 0; c < kMxBlockN / 64; ++c)
            tma_load_3d(&tmap_b, &full_bar[stage], sb + c * (64 * 128), kb * kMxBlockK, n0 + c * 64, 0);
          if (++stage == kMxStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    int stage = 0;
    uint32_t phase = 0, tile_phase = 0;
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem + kMxABytes), 16, 1024);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(acc_empty, tile_phase ^ 1, 52);
      tcgen05_fence_after();
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase, 53);
        mbar_wait(&sf_full[stage], phase, 54);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = a_desc0 + (uint64_t)stage * (kMxStageBytes >> 4);
          const uint64_t b_desc = b_desc0 + (uint64_t)stage * (kMxStageBytes >> 4);
          const uint32_t tsfa = tmem_base + kMxTmemSf + stage * kMxSfColsPerStage;
          const uint32_t tsfb = tsfa + 4;
#pragma unroll
          for (int k = 0; k < 4; ++k)     // 4 x K32; scale column k of the block via sf_id
            umma_mxf8(tmem_base + kMxTmemAcc, a_desc + k * 2, b_desc + k * 2, make_idesc_mx(kMxBlockM, kMxBlockN, k, k),
                      (kb > 0 || k > 0) ? 1u : 0u, tsfa, tsfb);
          umma_commit<1>(&empty_bar[stage]);
          if (kb == num_k_blocks - 1) umma_commit<1>(acc_full);
        }
        __syncwarp();
        if (++stage == kMxStages) { stage = 0; phase ^= 1; }
      }
      tile_phase ^= 1;
    }
  } else if (warp_idx < 6) {
    // ============================ epilogue ============================
    const int q = warp_idx % 4;
    using OutT = typename std::conditional<OUT_F32, float, __nv_bfloat16>::type;
    OutT* c_base = reinterpret_cast<OutT*>(c_ptr);
    uint32_t tile_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t % p.m_tiles) * kMxBlockM, n0 = (t / p.m_tiles) * kMxBlockN;
      const int row = m0 + q * 32 + lane;
      mbar_wait(acc_full, tile_phase, 55);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < kMxBlockN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + kMxTmemAcc + (uint32_t(q * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        if (row < p.M) {
          const int col0 = n0 + c * 32;
          OutT* dst = c_base + (long long)row * p.ldc + col0;
          const bool vec = (col0 + 32 <= p.N) && (p.ldc % 8 == 0);
          if (vec) {
            if constexpr (OUT_F32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]));
                v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
                v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
                v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                reinterpret_cast<uint4*>(dst)[j] = v;
              }
            }
          } else {
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < p.N) {
                if constexpr (OUT_F32) dst[j] = __uint_as_float(r[j]);
                else dst[j] = __float2bfloat16(__uint_as_float(r[j]));
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      tile_phase ^= 1;
    }
  } else {
    // ============================ scale-factor loaders (one warp per TMEM lane quadrant) ============================
    const int q = warp_idx % 4;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t % p.m_tiles) * kMxBlockM, n0 = (t / p.m_tiles) * kMxBlockN;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 56);      // the MMAs that read this stage's scale slot have retired
        tcgen05_fence_after();
        // word (row, kb) = the 4 scale bytes of k-blocks 4 kb .. 4 kb + 3 of that row; 0x7f7f7f7f (= 1.0) outside the matrix
        uint32_t w[12];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int row = m0 + c * 32 + lane;
          w[c] = (row < p.M) ? *reinterpret_cast<const uint32_t*>(p.sfa + (long long)row * p.ldsfa + kb * 4) : 0x7f7f7f7fu;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int row = n0 + c * 32 + lane;
          w[4 + c] = (row < p.N) ? *reinterpret_cast<const uint32_t*>(p.sfb + (long long)row * p.ldsfb + kb * 4) : 0x7f7f7f7fu;
        }
        const uint32_t base = tmem_base + kMxTmemSf + stage * kMxSfColsPerStage + lane_addr;
        tmem_st_32x32b_x4(base, w[0], w[1], w[2], w[3]);
        tmem_st_32x32b_x4(base + 4, w[4], w[5], w[6], w[7]);
        tmem_st_32x32b_x4(base + 8, w[8], w[9], w[10], w[11]);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sf_full[stage]);
        if (++stage == kMxStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace smx

extern "C" {

// A [M, K] e4m3 (row stride lda bytes), B [N, K] e4m3, sfa [M, K/32] / sfb [N, K/32] UE8M0 bytes; K % 128 == 0.
int smx_gemm_mxfp8(const void* A, const void* B, void* C, const void* sfa, const void* sfb, int M, int N, int K, int out_f32,
                   long long lda, long long ldb, long long ldc, long long ldsfa, long long ldsfb, int num_sms, void* stream_v) {
  using namespace smx;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % kMxBlockK || ldsfa % 4 || ldsfb % 4) return -16;
  CUtensorMap ta, tb;
  int rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, A, K, M, 1, lda, 0, kMxBlockK, kMxBlockM);
  if (rc) return rc;
  rc = make_tmap_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, B, K, N, 1, ldb, 0, kMxBlockK, 64);
  if (rc) return rc;
  MxParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.sfa = reinterpret_cast<const uint8_t*>(sfa); p.sfb = reinterpret_cast<const uint8_t*>(sfb);
  p.ldsfa = ldsfa; p.ldsfb = ldsfb;
  p.m_tiles = (M + kMxBlockM - 1) / kMxBlockM;
  p.n_tiles = (N + kMxBlockN - 1) / kMxBlockN;
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMxSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_mxfp8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMxSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (out_f32) gemm_mxfp8_kernel<true><<<grid, kMxThreads, kMxSmemBytes, stream>>>(ta, tb, C, p);
  else gemm_mxfp8_kernel<false><<<grid, kMxThreads, kMxSmemBytes, stream>>>(ta, tb, C, p);
  return (int)cudaGetLastError();
}

}  // extern "C"
