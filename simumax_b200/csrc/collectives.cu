// In-kernel collectives over NVLink 5 / NVSwitch for sm_100a: every transfer is issued by SM threads on
// symmetric (peer-mapped) memory -- plain ld/st.global on peer pointers, or multimem.* on the NVLS multicast
// address when one exists -- so a collective can live inside (or right behind) a compute kernel instead of being
// a separate NCCL launch.
//
//   all_gather      push: each rank stores its shard into every peer's output (multimem.st when NVLS: one store,
//                   the switch replicates)
//   reduce_scatter  pull: each rank reads slice[rank] of every peer and sums (multimem.ld_reduce when NVLS: the
//                   switch performs the reduction)
//   all_reduce      NVLS one-shot: ld_reduce own slice, multimem.st the result to everybody;  P2P: pull-reduce
//                   own slice then push it (two-shot inside one kernel)
//   all_to_all      push chunk[dst] of every rank into slot[src] of peer dst (EP dispatch/combine, Ulysses CP)
//
// Synchronisation: a per-(CTA, source-rank) flag word in a symmetric signal buffer, written with
// st.release.sys and polled with ld.acquire.sys; flags carry a monotonically increasing epoch so they never
// need resetting.  Every kernel does  barrier(epoch) -> data movement -> barrier(epoch+1): the first makes sure
// all peers have entered the collective (their inputs are final and their outputs may be overwritten), the
// second that all remote stores have landed before anybody returns.
//
// Replaces the reference's NCCL call sites (simu_tools/efficency_test/measure_comm_burst_window_worker.py:112-135).
#include "sm100_common.cuh"

namespace smx {

constexpr int kMaxRanks = 8;

struct PeerPtrs {
  void* data[kMaxRanks];       // symmetric data buffer of every rank (as mapped in this process)
  uint32_t* flags[kMaxRanks];  // symmetric signal buffer of every rank
  void* mc_data;               // NVLS multicast address of the data buffer (nullptr if unsupported)
  int rank, world;
};

SMX_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SMX_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs with the same blockIdx.x on all ranks meet.  Slot layout: flags[cta * kMaxRanks + src_rank].
SMX_DEVICE void cta_barrier_all_ranks(const PeerPtrs& pp, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < (unsigned)pp.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(pp.flags[peer] + blockIdx.x * kMaxRanks + pp.rank, epoch);
    const uint32_t* mine = pp.flags[pp.rank] + blockIdx.x * kMaxRanks + peer;
    const uint64_t t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (global_timer_ns() - t0 > 20000000000ull) {  // 20 s: peers may still be launching
        printf("simumax_b200: collective barrier timeout rank=%d peer=%d cta=%d epoch=%u\n", pp.rank, peer,
               (int)blockIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

SMX_DEVICE uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
SMX_DEVICE uint4 ld_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
SMX_DEVICE void st_v4(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
SMX_DEVICE uint4 multimem_ld_reduce_bf16x8(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
SMX_DEVICE uint4 multimem_ld_reduce_f32x4(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
SMX_DEVICE void multimem_st_v4(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
SMX_DEVICE uint4 add_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 fa = __bfloat1622float2(pa[i]), fb = __bfloat1622float2(pb[i]);
    pr[i] = __floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y);
  }
  return r;
}
SMX_DEVICE void acc_bf16x8(float (&acc)[8], uint4 a) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(pa[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}
SMX_DEVICE uint4 pack_bf16x8(const float (&acc)[8]) {
  uint4 r;
  r.x = pack_bf16x2(acc[0], acc[1]);
  r.y = pack_bf16x2(acc[2], acc[3]);
  r.z = pack_bf16x2(acc[4], acc[5]);
  r.w = pack_bf16x2(acc[6], acc[7]);
  return r;
}

// 16-byte granule reductions, bf16x8 (fp32 accumulate) or f32x4
template <bool F32>
SMX_DEVICE uint4 multimem_ld_reduce16(const void* mc) {
  if constexpr (F32) return multimem_ld_reduce_f32x4(mc);
  else return multimem_ld_reduce_bf16x8(mc);
}
template <bool F32>
SMX_DEVICE uint4 reduce16_peers(const PeerPtrs& pp, long long byte_off) {
  uint4 v[kMaxRanks];
#pragma unroll
  for (int s = 0; s < kMaxRanks; ++s)
    if (s < pp.world) v[s] = ld_v4(reinterpret_cast<const char*>(pp.data[(pp.rank + s) % pp.world]) + byte_off);
  if constexpr (F32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < kMaxRanks; ++s)
      if (s < pp.world) {
        acc.x += __uint_as_float(v[s].x); acc.y += __uint_as_float(v[s].y);
        acc.z += __uint_as_float(v[s].z); acc.w += __uint_as_float(v[s].w);
      }
    return make_uint4(__float_as_uint(acc.x), __float_as_uint(acc.y), __float_as_uint(acc.z), __float_as_uint(acc.w));
  } else {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < kMaxRanks; ++s)
      if (s < pp.world) acc_bf16x8(acc, v[s]);
    return pack_bf16x8(acc);
  }
}

// ----------------------------------------------------------------------------------------------
// Symmetric buffer layout (bytes): [ region A : in ][ region B : out ] -- offsets passed by the host.
// All sizes are multiples of 16 bytes.
// ----------------------------------------------------------------------------------------------

// all_gather: out[r * shard .. (r+1) * shard) on every rank <- in of rank r
__global__ void __launch_bounds__(512) all_gather_kernel(PeerPtrs pp, long long in_off, long long out_off,
                                                        long long shard_bytes, uint32_t epoch, int use_nvls) {
  cta_barrier_all_ranks(pp, epoch);
  const long long n16 = shard_bytes / 16;
  const char* src = reinterpret_cast<const char*>(pp.data[pp.rank]) + in_off;
  const long long dst_off = out_off + (long long)pp.rank * shard_bytes;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (use_nvls && pp.mc_data) {
    char* mc = reinterpret_cast<char*>(pp.mc_data) + dst_off;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += stride)
      multimem_st_v4(mc + i * 16, ld_nc_v4(src + i * 16));
  } else {
    // peers are visited in a rank-rotated order so that at any instant the 8 senders target 8 different receivers
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += stride) {
      const uint4 v = ld_nc_v4(src + i * 16);
#pragma unroll 1
      for (int s = 0; s < pp.world; ++s) {
        const int peer = (pp.rank + s) % pp.world;
        st_v4(reinterpret_cast<char*>(pp.data[peer]) + dst_off + i * 16, v);
      }
    }
  }
  cta_barrier_all_ranks(pp, epoch + 1);
}

// Sum granules [lo, hi) (16 B each, at byte offset src_off in every rank's buffer) and hand each result to `emit(i, v)`.
// Remote loads have microseconds of latency, so every thread keeps U granules x world peers in flight.
template <bool F32, int U, typename Emit>
SMX_DEVICE void reduce_granules(const PeerPtrs& pp, int use_nvls, long long src_off, long long lo, long long hi, Emit emit) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (use_nvls && pp.mc_data) {
    const char* mc = reinterpret_cast<const char*>(pp.mc_data) + src_off;
    for (long long i = first; i < hi; i += stride * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * stride < hi) v[u] = multimem_ld_reduce16<F32>(mc + (i + u * stride) * 16);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * stride < hi) emit(i + u * stride, v[u]);
    }
  } else {
    for (long long i = first; i < hi; i += stride * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * stride < hi) v[u] = reduce16_peers<F32>(pp, src_off + (i + u * stride) * 16);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * stride < hi) emit(i + u * stride, v[u]);
    }
  }
}

// reduce_scatter (sum; bf16 with fp32 accumulation, or fp32): out <- sum_r in_r[rank * shard .. (rank+1) * shard)
// 256 threads: (256 x <=128 regs) fits next to a resident 192-thread GEMM CTA, so an overlapped reduce-scatter does not
// evict compute CTAs from their SMs
template <bool F32, int U>
__global__ void __launch_bounds__(256) reduce_scatter_kernel(PeerPtrs pp, long long in_off, long long out_off,
                                                            long long shard_bytes, uint32_t epoch, int use_nvls) {
  cta_barrier_all_ranks(pp, epoch);
  char* dst = reinterpret_cast<char*>(pp.data[pp.rank]) + out_off;
  reduce_granules<F32, U>(pp, use_nvls, in_off + (long long)pp.rank * shard_bytes, 0, shard_bytes / 16,
                          [&](long long i, uint4 v) { st_v4(dst + i * 16, v); });
  cta_barrier_all_ranks(pp, epoch + 1);
}

// all_reduce (sum), result written to out on every rank: each rank reduces its slice and broadcasts it
template <bool F32, int U>
__global__ void __launch_bounds__(256) all_reduce_kernel(PeerPtrs pp, long long in_off, long long out_off,
                                                        long long total_bytes, uint32_t epoch, int use_nvls) {
  cta_barrier_all_ranks(pp, epoch);
  const long long n16_total = total_bytes / 16;
  // slice owned by this rank (16-byte granules, last rank takes the remainder)
  const long long per = (n16_total + pp.world - 1) / pp.world;
  const long long lo = per * pp.rank;
  const long long hi = min(n16_total, lo + per);
  if (use_nvls && pp.mc_data) {
    char* mc_out = reinterpret_cast<char*>(pp.mc_data) + out_off;
    reduce_granules<F32, U>(pp, 1, in_off, lo, hi, [&](long long i, uint4 v) { multimem_st_v4(mc_out + i * 16, v); });
  } else {
    reduce_granules<F32, U>(pp, 0, in_off, lo, hi, [&](long long i, uint4 v) {
#pragma unroll 1
      for (int s = 0; s < pp.world; ++s)
        st_v4(reinterpret_cast<char*>(pp.data[(pp.rank + s) % pp.world]) + out_off + i * 16, v);
    });
  }
  cta_barrier_all_ranks(pp, epoch + 1);
}

// all_to_all: out_dst[src * chunk ..] <- in_src[dst * chunk ..]
__global__ void __launch_bounds__(512) all_to_all_kernel(PeerPtrs pp, long long in_off, long long out_off,
                                                        long long chunk_bytes, uint32_t epoch) {
  cta_barrier_all_ranks(pp, epoch);
  const long long n16 = chunk_bytes / 16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const char* src_base = reinterpret_cast<const char*>(pp.data[pp.rank]) + in_off;
  for (int s = 0; s < pp.world; ++s) {
    const int peer = (pp.rank + s) % pp.world;
    const char* src = src_base + (long long)peer * chunk_bytes;
    char* dst = reinterpret_cast<char*>(pp.data[peer]) + out_off + (long long)pp.rank * chunk_bytes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += stride)
      st_v4(dst + i * 16, ld_nc_v4(src + i * 16));
  }
  cta_barrier_all_ranks(pp, epoch + 1);
}

// ----------------------------------------------------------------------------------------------------------------------
// Fused MoE dispatch = token permutation + expert-parallel all-to-all in ONE kernel (reference cost model: Permutation
// all2all fwd/bwd, simumax/core/transformer/moe_module.py:383-423).  Row r = (d, e, c) of the capacity-padded permuted order
// (d = destination EP rank, e = local expert on that rank, c = capacity slot) is gathered from local token row_map[r]
// (-1 = padding -> zeros), optionally scaled (backward: the routing probability), and stored straight into rank d's receive
// buffer at row (e * world + rank) * C + c -- expert-major [El][world * C], exactly what the grouped GEMM consumes -- so the
// permuted copy of the activations never exists in local memory.  One warp per row, destination ranks visited from rank + 1.
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) moe_dispatch_kernel(PeerPtrs pp, const __nv_bfloat16* __restrict__ x,
                                                          const int* __restrict__ row_map, const float* __restrict__ row_scale,
                                                          long long out_off, int H, int C, int El, uint32_t epoch) {
  cta_barrier_all_ranks(pp, epoch);
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long rows_per_dst = (long long)El * C;
  const long long total = rows_per_dst * pp.world;
  const int vec = H / 8;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += warps) {
    const int hop = (int)(w / rows_per_dst);
    const int d = (pp.rank + 1 + hop) % pp.world;
    const long long rem = w - (long long)hop * rows_per_dst;
    const int e = (int)(rem / C), c = (int)(rem - (long long)e * C);
    const long long r = (long long)d * rows_per_dst + rem;
    const int src = row_map[r];
    const float sc = (row_scale != nullptr && src >= 0) ? row_scale[r] : 1.f;
    char* dst = reinterpret_cast<char*>(pp.data[d]) + out_off + (((long long)e * pp.world + pp.rank) * C + c) * H * 2;
    const uint4* sp = reinterpret_cast<const uint4*>(x + (long long)(src < 0 ? 0 : src) * H);
    // four 16-byte loads in flight per lane (a row is 8-14 KB: one load per lane per trip leaves the warp latency bound)
    for (int i0 = lane; i0 < vec; i0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (src >= 0 && i0 + u * 32 < vec) v[u] = ld_nc_v4(sp + i0 + u * 32);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u * 32 >= vec) continue;
        if (row_scale != nullptr && src >= 0) {
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(h[j]);
            h[j] = __floats2bfloat162_rn(f.x * sc, f.y * sc);
          }
        }
        st_v4(dst + (long long)(i0 + u * 32) * 16, v[u]);
      }
    }
  }
  cta_barrier_all_ranks(pp, epoch + 1);
}

// Fused MoE combine (the way back): row (e, s, c) of this rank's expert-major buffer goes to rank s at row
// (rank * El + e) * C + c of its permuted-order buffer; the source rank then runs the local unpermute / weighted sum.
__global__ void __launch_bounds__(512) moe_combine_kernel(PeerPtrs pp, const __nv_bfloat16* __restrict__ y, long long out_off, int H,
                                                         int C, int El, uint32_t epoch) {
  cta_barrier_all_ranks(pp, epoch);
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)El * pp.world * C;
  const int vec = H / 8;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += warps) {
    // enumerate (hop, e, c) so that consecutive warps of one pass write to the same destination rank
    const long long per_dst = (long long)El * C;
    const int hop = (int)(w / per_dst);
    const int s = (pp.rank + 1 + hop) % pp.world;
    const long long rem = w - (long long)hop * per_dst;
    const int e = (int)(rem / C), c = (int)(rem - (long long)e * C);
    const uint4* sp = reinterpret_cast<const uint4*>(y + (((long long)e * pp.world + s) * C + c) * H);
    char* dst = reinterpret_cast<char*>(pp.data[s]) + out_off + (((long long)pp.rank * El + e) * C + c) * H * 2;
    for (int i0 = lane; i0 < vec; i0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * 32 < vec) v[u] = ld_nc_v4(sp + i0 + u * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * 32 < vec) st_v4(dst + (long long)(i0 + u * 32) * 16, v[u]);
    }
  }
  cta_barrier_all_ranks(pp, epoch + 1);
}

// point-to-point put (pipeline send): dst rank's buffer <- local buffer; flag protocol as above restricted to a pair
__global__ void __launch_bounds__(512) p2p_put_kernel(PeerPtrs pp, int dst_rank, long long in_off, long long out_off,
                                                     long long nbytes) {
  const long long n16 = nbytes / 16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const char* src = reinterpret_cast<const char*>(pp.data[pp.rank]) + in_off;
  char* dst = reinterpret_cast<char*>(pp.data[dst_rank]) + out_off;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += stride)
    st_v4(dst + i * 16, ld_nc_v4(src + i * 16));
}

// Chunked all-gather for the fused all-gather -> GEMM: CTA b pushes row chunk b of the local shard into every rank's
// gathered buffer and then raises flag[rank * gridDim.x + b] = epoch on every rank (st.release.sys after a system
// fence), so a consumer kernel can start on a chunk the moment it has landed.  No entry barrier: the consumer side is
// protected by the flags, the producer side by the caller's stream order (the previous consumer has finished).
__global__ void __launch_bounds__(512) all_gather_chunked_kernel(PeerPtrs pp, long long in_off, long long out_off,
                                                                long long flag_off, long long shard_bytes, uint32_t epoch) {
  const long long chunk = shard_bytes / gridDim.x;            // bytes per CTA (multiple of 16 by construction)
  const long long n16 = chunk / 16;
  const char* src = reinterpret_cast<const char*>(pp.data[pp.rank]) + in_off + (long long)blockIdx.x * chunk;
  const long long dst_off = out_off + (long long)pp.rank * shard_bytes + (long long)blockIdx.x * chunk;
  for (long long i = threadIdx.x; i < n16; i += blockDim.x) {
    const uint4 v = ld_nc_v4(src + i * 16);
#pragma unroll 1
    for (int s = 0; s < pp.world; ++s) {
      const int peer = (pp.rank + s) % pp.world;
      st_v4(reinterpret_cast<char*>(pp.data[peer]) + dst_off + i * 16, v);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < (unsigned)pp.world) {
    uint32_t* flag = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(pp.data[threadIdx.x]) + flag_off) +
                     (long long)pp.rank * gridDim.x + blockIdx.x;
    st_release_sys(flag, epoch);
  }
}

// cross-rank barrier only (one CTA): everything enqueued before it on every rank's stream is complete and visible
__global__ void __launch_bounds__(32) rank_barrier_kernel(PeerPtrs pp, uint32_t epoch) {
  cta_barrier_all_ranks(pp, epoch);
}

}  // namespace smx

extern "C" {

// ptrs: world data pointers then world flag pointers.  Returns cudaError.
static smx::PeerPtrs make_pp(void* const* data_ptrs, void* const* flag_ptrs, void* mc, int rank, int world) {
  smx::PeerPtrs pp;
  for (int i = 0; i < smx::kMaxRanks; ++i) {
    pp.data[i] = i < world ? data_ptrs[i] : nullptr;
    pp.flags[i] = i < world ? reinterpret_cast<uint32_t*>(flag_ptrs[i]) : nullptr;
  }
  pp.mc_data = mc;
  pp.rank = rank;
  pp.world = world;
  return pp;
}

int smx_coll_max_ctas() { return 128; }

int smx_all_gather(void* const* data_ptrs, void* const* flag_ptrs, void* mc, int rank, int world, long long in_off,
                   long long out_off, long long shard_bytes, unsigned epoch, int use_nvls, int ctas, void* stream) {
  if (world > smx::kMaxRanks || shard_bytes % 16) return -40;
  smx::all_gather_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      make_pp(data_ptrs, flag_ptrs, mc, rank, world), in_off, out_off, shard_bytes, epoch, use_nvls);
  return (int)cudaGetLastError();
}
int smx_reduce_scatter(void* const* data_ptrs, void* const* flag_ptrs, void* mc, int rank, int world, long long in_off,
                       long long out_off, long long shard_bytes, unsigned epoch, int use_nvls, int ctas, void* stream,
                       int f32) {
  if (world > smx::kMaxRanks || shard_bytes % 16) return -40;
  auto pp = make_pp(data_ptrs, flag_ptrs, mc, rank, world);
  auto st = reinterpret_cast<cudaStream_t>(stream);
#define SMX_RS(F, U) smx::reduce_scatter_kernel<F, U><<<ctas, 256, 0, st>>>(pp, in_off, out_off, shard_bytes, epoch, use_nvls)
  // loads in flight per thread = U (NVLS) or U * world (P2P): deeper unroll for small groups
  if (world <= 2) { if (f32) SMX_RS(true, 8); else SMX_RS(false, 8); }
  else if (world <= 4) { if (f32) SMX_RS(true, 4); else SMX_RS(false, 4); }
  else { if (f32) SMX_RS(true, 2); else SMX_RS(false, 2); }
#undef SMX_RS
  return (int)cudaGetLastError();
}
int smx_all_reduce(void* const* data_ptrs, void* const* flag_ptrs, void* mc, int rank, int world, long long in_off,
                   long long out_off, long long total_bytes, unsigned epoch, int use_nvls, int ctas, void* stream,
                   int f32) {
  if (world > smx::kMaxRanks || total_bytes % 16) return -40;
  auto pp = make_pp(data_ptrs, flag_ptrs, mc, rank, world);
  auto st = reinterpret_cast<cudaStream_t>(stream);
#define SMX_AR(F, U) smx::all_reduce_kernel<F, U><<<ctas, 256, 0, st>>>(pp, in_off, out_off, total_bytes, epoch, use_nvls)
  if (world <= 2) { if (f32) SMX_AR(true, 8); else SMX_AR(false, 8); }
  else if (world <= 4) { if (f32) SMX_AR(true, 4); else SMX_AR(false, 4); }
  else { if (f32) SMX_AR(true, 2); else SMX_AR(false, 2); }
#undef SMX_AR
  return (int)cudaGetLastError();
}
int smx_rank_barrier(void* const* data_ptrs, void* const* flag_ptrs, int rank, int world, unsigned epoch, void* stream) {
  if (world > smx::kMaxRanks) return -40;
  smx::rank_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(make_pp(data_ptrs, flag_ptrs, nullptr, rank, world),
                                                                              epoch);
  return (int)cudaGetLastError();
}
int smx_all_gather_chunked(void* const* data_ptrs, void* const* flag_ptrs, int rank, int world, long long in_off,
                           long long out_off, long long flag_off, long long shard_bytes, unsigned epoch, int chunks,
                           void* stream) {
  if (world > smx::kMaxRanks || chunks <= 0 || shard_bytes % (16LL * chunks)) return -40;
  smx::all_gather_chunked_kernel<<<chunks, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      make_pp(data_ptrs, flag_ptrs, nullptr, rank, world), in_off, out_off, flag_off, shard_bytes, epoch);
  return (int)cudaGetLastError();
}
int smx_all_to_all(void* const* data_ptrs, void* const* flag_ptrs, int rank, int world, long long in_off,
                   long long out_off, long long chunk_bytes, unsigned epoch, int ctas, void* stream) {
  if (world > smx::kMaxRanks || chunk_bytes % 16) return -40;
  smx::all_to_all_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      make_pp(data_ptrs, flag_ptrs, nullptr, rank, world), in_off, out_off, chunk_bytes, epoch);
  return (int)cudaGetLastError();
}
int smx_p2p_put(void* const* data_ptrs, void* const* flag_ptrs, int rank, int world, int dst_rank, long long in_off,
                long long out_off, long long nbytes, int ctas, void* stream) {
  if (world > smx::kMaxRanks || nbytes % 16) return -40;
  smx::p2p_put_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      make_pp(data_ptrs, flag_ptrs, nullptr, rank, world), dst_rank, in_off, out_off, nbytes);
  return (int)cudaGetLastError();
}

int smx_moe_dispatch(void* const* data_ptrs, void* const* flag_ptrs, int rank, int world, const void* x, const int* row_map,
                     const float* row_scale, long long out_off, int H, int C, int El, unsigned epoch, int ctas, void* stream) {
  if (world > smx::kMaxRanks || H % 8 || out_off % 16) return -40;
  smx::moe_dispatch_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      make_pp(data_ptrs, flag_ptrs, nullptr, rank, world), reinterpret_cast<const __nv_bfloat16*>(x), row_map, row_scale, out_off, H, C, El,
      epoch);
  return (int)cudaGetLastError();
}
int smx_moe_combine(void* const* data_ptrs, void* const* flag_ptrs, int rank, int world, const void* y, long long out_off, int H,
                    int C, int El, unsigned epoch, int ctas, void* stream) {
  if (world > smx::kMaxRanks || H % 8 || out_off % 16) return -40;
  smx::moe_combine_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      make_pp(data_ptrs, flag_ptrs, nullptr, rank, world), reinterpret_cast<const __nv_bfloat16*>(y), out_off, H, C, El, epoch);
  return (int)cudaGetLastError();
}

}  // extern "C"
