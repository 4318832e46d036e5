// Memory-bound kernels for sm_100a: RMSNorm fwd/bwd, SwiGLU fwd/bwd, RoPE (in place, fwd/bwd), fused
// cross-entropy fwd+bwd (online softmax, one pass over the logits, gradient written in place),
// fused Adam (fp32 master + bf16 model copy), squared-L2 norm, MoE token permute / unpermute.
//
// Design rules (B200: 148 SMs, ~6.6 TB/s measured HBM): 16-byte vector accesses, one row per CTA (or warp)
// with the row kept in registers between the reduction and the normalisation so every tensor is read once,
// grid sized to the row count (>> 148 x resident CTAs), fp32 math, bf16 I/O.
//
// These replace the reference's library calls in the calibration suite (TE fused CE, TE permute:
// /root/reference/simu_tools/efficency_test/test_ce_permute_efficiency.py:216-293) and are the building
// blocks of the measured training step.
#include "sm100_common.cuh"

namespace smx {

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

SMX_DEVICE void load8(const __nv_bfloat16* p, float (&f)[8]) {
  bf16x8 r = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(r.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
SMX_DEVICE void store8(__nv_bfloat16* p, const float (&f)[8]) {
  bf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<bf16x8*>(p) = r;
}

template <int THREADS>
SMX_DEVICE float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  if (l == 0) smem[w] = v;
  __syncthreads();
  float t = (threadIdx.x < THREADS / 32) ? smem[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) smem[0] = t;
  }
  __syncthreads();
  t = smem[0];
  __syncthreads();
  return t;
}
template <int THREADS>
SMX_DEVICE float block_max(float v, float* smem) {
  v = warp_max(v);
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  if (l == 0) smem[w] = v;
  __syncthreads();
  float t = (threadIdx.x < THREADS / 32) ? smem[threadIdx.x] : -INFINITY;
  if (w == 0) {
    t = warp_max(t);
    if (l == 0) smem[0] = t;
  }
  __syncthreads();
  t = smem[0];
  __syncthreads();
  return t;
}

// ----------------------------------------------------------------------------------------------
// RMSNorm.  One CTA per row, H <= THREADS * 8 * VPT elements kept in registers.
// ----------------------------------------------------------------------------------------------
template <int THREADS, int VPT>
__global__ void __launch_bounds__(THREADS) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ gamma,
                                                             __nv_bfloat16* __restrict__ y, float* __restrict__ rstd,
                                                             int H, float eps) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * H;
  float v[VPT][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < H) {
      load8(xr + c, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
    }
  }
  ss = block_sum<THREADS>(ss, red);
  const float r = rsqrtf(ss / H + eps);
  if (threadIdx.x == 0 && rstd) rstd[row] = r;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < H) {
      float g[8], o[8];
      load8(gamma + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[i][j] * r * g[j];
      store8(y + row * H + c, o);
    }
  }
}

// dx = rstd * (g*dy - xhat * mean(g*dy*xhat)); dgamma partials accumulated per CTA-stripe with atomics on fp32.
template <int THREADS, int VPT>
__global__ void __launch_bounds__(THREADS) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ gamma,
                                                             const __nv_bfloat16* __restrict__ dy,
                                                             const float* __restrict__ rstd,
                                                             __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                             int rows, int H, int add_to_dx) {
  __shared__ float red[32];
  float dg[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i][j] = 0.f;
  float g[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < H) load8(gamma + c, g[i]);
  }
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float r = rstd[row];
    float xv[VPT][8], gy[VPT][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * THREADS + threadIdx.x) * 8;
      if (c < H) {
        float d[8];
        load8(x + row * H + c, xv[i]);
        load8(dy + row * H + c, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = xv[i][j] * r;
          dg[i][j] += d[j] * xh;
          gy[i][j] = d[j] * g[i][j];
          dot += gy[i][j] * xh;
          xv[i][j] = xh;
        }
      }
    }
    dot = block_sum<THREADS>(dot, red) / H;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * THREADS + threadIdx.x) * 8;
      if (c < H) {
        float o[8];
        if (add_to_dx) load8(dx + row * H + c, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = r * (gy[i][j] - xv[i][j] * dot);
          o[j] = add_to_dx ? o[j] + t : t;
        }
        store8(dx + row * H + c, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < H) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dgamma + c + j, dg[i][j]);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Owner-side half of the fused GEMM -> reduce-scatter (gemm_2cta.cu, RS epilogue), fused with what follows a row-parallel
// linear in a transformer block:   x2 = residual + sum_s slots[s];   h = rmsnorm(x2) * gamma.
// The kernel itself waits (thread 0 of every CTA, ld.acquire.sys) until this rank's completion counter says that every
// CTA of every rank's GEMM has landed its stores -- no separate barrier launch.  One CTA per row.
// ----------------------------------------------------------------------------------------------
template <int THREADS, int VPT>
__global__ void __launch_bounds__(THREADS) rs_consume_kernel(const __nv_bfloat16* __restrict__ slots, int nslots,
                                                            long long slot_stride, const __nv_bfloat16* __restrict__ residual,
                                                            __nv_bfloat16* __restrict__ sum_out,
                                                            const __nv_bfloat16* __restrict__ gamma,
                                                            __nv_bfloat16* __restrict__ normed, float* __restrict__ rstd,
                                                            int H, float eps, const uint32_t* counter, uint32_t target) {
  __shared__ float red[32];
  if (counter != nullptr) {
    if (threadIdx.x == 0) {
      const uint64_t t0 = global_timer_ns();
      uint32_t v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if ((int32_t)(v - target) < 0 && global_timer_ns() - t0 > 20000000000ull) {
          printf("simumax_b200: reduce-scatter consumer timed out (counter %u, want %u)\n", v, target);
          __trap();
        }
      } while ((int32_t)(v - target) < 0);
    }
    __syncthreads();
  }
  const long long row = blockIdx.x;
  float v[VPT][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < H) {
      if (residual) load8(residual + row * H + c, v[i]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
      for (int s = 0; s < nslots; ++s) {
        float t[8];
        load8(slots + s * slot_stride + row * H + c, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += t[j];
      }
      // round to bf16 first: the normalisation must see exactly the residual stream that is stored (and re-read in backward)
      bf16x8 r;
#pragma unroll
      for (int j = 0; j < 4; ++j) r.v[j] = __floats2bfloat162_rn(v[i][2 * j], v[i][2 * j + 1]);
      if (sum_out) *reinterpret_cast<bf16x8*>(sum_out + row * H + c) = r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(r.v[j]);
        v[i][2 * j] = f.x;
        v[i][2 * j + 1] = f.y;
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
  if (gamma == nullptr) return;
  ss = block_sum<THREADS>(ss, red);
  const float r = rsqrtf(ss / H + eps);
  if (threadIdx.x == 0 && rstd) rstd[row] = r;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < H) {
      float g[8], o[8];
      load8(gamma + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[i][j] * r * g[j];
      store8(normed + row * H + c, o);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Persistent, software-pipelined row kernels (RMSNorm forward, fused residual add + RMSNorm, RMSNorm backward).
// One CTA walks rows blockIdx.x, blockIdx.x + gridDim.x, ...; the raw bf16 vectors of the NEXT row are already in flight while the
// current row is reduced (one __syncthreads per row, double-buffered partial sums), so the load -> reduce -> store chain of a row
// no longer leaves the memory pipe idle: the one-row-per-CTA versions above reached 59 % of the measured copy bandwidth.
// ----------------------------------------------------------------------------------------------
SMX_DEVICE void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
SMX_DEVICE uint4 pack8(const float (&f)[8]) {
  uint4 r;
  r.x = pack_bf16x2(f[0], f[1]); r.y = pack_bf16x2(f[2], f[3]); r.z = pack_bf16x2(f[4], f[5]); r.w = pack_bf16x2(f[6], f[7]);
  return r;
}
// all threads get the block-wide sum; `red` is this iteration's buffer (callers alternate between two)
template <int THREADS>
SMX_DEVICE float block_sum_1sync(float v, float* red) {
  v = warp_sum(v);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) t += red[w];
  return t;
}

// y = rmsnorm(x (+ delta)) * gamma; with ADD also sum_out = bf16(x + delta) (the residual stream the backward re-reads)
template <int THREADS, int VPT, bool ADD>
__global__ void __launch_bounds__(THREADS) rmsnorm_fwd_pipe_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                                                                  const __nv_bfloat16* __restrict__ gamma, __nv_bfloat16* __restrict__ sum_out,
                                                                  __nv_bfloat16* __restrict__ y, float* __restrict__ rstd, long long rows,
                                                                  int H, float eps) {
  __shared__ float red[2][32];
  uint4 g[VPT], cur[VPT], cur_d[ADD ? VPT : 1], nxt[VPT], nxt_d[ADD ? VPT : 1];
  bool on[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    on[i] = c < H;
    if (on[i]) g[i] = *reinterpret_cast<const uint4*>(gamma + c);
  }
  long long row = blockIdx.x;
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (on[i]) {
        const long long o = row * H + (i * THREADS + threadIdx.x) * 8;
        cur[i] = *reinterpret_cast<const uint4*>(x + o);
        if constexpr (ADD) cur_d[i] = *reinterpret_cast<const uint4*>(delta + o);
      }
  }
  for (int it = 0; row < rows; row += gridDim.x, ++it) {
    const long long nrow = row + gridDim.x;
    if (nrow < rows) {
#pragma unroll
      for (int i = 0; i < VPT; ++i)
        if (on[i]) {
          const long long o = nrow * H + (i * THREADS + threadIdx.x) * 8;
          nxt[i] = *reinterpret_cast<const uint4*>(x + o);
          if constexpr (ADD) nxt_d[i] = *reinterpret_cast<const uint4*>(delta + o);
        }
    }
    float v[VPT][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (on[i]) {
        unpack8(cur[i], v[i]);
        if constexpr (ADD) {
          float d[8];
          unpack8(cur_d[i], d);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += d[j];
          const uint4 r = pack8(v[i]);       // round first: the norm must see exactly the stored residual stream
          *reinterpret_cast<uint4*>(sum_out + row * H + (i * THREADS + threadIdx.x) * 8) = r;
          unpack8(r, v[i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
      }
    ss = block_sum_1sync<THREADS>(ss, red[it & 1]);
    const float r = rsqrtf(ss / H + eps);
    if (threadIdx.x == 0 && rstd) rstd[row] = r;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (on[i]) {
        float gf[8], o[8];
        unpack8(g[i], gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[i][j] * r * gf[j];
        *reinterpret_cast<uint4*>(y + row * H + (i * THREADS + threadIdx.x) * 8) = pack8(o);
      }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      cur[i] = nxt[i];
      if constexpr (ADD) cur_d[i] = nxt_d[i];
    }
  }
}

// dx = rstd * (g*dy - xhat * mean(g*dy*xhat)) (optionally += into dx); dgamma partial sums per CTA, one atomic pass at the end
template <int THREADS, int VPT>
__global__ void __launch_bounds__(THREADS) rmsnorm_bwd_pipe_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                                                  const __nv_bfloat16* __restrict__ dy, const float* __restrict__ rstd,
                                                                  __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, long long rows,
                                                                  int H, int add_to_dx, float* __restrict__ partial) {
  __shared__ float red[2][32];
  float dg[VPT][8], g[VPT][8];
  uint4 cx[VPT], cd[VPT], co[VPT], nx[VPT], nd[VPT], no[VPT];
  bool on[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    on[i] = c < H;
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i][j] = 0.f;
    if (on[i]) load8(gamma + c, g[i]);
  }
  long long row = blockIdx.x;
  float r_cur = 0.f, r_nxt = 0.f;
  if (row < rows) {
    r_cur = rstd[row];
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (on[i]) {
        const long long o = row * H + (i * THREADS + threadIdx.x) * 8;
        cx[i] = *reinterpret_cast<const uint4*>(x + o);
        cd[i] = *reinterpret_cast<const uint4*>(dy + o);
        if (add_to_dx) co[i] = *reinterpret_cast<const uint4*>(dx + o);
      }
  }
  for (int it = 0; row < rows; row += gridDim.x, ++it) {
    const long long nrow = row + gridDim.x;
    if (nrow < rows) {
      r_nxt = rstd[nrow];
#pragma unroll
      for (int i = 0; i < VPT; ++i)
        if (on[i]) {
          const long long o = nrow * H + (i * THREADS + threadIdx.x) * 8;
          nx[i] = *reinterpret_cast<const uint4*>(x + o);
          nd[i] = *reinterpret_cast<const uint4*>(dy + o);
          if (add_to_dx) no[i] = *reinterpret_cast<const uint4*>(dx + o);
        }
    }
    float xh[VPT][8], gy[VPT][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (on[i]) {
        float d[8];
        unpack8(cx[i], xh[i]);
        unpack8(cd[i], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] *= r_cur;
          dg[i][j] += d[j] * xh[i][j];
          gy[i][j] = d[j] * g[i][j];
          dot += gy[i][j] * xh[i][j];
        }
      }
    dot = block_sum_1sync<THREADS>(dot, red[it & 1]) / H;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (on[i]) {
        float o[8];
        if (add_to_dx) unpack8(co[i], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = r_cur * (gy[i][j] - xh[i][j] * dot);
          o[j] = add_to_dx ? o[j] + t : t;
        }
        *reinterpret_cast<uint4*>(dx + row * H + (i * THREADS + threadIdx.x) * 8) = pack8(o);
      }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      cx[i] = nx[i]; cd[i] = nd[i]; co[i] = no[i];
    }
    r_cur = r_nxt;
  }
  // dgamma: with a scratch buffer every CTA stores its partial row (coalesced, no atomics) and dgamma_reduce_kernel sums the rows --
  // 592 CTAs x H atomics on H addresses cost as much as the whole streaming pass
#pragma unroll
  for (int i = 0; i < VPT; ++i)
    if (on[i]) {
      const int c = (i * THREADS + threadIdx.x) * 8;
      if (partial) {
        float4* dst = reinterpret_cast<float4*>(partial + (long long)blockIdx.x * H + c);
        dst[0] = make_float4(dg[i][0], dg[i][1], dg[i][2], dg[i][3]);
        dst[1] = make_float4(dg[i][4], dg[i][5], dg[i][6], dg[i][7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dgamma + c + j, dg[i][j]);
      }
    }
}

// dgamma[c] += sum over the partial rows (one thread per column, coalesced)
__global__ void dgamma_reduce_kernel(const float* __restrict__ partial, int nrows, int H, float* __restrict__ dgamma) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int r = 0;
  for (; r + 3 < nrows; r += 4) {
    acc0 += partial[(long long)r * H + c];
    acc1 += partial[(long long)(r + 1) * H + c];
    acc2 += partial[(long long)(r + 2) * H + c];
    acc3 += partial[(long long)(r + 3) * H + c];
  }
  for (; r < nrows; ++r) acc0 += partial[(long long)r * H + c];
  dgamma[c] += (acc0 + acc1) + (acc2 + acc3);
}

// ----------------------------------------------------------------------------------------------
// SwiGLU: y = silu(a) * b with [a | b] = x[:, :F], x[:, F:]
// ----------------------------------------------------------------------------------------------
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows,
                                  int F) {
  const long long vec_per_row = F / 8;
  const long long total = rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row) * 8;
    float a[8], b[8], o[8];
    load8(x + r * 2 * F + c, a);
    load8(x + r * 2 * F + F + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] / (1.f + __expf(-a[j])) * b[j];
    store8(y + r * F + c, o);
  }
}
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                  __nv_bfloat16* __restrict__ dx, long long rows, int F) {
  const long long vec_per_row = F / 8;
  const long long total = rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row) * 8;
    float a[8], b[8], d[8], da[8], db[8];
    load8(x + r * 2 * F + c, a);
    load8(x + r * 2 * F + F + c, b);
    load8(dy + r * F + c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = 1.f / (1.f + __expf(-a[j]));
      const float silu = a[j] * s;
      db[j] = d[j] * silu;
      da[j] = d[j] * b[j] * (s + silu * (1.f - s));
    }
    store8(dx + r * 2 * F + c, da);
    store8(dx + r * 2 * F + F + c, db);
  }
}

// ----------------------------------------------------------------------------------------------
// RoPE (rotate-half, Llama convention) applied in place to the q and k heads of a packed qkv row.
// qkv row layout: [q heads (nq*d) | k heads (nk*d) | v heads].  One thread handles 8 (i, i+d/2) pairs.
// ----------------------------------------------------------------------------------------------
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ cos_t,
                            const float* __restrict__ sin_t, long long tokens, int seq_len, long long row_stride,
                            int n_rot_heads, int d, float sign, int head_stride) {
  const int half = d / 2;
  const int vec_per_head = half / 8;
  const long long total = tokens * n_rot_heads * vec_per_head;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_head);
    const long long th = i / vec_per_head;
    const int h = (int)(th % n_rot_heads);
    const long long t = th / n_rot_heads;
    const int pos = (int)(t % seq_len);
    __nv_bfloat16* p = qkv + t * row_stride + (long long)h * head_stride + v * 8;
    float x1[8], x2[8], o1[8], o2[8];
    load8(p, x1);
    load8(p + half, x2);
    const float* c = cos_t + (long long)pos * half + v * 8;
    const float* s = sin_t + (long long)pos * half + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cs = c[j], sn = s[j] * sign;
      o1[j] = x1[j] * cs - x2[j] * sn;
      o2[j] = x2[j] * cs + x1[j] * sn;
    }
    store8(p, o1);
    store8(p + half, o2);
  }
}

// ----------------------------------------------------------------------------------------------
// Fused cross entropy: per row loss = logsumexp(z) - z[target]; writes dlogits = (softmax - onehot) * scale
// in place over the logits (bf16).  One CTA per row; two passes over the row (max+sum online, then grad),
// the second pass usually hits L2 (row = V * 2 B <= 256 KB).
// ----------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) cross_entropy_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits,
                                                                       const long long* __restrict__ targets,
                                                                       float* __restrict__ loss, int V, long long ld,
                                                                       float grad_scale, int write_grad) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  __nv_bfloat16* z = logits + row * ld;
  const int nvec = V / 8;
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += THREADS) {
    float f[8];
    load8(z + i * 8, f);
    float lm = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) lm = fmaxf(lm, f[j]);
    const float nm = fmaxf(m, lm);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __expf(f[j] - nm);
    s = s * __expf(m - nm) + acc;
    m = nm;
  }
  for (int i = nvec * 8 + threadIdx.x; i < V; i += THREADS) {
    const float f = __bfloat162float(z[i]);
    const float nm = fmaxf(m, f);
    s = s * __expf(m - nm) + __expf(f - nm);
    m = nm;
  }
  const float gm = block_max<THREADS>(m, red);
  s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
  const float gs = block_sum<THREADS>(s, red);
  const long long tgt = targets[row];
  const float lse = gm + __logf(gs);
  if (threadIdx.x == 0) loss[row] = (tgt >= 0 && tgt < V) ? lse - __bfloat162float(z[tgt]) : 0.f;
  if (!write_grad) return;
  __syncthreads();  // z[tgt] must be read before any thread overwrites the row with the gradient
  const float inv = 1.f / gs;
  const float sc = (tgt >= 0 && tgt < V) ? grad_scale : 0.f;
  for (int i = threadIdx.x; i < nvec; i += THREADS) {
    float f[8];
    load8(z + i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = __expf(f[j] - gm) * inv;
      if (i * 8 + j == tgt) p -= 1.f;
      f[j] = p * sc;
    }
    store8(z + i * 8, f);
  }
  for (int i = nvec * 8 + threadIdx.x; i < V; i += THREADS) {
    float p = __expf(__bfloat162float(z[i]) - gm) * inv;
    if (i == tgt) p -= 1.f;
    z[i] = __float2bfloat16(p * sc);
  }
}

// Vocab-parallel cross-entropy (tensor-parallel LM head, Megatron vocab_parallel_cross_entropy): every rank holds V_local
// columns [v0, v0 + V_local) of the logits.  Phase 1 writes per-row (max, sum exp(x - max), target logit or 0, has_target) of
// the local shard; the tiny [rows, 4] fp32 statistics are all-gathered over the TP group; phase 2 combines the `world`
// statistics per row, emits the loss and overwrites the local logits with their gradient -- each logit is read twice and
// written once, exactly like the single-rank fused kernel.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) ce_vp_stats_kernel(const __nv_bfloat16* __restrict__ logits,
                                                             const long long* __restrict__ targets, float4* __restrict__ stats,
                                                             int V, long long ld, long long v0) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* z = logits + row * ld;
  const int nvec = V / 8;
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += THREADS) {
    float f[8];
    load8(z + i * 8, f);
    float lm = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) lm = fmaxf(lm, f[j]);
    const float nm = fmaxf(m, lm);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __expf(f[j] - nm);
    s = s * __expf(m - nm) + acc;
    m = nm;
  }
  const float gm = block_max<THREADS>(m, red);
  s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
  const float gs = block_sum<THREADS>(s, red);
  if (threadIdx.x == 0) {
    const long long t = targets[row] - v0;
    const bool mine = t >= 0 && t < V;
    stats[row] = make_float4(gm, gs, mine ? __bfloat162float(z[t]) : 0.f, mine ? 1.f : 0.f);
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) ce_vp_grad_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ targets,
                                                            const float4* __restrict__ all_stats, int world, long long rows,
                                                            float* __restrict__ loss, int V, long long ld, long long v0,
                                                            float grad_scale) {
  const long long row = blockIdx.x;
  __nv_bfloat16* z = logits + row * ld;
  float gm = -INFINITY;
  for (int r = 0; r < world; ++r) gm = fmaxf(gm, all_stats[r * rows + row].x);
  float gs = 0.f, tl = 0.f, has = 0.f;
  for (int r = 0; r < world; ++r) {
    const float4 st = all_stats[r * rows + row];
    gs += st.y * __expf(st.x - gm);
    tl += st.z;
    has += st.w;
  }
  if (threadIdx.x == 0) loss[row] = has > 0.f ? gm + __logf(gs) - tl : 0.f;
  const float inv = 1.f / gs;
  const float sc = has > 0.f ? grad_scale : 0.f;
  const long long tgt = targets[row] - v0;
  const int nvec = V / 8;
  for (int i = threadIdx.x; i < nvec; i += THREADS) {
    float f[8];
    load8(z + i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = __expf(f[j] - gm) * inv;
      if (i * 8 + j == tgt) p -= 1.f;
      f[j] = p * sc;
    }
    store8(z + i * 8, f);
  }
}

// ----------------------------------------------------------------------------------------------
// Optimizer: squared L2 norm of fp32 grads (per-CTA partials + one atomic), fused AdamW step on
// flat fp32 master / m / v with bf16 model-param write-back and grad clipping scale.
// ----------------------------------------------------------------------------------------------
__global__ void sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long nv = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = nv * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    acc += g[i] * g[i];
  acc = block_sum<256>(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void adamw_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                             float* __restrict__ grad, __nv_bfloat16* __restrict__ param, long long n, float lr,
                             float beta1, float beta2, float eps, float wd, float bc1, float bc2,
                             const float* __restrict__ sqnorm, float max_norm, float grad_prescale, int zero_grad) {
  float clip = 1.f;
  if (sqnorm) {
    const float nrm = sqrtf(*sqnorm) * grad_prescale;
    if (max_norm > 0.f && nrm > max_norm) clip = max_norm / (nrm + 1e-6f);
  }
  const float gscale = grad_prescale * clip;
  const long long nv = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float4 g4 = reinterpret_cast<float4*>(grad)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    float4 w4 = reinterpret_cast<float4*>(master)[i];
    float gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w},
          ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g = gg[j] * gscale;
      mm[j] = beta1 * mm[j] + (1.f - beta1) * g;
      vv[j] = beta2 * vv[j] + (1.f - beta2) * g * g;
      const float upd = (mm[j] / bc1) / (sqrtf(vv[j] / bc2) + eps);
      ww[j] = ww[j] * (1.f - lr * wd) - lr * upd;
    }
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(master)[i] = make_float4(ww[0], ww[1], ww[2], ww[3]);
    if (param) {
      uint2 pk;
      pk.x = pack_bf16x2(ww[0], ww[1]);
      pk.y = pack_bf16x2(ww[2], ww[3]);
      reinterpret_cast<uint2*>(param)[i] = pk;
    }
    if (zero_grad) reinterpret_cast<float4*>(grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ----------------------------------------------------------------------------------------------
// MoE permute / unpermute.  `row_map[i]` = source token of permuted row i (permute: gather) ;
// unpermute = weighted scatter-add back: out[token] += prob[i] * x[i] (topk rows per token, done as a gather
// over the inverse map so no atomics are needed).
// ----------------------------------------------------------------------------------------------
__global__ void permute_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                    const int* __restrict__ row_map, long long n_rows, int H) {
  const int vec_per_row = H / 8;
  const long long total = n_rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row) * 8;
    const int s = row_map[r];
    bf16x8 v;
    if (s >= 0) v = *reinterpret_cast<const bf16x8*>(src + (long long)s * H + c);
    else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v.v[j] = __floats2bfloat162_rn(0.f, 0.f);
    }
    *reinterpret_cast<bf16x8*>(dst + r * H + c) = v;
  }
}
// out[t] = sum_k prob[t,k] * x[inv_map[t,k]]
__global__ void unpermute_combine_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                         const int* __restrict__ inv_map, const float* __restrict__ probs,
                                         long long tokens, int topk, int H) {
  const int vec_per_row = H / 8;
  const long long total = tokens * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / vec_per_row;
    const int c = (int)(i - t * vec_per_row) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < topk; ++k) {
      const int r = inv_map[t * topk + k];
      if (r < 0) continue;
      const float p = probs ? probs[t * topk + k] : 1.f;
      float f[8];
      load8(x + (long long)r * H + c, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += p * f[j];
    }
    store8(out + t * H + c, acc);
  }
}

// d_probs[t, k] = < dy[t, :], y_perm[inv_map[t, k], :] >  (0 for dropped slots): gradient of the MoE combine weights.
__global__ void moe_prob_grad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y_perm,
                                     const int* __restrict__ inv_map, float* __restrict__ d_probs, long long slots, int topk, int H) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long sidx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sidx < slots; sidx += warps) {
    const int r = inv_map[sidx];
    float acc = 0.f;
    if (r >= 0) {
      const __nv_bfloat16* a = dy + (sidx / topk) * H;
      const __nv_bfloat16* b = y_perm + (long long)r * H;
      for (int c = lane * 8; c < H; c += 256) {
        float fa[8], fb[8];
        load8(a + c, fa);
        load8(b + c, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += fa[j] * fb[j];
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) d_probs[sidx] = acc;
  }
}

// residual add: y = a + b (bf16) -- used by the training harness between sub-layers
__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           __nv_bfloat16* __restrict__ y, long long n) {
  const long long nv = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float x[8], z[8];
    load8(a + i * 8, x);
    load8(b + i * 8, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += z[j];
    store8(y + i * 8, x);
  }
}

// ----------------------------------------------------------------------------------------------
// Embedding: out[t, :] = table[tok[t] - v0, :] (zeros when the token is not in this rank's vocabulary shard [v0, v0 + V)),
// and its backward grad[tok[t] - v0, :] += dy[t, :] into the fp32 main-grad table (red.global.add.v4.f32: no return trip).
// One warp per token row, 16-byte accesses.
// ----------------------------------------------------------------------------------------------
__global__ void embedding_fwd_kernel(const long long* __restrict__ tok, const __nv_bfloat16* __restrict__ table,
                                     __nv_bfloat16* __restrict__ out, long long rows, int H, long long v0, long long V) {
  const int lane = threadIdx.x % 32;
  const long long warps = (long long)gridDim.x * (blockDim.x / 32);
  for (long long r = blockIdx.x * (long long)(blockDim.x / 32) + threadIdx.x / 32; r < rows; r += warps) {
    const long long id = tok[r] - v0;
    const bool mine = id >= 0 && id < V;
    const uint4* src = reinterpret_cast<const uint4*>(table + (mine ? id : 0) * H);
    uint4* dst = reinterpret_cast<uint4*>(out + r * H);
    for (int c = lane; c < H / 8; c += 32) dst[c] = mine ? src[c] : make_uint4(0, 0, 0, 0);
  }
}

__global__ void embedding_bwd_kernel(const long long* __restrict__ tok, const __nv_bfloat16* __restrict__ dy,
                                     float* __restrict__ grad, long long rows, int H, long long v0, long long V) {
  const int lane = threadIdx.x % 32;
  const long long warps = (long long)gridDim.x * (blockDim.x / 32);
  for (long long r = blockIdx.x * (long long)(blockDim.x / 32) + threadIdx.x / 32; r < rows; r += warps) {
    const long long id = tok[r] - v0;
    if (id < 0 || id >= V) continue;
    float* g = grad + id * H;
    for (int c = lane; c < H / 8; c += 32) {
      float f[8];
      load8(dy + r * H + c * 8, f);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + c * 8), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + c * 8 + 4), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
    }
  }
}

// ----------------------------------------------------------------------------------------------
// MLA key assembly (DeepSeek-V2/V3; reference MLAAttention, dense_module.py:2569-2853): the kv_up GEMM writes [k_nope | v] per head,
// the rotary key slice is shared by all heads.  k[t, h, :dn] = kv[t, h, :dn];  k[t, h, dn:dn+dr] = k_rope[t, :].
// And its backward: d_kv[t, h, :dn] = dk[t, h, :dn];  d_k_rope[t, :] = sum_h dk[t, h, dn:].
// ----------------------------------------------------------------------------------------------
__global__ void mla_k_assemble_kernel(const __nv_bfloat16* __restrict__ kv, long long kv_row, int kv_head,
                                      const __nv_bfloat16* __restrict__ krope, long long kr_row,
                                      __nv_bfloat16* __restrict__ k, long long tokens, int heads, int dn, int dr) {
  const int vph = (dn + dr) / 8;
  const long long total = tokens * heads * vph;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vph);
    const long long th = i / vph;
    const int h = (int)(th % heads);
    const long long t = th / heads;
    const uint4* src = (v * 8 < dn) ? reinterpret_cast<const uint4*>(kv + t * kv_row + (long long)h * kv_head + v * 8)
                                    : reinterpret_cast<const uint4*>(krope + t * kr_row + (v * 8 - dn));
    reinterpret_cast<uint4*>(k)[i] = *src;
  }
}

__global__ void mla_k_split_bwd_kernel(const __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ d_kv, long long kv_row,
                                       int kv_head, __nv_bfloat16* __restrict__ d_krope, long long kr_row, long long tokens,
                                       int heads, int dn, int dr) {
  // one warp per token; lane l owns vector l of every head (dn / 8 + dr / 8 <= 32 vectors per head)
  const int lane = threadIdx.x % 32;
  const int vn = dn / 8, vr = dr / 8;
  const long long warps = (long long)gridDim.x * (blockDim.x / 32);
  for (long long t = blockIdx.x * (long long)(blockDim.x / 32) + threadIdx.x / 32; t < tokens; t += warps) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int h = 0; h < heads; ++h) {
      const __nv_bfloat16* src = dk + (t * heads + h) * (long long)(dn + dr);
      if (lane < vn) {
        *reinterpret_cast<uint4*>(d_kv + t * kv_row + (long long)h * kv_head + lane * 8) = *reinterpret_cast<const uint4*>(src + lane * 8);
      } else if (lane < vn + vr) {
        float f[8];
        load8(src + lane * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    if (lane >= vn && lane < vn + vr) store8(d_krope + t * kr_row + (lane - vn) * 8, acc);
  }
}

// 2-D strided copy in 16-byte vectors: dst[r, :width] = src[r, :width]
__global__ void copy2d_kernel(const unsigned char* __restrict__ src, long long src_ld, unsigned char* __restrict__ dst, long long dst_ld,
                              long long rows, int vec_per_row) {
  const long long total = rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int v = (int)(i % vec_per_row);
    *reinterpret_cast<uint4*>(dst + r * dst_ld + v * 16) = *reinterpret_cast<const uint4*>(src + r * src_ld + v * 16);
  }
}

// dst[a][b][:] = src[b][a][:]  (inner = contiguous run of `vec` 16-byte vectors): packs / unpacks the per-destination column blocks
// around a context-parallel (Ulysses) all-to-all -- [tokens, ranks, width] <-> [ranks, tokens, width].
__global__ void transpose01_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long A, long long B, int vec) {
  const long long total = A * B * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long ab = i / vec;
    const long long b = ab % B, a = ab / B;
    dst[i] = src[(b * A + a) * vec + v];
  }
}

// streaming copy (the "default" memory-bandwidth calibration op)
__global__ void copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace smx

// ----------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------
using bf16 = __nv_bfloat16;
static inline int grid_for(long long work_items, int threads, int cap = 148 * 16) {
  long long g = (work_items + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}
#define SMX_STREAM reinterpret_cast<cudaStream_t>(stream)

namespace smx {
// out[i] = sum_s slots[s * n + i]   (bf16 in, fp32 accumulate, bf16 out): the owner-side half of the fused GEMM + reduce-scatter
__global__ void sum_slots_kernel(const __nv_bfloat16* __restrict__ slots, __nv_bfloat16* __restrict__ out, long long n, int nslots) {
  const long long i8 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  if (i8 >= n) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < nslots; ++s) {
    const uint4 v = *reinterpret_cast<const uint4*>(slots + (long long)s * n + i8);
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(p2[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  }
  uint4 r;
  r.x = pack_bf16x2(acc[0], acc[1]); r.y = pack_bf16x2(acc[2], acc[3]); r.z = pack_bf16x2(acc[4], acc[5]); r.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + i8) = r;
}
}  // namespace smx

extern "C" {

int smx_rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, long long rows, int H, float eps,
                    void* stream) {
  if (H % 8) return -20;
  const int grid = (int)(rows < 148 * 8 ? rows : 148 * 8);
#define SMX_NF(T, V) smx::rmsnorm_fwd_pipe_kernel<T, V, false><<<grid, T, 0, SMX_STREAM>>>((const bf16*)x, nullptr, (const bf16*)gamma, nullptr, (bf16*)y, rstd, rows, H, eps)
  if (H <= 2048) SMX_NF(256, 1);
  else if (H <= 4096) SMX_NF(256, 2);
  else if (H <= 8192) SMX_NF(512, 2);
  else if (H <= 16384) SMX_NF(512, 4);
  else return -21;
#undef SMX_NF
  return (int)cudaGetLastError();
}

// slots: [nslots][rows, H] bf16 (this rank's staging buffer).  residual / sum_out / gamma (+ normed, rstd) / counter optional.
int smx_rs_consume(const void* slots, int nslots, long long rows, int H, const void* residual, void* sum_out,
                   const void* gamma, void* normed, float* rstd, float eps, const void* counter, unsigned target,
                   void* stream) {
  if (H % 8) return -20;
  if (gamma && !normed) return -22;
  if (nslots == 1 && counter == nullptr && residual && sum_out && gamma) {      // plain fused residual add + RMSNorm
    const int grid = (int)(rows < 148 * 8 ? rows : 148 * 8);
#define SMX_AN(T, V) smx::rmsnorm_fwd_pipe_kernel<T, V, true><<<grid, T, 0, SMX_STREAM>>>((const bf16*)residual, (const bf16*)slots, (const bf16*)gamma, (bf16*)sum_out, (bf16*)normed, rstd, rows, H, eps)
    if (H <= 2048) SMX_AN(256, 1);
    else if (H <= 4096) SMX_AN(256, 2);
    else if (H <= 8192) SMX_AN(512, 2);
    else if (H <= 16384) SMX_AN(512, 4);
    else return -21;
#undef SMX_AN
    return (int)cudaGetLastError();
  }
  const long long ss = rows * H;
#define SMX_RSC(T, V) smx::rs_consume_kernel<T, V><<<(unsigned)rows, T, 0, SMX_STREAM>>>((const bf16*)slots, nslots, ss, \
    (const bf16*)residual, (bf16*)sum_out, (const bf16*)gamma, (bf16*)normed, rstd, H, eps, (const uint32_t*)counter, target)
  if (H <= 2048) SMX_RSC(256, 1);
  else if (H <= 4096) SMX_RSC(256, 2);
  else if (H <= 8192) SMX_RSC(512, 2);
  else if (H <= 16384) SMX_RSC(512, 4);
  else return -21;
#undef SMX_RSC
  return (int)cudaGetLastError();
}

// scratch: optional [148 * 4, H] fp32 for the per-CTA dgamma partials (nullptr: atomics straight into dgamma)
int smx_rmsnorm_bwd(const void* x, const void* gamma, const void* dy, const float* rstd, void* dx, float* dgamma,
                    long long rows, int H, int add_to_dx, float* scratch, void* stream) {
  if (H % 8) return -20;
  int grid = rows < 148 * 4 ? (int)rows : 148 * 4;
  if (H > 8192) scratch = nullptr;
#define SMX_NB(T, V) smx::rmsnorm_bwd_pipe_kernel<T, V><<<grid, T, 0, SMX_STREAM>>>((const bf16*)x, (const bf16*)gamma, (const bf16*)dy, rstd, (bf16*)dx, dgamma, rows, H, add_to_dx, scratch)
  if (H <= 2048) SMX_NB(256, 1);
  else if (H <= 4096) SMX_NB(256, 2);
  else if (H <= 8192) SMX_NB(512, 2);
  else if (H <= 16384)      // 4 x 3 prefetched vectors per thread would spill: the non-pipelined kernel keeps this width
    smx::rmsnorm_bwd_kernel<512, 4><<<grid, 512, 0, SMX_STREAM>>>((const bf16*)x, (const bf16*)gamma, (const bf16*)dy, rstd, (bf16*)dx, dgamma, (int)rows, H, add_to_dx);
  else return -21;
#undef SMX_NB
  if (scratch) smx::dgamma_reduce_kernel<<<(H + 255) / 256, 256, 0, SMX_STREAM>>>(scratch, grid, H, dgamma);
  return (int)cudaGetLastError();
}

int smx_swiglu_fwd(const void* x, void* y, long long rows, int F, void* stream) {
  if (F % 8) return -20;
  smx::swiglu_fwd_kernel<<<grid_for(rows * (F / 8), 256), 256, 0, SMX_STREAM>>>((const bf16*)x, (bf16*)y, rows, F);
  return (int)cudaGetLastError();
}
int smx_swiglu_bwd(const void* x, const void* dy, void* dx, long long rows, int F, void* stream) {
  if (F % 8) return -20;
  smx::swiglu_bwd_kernel<<<grid_for(rows * (F / 8), 256), 256, 0, SMX_STREAM>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, rows, F);
  return (int)cudaGetLastError();
}

// sign = +1 forward rotation, -1 inverse rotation (= backward of RoPE)
int smx_rope(void* qkv, const float* cos_t, const float* sin_t, long long tokens, int seq_len, int row_stride,
             int n_rot_heads, int d, float sign, void* stream) {
  if ((d / 2) % 8) return -20;
  smx::rope_kernel<<<grid_for(tokens * n_rot_heads * (d / 16), 256), 256, 0, SMX_STREAM>>>((bf16*)qkv, cos_t, sin_t, tokens, seq_len, row_stride, n_rot_heads, d, sign, d);
  return (int)cudaGetLastError();
}
// general form: ``n_rot_heads`` rotary slices of width d per row, ``head_stride`` elements apart, starting at ``base`` (MLA: the last 64
// of every 192-wide q head; the 64-wide shared key slice at the end of the kv_down output)
int smx_rope_strided(void* base, const float* cos_t, const float* sin_t, long long tokens, int seq_len, long long row_stride,
                     int n_rot_heads, int d, int head_stride, float sign, void* stream) {
  if ((d / 2) % 8 || head_stride % 8 || row_stride % 8) return -20;
  smx::rope_kernel<<<grid_for(tokens * n_rot_heads * (d / 16), 256), 256, 0, SMX_STREAM>>>((bf16*)base, cos_t, sin_t, tokens, seq_len, row_stride, n_rot_heads, d, sign, head_stride);
  return (int)cudaGetLastError();
}

int smx_cross_entropy(void* logits, const long long* targets, float* loss, long long rows, int V, long long ld,
                      float grad_scale, int write_grad, void* stream) {
  if (ld % 8) return -20;
  smx::cross_entropy_fwd_bwd_kernel<1024><<<(unsigned)rows, 1024, 0, SMX_STREAM>>>((bf16*)logits, targets, loss, V, ld, grad_scale, write_grad);
  return (int)cudaGetLastError();
}

int smx_ce_vp_stats(const void* logits, const long long* targets, void* stats, long long rows, int V, long long ld,
                    long long v0, void* stream) {
  if (V % 8 || ld % 8) return -20;
  smx::ce_vp_stats_kernel<1024><<<(unsigned)rows, 1024, 0, SMX_STREAM>>>((const bf16*)logits, targets, (float4*)stats, V, ld, v0);
  return (int)cudaGetLastError();
}
int smx_ce_vp_grad(void* logits, const long long* targets, const void* all_stats, int world, float* loss, long long rows, int V,
                   long long ld, long long v0, float grad_scale, void* stream) {
  if (V % 8 || ld % 8) return -20;
  smx::ce_vp_grad_kernel<1024><<<(unsigned)rows, 1024, 0, SMX_STREAM>>>((bf16*)logits, targets, (const float4*)all_stats, world, rows,
                                                                      loss, V, ld, v0, grad_scale);
  return (int)cudaGetLastError();
}

int smx_sqnorm(const float* g, long long n, float* out, void* stream) {
  smx::sqnorm_kernel<<<grid_for(n / 4, 256, 148 * 8), 256, 0, SMX_STREAM>>>(g, n, out);
  return (int)cudaGetLastError();
}
int smx_adamw(float* master, float* m, float* v, float* grad, void* param, long long n, float lr, float beta1,
              float beta2, float eps, float wd, float bc1, float bc2, const float* sqnorm, float max_norm,
              float grad_prescale, int zero_grad, void* stream) {
  if (n % 4) return -20;
  smx::adamw_kernel<<<grid_for(n / 4, 256, 148 * 8), 256, 0, SMX_STREAM>>>(master, m, v, grad, (bf16*)param, n, lr, beta1, beta2, eps, wd, bc1, bc2, sqnorm, max_norm, grad_prescale, zero_grad);
  return (int)cudaGetLastError();
}

int smx_permute_rows(const void* src, void* dst, const int* row_map, long long n_rows, int H, void* stream) {
  if (H % 8) return -20;
  smx::permute_rows_kernel<<<grid_for(n_rows * (H / 8), 256), 256, 0, SMX_STREAM>>>((const bf16*)src, (bf16*)dst, row_map, n_rows, H);
  return (int)cudaGetLastError();
}
int smx_unpermute_combine(const void* x, void* out, const int* inv_map, const float* probs, long long tokens, int topk,
                          int H, void* stream) {
  if (H % 8) return -20;
  smx::unpermute_combine_kernel<<<grid_for(tokens * (H / 8), 256), 256, 0, SMX_STREAM>>>((const bf16*)x, (bf16*)out, inv_map, probs, tokens, topk, H);
  return (int)cudaGetLastError();
}
int smx_moe_prob_grad(const void* dy, const void* y_perm, const int* inv_map, float* d_probs, long long tokens, int topk, int H,
                      void* stream) {
  if (H % 8) return -20;
  smx::moe_prob_grad_kernel<<<grid_for(tokens * topk * 32, 256), 256, 0, SMX_STREAM>>>((const bf16*)dy, (const bf16*)y_perm, inv_map, d_probs,
                                                                                      tokens * topk, topk, H);
  return (int)cudaGetLastError();
}
int smx_embedding_fwd(const long long* tok, const void* table, void* out, long long rows, int H, long long v0, long long V, void* stream) {
  if (H % 8) return -20;
  smx::embedding_fwd_kernel<<<grid_for(rows, 8), 256, 0, SMX_STREAM>>>(tok, (const bf16*)table, (bf16*)out, rows, H, v0, V);
  return (int)cudaGetLastError();
}
int smx_embedding_bwd(const long long* tok, const void* dy, float* grad, long long rows, int H, long long v0, long long V, void* stream) {
  if (H % 8) return -20;
  smx::embedding_bwd_kernel<<<grid_for(rows, 8), 256, 0, SMX_STREAM>>>(tok, (const bf16*)dy, grad, rows, H, v0, V);
  return (int)cudaGetLastError();
}
int smx_mla_k_assemble(const void* kv, long long kv_row, int kv_head, const void* krope, long long kr_row, void* k, long long tokens,
                       int heads, int dn, int dr, void* stream) {
  if (dn % 8 || dr % 8 || kv_row % 8 || kv_head % 8 || kr_row % 8) return -20;
  smx::mla_k_assemble_kernel<<<grid_for(tokens * heads * ((dn + dr) / 8), 256), 256, 0, SMX_STREAM>>>((const bf16*)kv, kv_row, kv_head, (const bf16*)krope, kr_row, (bf16*)k, tokens, heads, dn, dr);
  return (int)cudaGetLastError();
}
int smx_mla_k_split_bwd(const void* dk, void* d_kv, long long kv_row, int kv_head, void* d_krope, long long kr_row, long long tokens,
                        int heads, int dn, int dr, void* stream) {
  if (dn % 8 || dr % 8 || kv_row % 8 || kv_head % 8 || kr_row % 8 || (dn + dr) / 8 > 32) return -20;
  smx::mla_k_split_bwd_kernel<<<grid_for(tokens, 8), 256, 0, SMX_STREAM>>>((const bf16*)dk, (bf16*)d_kv, kv_row, kv_head, (bf16*)d_krope, kr_row, tokens, heads, dn, dr);
  return (int)cudaGetLastError();
}
int smx_copy2d(const void* src, long long src_ld_bytes, void* dst, long long dst_ld_bytes, long long rows, long long width_bytes, void* stream) {
  if (width_bytes % 16 || src_ld_bytes % 16 || dst_ld_bytes % 16) return -20;
  smx::copy2d_kernel<<<grid_for(rows * (width_bytes / 16), 256), 256, 0, SMX_STREAM>>>((const unsigned char*)src, src_ld_bytes, (unsigned char*)dst, dst_ld_bytes, rows, (int)(width_bytes / 16));
  return (int)cudaGetLastError();
}
int smx_transpose01(const void* src, void* dst, long long A, long long B, long long inner_bytes, void* stream) {
  if (inner_bytes % 16) return -20;
  smx::transpose01_kernel<<<grid_for(A * B * (inner_bytes / 16), 256), 256, 0, SMX_STREAM>>>((const uint4*)src, (uint4*)dst, A, B, (int)(inner_bytes / 16));
  return (int)cudaGetLastError();
}
int smx_add(const void* a, const void* b, void* y, long long n, void* stream) {
  if (n % 8) return -20;
  smx::add_kernel<<<grid_for(n / 8, 256), 256, 0, SMX_STREAM>>>((const bf16*)a, (const bf16*)b, (bf16*)y, n);
  return (int)cudaGetLastError();
}
int smx_copy(const void* src, void* dst, long long nbytes, void* stream) {
  if (nbytes % 16) return -20;
  smx::copy_kernel<<<grid_for(nbytes / 16, 256, 148 * 8), 256, 0, SMX_STREAM>>>((const uint4*)src, (uint4*)dst, nbytes / 16);
  return (int)cudaGetLastError();
}

int smx_sum_slots(const void* slots, void* out, long long n, int nslots, void* stream) {
  if (n % 8) return -50;
  const long long threads = n / 8;
  smx::sum_slots_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(slots), reinterpret_cast<__nv_bfloat16*>(out), n, nslots);
  return (int)cudaGetLastError();
}

}  // extern "C"
