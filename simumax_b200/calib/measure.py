"""Single-GPU operator sweeps: time this repo's sm_100a kernels for every shape the model asks about and turn the
times into the efficiency numbers the system config stores.

    efficiency(key) = model_flops(key) / (measured_seconds * nominal_tflops)          compute ops
    efficient_factor(op) = bytes_moved / (measured_seconds * nominal_gbps)             memory-bound ops

Timing rules (B200_PROFILING.md): CUDA events on the launching stream, warm-up first, operands rotated through a ring
of buffers larger than L2 (126 MB) so no iteration reads what the previous one left in cache.

(reference: simu_tools/efficency_test/test_{gemm,grouped_gemm,fa,ce,permutation}_efficiency.py — those time cuBLAS /
TransformerEngine / cuDNN; here the timed kernels are csrc/gemm_tcgen05.cu, csrc/fmha_{fwd,bwd}.cu, csrc/elementwise.cu.)
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

from .shapes import ShapeRequest

L2_BYTES = 126 * 1024 * 1024


SUSTAINED_MS = float(__import__("os").environ.get("SMX_CALIB_SUSTAINED_MS", "0"))
LIBRARY_COLUMN = False          # CLI ``ops --library-column``: time cuBLAS (torch.matmul) on the same operands next to every bf16 GEMM key


def _time(fn: Callable[[int], None], iters: int = 8, warmup: int = 3) -> float:
    """Median-of-groups device time per call in ms. ``fn(i)`` receives the iteration index (to rotate buffers).

    Sustained mode (``SUSTAINED_MS`` > 0, CLI ``ops --sustained-ms``): a training step keeps the GPU at its power cap (SM clock
    1.4-1.6 GHz on this box) while a handful of back-to-back launches run at the 1.9 GHz boost clock, so burst-calibrated GEMM
    efficiencies make the model optimistic.  In this mode the op is first run for ``SUSTAINED_MS`` of device time to bring the
    part to its steady clock and every timed group lasts at least a third of that."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    base = 0
    if SUSTAINED_MS > 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        per_call = max(e0.elapsed_time(e1) / iters, 1e-3)
        heat = int(math.ceil(SUSTAINED_MS / per_call))
        for i in range(heat):
            fn(i)
        iters = max(iters, int(math.ceil(SUSTAINED_MS / 3.0 / per_call)))
        base = heat
    times = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(base + rep * iters + i)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / iters)
    return sorted(times)[1]


def _ring(nbytes_per_set: int) -> int:
    """How many operand sets are needed so that consecutive iterations never hit L2-resident data."""
    return max(2, min(16, int(math.ceil(2.0 * L2_BYTES / max(nbytes_per_set, 1))) + 1))


def _rand(shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device="cuda", dtype=torch.float32) * scale).to(dtype)


# ------------------------------------------------------------------------------------------------------------------
def measure_matmul(req: ShapeRequest, fp8: bool = False) -> Dict[str, float]:
    """key: b, m, k, n, layout, accumulate, out_dtype  (out[m,n]; TN: x[m,k] w[n,k]; NN: dy[m,k] w[k,n]; NT: dy[k,m] x[k,n])."""
    from ..ops.gemm import gemm
    b, m, k, n = (req.geti(x) for x in ("b", "m", "k", "n"))
    layout = req.fields["layout"]
    acc = req.getb("accumulate")
    out_dtype = torch.float32 if req.fields.get("out_dtype", "bf16") == "fp32" else torch.bfloat16
    m_eff = b * m if layout in ("TN", "NN") else m          # batch folds into rows for activations
    k_eff = k if layout in ("TN", "NN") else b * k
    dt = torch.float8_e4m3fn if fp8 else torch.bfloat16
    if fp8:
        # fp8 tensor-core operands must be K-major: dgrad / wgrad multiply pre-transposed fp8 copies (what TE's
        # Float8Tensor keeps around); the transposition itself belongs to the quantiser op, not to the GEMM key
        layout = "TN"
    a_shape = {"TN": (m_eff, k_eff), "NN": (m_eff, k_eff), "NT": (k_eff, m_eff), "TT": (k_eff, m_eff)}[layout]
    b_shape = {"TN": (n, k_eff), "NN": (k_eff, n), "NT": (k_eff, n), "TT": (n, k_eff)}[layout]
    per_set = (a_shape[0] * a_shape[1] + b_shape[0] * b_shape[1]) * (1 if fp8 else 2) + m_eff * n * out_dtype.itemsize
    ring = _ring(per_set)
    As = [_rand(a_shape, dt, 0.5) for _ in range(ring)]
    Bs = [_rand(b_shape, dt, 0.5) for _ in range(ring)]
    Cs = [torch.zeros(m_eff, n, device="cuda", dtype=out_dtype) for _ in range(ring)]
    ms = _time(lambda i: gemm(As[i % ring], Bs[i % ring], layout, out=Cs[i % ring], accumulate=acc))
    flops = 2.0 * m_eff * n * k_eff
    res = {"ms": ms, "tflops": flops / ms / 1e9, "flops": flops}
    if LIBRARY_COLUMN and not fp8:
        lib_ms = _time(lambda i: _library_gemm(As[i % ring], Bs[i % ring], layout, Cs[i % ring], acc))
        res.update(cublas_ms=lib_ms, cublas_tflops=flops / lib_ms / 1e9, vs_cublas=lib_ms / ms)
    return res


def _library_gemm(a, b, layout, c, accumulate):
    """The same product through cuBLAS as stock PyTorch issues it (the bar the sweep reports against): bf16 outputs write ``c``
    directly; fp32 outputs use the mixed-precision ``mm(out_dtype=fp32)`` and, when accumulating, one ``add_`` (what a library
    main-grad wgrad without a fused accumulate costs)."""
    x = a.t() if layout in ("NT", "TT") else a
    y = b.t() if layout in ("TN", "TT") else b
    if c.dtype == torch.bfloat16 and not accumulate:
        return torch.matmul(x, y, out=c)
    if c.dtype == torch.bfloat16:
        return c.addmm_(x, y)
    t = torch.mm(x, y, out_dtype=torch.float32)
    return c.add_(t) if accumulate else c.copy_(t)


def measure_matmul_mxfp8(req: ShapeRequest) -> Dict[str, float]:
    """Same keys as ``fp8_matmul`` through the block-scaled kernel (e4m3 payload + UE8M0 scale per 32 K elements, scaling inside
    ``tcgen05.mma.kind::mxf8f6f4.block_scale``).  Operands are K-major, so all three layouts run as TN on pre-transposed copies;
    fp32 accumulate-into-C is not in this kernel (a main-grad wgrad adds one fp32 read-modify-write, charged by the caller)."""
    from ..ops.gemm import gemm_mxfp8
    b, m, k, n = (req.geti(x) for x in ("b", "m", "k", "n"))
    layout = req.fields["layout"]
    out_dtype = torch.float32 if req.fields.get("out_dtype", "bf16") == "fp32" else torch.bfloat16
    m_eff = b * m if layout in ("TN", "NN") else m
    k_eff = k if layout in ("TN", "NN") else b * k
    if k_eff % 128:
        raise NotImplementedError(f"mxfp8 kernel needs K % 128 == 0, got {k_eff}")
    per_set = (m_eff + n) * k_eff * 33 // 32 + m_eff * n * out_dtype.itemsize
    ring = _ring(per_set)
    As = [_rand((m_eff, k_eff), torch.float8_e4m3fn, 0.5) for _ in range(ring)]
    Bs = [_rand((n, k_eff), torch.float8_e4m3fn, 0.5) for _ in range(ring)]
    SAs = [torch.full((m_eff, k_eff // 32), 127, device="cuda", dtype=torch.uint8) for _ in range(ring)]
    SBs = [torch.full((n, k_eff // 32), 127, device="cuda", dtype=torch.uint8) for _ in range(ring)]
    Cs = [torch.zeros(m_eff, n, device="cuda", dtype=out_dtype) for _ in range(ring)]
    ms = _time(lambda i: gemm_mxfp8(As[i % ring], SAs[i % ring], Bs[i % ring], SBs[i % ring], out=Cs[i % ring]))
    flops = 2.0 * m_eff * n * k_eff
    return {"ms": ms, "tflops": flops / ms / 1e9, "flops": flops}


def measure_group_matmul(req: ShapeRequest, fp8: bool = False) -> Dict[str, float]:
    """key: ng, M (tokens per expert), N, K, stage in {fwd, bwd_grad_act, bwd_grad_w}."""
    from ..ops.gemm import gemm
    ng, M, N, K = (req.geti(x) for x in ("ng", "M", "N", "K"))
    stage = req.fields["stage"]
    dt = torch.float8_e4m3fn if fp8 else torch.bfloat16
    if fp8 and stage != "fwd":     # K-major operands only (see measure_matmul): same flops through the TN kernel
        if stage == "bwd_grad_act":
            M, N, K, stage = M, K, N, "fwd"                 # dx[g, M, K] = dy[g, M, N] w^T[g, K, N]^T
        else:
            ng, M, N, K = ng, N, K, M                       # dw[g, N, K] = dy^T[g, N, M] x^T[g, K, M]^T
            stage = "fwd_f32"
    if stage == "fwd_f32":
        a_shape, b_shape, layout, out_shape, out_dtype, acc = (ng, M, K), (ng, N, K), "TN", (ng, M, N), torch.float32, True
    elif stage == "fwd":            # y[g, M, N] = x[g, M, K] w[g, N, K]^T
        a_shape, b_shape, layout, out_shape, out_dtype, acc = (ng, M, K), (ng, N, K), "TN", (ng, M, N), torch.bfloat16, False
    elif stage == "bwd_grad_act":  # dx[g, M, K] = dy[g, M, N] w[g, N, K]
        a_shape, b_shape, layout, out_shape, out_dtype, acc = (ng, M, N), (ng, N, K), "NN", (ng, M, K), torch.bfloat16, False
    else:                          # dw[g, N, K] += dy[g, M, N]^T x[g, M, K]   (fp32 main grad)
        a_shape, b_shape, layout, out_shape, out_dtype, acc = (ng, M, N), (ng, M, K), "NT", (ng, N, K), torch.float32, True
    numel = lambda s: s[0] * s[1] * s[2]
    per_set = (numel(a_shape) + numel(b_shape)) * (1 if fp8 else 2) + numel(out_shape) * out_dtype.itemsize
    ring = _ring(per_set)
    As = [_rand(a_shape, dt, 0.5) for _ in range(ring)]
    Bs = [_rand(b_shape, dt, 0.5) for _ in range(ring)]
    Cs = [torch.zeros(*out_shape, device="cuda", dtype=out_dtype) for _ in range(ring)]
    ms = _time(lambda i: gemm(As[i % ring], Bs[i % ring], layout, out=Cs[i % ring], accumulate=acc))
    flops = 2.0 * ng * M * N * K
    return {"ms": ms, "tflops": flops / ms / 1e9, "flops": flops}


def measure_sdp(req: ShapeRequest, backward: bool) -> Dict[str, float]:
    """key: batch, seq_len, head_num, kv_head_num, qk_head_dim, v_head_dim, qkv_contiguous (causal, as Megatron trains)."""
    from ..ops.attention import attention_flops, fmha_bwd, fmha_fwd
    B, S, H, Hk, Dqk, Dv = (req.geti(x) for x in ("batch", "seq_len", "head_num", "kv_head_num", "qk_head_dim", "v_head_dim"))
    packed = req.getb("qkv_contiguous") and Dqk == Dv
    ring = _ring(B * S * (H * (Dqk + Dv) + Hk * (Dqk + Dv)) * 2)
    sets = []
    for _ in range(ring):
        if packed:
            qkv = _rand((B, S, (H + 2 * Hk) * Dqk))
            q = qkv[..., : H * Dqk].view(B, S, H, Dqk)
            k = qkv[..., H * Dqk:(H + Hk) * Dqk].view(B, S, Hk, Dqk)
            v = qkv[..., (H + Hk) * Dqk:].view(B, S, Hk, Dv)
        else:
            q, k, v = _rand((B, S, H, Dqk)), _rand((B, S, Hk, Dqk)), _rand((B, S, Hk, Dv))
        o = torch.empty(B, S, H, Dv, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B, H, S, device="cuda", dtype=torch.float32)
        sets.append([q, k, v, o, lse])
    if not backward:
        ms = _time(lambda i: fmha_fwd(*sets[i % ring][:3], causal=True, out=sets[i % ring][3], lse=sets[i % ring][4]))
    else:
        if Dqk not in (128, 192) or Dv != 128:
            raise NotImplementedError("fmha_bwd supports qk head dim 128 / 192 with v head dim 128")
        for s in sets:
            fmha_fwd(s[0], s[1], s[2], causal=True, out=s[3], lse=s[4])
            s.append(_rand((B, S, H, Dv)))
            s.extend([torch.empty_like(s[0]), torch.empty_like(s[1]), torch.empty_like(s[2])])
        ms = _time(lambda i: fmha_bwd(sets[i % ring][5], *sets[i % ring][:5], *sets[i % ring][6:9], causal=True))
    useful = attention_flops(B, S, S, H, Dqk, Dv, True, backward)
    return {"ms": ms, "tflops": useful / ms / 1e9, "flops": useful}


# ------------------------------------------------------------------------------------------------------------------
def measure_bandwidth_ops(hidden: int = 4096, tokens: int = 16384, vocab: int = 128256, topk: int = 6) -> Dict[str, Dict[str, float]]:
    """Achieved GB/s of the memory-bound kernels the model prices by bytes: ``default`` (elementwise streams),
    ``permute_fwd`` / ``permute_bwd`` (MoE dispatch / combine), ``ce`` and ``ce_fusion`` (vocab-parallel cross-entropy)."""
    from ..ops import elementwise as E
    res = {}
    # default: residual add (2 reads + 1 write) and RMSNorm (1 read + 1 write) streams
    ring = _ring(3 * tokens * hidden * 2)
    xs = [_rand((tokens, hidden)) for _ in range(ring)]
    ys = [_rand((tokens, hidden)) for _ in range(ring)]
    zs = [torch.empty_like(xs[0]) for _ in range(ring)]
    ms = _time(lambda i: E.add(xs[i % ring], ys[i % ring], out=zs[i % ring]))
    add_gbs = 3 * tokens * hidden * 2 / ms / 1e6
    gamma = torch.ones(hidden, device="cuda", dtype=torch.bfloat16)
    ms = _time(lambda i: E.rmsnorm_fwd(xs[i % ring], gamma, out=zs[i % ring]))
    norm_gbs = 2 * tokens * hidden * 2 / ms / 1e6
    ms = _time(lambda i: E.swiglu_fwd(xs[i % ring], out=zs[i % ring][:, : hidden // 2]))
    res["default"] = {"gbs": min(add_gbs, norm_gbs), "add_gbs": add_gbs, "rmsnorm_gbs": norm_gbs,
                      "swiglu_gbs": 1.5 * tokens * hidden * 2 / ms / 1e6}
    # RMSNorm backward (x, dy read; dx written; dgamma through per-CTA partial rows + a reduce kernel): register-limited, its own entry
    rstd = torch.rand(tokens, device="cuda") + 0.5
    dgam = torch.zeros(hidden, device="cuda")
    ms = _time(lambda i: E.rmsnorm_bwd(xs[i % ring], gamma, ys[i % ring], rstd, dgam, dx=zs[i % ring]))
    res["norm_bwd"] = {"gbs": 3 * tokens * hidden * 2 / ms / 1e6, "ms": ms}
    # permute: gather rows (token -> expert-sorted copy, topk expansion); combine: weighted gather-sum back
    n_out = tokens * topk
    row_map = torch.randint(0, tokens, (n_out,), device="cuda", dtype=torch.int32)
    dst = [torch.empty(n_out, hidden, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
    ms = _time(lambda i: E.permute_rows(xs[i % ring], row_map, out=dst[i % 2]))
    res["permute_fwd"] = {"gbs": (n_out * hidden * 2 * 2) / ms / 1e6, "ms": ms}
    inv = torch.randperm(n_out, device="cuda", dtype=torch.int32).view(tokens, topk).contiguous()
    probs = torch.rand(tokens, topk, device="cuda", dtype=torch.float32)
    ms = _time(lambda i: E.unpermute_combine(dst[i % 2], inv, probs, tokens, topk, out=zs[i % ring]))
    res["permute_bwd"] = {"gbs": (n_out * hidden * 2 + tokens * hidden * 2) / ms / 1e6, "ms": ms}
    # cross-entropy: fused loss + in-place gradient over [rows, vocab] bf16 logits
    rows = 4096
    lg = [_rand((rows, vocab)) for _ in range(2)]
    tgt = torch.randint(0, vocab, (rows,), device="cuda")
    ms = _time(lambda i: E.cross_entropy_(lg[i % 2], tgt, grad_scale=1.0 / rows))
    res["ce_fusion"] = {"gbs": 2 * rows * vocab * 2 / ms / 1e6, "ms": ms}
    ms = _time(lambda i: E.cross_entropy_(lg[i % 2], tgt, write_grad=False))
    res["ce"] = {"gbs": rows * vocab * 2 / ms / 1e6, "ms": ms}
    del lg, xs, ys, zs, dst
    # optimizer: grad-norm pass (4 B / parameter) + fused AdamW (reads grad + master / m / v, writes master / m / v + the bf16
    # parameter, clears the grad: 34 B / parameter) -- long pure streams, faster than the short elementwise kernels above
    n = 128 << 20
    master, m_, v_, grad = (torch.zeros(n, device="cuda") for _ in range(4))
    param = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    sq = torch.ones(1, device="cuda")
    step = [0]

    def opt(i):
        step[0] += 1
        E.sqnorm(grad, out=sq)
        E.adamw_(master, m_, v_, grad, param, 1e-4, 0.9, 0.95, 1e-8, 0.1, step[0], sqnorm_t=sq, max_norm=1.0, zero_grad=True)
    ms = _time(opt)
    res["optimizer"] = {"gbs": 38.0 * n / ms / 1e6, "ms": ms}
    return res


def measure_launch_latency_us(n: int = 256) -> float:
    """Device-side fixed cost of a memory-bound kernel: ``n`` back-to-back launches of the residual add on an 8 KiB tensor (fits in
    L1 / L2) captured in a CUDA graph and replayed, so the host launch rate of a Python loop (6-7 us per launch here) is out of the
    picture and what remains is launch latency + prologue + tail of one kernel.  This is the ``latency_us`` the model adds to every
    bandwidth-priced op (the reference hard-codes 30 us for its library kernels)."""
    from ..ops import elementwise as E
    x = torch.zeros(4096, device="cuda", dtype=torch.bfloat16)
    y = torch.zeros_like(x)
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for _ in range(8):
            E.add(x, y, out=y)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(n):
                E.add(x, y, out=y)
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


# ------------------------------------------------------------------------------------------------------------------
def sweep(requests: Dict[str, Dict[str, ShapeRequest]], nominal_tflops: Dict[str, float], log=print,
          limit_per_op: Optional[int] = None) -> Dict[str, Dict[str, dict]]:
    """Measure every request; returns op -> key -> {ms, tflops, efficiency}."""
    runners = {
        "matmul": lambda r: measure_matmul(r, False), "fp8_matmul": lambda r: measure_matmul(r, True),
        "mxfp8_matmul": measure_matmul_mxfp8,
        "group_matmul": lambda r: measure_group_matmul(r, False), "fp8_group_matmul": lambda r: measure_group_matmul(r, True),
        "sdp_fwd": lambda r: measure_sdp(r, False), "sdp_bwd": lambda r: measure_sdp(r, True),
    }
    out: Dict[str, Dict[str, dict]] = {}
    for op, table in requests.items():
        run = runners.get(op)
        if run is None:
            log(f"[calib] no kernel for op '{op}' ({len(table)} keys) - left at the default efficiency")
            continue
        for n, (key, req) in enumerate(table.items()):
            if limit_per_op is not None and n >= limit_per_op:
                log(f"[calib] {op}: stopped after {limit_per_op} of {len(table)} keys (limit_per_op)")
                break
            try:
                m = run(req)
            except (NotImplementedError, RuntimeError, ValueError, TypeError, AssertionError) as e:
                log(f"[calib] {op} {key}: skipped ({e})")
                continue
            finally:
                torch.cuda.empty_cache()
            m["efficiency"] = req.flops / (m["ms"] * 1e-3) / (nominal_tflops[op] * 1e12)
            out.setdefault(op, {})[key] = m
            log(f"[calib] {op:13s} {m['ms']:8.4f} ms  {m['tflops']:7.1f} TF  eff {m['efficiency']:.4f}  {key}")
    return out
