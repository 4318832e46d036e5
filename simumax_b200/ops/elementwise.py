"""Python front-ends of the memory-bound sm_100a kernels (csrc/elementwise.cu).

Every wrapper validates shapes/dtypes, launches on the current CUDA stream and counts the launch
(``_lib.launches()`` feeds bench.py's ``gpu_launches``).  There is no PyTorch fallback: on a GPU box a
missing kernel library is an error.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

_c = ctypes
_SIGS = {
    "smx_rmsnorm_fwd": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_float, _c.c_void_p],
    "smx_rmsnorm_bwd": [_c.c_void_p] * 6 + [_c.c_longlong, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p],
    "smx_swiglu_fwd": [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_void_p],
    "smx_swiglu_bwd": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_void_p],
    "smx_rope": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_void_p],
    "smx_cross_entropy": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_longlong, _c.c_float, _c.c_int, _c.c_void_p],
    "smx_ce_vp_stats": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_void_p],
    "smx_ce_vp_grad": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_longlong, _c.c_longlong,
                       _c.c_float, _c.c_void_p],
    "smx_sqnorm": [_c.c_void_p, _c.c_longlong, _c.c_void_p, _c.c_void_p],
    "smx_adamw": [_c.c_void_p] * 5 + [_c.c_longlong] + [_c.c_float] * 7 + [_c.c_void_p, _c.c_float, _c.c_float, _c.c_int, _c.c_void_p],
    "smx_permute_rows": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_void_p],
    "smx_unpermute_combine": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_void_p],
    "smx_moe_prob_grad": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_void_p],
    "smx_add": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_void_p],
    "smx_rope_strided": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_void_p],
    "smx_mla_k_assemble": [_c.c_void_p, _c.c_longlong, _c.c_int, _c.c_void_p, _c.c_longlong, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p],
    "smx_mla_k_split_bwd": [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_void_p, _c.c_longlong, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p],
    "smx_transpose01": [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_longlong, _c.c_longlong, _c.c_void_p],
    "smx_copy2d": [_c.c_void_p, _c.c_longlong, _c.c_void_p, _c.c_longlong, _c.c_longlong, _c.c_longlong, _c.c_void_p],
    "smx_embedding_fwd": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_void_p],
    "smx_embedding_bwd": [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_void_p],
    "smx_rs_consume": [_c.c_void_p, _c.c_int, _c.c_longlong, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                       _c.c_float, _c.c_void_p, _c.c_uint, _c.c_void_p],
    "smx_copy": [_c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_void_p],
}
_fns = {}
_DGAMMA_SCRATCH = {}


def _fn(name):
    f = _fns.get(name)
    if f is None:
        f = getattr(_lib.lib(), name)
        f.restype = _c.c_int
        f.argtypes = _SIGS[name]
        _fns[name] = f
    return f


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _call(name, *args):
    rc = _fn(name)(*args)
    _lib.check(rc, name)
    _lib.count_launch()


def _bf16c(t: torch.Tensor, what: str):
    if t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise TypeError(f"{what} must be a contiguous bf16 tensor, got {t.dtype} contiguous={t.is_contiguous()}")


# ------------------------------------------------------------------------------------------------
def rmsnorm_fwd(x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None,
                rstd: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """y = x * rsqrt(mean(x^2) + eps) * gamma over the last dim; returns (y, rstd[rows] fp32)."""
    _bf16c(x, "x"); _bf16c(gamma, "gamma")
    H = x.shape[-1]
    rows = x.numel() // H
    y = torch.empty_like(x) if out is None else out
    if rstd is None:
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("smx_rmsnorm_fwd", x.data_ptr(), gamma.data_ptr(), y.data_ptr(), rstd.data_ptr(), rows, H, eps, _stream(x))
    return y, rstd


def rmsnorm_bwd(x, gamma, dy, rstd, dgamma_fp32: torch.Tensor, dx: Optional[torch.Tensor] = None,
                accumulate_dx: bool = False) -> torch.Tensor:
    """dx (optionally += into an existing residual gradient) and dgamma += (fp32 main-grad accumulate)."""
    _bf16c(x, "x"); _bf16c(dy, "dy")
    H = x.shape[-1]
    rows = x.numel() // H
    if dx is None:
        assert not accumulate_dx
        dx = torch.empty_like(x)
    assert dgamma_fp32.dtype == torch.float32 and dgamma_fp32.numel() == H
    key = (x.device, H)
    scratch = _DGAMMA_SCRATCH.get(key)
    if scratch is None:          # per-CTA dgamma partials, summed by a second tiny kernel (no atomics on H hot addresses)
        scratch = _DGAMMA_SCRATCH[key] = torch.empty(148 * 4 * H, device=x.device, dtype=torch.float32)
    _call("smx_rmsnorm_bwd", x.data_ptr(), gamma.data_ptr(), dy.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
          dgamma_fp32.data_ptr(), rows, H, 1 if accumulate_dx else 0, scratch.data_ptr(), _stream(x))
    _lib.count_launch(2 if H <= 8192 else 1)        # main kernel + dgamma reduce (csrc/elementwise.cu: no scratch path above H = 8192)
    return dx


def add_rmsnorm_fwd(residual: torch.Tensor, delta: torch.Tensor, gamma: torch.Tensor, eps: float, out_sum: torch.Tensor,
                    out_normed: torch.Tensor, rstd: torch.Tensor) -> None:
    """One pass: out_sum = residual + delta (rounded to bf16, what the backward re-reads);  out_normed = rmsnorm(out_sum) * gamma.
    (the owner-side kernel of the fused GEMM -> reduce-scatter with a single slot and no completion counter)"""
    _bf16c(residual, "residual"); _bf16c(delta, "delta"); _bf16c(gamma, "gamma")
    H = residual.shape[-1]
    rows = residual.numel() // H
    _call("smx_rs_consume", delta.data_ptr(), 1, rows, H, residual.data_ptr(), out_sum.data_ptr(), gamma.data_ptr(),
          out_normed.data_ptr(), rstd.data_ptr(), float(eps), None, 0, _stream(residual))


def swiglu_fwd(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x = [gate | up] along the last dim -> silu(gate) * up."""
    _bf16c(x, "x")
    F = x.shape[-1] // 2
    rows = x.numel() // (2 * F)
    y = torch.empty(*x.shape[:-1], F, device=x.device, dtype=x.dtype) if out is None else out
    _call("smx_swiglu_fwd", x.data_ptr(), y.data_ptr(), rows, F, _stream(x))
    return y


def swiglu_bwd(x: torch.Tensor, dy: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _bf16c(x, "x"); _bf16c(dy, "dy")
    F = x.shape[-1] // 2
    rows = x.numel() // (2 * F)
    dx = torch.empty_like(x) if out is None else out
    _call("smx_swiglu_bwd", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), rows, F, _stream(x))
    return dx


def rope_tables(seq_len: int, head_dim: int, theta: float = 500000.0, device="cuda"):
    half = head_dim // 2
    inv = 1.0 / (theta ** (torch.arange(0, half, device=device, dtype=torch.float32) / half))
    ang = torch.arange(seq_len, device=device, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().contiguous(), ang.sin().contiguous()


def rope_inplace(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq_len: int, n_rot_heads: int,
                 head_dim: int, inverse: bool = False) -> torch.Tensor:
    """Rotate the first ``n_rot_heads`` heads (q then k) of each packed [tokens, (nq+2nk)*d] row in place."""
    _bf16c(qkv, "qkv")
    row = qkv.shape[-1]
    tokens = qkv.numel() // row
    _call("smx_rope", qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), tokens, seq_len, row, n_rot_heads, head_dim,
          -1.0 if inverse else 1.0, _stream(qkv))
    return qkv


def cross_entropy_(logits: torch.Tensor, targets: torch.Tensor, grad_scale: float = 1.0, write_grad: bool = True
                   ) -> torch.Tensor:
    """Fused CE: returns per-row loss (fp32) and overwrites ``logits`` with d loss / d logits * grad_scale."""
    assert logits.dtype == torch.bfloat16 and logits.stride(-1) == 1 and logits.dim() == 2
    assert targets.dtype == torch.int64 and targets.is_contiguous()
    rows, V = logits.shape
    loss = torch.empty(rows, device=logits.device, dtype=torch.float32)
    _call("smx_cross_entropy", logits.data_ptr(), targets.data_ptr(), loss.data_ptr(), rows, V, logits.stride(0),
          float(grad_scale), 1 if write_grad else 0, _stream(logits))
    return loss


def ce_vocab_parallel_stats(logits: torch.Tensor, targets: torch.Tensor, v0: int, stats: torch.Tensor) -> torch.Tensor:
    """Phase 1 of the vocab-parallel CE: ``stats[row] = (max, sum exp(x - max), target logit | 0, has_target)`` of this rank's
    ``[rows, V_local]`` logits whose first column is vocabulary id ``v0`` (reference cost model: ParallelCE,
    simumax/core/transformer/dense_module.py:2097-2215)."""
    assert logits.dtype == torch.bfloat16 and logits.stride(-1) == 1 and logits.dim() == 2
    assert stats.dtype == torch.float32 and stats.numel() >= logits.shape[0] * 4 and stats.is_contiguous()
    rows, V = logits.shape
    _call("smx_ce_vp_stats", logits.data_ptr(), targets.data_ptr(), stats.data_ptr(), rows, V, logits.stride(0), int(v0), _stream(logits))
    return stats


def ce_vocab_parallel_grad_(logits: torch.Tensor, targets: torch.Tensor, v0: int, all_stats: torch.Tensor, world: int,
                            grad_scale: float = 1.0) -> torch.Tensor:
    """Phase 2: combine the gathered ``[world, rows, 4]`` statistics, return the per-row loss and overwrite the local logits
    with ``d loss / d logits * grad_scale``."""
    rows, V = logits.shape
    assert all_stats.dtype == torch.float32 and all_stats.numel() >= world * rows * 4
    loss = torch.empty(rows, device=logits.device, dtype=torch.float32)
    _call("smx_ce_vp_grad", logits.data_ptr(), targets.data_ptr(), all_stats.data_ptr(), world, loss.data_ptr(), rows, V,
          logits.stride(0), int(v0), float(grad_scale), _stream(logits))
    return loss


def sqnorm(g: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out += sum(g^2) (fp32, device scalar)."""
    assert g.dtype == torch.float32 and g.is_contiguous()
    if out is None:
        out = torch.zeros(1, device=g.device, dtype=torch.float32)
    _call("smx_sqnorm", g.data_ptr(), g.numel(), out.data_ptr(), _stream(g))
    return out


def adamw_(master, m, v, grad, param_bf16, lr, beta1, beta2, eps, weight_decay, step, sqnorm_t=None,
           max_norm: float = 0.0, grad_prescale: float = 1.0, zero_grad: bool = True) -> None:
    """One fused AdamW step on flat fp32 buffers; writes the bf16 model copy and (optionally) zeroes the grads."""
    n = master.numel()
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    _call("smx_adamw", master.data_ptr(), m.data_ptr(), v.data_ptr(), grad.data_ptr(),
          param_bf16.data_ptr() if param_bf16 is not None else None, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
          sqnorm_t.data_ptr() if sqnorm_t is not None else None, float(max_norm), float(grad_prescale),
          1 if zero_grad else 0, _stream(master))


def permute_rows(src: torch.Tensor, row_map: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dst[i] = src[row_map[i]] (row_map int32, -1 = zero row): MoE dispatch permutation."""
    _bf16c(src, "src")
    assert row_map.dtype == torch.int32
    H = src.shape[-1]
    n = row_map.numel()
    dst = torch.empty(n, H, device=src.device, dtype=src.dtype) if out is None else out
    _call("smx_permute_rows", src.data_ptr(), dst.data_ptr(), row_map.data_ptr(), n, H, _stream(src))
    return dst


def unpermute_combine(x: torch.Tensor, inv_map: torch.Tensor, probs: Optional[torch.Tensor], tokens: int, topk: int,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[t] = sum_k probs[t,k] * x[inv_map[t,k]]: MoE combine."""
    _bf16c(x, "x")
    H = x.shape[-1]
    y = torch.empty(tokens, H, device=x.device, dtype=x.dtype) if out is None else out
    _call("smx_unpermute_combine", x.data_ptr(), y.data_ptr(), inv_map.data_ptr(),
          probs.data_ptr() if probs is not None else None, tokens, topk, H, _stream(x))
    return y


def moe_prob_grad(dy: torch.Tensor, y_perm: torch.Tensor, inv_map: torch.Tensor, topk: int) -> torch.Tensor:
    """d_probs[t, k] = <dy[t], y_perm[inv_map[t, k]]> (0 where the slot was dropped): gradient of the MoE combine weights."""
    _bf16c(dy, "dy"); _bf16c(y_perm, "y_perm")
    tokens, H = dy.shape
    out = torch.empty(tokens, topk, device=dy.device, dtype=torch.float32)
    _call("smx_moe_prob_grad", dy.data_ptr(), y_perm.data_ptr(), inv_map.data_ptr(), out.data_ptr(), tokens, topk, H, _stream(dy))
    return out


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _bf16c(a, "a"); _bf16c(b, "b")
    y = torch.empty_like(a) if out is None else out
    _call("smx_add", a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream(a))
    return y


def copy_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    nbytes = src.numel() * src.element_size()
    _call("smx_copy", src.data_ptr(), dst.data_ptr(), nbytes, _stream(src))
    return dst


def embedding_fwd(tokens: torch.Tensor, table: torch.Tensor, out: torch.Tensor, v0: int = 0) -> torch.Tensor:
    """out[t] = table[tokens[t] - v0]; rows whose token is outside this vocabulary shard ``[v0, v0 + table.shape[0])`` are zeroed
    (vocab-parallel embedding, reference dense_module.py Embedding)."""
    _bf16c(table, "table"); _bf16c(out, "out")
    assert tokens.dtype == torch.int64 and tokens.is_contiguous()
    _call("smx_embedding_fwd", tokens.data_ptr(), table.data_ptr(), out.data_ptr(), tokens.numel(), table.shape[1], v0, table.shape[0],
          _stream(out))
    return out


def embedding_bwd(tokens: torch.Tensor, dy: torch.Tensor, grad_fp32: torch.Tensor, v0: int = 0) -> None:
    """grad[tokens[t] - v0] += dy[t] (fp32 main grad, vector reductions; tokens outside the shard are skipped)."""
    _bf16c(dy, "dy")
    assert grad_fp32.dtype == torch.float32 and grad_fp32.is_contiguous() and tokens.dtype == torch.int64
    _call("smx_embedding_bwd", tokens.data_ptr(), dy.data_ptr(), grad_fp32.data_ptr(), tokens.numel(), grad_fp32.shape[1], v0,
          grad_fp32.shape[0], _stream(dy))


def rope_strided_(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq_len: int, inverse: bool = False) -> torch.Tensor:
    """In-place RoPE on a strided view ``t[tokens, heads, d]`` (last dim contiguous; any token / head strides that are multiples of
    8 elements): MLA rotates the last 64 of every 192-wide query head and the shared 64-wide key slice."""
    assert t.dtype == torch.bfloat16 and t.dim() == 3 and t.stride(2) == 1
    tokens, heads, d = t.shape
    _call("smx_rope_strided", t.data_ptr(), cos.data_ptr(), sin.data_ptr(), tokens, seq_len, t.stride(0), heads, d,
          t.stride(1) if heads > 1 else d, -1.0 if inverse else 1.0, _stream(t))
    return t


def mla_k_assemble(kv: torch.Tensor, k_rope: torch.Tensor, out: torch.Tensor, dn: int) -> torch.Tensor:
    """out[t, h, :dn] = kv[t, h, :dn] (k_nope half of the kv_up output), out[t, h, dn:] = k_rope[t] (shared rotary key slice).
    ``kv`` [T, heads, dn + dv] and ``k_rope`` [T, dr] may be strided views; ``out`` [T, heads, dn + dr] is contiguous."""
    T, heads, _ = kv.shape
    dr = k_rope.shape[-1]
    assert out.is_contiguous() and out.shape == (T, heads, dn + dr) and kv.stride(2) == 1 and k_rope.stride(1) == 1
    _call("smx_mla_k_assemble", kv.data_ptr(), kv.stride(0), kv.stride(1), k_rope.data_ptr(), k_rope.stride(0), out.data_ptr(), T, heads,
          dn, dr, _stream(out))
    return out


def mla_k_split_bwd(dk: torch.Tensor, d_kv: torch.Tensor, d_k_rope: torch.Tensor, dn: int) -> None:
    """Backward of :func:`mla_k_assemble`: d_kv[t, h, :dn] = dk[t, h, :dn];  d_k_rope[t] = sum over heads of dk[t, h, dn:]."""
    T, heads, w = dk.shape
    assert dk.is_contiguous() and d_kv.stride(2) == 1 and d_k_rope.stride(1) == 1
    _call("smx_mla_k_split_bwd", dk.data_ptr(), d_kv.data_ptr(), d_kv.stride(0), d_kv.stride(1), d_k_rope.data_ptr(), d_k_rope.stride(0),
          T, heads, dn, w - dn, _stream(dk))


def copy2d_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst[r, :] = src[r, :] for 2-D views with a contiguous last dim (row strides free): column slices in and out of packed buffers."""
    assert dst.shape == src.shape and dst.dim() == 2 and dst.stride(1) == 1 and src.stride(1) == 1 and dst.dtype == src.dtype
    es = dst.element_size()
    _call("smx_copy2d", src.data_ptr(), src.stride(0) * es, dst.data_ptr(), dst.stride(0) * es, dst.shape[0], dst.shape[1] * es, _stream(dst))
    return dst


def transpose01_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst[a, b, :] = src[b, a, :] for contiguous 3-D tensors (``src`` [B, A, w], ``dst`` [A, B, w]): the pack / unpack step around a
    context-parallel all-to-all (tokens x destination-rank column blocks <-> destination rank x tokens)."""
    assert src.is_contiguous() and dst.is_contiguous() and src.dim() == 3 and dst.dim() == 3 and src.dtype == dst.dtype
    B, A, w = src.shape
    assert dst.shape == (A, B, w)
    _call("smx_transpose01", src.data_ptr(), dst.data_ptr(), A, B, w * src.element_size(), _stream(dst))
    return dst
