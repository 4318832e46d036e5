"""Layer-by-layer comparison of LlamaTrainer's activations / gradients with an fp32 autograd model (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer

dims = LlamaDims(hidden=256, heads=2, kv_heads=1, head_dim=128, ffn=512, layers=2, vocab=512)
tr = LlamaTrainer(dims, micro_batch=2, seq_len=128, attention_bwd=sys.argv[1] if len(sys.argv) > 1 else "auto", max_grad_norm=0.0)
for k_ in ('embed', 'lm_head'):
    tr.w[k_].mul_(4.0)
for k_ in ('wqkv', 'wo', 'wfc1', 'wfc2'):
    for t_ in tr.w[k_]:
        t_.mul_(4.0)
d = dims; B, S, T = tr.B, tr.S, tr.T
g = torch.Generator(device="cuda").manual_seed(7)
tokens = torch.randint(0, d.vocab, (T,), device="cuda", generator=g)
targets = torch.randint(0, d.vocab, (T,), device="cuda", generator=g)
W = {k: ([t.detach().float().requires_grad_(True) for t in v] if isinstance(v, list) else v.detach().float().requires_grad_(True)) for k, v in tr.w.items()}
cos, sin = tr.cos, tr.sin
rms = lambda x, gm: x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + d.eps) * gm
def rope(t):
    h = d.head_dim // 2; a, b = t[..., :h], t[..., h:]; c, s = cos[None, :, None, :], sin[None, :, None, :]
    return torch.cat([a * c - b * s, b * c + a * s], -1)
keep = {}
def K(name, t):
    t.retain_grad(); keep[name] = t; return t
x = K("x0", W["embed"][tokens])
for i in range(d.layers):
    h1 = K(f"h1_{i}", rms(x, W["ln1"][i]))
    qkv = K(f"qkvpre_{i}", h1 @ W["wqkv"][i].t()).view(B, S, -1)
    nq, nk, hd = d.heads, d.kv_heads, d.head_dim
    q = rope(qkv[..., : nq * hd].reshape(B, S, nq, hd)); k = rope(qkv[..., nq * hd:(nq + nk) * hd].reshape(B, S, nk, hd))
    v = qkv[..., (nq + nk) * hd:].reshape(B, S, nk, hd)
    K(f"qkv_{i}", torch.cat([q.reshape(B, S, -1), k.reshape(B, S, -1), v.reshape(B, S, -1)], -1).reshape(T, -1))
    o = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, enable_gqa=True)
    o = K(f"o_{i}", o.transpose(1, 2).reshape(T, nq * hd))
    x = K(f"x2_{i}", x + o @ W["wo"][i].t())
    h2 = K(f"h2_{i}", rms(x, W["ln2"][i])); u = K(f"u_{i}", h2 @ W["wfc1"][i].t())
    gg = K(f"g_{i}", torch.nn.functional.silu(u[:, :d.ffn]) * u[:, d.ffn:])
    x = K(f"x{i + 1}", x + gg @ W["wfc2"][i].t())
xf = K("xf", rms(x, W["final_norm"])); logits = K("logits", xf @ W["lm_head"].t())
loss = torch.nn.functional.cross_entropy(logits, targets); loss.backward()

def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
orig_ce = None
import simumax_b200.ops.elementwise as E
snap = {}
real_ce = E.cross_entropy_
def ce_spy(lg, tg, **kw):
    snap["logits"] = lg.clone(); r = real_ce(lg, tg, **kw); snap["dlogits"] = lg.clone(); return r
E.cross_entropy_ = ce_spy
dxs = {}
tr.debug_hook = lambda i, t: dxs.__setitem__(i, t.dx.clone())
rows = tr.forward(tokens, targets, loss_scale=1.0 / T)
a = tr.act
print("loss", rows.mean().item(), loss.item())
for i in range(d.layers):
    print(i, "x", rel(a["x"][i], keep[f"x{i}"]), "h1", rel(a["h1"][i], keep[f"h1_{i}"]), "qkv", rel(a["qkv"][i], keep[f"qkv_{i}"]),
          "o", rel(a["o"][i], keep[f"o_{i}"]), "x2", rel(a["x2"][i], keep[f"x2_{i}"]), "h2", rel(a["h2"][i], keep[f"h2_{i}"]),
          "u", rel(a["u"][i], keep[f"u_{i}"]), "g", rel(a["g"][i], keep[f"g_{i}"]))
print("xf", rel(a["xf"], keep["xf"]), "logits", rel(snap["logits"], keep["logits"]), "dlogits", rel(snap["dlogits"], keep["logits"].grad))
# backward with spies on scratch buffers
tr.backward(tokens)
for i in reversed(range(d.layers)):
    print(i, "dx_in(layer output grad -> after layer)", rel(dxs[i], keep[f"x{i}"].grad))
print("d_qkv(last computed = layer0, post inverse rope) vs qkvpre grad", rel(tr.d_qkv, keep["qkvpre_0"].grad))
print("d_o layer0", rel(tr.d_o, keep["o_0"].grad))
print("d_big layer0 (du)", rel(tr.d_big, keep["u_0"].grad), "d_f (dg)", rel(tr.d_f, keep["g_0"].grad))
for name in ("lm_head", "final_norm", "embed"):
    print(name, rel(tr.g[name], W[name].grad))
for name in ("wfc2", "wfc1", "wo", "wqkv", "ln1", "ln2"):
    for i in range(d.layers):
        print(name, i, rel(tr.g[name][i], W[name][i].grad))
