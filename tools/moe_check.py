"""Expert-parallel MoE block (fused permute + all-to-all dispatch / combine kernels, grouped tcgen05 GEMMs) vs an fp32 PyTorch
reference with the same routing, under torchrun on >= 2 GPUs: output, input gradient, router and expert weight gradients;
then the DeepSeek-V2-shaped block is timed and the fused dispatch is compared with permute kernel + NCCL all_to_all.
Writes gpurun_out/moe_check_<world>.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F_

from simumax_b200.ops import elementwise as E
from simumax_b200.runtime.moe_layer import EPMoELayer


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def reference(layer, x, dy, w1_full, w2_full):
    """fp32 autograd reference using the layer's own routing decisions (indices + drops)."""
    T, k, Etot, Fh = layer.T, layer.k, layer.E, layer.F
    xf = x.float().requires_grad_(True)
    wr = layer.w_router.float().requires_grad_(True)
    w1 = w1_full.float().requires_grad_(True)
    w2 = w2_full.float().requires_grad_(True)
    scores = torch.softmax(xf @ wr.t(), dim=-1)
    probs = scores.gather(1, layer.idx)
    keep = (layer.inv_map.view(T, k) >= 0)
    y = torch.zeros(T, layer.H, device=x.device)
    for e in range(Etot):
        sel = (layer.idx == e) & keep
        if not sel.any():
            continue
        t_idx, j_idx = sel.nonzero(as_tuple=True)
        h = xf[t_idx] @ w1[e].t()
        a = F_.silu(h[:, :Fh]) * h[:, Fh:]
        o = a @ w2[e].t()
        y = y.index_add(0, t_idx, o * probs[t_idx, j_idx].unsqueeze(1))
    (y * dy.float()).sum().backward()
    return y.detach(), xf.grad, wr.grad, w1.grad, w2.grad


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {"world": world}
    ok = True
    for balanced in (False, True):
        T, H, Fh, Etot, k = 1024, 1024, 512, 4 * world, 2
        layer = EPMoELayer(H, Fh, Etot, k, T, force_balanced=balanced, seed=3)
        g = torch.Generator(device="cuda").manual_seed(50 + rank)
        x = (torch.randn(T, H, device="cuda", generator=g)).bfloat16()
        dy = (torch.randn(T, H, device="cuda", generator=g) * 0.1).bfloat16()
        y = layer.forward(x)
        dx = layer.backward(dy)
        torch.cuda.synchronize()
        w1s = [torch.empty_like(layer.w1) for _ in range(world)]
        w2s = [torch.empty_like(layer.w2) for _ in range(world)]
        dist.all_gather(w1s, layer.w1); dist.all_gather(w2s, layer.w2)
        y_ref, dx_ref, gwr_ref, gw1_ref, gw2_ref = reference(layer, x, dy, torch.cat(w1s), torch.cat(w2s))
        dist.all_reduce(gw1_ref); dist.all_reduce(gw2_ref)
        sl = slice(rank * layer.El, (rank + 1) * layer.El)
        checks = {"y": rel(y, y_ref), "dx": rel(dx, dx_ref), "g_router": rel(layer.g_router, gwr_ref),
                  "g_w1": rel(layer.g_w1, gw1_ref[sl]), "g_w2": rel(layer.g_w2, gw2_ref[sl]),
                  "dropped_slots": int((layer.inv_map < 0).sum())}
        worst = torch.tensor([max(v for n, v in checks.items() if n != "dropped_slots")], device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        checks["max_rel_err_all_ranks"] = float(worst)
        ok = ok and float(worst) < 3e-2
        out["balanced" if balanced else "routed"] = checks
        del layer
    # ---- DeepSeek-V2-shaped block: hidden 5120, moe_ffn 1536, 160 experts, top-6, 4096 tokens per rank, balanced routing
    T, H, Fh, Etot, k = 4096, 5120, 1536, 160, 6
    if Etot % world == 0:
        layer = EPMoELayer(H, Fh, Etot, k, T, force_balanced=True, seed=5)
        x = torch.randn(T, H, device="cuda").bfloat16()
        dy = (torch.randn(T, H, device="cuda") * 0.1).bfloat16()
        def step():
            layer.forward(x)
            layer.backward(dy)
        ms = timed(step, 5)
        t_disp = timed(lambda: layer.ws.moe_dispatch(x, layer.row_map, layer.off_recv, layer.C, layer.El))
        perm = torch.empty(layer.rows, H, device="cuda", dtype=torch.bfloat16)
        recv = torch.empty_like(perm)
        def unfused():
            E.permute_rows(x, layer.row_map, out=perm)
            dist.all_to_all_single(recv, perm)
        t_unf = timed(unfused)
        nv_bytes = (world - 1) / world * layer.rows * H * 2
        out["deepseek_v2_block"] = {"tokens": T, "hidden": H, "moe_ffn": Fh, "experts": Etot, "topk": k, "capacity": layer.C,
                                    "fwd_bwd_ms": ms, "tflops_per_gpu": layer.flops() / ms / 1e9,
                                    "dispatch_fused_ms": t_disp, "permute_plus_nccl_a2a_ms": t_unf,
                                    "dispatch_nvlink_GBps": nv_bytes / t_disp / 1e6,
                                    "dispatch_roofline_fraction": (nv_bytes / 900e9 * 1e3) / t_disp}
    out["ok"] = bool(ok)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/moe_check_{world}.json", "w"), indent=1)
        print(json.dumps(out))
        print("MOE_OK" if ok else "MOE_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
