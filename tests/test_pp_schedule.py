"""1F1B schedule + point-to-point matching of ``PPLlamaTrainer`` on CPU: gloo, 2-4 processes, a small autograd stage plugged in
through ``stage_factory``.  Every (pp, micro-batch count) combination must finish (no deadlock) and reproduce the losses and
weights of the same network trained in one process."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

H, T, V, LAYERS, LR, STEPS = 16, 8, 32, 7, 0.5, 3


def _weights():
    g = torch.Generator().manual_seed(0)
    return {"embed": torch.randn(V, H, generator=g), "layers": [torch.randn(H, H, generator=g) * 0.4 for _ in range(LAYERS)],
            "head": torch.randn(V, H, generator=g) * 0.4}


def _batches(M):
    g = torch.Generator().manual_seed(1)
    return [(torch.randint(0, V, (T,), generator=g), torch.randint(0, V, (T,), generator=g)) for _ in range(M)]


def _reference(M):
    w = _weights()
    params = [w["embed"], *w["layers"], w["head"]]
    for p in params:
        p.requires_grad_(True)
    losses = []
    for _ in range(STEPS):
        total = 0.0
        for tok, tgt in _batches(M):
            x = w["embed"][tok]
            for W in w["layers"]:
                x = torch.tanh(x @ W.t())
            total = total + torch.nn.functional.cross_entropy(x @ w["head"].t(), tgt, reduction="sum") / (M * T)
        total.backward()
        with torch.no_grad():
            for p in params:
                p -= LR * p.grad
                p.grad = None
        losses.append(float(total))
    return losses, w


class _Stage:
    """What PPLlamaTrainer needs from a stage, on top of autograd."""

    def __init__(self, lo, hi, first, last, in_flight):
        w = _weights()
        self.first, self.last, self.T = first, last, T
        self.layers = [w["layers"][i].clone().requires_grad_(True) for i in range(lo, hi)]
        self.embed = w["embed"].clone().requires_grad_(True) if first else None
        self.head = w["head"].clone().requires_grad_(True) if last else None
        self.slots = [dict() for _ in range(in_flight)]
        self.cur = self.slots[0]
        self.max_live = 0

    def use_slot(self, k):
        self.cur = self.slots[k % len(self.slots)]

    def forward(self, x_or_tok, targets, loss_scale):
        assert "y" not in self.cur, "activation slot reused before its backward ran"
        x = self.embed[x_or_tok] if self.first else x_or_tok.clone().requires_grad_(True)
        self.cur["x"] = x
        y = x
        for W in self.layers:
            y = torch.tanh(y @ W.t())
        self.max_live = max(self.max_live, sum("y" in s for s in self.slots) + 1)
        if self.last:
            rows = torch.nn.functional.cross_entropy(y @ self.head.t(), targets, reduction="none")
            self.cur["y"] = rows.sum() * loss_scale
            return rows.detach()
        self.cur["y"] = y
        return y.detach()

    def backward(self, tokens, last_microbatch=False, d_out=None):
        y = self.cur.pop("y")
        y.backward() if self.last else y.backward(d_out)
        x = self.cur.pop("x")
        return None if self.first else x.grad

    def optimizer_step(self):
        with torch.no_grad():
            for p in [*self.layers, self.embed, self.head]:
                if p is not None:
                    p -= LR * p.grad
                    p.grad = None


def _worker(rank, world, port, M, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simumax_b200.runtime.pp_trainer import PPLlamaTrainer, split_layers
    from simumax_b200.runtime.trainer import LlamaDims
    lo, hi = split_layers(LAYERS, world)[rank]
    dims = LlamaDims(hidden=H, heads=1, kv_heads=1, head_dim=H, ffn=H, layers=LAYERS, vocab=V)
    pp = PPLlamaTrainer(dims, 1, T, device="cpu", act_dtype=torch.float32,
                        stage_factory=lambda local, first, last, n: _Stage(lo, hi, first, last, n))
    assert (pp.lo, pp.hi) == (lo, hi) and pp.tr.first == (rank == 0) and pp.tr.last == (rank == world - 1)
    losses = [float(pp.run_microbatches(_batches(M))) for _ in range(STEPS)]
    ref_losses, ref_w = _reference(M)
    ok = all(abs(a - b) < 1e-5 * max(1.0, abs(b)) for a, b in zip(losses, ref_losses))
    for j, i in enumerate(range(lo, hi)):
        ok = ok and torch.allclose(pp.tr.layers[j], ref_w["layers"][i], atol=1e-5)
    if rank == 0:
        ok = ok and torch.allclose(pp.tr.embed, ref_w["embed"], atol=1e-5)
    if rank == world - 1:
        ok = ok and torch.allclose(pp.tr.head, ref_w["head"], atol=1e-5)
    # 1F1B keeps at most pp - stage micro-batches alive on a stage
    ok = ok and pp.tr.max_live <= world - rank
    out[rank] = (bool(ok), losses, ref_losses, pp.tr.max_live)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,M", [(2, 1), (2, 2), (2, 5), (3, 2), (3, 3), (3, 7), (4, 4), (4, 9)])
def test_1f1b_matches_single_process(world, M):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), M, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok, losses, ref, live = out[rank]
        assert ok, (rank, losses, ref, live)
