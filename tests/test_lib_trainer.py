"""The library arm of the bench (baseline/lib_trainer.py: stock PyTorch on cuBLAS / cuDNN / NCCL) on CPU with gloo: ZeRO-1 data parallel
over 2 processes reproduces the single-process step on the same global batch; tensor + sequence parallel ranks agree on the loss and
the loss goes down.  (The arm's job is to be a correct Megatron-style step on vendor libraries -- nothing of simumax_b200 is imported.)"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
SEQ, STEPS = 16, 3


def _dims():
    from lib_trainer import Dims
    return Dims(hidden=32, heads=4, kv_heads=2, head_dim=8, ffn=64, layers=2, vocab=64)


def _batches(n):
    g = torch.Generator().manual_seed(11)
    return [(torch.randint(0, 64, (SEQ,), generator=g), torch.randint(0, 64, (SEQ,), generator=g)) for _ in range(n)]


def _single():
    from lib_trainer import LibLlamaTrainer
    tr = LibLlamaTrainer(_dims(), 1, SEQ, device=torch.device("cpu"), lr=1e-2)
    losses = [float(tr.run_microbatches(_batches(4))) for _ in range(STEPS)]
    return losses, tr.params.float().clone()


def _worker(rank, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    from lib_trainer import LibLlamaTrainer
    try:
        if mode == "dp":
            tr = LibLlamaTrainer(_dims(), 1, SEQ, dp_group=dist.group.WORLD, device=torch.device("cpu"), lr=1e-2)
            mine = _batches(4)[rank::2]
            losses = []
            for _ in range(STEPS):
                l = tr.run_microbatches(mine).clone()
                dist.all_reduce(l)
                losses.append(float(l) / 2)
            if rank == 0:
                out.put((losses, tr.params.float().clone()))
        else:
            tr = LibLlamaTrainer(_dims(), 1, SEQ, tp_group=dist.group.WORLD, device=torch.device("cpu"), lr=1e-2)
            losses = [float(tr.run_microbatches(_batches(2))) for _ in range(6)]
            both = [None, None]
            dist.all_gather_object(both, losses)
            if rank == 0:
                out.put(both)
    finally:
        dist.destroy_process_group()


def _spawn(mode):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return res


@pytest.mark.timeout(300)
def test_zero1_data_parallel_matches_single_process():
    ref_losses, ref_params = _single()
    losses, params = _spawn("dp")
    assert losses == pytest.approx(ref_losses, rel=2e-2)
    assert ref_losses[-1] < ref_losses[0]
    n = min(params.numel(), ref_params.numel())            # the flat buffer is padded to a multiple of the ZeRO shard count
    assert (params[:n] - ref_params[:n]).abs().max() <= 2e-2 * ref_params.abs().max()


@pytest.mark.timeout(300)
def test_tensor_sequence_parallel_ranks_agree_and_learn():
    a, b = _spawn("tp")
    assert a == pytest.approx(b, rel=1e-5)
    assert all(x == x for x in a) and a[-1] < a[0]
