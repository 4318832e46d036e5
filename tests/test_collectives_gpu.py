"""In-kernel NVLink collectives vs NCCL on 2+ GPUs (skipped on single-GPU boxes)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_collectives_match_nccl(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    env = dict(os.environ, COLL_MAX_BYTES=str(16 << 20), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "coll_bench.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(os.path.join(ROOT, "gpurun_out", f"coll_bench_{world}.json")))
    assert res["all_ranks_correct"], res["correct"]


@pytest.mark.timeout(300)
def test_zero1_step_smx_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"], line


@pytest.mark.timeout(300)
def test_fused_gemm_reduce_scatter():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tools", "fused_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and "FUSED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.timeout(300)
def test_tensor_parallel_trainer_matches_single_gpu():
    """TP + SP training step on the fused AG->GEMM / GEMM->RS kernels vs the 1-GPU trainer with the same weights."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29536", os.path.join(ROOT, "tools", "tp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and "TP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _torchrun(script, port, marker, nproc=2):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs >= {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and marker in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.timeout(300)
def test_expert_parallel_moe_block_matches_fp32_reference():
    """Fused permute + all-to-all dispatch / combine, grouped GEMMs, router backward vs an fp32 autograd reference."""
    _torchrun("moe_check.py", 29537, "MOE_OK")


@pytest.mark.timeout(300)
def test_pipeline_parallel_trainer_matches_single_gpu():
    """1F1B over NCCL p2p on a side stream vs the 1-GPU trainer: loss trajectory and parameters after 4 steps (validated on 2, 4 and
    8 GPUs in round 2: profiles/pp_check_*.json)."""
    _torchrun("pp_check.py", 29538, "PP_OK")


@pytest.mark.timeout(300)
def test_tp_recompute_and_collectives_self_checks():
    """runtime/selfcheck.py in one process group: collectives vs NCCL, TP trainer vs single GPU without / with selective / full recompute."""
    _torchrun("multi_check.py", 29539, "MULTI_OK")


@pytest.mark.timeout(300)
def test_context_parallel_trainer_matches_single_gpu():
    """Ulysses all-to-all context parallelism vs the single-GPU trainer on the full sequence (also with selective recompute)."""
    _torchrun("cp_check.py", 29540, "CP_OK")


@pytest.mark.timeout(300)
def test_deepseek_trainer_expert_parallel_matches_fp32_reference():
    """MLA + MoE (EP = DP = 2) training step vs an fp32 autograd model with the same routing."""
    _torchrun("deepseek_check.py", 29541, "DEEPSEEK_OK")
