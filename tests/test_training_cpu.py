"""Host logic of the data-parallel training step (SURVEY.md 8e "training DP"): the bucket plan is a pure function every
rank derives identically, and `GradientBuckets` averages gradients over a world-size-2 gloo group the way DDP does for the
reference (scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1165,1470).  The kernels themselves need a GPU
(tests/test_training_gpu.py); here only torch.distributed plumbing and index arithmetic run."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_bucket_plan_is_reverse_order_and_bounded():
    from posetraj_b200.training import bucket_plan
    sizes = [10, 200, 30, 5, 400, 50]
    plan = bucket_plan(sizes, 250)
    flat = [i for b in plan for i in b]
    assert flat == list(reversed(range(len(sizes))))               # backward order: last parameters first
    for b in plan:
        assert sum(sizes[i] for i in b) <= 250 or len(b) == 1       # an oversized parameter gets its own bucket
    assert bucket_plan(sizes, 10 ** 9) == [list(reversed(range(len(sizes))))]
    assert bucket_plan([], 10) == []


def test_controlnet_bucket_plan_at_full_size():
    """The 682 M-parameter ControlNet in 100 MB fp32 buckets: every parameter exactly once, ~27 buckets."""
    from posetraj_b200.config import SVDConfig, controlnet_param_shapes
    from posetraj_b200.training import bucket_plan
    import math
    sizes = [int(math.prod(s)) for s in controlnet_param_shapes(SVDConfig(), cam=True, bbox=False).values()]
    plan = bucket_plan(sizes, int(100 * (1 << 20) / 4))
    assert sorted(i for b in plan for i in b) == list(range(len(sizes)))
    assert 20 <= len(plan) <= 40 and sum(sizes) > 680e6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    dist.init_process_group("gloo", rank=rank, world_size=2)
    from posetraj_b200.training import GradientBuckets
    sizes = [7, 1000, 33, 512]
    gb = GradientBuckets(sizes, "cpu", bucket_mb=0.002)          # 524 elements per bucket: several buckets
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(n, generator=g) for n in sizes]
    for i in reversed(range(len(sizes))):                        # the order a backward pass produces them in
        gb.view(i).copy_(grads[i])
        gb.ready(i)
    gb.finish()
    q.put((rank, [gb.view(i).clone() / gb.world for i in range(len(sizes))], grads, len(gb.buckets)))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_buckets_average_over_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, avg, local, nb = q.get(timeout=400)
        got[rank] = (avg, local, nb)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert got[0][2] == got[1][2] >= 3
    for i in range(4):
        want = (got[0][1][i] + got[1][1][i]) / 2
        assert torch.allclose(got[0][0][i], want, atol=1e-6) and torch.allclose(got[1][0][i], want, atol=1e-6)


def test_training_module_refuses_cpu_tensors():
    from posetraj_b200 import training as T
    with pytest.raises(RuntimeError):
        T.colsum(torch.zeros(4, 32, dtype=torch.bfloat16))
