"""
Sharded-env logging over two processes (gloo, CPU): the per-rank accumulators written by the
finalize kernel are summed with an all-reduce and turned into means by `combine_logging`, which must
reproduce what one process over all envs would log.  This is the only cross-rank exchange of the
path; the kernels themselves are covered by the -m gpu tests.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genesis_forge_b200.fused import combine_logging

N_R, N_T = 3, 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_accumulator(rank: int, envs_per_rank: int):
    """Synthetic per-env episode quotients / fire flags for one shard, and the accumulator a rank builds."""
    g = torch.Generator().manual_seed(100 + rank)
    reset = torch.rand(envs_per_rank, generator=g) < 0.3
    quot = torch.randn(N_R, envs_per_rank, generator=g, dtype=torch.float64)
    fired = torch.rand(N_T, envs_per_rank, generator=g) < 0.2
    acc = torch.zeros(N_R + N_T + 1, dtype=torch.float64)
    acc[:N_R] = (quot * reset).sum(dim=1)
    acc[N_R:N_R + N_T] = fired.sum(dim=1).double()
    acc[-1] = reset.sum()
    return acc, quot, fired, reset


def _worker(rank: int, world: int, port: int, envs_per_rank: int, out_path: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc, _, _, _ = _rank_accumulator(rank, envs_per_rank)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    out = combine_logging(acc, N_R, N_T, envs_per_rank * world)
    if rank == 0:
        torch.save({"out": out, "acc": acc}, out_path)
    dist.destroy_process_group()


def test_two_rank_logging_equals_single_process(tmp_path):
    world, envs = 2, 257
    out_path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(world, _free_port(), envs, out_path), nprocs=world, join=True)
    got = torch.load(out_path)
    parts = [_rank_accumulator(r, envs) for r in range(world)]
    quot = torch.cat([p[1] for p in parts], dim=1)
    fired = torch.cat([p[2] for p in parts], dim=1)
    reset = torch.cat([p[3] for p in parts])
    want_rewards = (quot * reset).sum(dim=1) / reset.sum()
    want_terms = fired.sum(dim=1).double() / (envs * world)
    assert torch.allclose(got["out"][:N_R].double(), want_rewards, rtol=1e-6)
    assert torch.allclose(got["out"][N_R:].double(), want_terms, rtol=1e-6)
    assert got["acc"][-1] == reset.sum()
