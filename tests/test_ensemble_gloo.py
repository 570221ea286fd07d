"""Multi-rank host logic on CPU: interleaved sharding + all-gather, world_size 2 over gloo."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, K, d, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "code-adaptive-prob-ode-solvers_b200"))
    from odecheckpts_b200 import ensemble

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = ensemble.shard_indices(B, rank, world)
    # stand-in for the per-rank solve: results are a deterministic function of the member id
    ids = torch.as_tensor(idx, dtype=torch.float64)
    local = {
        "u": (ids[:, None, None] * 10 + torch.arange(K, dtype=torch.float64)[None, :, None]).expand(-1, -1, d).contiguous(),
        "n_accepted": (torch.as_tensor(idx)[:, None] * 100 + torch.arange(K)[None, :]).to(torch.int64),
        "status": (torch.as_tensor(idx) % 3).to(torch.int32),
    }
    out = ensemble.all_gather_results(local, B)
    member = torch.arange(B, dtype=torch.float64)
    ok = bool(torch.equal(out["u"][:, :, 0], member[:, None] * 10 + torch.arange(K, dtype=torch.float64)[None, :]))
    ok &= bool(torch.equal(out["n_accepted"], (torch.arange(B)[:, None] * 100 + torch.arange(K)[None, :])))
    ok &= bool(torch.equal(out["status"], (torch.arange(B) % 3).to(torch.int32)))
    # the packed path: results are written into views of ONE buffer, one collective, strided member-order views
    packed = ensemble.PackedResults(B, K, d, world, torch.device("cpu"))
    mine = packed.local(len(idx))
    mine["u"].copy_(local["u"])
    mine["u_std"].copy_(local["u"] * 0.5)
    mine["n_accepted"].copy_(local["n_accepted"])
    mine["n_rejected"].copy_(torch.as_tensor(idx) * 7)
    mine["status"].copy_(local["status"])
    g = packed.all_gather()
    for i in range(packed.capacity):
        for r in range(world):
            b = i * world + r
            if b >= B:
                continue
            ok &= bool(g["u"][i, r, 0, 0].item() == b * 10) and bool(g["u_std"][i, r, K - 1, d - 1].item() == 0.5 * (b * 10 + K - 1))
            ok &= bool(g["n_accepted"][i, r, 1].item() == b * 100 + 1) and bool(g["n_rejected"][i, r].item() == 7 * b)
            ok &= bool(g["status"][i, r].item() == b % 3)
    ok &= bool(torch.equal(packed.member_order(g["n_rejected"]), torch.arange(B) * 7))
    ok &= g["u"].data_ptr() == packed._gathered.data_ptr()  # a view of the gathered buffer, not a copy
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if t.item() != 1.0:
        raise SystemExit(3)


@pytest.mark.parametrize("B", [10, 11])  # even and ragged shards
def test_interleaved_shards_all_gather_restores_member_order(B):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, 4, 2, 1), nprocs=world, join=True)


def test_shard_arithmetic():
    from odecheckpts_b200 import ensemble

    for B in (0, 1, 7, 8, 65536):
        for G in (1, 2, 4, 8):
            sizes = ensemble.shard_sizes(B, G)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
            cat = np.concatenate([ensemble.shard_indices(B, r, G) for r in range(G)]) if B else np.zeros(0, int)
            assert sorted(cat.tolist()) == list(range(B))
            if B:
                np.testing.assert_array_equal(cat[ensemble.unshard_order(B, G)], np.arange(B))


def test_bench_inputs_are_the_interleaved_shards_of_one_global_ensemble():
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    full, _ = bench.ensemble_inputs(0, 64)
    for world in (2, 4):
        for r in range(world):
            part, _ = bench.ensemble_inputs(r, 64 // world, world)
            np.testing.assert_array_equal(part, full[r::world])
