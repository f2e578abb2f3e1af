"""world_size-2 gloo test of the N>1 host logic: shard the evidence batch, run each shard independently,
gather the posteriors. The per-rank compute stand-in is the oracle's plan interpreter (tests may use it;
the product ranks run the CUDA engine)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pgmpy_b200 as px
from oracle.plan_exec import run_plan
from pgmpy_b200.distributed import gather_posteriors, gather_posteriors_to_root, shard_range, shard_rows
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.planner import JTStructure, compile_jt_plan


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 7, 8, 1000, 131072):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = px.get_example_model("asia")
    ev_vars, states = sample_evidence(m, n, 2, seed=3)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars)
    mine = shard_rows(states, world, rank)
    local = torch.from_numpy(run_plan(plan.pool, plan.const_blob, mine))
    full = gather_posteriors(local, total_rows=n)
    if rank == 0:
        np.save(out_path, full.numpy())
    if n % world == 0:
        # gather-to-root (what bench.py times at N > 1): only rank 0 receives, in rank order
        rooted = gather_posteriors_to_root(local, dst=0)
        assert (rooted is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(rooted, full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 37])
def test_two_rank_shard_and_gather_matches_single_process(tmp_path, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, n, out_path), nprocs=2, join=True)
    m = px.get_example_model("asia")
    ev_vars, states = sample_evidence(m, n, 2, seed=3)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars)
    want = run_plan(plan.pool, plan.const_blob, states)
    np.testing.assert_array_equal(np.load(out_path), want)
