"""CPU, world_size 2 over gloo: the row-sharding host logic of the N>1 path.  The oracle stands in for
the kernel (it is the checker here, not the product) so the test can run without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from repo_b200 import parallel


def test_shard_rows_partition():
    for n in (0, 1, 7, 50, 2450, 65536):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    assert [parallel.shard_rows(50, r, 8)[1] for r in range(8)] == [7, 7, 6, 6, 6, 6, 6, 6]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rssm_oracle as O
    N, H = 9, 4  # uneven split: 5 + 4
    P = O.make_transition_params(3)
    PA = O.make_mlp_params(4, 230, 200, 12, 4)
    PV = O.make_mlp_params(5, 230, 200, 1, 3)
    x = O.make_imagine_inputs(6, N, H)
    start, count = parallel.shard_rows(N, rank, world)
    sl = slice(start, start + count)
    outs = O.imagine(P, PA, x["belief"][sl], x["state"][sl], parallel.shard_time_major(x["eps_action"], rank, world),
                     parallel.shard_time_major(x["eps_prior"], rank, world), H)
    val = O.head_forward(PV, outs[0].flatten(0, 1), outs[1].flatten(0, 1)).reshape(H - 1, count)
    # exact global mean despite uneven shards
    gm = parallel.global_mean(val)
    # gather shards back (padding to the max shard) to compare with the unsharded run on rank 0
    pad = torch.zeros(H - 1, 5, 200)
    pad[:, :count] = outs[0]
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    # flat gradient bucket all-reduce
    g1, g2 = torch.full((3, 2), float(rank + 1)), torch.full((5,), float(10 * (rank + 1)))
    parallel.allreduce_flat([g1, g2])
    t = parallel.max_over_ranks(1.0 + rank, torch.device("cpu"))
    if rank == 0:
        full = O.imagine(P, PA, x["belief"], x["state"], x["eps_action"], x["eps_prior"], H)
        fval = O.head_forward(PV, full[0].flatten(0, 1), full[1].flatten(0, 1))
        beliefs = torch.cat([gathered[r][:, :parallel.shard_rows(N, r, world)[1]] for r in range(world)], 1)
        out["beliefs_equal"] = bool(torch.allclose(beliefs, full[0], rtol=1e-5, atol=1e-6))  # MKL blocks by batch size
        out["mean_err"] = abs(gm.item() - fval.mean().item())
        out["bucket"] = (g1.flatten().tolist(), g2.tolist())
        out["tmax"] = t
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_row_sharding_matches_single_process():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["beliefs_equal"]                      # shards reproduce the unsharded run
    assert out["mean_err"] < 1e-6
    assert out["bucket"] == ([3.0] * 6, [30.0] * 5)
    assert out["tmax"] == 2.0
