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


def _heads_worker(rank, world, port, out):
    """Data-parallel wiring of the optional heads: each rank takes a block of batch columns, weights its loss by
    kept_local / kept_global (Agent._head_step, one all-reduce of the row counts) and the SUM of the per-rank gradients
    must equal the single-process gradient.  The CUDA GEMM op is replaced by torch on CPU — host logic only."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import repo_b200.autograd as ag
    from repo_b200 import synth
    from repo_b200.trainer import Agent, Config

    class TorchLinear:
        @staticmethod
        def apply(x, w, b):
            return torch.nn.functional.linear(x, w, b)

    ag.LinearFn = TorchLinear
    T, B = 9, 37
    x = synth.make_head_rollout(702, T, B)
    cols = slice(0, 20) if rank == 0 else slice(20, B)          # uneven shards, different numbers of kept rows
    torch.manual_seed(0)
    agent = Agent(Config(disag_model=True, inv_dynamics=True), 6, algo="dreamer", device="cpu")
    agent.disag_model.load_state_dict(synth.make_ensemble_params(700, 236, 200, 200, 6))
    agent.train_disag(x["beliefs"][:, cols], x["states"][:, cols], x["actions"][:, cols], x["nonterms"][:, cols], step=False)
    grads = [p.grad.clone() for p in agent.disag_model.parameters()]
    parallel.allreduce_flat(grads)
    loss = agent.reduced_logs()["train/disag_loss"]
    if rank == 0:
        dist.barrier()
        # single-process run of the same call on the whole batch (no process-group weighting: compare by hand)
        keep = x["nonterms"][1:-1].flatten() == 1
        a, b, s, bn = [t.flatten(0, 1)[keep] for t in (x["actions"][1:-1], x["beliefs"][:-1], x["states"][:-1], x["beliefs"][1:])]
        ref = Agent(Config(disag_model=True), 6, algo="dreamer", device="cpu")
        ref.disag_model.load_state_dict(synth.make_ensemble_params(700, 236, 200, 200, 6))
        pred = ref.disag_model(b, s, a)
        full = (0.5 * (pred - bn.unsqueeze(0)) ** 2 + 0.5 * np.log(2 * np.pi)).sum(2).sum(0).mean()
        full.backward()
        out["loss_err"] = abs(loss.item() - full.item()) / abs(full.item())
        out["grad_err"] = max(float((g - p.grad).abs().max() / (p.grad.abs().max() + 1e-12))
                              for g, p in zip(grads, ref.disag_model.parameters()))
    else:
        dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_optional_heads_weighting():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_heads_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["loss_err"] < 1e-5
    assert out["grad_err"] < 1e-4
