"""GPU, 2 ranks over NCCL: data-parallel gradients of the RSSM path.  Each rank owns a row shard,
weights its loss by rows_local / rows_global, all-reduces ONE flat bucket per parameter group, and must end up
with the single-process gradients.  Skipped when fewer than two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _models(dev, seed=500):
    from oracle import rssm_oracle as O
    from repo_b200.models import ActorModel
    from repo_b200.rssm import TransitionModel
    tm = TransitionModel(200, 30, 6, 200, 1024, "elu").to(dev)
    tm.load_state_dict(O.make_transition_params(seed))
    actor = ActorModel(200, 30, 200, 6, "elu").to(dev)
    actor.load_state_dict(O.make_mlp_params(seed + 1, 230, 200, 12, 4))
    return tm, actor


def _losses(tm, actor, xo, xi, cols, rows, n_cols, n_rows, dev):
    """world-model-like loss on observe columns `cols` + actor-like loss on imagine rows `rows` (means over rows)."""
    g = lambda k: xo[k][:, cols].to(dev) if xo[k].dim() == 3 else xo[k][cols].to(dev)
    outs = tm.observe(g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"),
                      eps_prior=g("eps_prior"), eps_post=g("eps_post"))
    wm = (tm.last_kl.detach() * 0 + outs[0].pow(2).sum(2) + outs[4].sum(2) + outs[6].sum(2) + outs[3].sum(2)).sum() / (outs[0].shape[0] * n_cols)
    traj = tm.imagine(xi["belief"][rows].to(dev), xi["state"][rows].to(dev), actor, 5,
                      eps_action=xi["eps_action"][:, rows].to(dev), eps_prior=xi["eps_prior"][:, rows].to(dev))
    ac = (traj[0].sum(2) + traj[1].pow(2).sum(2)).sum() / (4 * n_rows)
    return wm + ac


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import rssm_oracle as O
    from repo_b200 import parallel
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    tm, actor = _models(dev)
    B, N = 7, 45  # uneven shards: 4+3 columns, 23+22 rows
    xo, xi = O.make_observe_inputs(510, 6, B), O.make_imagine_inputs(511, N, 5)
    c0, cn = parallel.shard_rows(B, rank, world)
    r0, rn = parallel.shard_rows(N, rank, world)
    _losses(tm, actor, xo, xi, slice(c0, c0 + cn), slice(r0, r0 + rn), B, N, dev).backward()
    parallel.allreduce_flat([p.grad for p in tm.parameters()])
    parallel.allreduce_flat([p.grad for p in actor.parameters()])
    if rank == 0:
        tm1, actor1 = _models(dev)
        _losses(tm1, actor1, xo, xi, slice(0, B), slice(0, N), B, N, dev).backward()
        worst = 0.0
        for (k, p), p1 in zip(list(tm.named_parameters()) + list(actor.named_parameters()),
                              list(tm1.parameters()) + list(actor1.parameters())):
            scale = float(p1.grad.abs().max()) + 1e-12
            worst = max(worst, float((p.grad - p1.grad).abs().max()) / scale)
        out["worst_rel"] = worst
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_gradients_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["worst_rel"] < 2e-4, out["worst_rel"]


def _agent(dev, algo, B):
    from oracle import rssm_oracle as O
    from repo_b200.trainer import Agent, Config
    torch.manual_seed(0)
    agent = Agent(Config(batch_size=B, chunk_size=5, free_nats=0.1, init_beta=0.3), 6, algo=algo, device=dev)
    agent.transition_model.load_state_dict(O.make_transition_params(600))
    agent.reward_model.load_state_dict(O.make_mlp_params(602, 230, 200, 1, 3))
    agent.encoder.load_state_dict(O.make_conv_params("encoder", 604))
    agent.obs_model.load_state_dict(O.make_conv_params("decoder", 605))
    agent.actor_model.load_state_dict(O.make_mlp_params(601, 230, 200, 12, 4))
    agent.value_model.load_state_dict(O.make_mlp_params(603, 230, 200, 1, 3))
    return agent


def _agent_step(agent, dev, cols, B, T=5):
    """one train_dynamics + train_actor_critic (with optimiser steps) on batch columns `cols` of a fixed global batch"""
    from oracle import rssm_oracle as O
    batch = O.make_train_batch(610, T, B, 6)
    eps = O.make_observe_inputs(611, T, B)
    g = lambda x: x[:, cols].contiguous().to(dev)
    b, s = agent.train_dynamics(g(batch["obs"]), g(batch["actions"]), g(batch["rewards"]), g(batch["nonterms"]),
                                eps_prior=g(eps["eps_prior"]), eps_post=g(eps["eps_post"]))
    n = (T - 1) * B
    xi = O.make_imagine_inputs(612, n, agent.c.horizon)
    rows = torch.arange(n).reshape(T - 1, B)[:, cols].reshape(-1)  # time-major flatten of this rank's columns
    ent = torch.from_numpy(np.random.RandomState(613).standard_normal((100, (agent.c.horizon - 1), n, 6)).astype(np.float32))
    agent.train_actor_critic(b.flatten(0, 1), s.flatten(0, 1), eps_action=xi["eps_action"][:, rows].to(dev),
                             eps_prior=xi["eps_prior"][:, rows].to(dev),
                             eps_entropy=ent[:, :, rows].reshape(100, -1, 6).contiguous().to(dev))


def _agent_worker(rank, world, port, out, algo):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from repo_b200 import parallel
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 5  # uneven shards: 3 + 2 batch columns
    agent = _agent(dev, algo, B)
    c0, cn = parallel.shard_rows(B, rank, world)
    _agent_step(agent, dev, slice(c0, c0 + cn), B)
    logs = agent.reduced_logs()
    params = torch.cat([p.detach().reshape(-1) for p in agent.model_params + list(agent.actor_model.parameters())
                        + list(agent.value_model.parameters())] + [agent.log_beta.detach().reshape(1)])
    dist.barrier()
    dist.destroy_process_group()  # the single-process run below must not see a process group
    if rank == 0:
        ref = _agent(dev, algo, B)
        start = torch.cat([p.detach().reshape(-1) for p in ref.model_params + list(ref.actor_model.parameters())
                           + list(ref.value_model.parameters())] + [ref.log_beta.detach().reshape(1)]).clone()
        _agent_step(ref, dev, slice(0, B), B)
        want = torch.cat([p.detach().reshape(-1) for p in ref.model_params + list(ref.actor_model.parameters())
                          + list(ref.value_model.parameters())] + [ref.log_beta.detach().reshape(1)])
        # compare the UPDATES (Adam's first step is +-lr per element: sign agreement is the real check)
        du, dw = params - start, want - start
        moved = dw.abs() > 1e-7
        out["update_mismatch"] = float(((du - dw).abs()[moved] > 0.02 * dw.abs()[moved] + 1e-9).float().mean())
        out["log_err"] = max(abs(float(logs[k]) - float(ref.logs[k])) / (abs(float(ref.logs[k])) + 1e-6) for k in ref.logs)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("algo", ["dreamer", "repo"])
def test_two_gpu_agent_update_matches_single_gpu(algo):
    """Config 3 in miniature: Agent.train_dynamics + train_actor_critic with the batch sharded 3+2 over two ranks
    (weighted losses, flat-bucket all-reduce inside FlatAdam) must move every parameter like the single-GPU update."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_agent_worker, args=(2, _free_port(), out, algo), nprocs=2, join=True)
    assert out["log_err"] < 1e-3, dict(out)
    assert out["update_mismatch"] < 2e-3, dict(out)  # fraction of elements whose Adam step differs (near-zero gradients flip sign)
