"""Kernel-level breakdown (torch.profiler) of one Agent.train_dynamics and one Agent.train_actor_critic at the
RePo default shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from repo_b200 import synth
from repo_b200.trainer import Agent, Config

dev = torch.device("cuda:0")
algo = sys.argv[1] if len(sys.argv) > 1 else "repo"
agent = Agent(Config(batch_size=7), 6, algo=algo, device=dev)
agent.transition_model.load_state_dict(synth.make_transition_params(1))
agent.optimizers()
batch = {k: v.to(dev) for k, v in synth.make_train_batch(7, 50, 7, 6).items()}
st = {}


def wm():
    st["b"], st["s"] = agent.train_dynamics(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])


def ac():
    agent.train_actor_critic(st["b"].flatten(0, 1), st["s"].flatten(0, 1))


for name, fn in (("train_dynamics", wm), ("train_actor_critic", ac)):
    fn(); fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn(); torch.cuda.synchronize()
    print("=====", name)
    evs = [e for e in prof.key_averages() if e.device_type.name == "CUDA" or e.self_device_time_total > 0]
    evs = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:22]
    tot = sum(e.self_device_time_total for e in prof.key_averages())
    for e in evs:
        if e.self_device_time_total > 0:
            print(f"{e.self_device_time_total/1e3:9.3f} ms  x{e.count:<4d} {e.key[:110]}")
    ksum = sum(e.self_device_time_total for e in prof.key_averages() if e.self_cpu_time_total == 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{ksum/1e3:9.3f} ms sum of kernel durations; {e0.elapsed_time(e1)/5:9.3f} ms wall per call (no profiler)")
