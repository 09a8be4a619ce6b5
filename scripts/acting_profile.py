"""Kernel-level breakdown (torch.profiler) of one acting step: Agent.update_latent_and_select_action on one frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from repo_b200 import synth
from repo_b200.trainer import Agent, Config

dev = torch.device("cuda:0")
agent = Agent(Config(), 6, algo="repo", device=dev)
agent.transition_model.load_state_dict(synth.make_transition_params(1))
batch = {k: v.to(dev) for k, v in synth.make_train_batch(7, 4, 2, 6).items()}
frame = batch["obs"][0, :1].contiguous()
lat = list(agent.init_latent_and_action())


def step():
    lat[0], lat[1], lat[2] = agent.update_latent_and_select_action(lat[0], lat[1], lat[2], frame)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
evs = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:25]
for e in evs:
    if e.self_device_time_total > 0:
        print(f"{e.self_device_time_total:9.1f} us  x{e.count:<4d} {e.key[:100]}")
print(f"{sum(e.self_device_time_total for e in prof.key_averages() if e.self_cpu_time_total == 0):9.1f} us sum of kernel durations")
