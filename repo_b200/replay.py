"""`SequenceReplayBuffer` with a device-side sample path (reference: common/buffers.py:128-202).

Host side is the reference's numpy ring, field for field (so `save`/`load` keep the `buffer.npz`
format, buffers.py:193-202) and `sample()` consumes `np.random` exactly like the reference.
`sample_device()` draws the same start indices on the host, then does the index bookkeeping
(time-major order, `(idx + pos) % len` wrap), the gather, `preprocess` (x/255*2-1) and
`nonterms = 1 - dones` in one kernel on a device mirror of the ring — uploading 12 KB uint8 frames as
they are pushed instead of 122.9 MB of float32 per training step."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

_FIELDS = ("capacity", "observations", "actions", "rewards", "dones", "pos", "full")


class SequenceReplayBuffer:
    def __init__(self, capacity, obs_shape, act_shape, obs_type=np.float32, act_type=np.float32):
        self.capacity = capacity
        self.observations = np.zeros((self.capacity,) + tuple(obs_shape), dtype=obs_type)
        self.actions = np.zeros((self.capacity,) + tuple(act_shape), dtype=act_type)
        self.rewards = np.zeros((self.capacity, 1), dtype=np.float32)
        self.dones = np.zeros((self.capacity, 1), dtype=np.float32)
        self.pos = 0
        self.full = False
        self._dev: Optional[dict] = None
        self._dirty: list = []  # ring slots written since the last device sync

    def __len__(self):
        return self.capacity if self.full else self.pos

    def push(self, obs, act, rew, done):
        self.observations[self.pos] = np.array(obs).copy()
        self.actions[self.pos] = np.array(act).copy()
        self.rewards[self.pos] = np.array(rew).copy()
        self.dones[self.pos] = np.array(done).copy()
        if self._dev is not None:   # without a device mirror the first sync uploads everything anyway
            self._dirty.append(self.pos)
        self.pos += 1
        if self.pos == self.capacity:
            self.pos = 0
            self.full = True

    # ------------------------------------------------------------------ reference (host) path
    def _batch_inds(self, start_inds, seq_len):
        batch_inds = np.stack([np.arange(s, s + seq_len) for s in start_inds], 0)
        batch_inds = batch_inds.transpose().reshape(-1)
        if self.full:
            batch_inds = (batch_inds + self.pos) % len(self)  # never straddles the write head
        return batch_inds

    def sample(self, batch_size, seq_len):
        start_inds = np.random.choice(len(self) - seq_len, size=batch_size)
        batch = self._get_samples(self._batch_inds(start_inds, seq_len))
        return tuple(data.reshape(seq_len, batch_size, *data.shape[1:]) for data in batch)

    def iterate(self, batch_size, seq_len):
        all_start_inds = np.arange(0, len(self) - seq_len, seq_len)
        if self.full:
            all_start_inds = (all_start_inds + self.pos) % len(self)
        np.random.shuffle(all_start_inds)
        for i in range(0, len(all_start_inds) - batch_size, batch_size):
            batch = self._get_samples(self._batch_inds(all_start_inds[i:i + batch_size], seq_len))
            yield [data.reshape(seq_len, batch_size, *data.shape[1:]) for data in batch]

    def _get_samples(self, batch_inds):
        return (self.observations[batch_inds], self.actions[batch_inds], self.rewards[batch_inds], self.dones[batch_inds])

    def save(self, path):
        np.savez(path, **{k: getattr(self, k) for k in _FIELDS})

    def load(self, path):
        with np.load(path) as buffer:
            for key in _FIELDS:
                setattr(self, key, buffer[key])
        self.capacity, self.pos, self.full = int(self.capacity), int(self.pos), bool(self.full)
        if self.pos > 0 or self.full:
            self.dones[self.pos - 1] = 1  # buffers.py:200-202
        self._dev, self._dirty = None, []

    # ------------------------------------------------------------------ device path
    def _sync_device(self, device):
        if self.observations.dtype != np.uint8:
            raise RuntimeError("sample_device: the fused gather+preprocess kernel takes uint8 pixel observations")
        if self._dev is None or self._dev["device"] != device:
            self._dev = dict(device=device,
                             obs=torch.from_numpy(self.observations.reshape(self.capacity, -1)).to(device),
                             act=torch.from_numpy(self.actions.reshape(self.capacity, -1).astype(np.float32)).to(device),
                             rew=torch.from_numpy(self.rewards.reshape(-1)).to(device),
                             done=torch.from_numpy(self.dones.reshape(-1)).to(device))
            self._dirty = []
            return
        if self._dirty:
            idx = np.unique(np.asarray(self._dirty, dtype=np.int64))
            t = torch.from_numpy(idx).to(device)
            d = self._dev
            d["obs"][t] = torch.from_numpy(self.observations[idx].reshape(len(idx), -1)).to(device)
            d["act"][t] = torch.from_numpy(self.actions[idx].reshape(len(idx), -1).astype(np.float32)).to(device)
            d["rew"][t] = torch.from_numpy(self.rewards[idx].reshape(-1)).to(device)
            d["done"][t] = torch.from_numpy(self.dones[idx].reshape(-1)).to(device)
            self._dirty = []

    def sample_device(self, batch_size, seq_len, device="cuda", return_indices=False):
        """(obs, actions, rewards, nonterms) as the trainer wants them after dreamer.py:385-391:
        obs float32 preprocessed (L,B,*obs_shape), actions (L,B,A), rewards (L,B,1), nonterms (L,B,1)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("sample_device needs a CUDA device (no CPU fallback); use sample() on the host")
        start_inds = np.random.choice(len(self) - seq_len, size=batch_size)  # same RNG draw as sample()
        self._sync_device(device)
        d = self._dev
        L = _lib.lib()
        frame = d["obs"].shape[1]
        A = d["act"].shape[1]
        starts = torch.from_numpy(start_inds.astype(np.int64)).to(device)
        obs = torch.empty((seq_len, batch_size) + self.observations.shape[1:], dtype=torch.float32, device=device)
        act = torch.empty(seq_len, batch_size, A, dtype=torch.float32, device=device)
        rew = torch.empty(seq_len, batch_size, 1, dtype=torch.float32, device=device)
        nt = torch.empty(seq_len, batch_size, 1, dtype=torch.float32, device=device)
        inds = torch.empty(seq_len * batch_size, dtype=torch.int64, device=device) if return_indices else None
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = L.repo_b200_replay_gather(p(d["obs"]), p(d["act"]), p(d["rew"]), p(d["done"]), p(starts), batch_size, seq_len,
                                       int(self.pos), int(self.full), int(len(self)), frame, A, p(obs), p(act), p(rew), p(nt),
                                       p(inds), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "repo_b200_replay_gather")
        act = act.reshape((seq_len, batch_size) + self.actions.shape[1:])
        return (obs, act, rew, nt, inds) if return_indices else (obs, act, rew, nt)
