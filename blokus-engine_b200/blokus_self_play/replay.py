"""Device-resident replay buffer for the training tuples: the consumer side of
`model/training.py` — `ReplayBuffer(storage=LazyTensorStorage(capacity), batch_size=...)` with `buffer.extend(data)`
in `save()` (training.py:114-119) and `buffer.sample()` in `train()` (training.py:126-129) — kept in HBM, so the
tensors `bk_selfplay_training_tensors` builds on the device never visit the host.

Plumbing only (PyTorch tensors as device memory): a ring of `capacity` samples, uniform sampling with replacement like
torchrl's default sampler.  States are stored as bytes (they are 0/1 planes) and handed out as float32, so a million
positions take 2 GB + 1.6 GB (policies) + 16 MB (scores) of the 180 GB.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class Batch(dict):
    """What `buffer.sample()` returns: `batch.get("states")`, `.get("policies")`, `.get("scores")` (training.py:127-129)."""


class DeviceReplayBuffer:
    def __init__(self, capacity: int, batch_size: int, device="cuda", seed: Optional[int] = None):
        self.capacity, self.batch_size = int(capacity), int(batch_size)
        self.device = torch.device(device)
        self.states = torch.zeros((self.capacity, 5, 20, 20), dtype=torch.uint8, device=self.device)
        self.policies = torch.zeros((self.capacity, 400), dtype=torch.float32, device=self.device)
        self.scores = torch.zeros((self.capacity, 4), dtype=torch.float32, device=self.device)
        self.size = 0            # samples held
        self.cursor = 0          # next slot to write (ring)
        self.gen = torch.Generator(device=self.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))

    def __len__(self) -> int:
        return self.size

    def extend(self, data: Dict[str, torch.Tensor]) -> None:
        """`buffer.extend(Data(states=..., policies=..., scores=...))` (training.py:112-119); the oldest samples are
        overwritten once the ring is full."""
        st, po, sc = data["states"], data["policies"], data["scores"]
        n = int(st.shape[0])
        if not (po.shape[0] == n and sc.shape[0] == n):
            raise ValueError("states / policies / scores must have the same number of samples")
        if n > self.capacity:                                  # only the newest `capacity` samples can stay
            st, po, sc, n = st[-self.capacity:], po[-self.capacity:], sc[-self.capacity:], self.capacity
        first = min(n, self.capacity - self.cursor)
        for dst, src in ((self.states, st), (self.policies, po), (self.scores, sc)):
            src = src.to(device=self.device, dtype=dst.dtype)
            dst[self.cursor:self.cursor + first] = src[:first]
            if n > first:
                dst[:n - first] = src[first:]
        self.cursor = (self.cursor + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def extend_from(self, selfplay) -> int:
        """`for game in games: save(game, buffer)` for every finished game of a SelfPlay batch, on the device."""
        states, policies, values, _ = selfplay.training_tensors()
        self.extend({"states": states, "policies": policies, "scores": values})
        return int(states.shape[0])

    def sample(self, batch_size: Optional[int] = None) -> Batch:
        if self.size == 0:
            raise RuntimeError("the replay buffer is empty")
        b = int(batch_size or self.batch_size)
        idx = torch.randint(0, self.size, (b,), device=self.device, generator=self.gen)
        return Batch(states=self.states[idx].to(torch.float32), policies=self.policies[idx], scores=self.scores[idx])
