"""Leaf evaluator for BASELINE.json config 4: the reference's policy/value network, consumed as a LIBRARY
model (PyTorch / cuDNN — the one dense contraction on the path; SURVEY.md §8 row a25, "next" row f2 is a
hand-written tcgen05 version).  Architecture and parameter names follow model/resnet.py:8-94 so the
reference's checkpoints (weights/model_*.pt, `ResNet(blocks, width)`) load with load_state_dict:

    input: conv3x3(5 -> W)                                   (no BN / ReLU after it, as in the reference)
    res_blocks[i]: conv1-bn1-ReLU-conv2-bn2, + skip, ReLU
    policy_head: conv1x1(W -> 1) - BN - ReLU - flatten, then softmax over LEGAL tiles only
                 (logits * mask, -1e9 elsewhere), result * mask  -> 0 on illegal tiles
    value_head:  conv1x1(W -> 1) - BN - ReLU - flatten - Linear(400, 4) - tanh, then softmax over the 4 seats
"""
from __future__ import annotations

import torch
from torch import nn

BOARD = 20


class ResidualBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(self.bn2(self.conv2(h)) + x)


def _head(width: int, tail=()) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(width, 1, 1), nn.BatchNorm2d(1), nn.ReLU(), nn.Flatten(), *tail)


class ResNet(nn.Module):
    """ResNet(blocks, width): boards[B,5,20,20] -> (policy[B,400], value[B,4]); argument order of model/resnet.py:44."""

    def __init__(self, blocks: int, width: int, custom_filters: bool = False):
        super().__init__()
        self.blocks, self.width = blocks, width
        self.custom_filters = custom_filters      # accepted and unused, as in the reference (model/resnet.py:44-49; training.py:169)
        self.input = nn.Conv2d(5, width, 3, padding=1)
        self.res_blocks = nn.ModuleList(ResidualBlock(width, width) for _ in range(blocks))
        self.policy_head = _head(width)
        self.value_head = _head(width, (nn.Linear(BOARD * BOARD, 4), nn.Tanh()))

    def forward(self, boards: torch.Tensor):
        x = self.input(boards)
        for blk in self.res_blocks:
            x = blk(x)
        legal = boards[:, 4].reshape(boards.shape[0], -1).to(x.dtype)
        logits = self.policy_head(x)
        policy = torch.softmax(logits * legal + (1 - legal) * -1e9, dim=1) * legal
        value = torch.softmax(self.value_head(x), dim=1)
        return policy, value


class LeafEvaluator:
    """Callable evaluator for SelfPlay.run_evaluator: one forward pass per batch of pending leaves.
    `.eval()` mode (the reference's trainer leaves BatchNorm in training mode during self-play,
    model/training.py:59-60 — a deviation SURVEY.md §8d asks to state); optional bf16 autocast + channels_last."""

    def __init__(self, model: ResNet, bf16: bool = False):
        self.model = model.eval()
        self.bf16 = bf16
        if bf16:
            self.model = self.model.to(memory_format=torch.channels_last)

    @torch.no_grad()
    def __call__(self, planes: torch.Tensor):
        if self.bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                p, v = self.model(planes.contiguous(memory_format=torch.channels_last))
            return p.float(), v.float()
        return self.model(planes)
