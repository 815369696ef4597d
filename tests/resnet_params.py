"""Deterministic parameters for golden vectors of LARGE networks (ResNet(20,256) has 23.6 M parameters = 94 MB of
fp32: too big to commit).  Every tensor of the state dict is regenerated from (seed, its name) with numpy's PCG64,
so the container that writes the golden outputs (importing the reference's model/resnet.py) and the GPU box that
checks them build bit-identical weights without shipping them.  Distributions follow PyTorch's default initialisers
(uniform +-1/sqrt(fan_in) for convolution / linear weights and biases), with non-trivial BatchNorm statistics so the
eval-mode folding is exercised."""
import zlib

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def make_state_dict(reference_state_dict, seed: int):
    """reference_state_dict: any state dict with the target names/shapes (values are ignored)."""
    out = {}
    for name, t in reference_state_dict.items():
        shape = tuple(t.shape)
        r = _rng(seed, name)
        if name.endswith("num_batches_tracked"):
            v = np.zeros(shape, dtype=np.int64)
        elif name.endswith("running_var"):
            v = r.uniform(0.8, 1.25, size=shape).astype(np.float32)
        elif name.endswith("running_mean"):
            v = r.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        elif ".bn" in name or "head.1." in name:                     # BatchNorm affine
            v = (r.uniform(0.8, 1.25, size=shape) if name.endswith("weight") else r.uniform(-0.1, 0.1, size=shape)).astype(np.float32)
        else:                                                         # Conv2d / Linear
            base = name.rsplit(".", 1)[0] + ".weight"
            wshape = tuple(reference_state_dict[base].shape)
            fan_in = int(np.prod(wshape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            v = r.uniform(-b, b, size=shape).astype(np.float32)
        out[name] = torch.from_numpy(v)
    return out


def unpack_planes(bits: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(bits)[: n * 2000].reshape(n, 5, 20, 20).astype(np.float32)
