"""Generates tests/golden/resnet_20x256.npz by IMPORTING the reference's model/resnet.py in this container: the
network of BASELINE.json config 4, ResNet(20, 256), eval mode, fp32 on the CPU, with the deterministic parameters of
tests/resnet_params.py (regenerated on the GPU box; only inputs and fp32 OUTPUTS are stored).

Stored: 256 positions from oracle games (bit-packed planes), and from the reference model: the policy head's
pre-softmax logits [256,400], the value head's tanh output [256,4], policy [256,400], value [256,4].

    python tests/golden/make_resnet20_golden.py      (from the repo root; ~1 min of CPU)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/model")
import resnet as ref_resnet  # noqa: E402  (the reference's own file)
from oracle import oracle as orc  # noqa: E402
from resnet_params import make_state_dict  # noqa: E402

SEED = 2026
model = ref_resnet.ResNet(20, 256)
model.load_state_dict(make_state_dict(model.state_dict(), SEED), strict=True)
model.eval()

rng = np.random.default_rng(9)
planes = []
for game in range(16):                     # 16 positions from each of 16 random games, every 17th ply
    g = orc.Game()
    ply = 0
    while not g.is_terminal() and len(planes) < 16 * (game + 1):
        lt = g.legal_tiles()
        if ply % 17 == 3:
            planes.append(g.board_state().astype(np.uint8))
        g.apply(int(lt[rng.integers(len(lt))]))
        ply += 1
x8 = np.stack(planes)
x = torch.from_numpy(x8.astype(np.float32))
# A random-init head gives near-constant logits (range 0..0.1): nothing for a parity test to bite on.  Calibrate the two
# head BatchNorms the way training would — running statistics = the statistics of their input on these positions — and
# give them an affine that spreads the policy logits over several units.  The eight scalars are stored with the outputs.
overrides = {}
with torch.no_grad():
    h = model.input(x)
    for blk in model.res_blocks:
        h = blk(h)
    for head, (gamma, beta) in (("policy_head", (2.0, 0.5)), ("value_head", (1.0, 0.5))):
        z = getattr(model, head)[0](h)
        overrides[head + ".1.running_mean"] = z.mean().reshape(1)
        overrides[head + ".1.running_var"] = z.var(unbiased=False).reshape(1)
        overrides[head + ".1.weight"] = torch.tensor([gamma])
        overrides[head + ".1.bias"] = torch.tensor([beta])
sd = model.state_dict()
sd.update({k: v.float() for k, v in overrides.items()})
model.load_state_dict(sd, strict=True)
with torch.no_grad():
    h = model.input(x)
    for blk in model.res_blocks:
        h = blk(h)
    logits = model.policy_head(h)
    vtanh = model.value_head(h)
    policy, value = model(x)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resnet_20x256.npz"), seed=SEED, n=len(planes),
                    planes_bits=np.packbits(x8.reshape(-1)), logits=logits.numpy(), vtanh=vtanh.numpy(),
                    policy=policy.numpy(), value=value.numpy(), **{"override/" + k: v.numpy() for k, v in overrides.items()}, trunk_absmax=float(h.abs().max()), trunk_mean=float(h.mean()))
print("wrote", len(planes), "positions; trunk |max|", float(h.abs().max()), "mean", float(h.mean()),
      "logit range", float(logits.min()), float(logits.max()), "legal counts", x8[:, 4].reshape(len(planes), -1).sum(1).tolist())
