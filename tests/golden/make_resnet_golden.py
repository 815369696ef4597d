"""Generates tests/golden/resnet_2x16.npz by IMPORTING the reference's model/resnet.py in this container
(run once here; /root/reference does not exist on the GPU box).  Random-init ResNet(2,16) under a fixed
seed with perturbed BatchNorm statistics, eval mode, fp32, on planes taken from oracle games."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/model")
import resnet as ref_resnet  # noqa: E402  (the reference's own file)
from oracle import oracle as orc  # noqa: E402

torch.manual_seed(1234)
model = ref_resnet.ResNet(2, 16)
with torch.no_grad():
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.3)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.uniform_(0.5, 1.5)
            m.bias.normal_(0, 0.2)
model.eval()

rng = np.random.default_rng(5)
planes = []
g = orc.Game()
ply = 0
while not g.is_terminal() and len(planes) < 12:
    lt = g.legal_tiles()
    if ply % 23 == 0:
        planes.append(g.board_state().astype(np.float32))
    g.apply(int(lt[rng.integers(len(lt))]))
    ply += 1
x = torch.from_numpy(np.stack(planes))
with torch.no_grad():
    policy, value = model(x)
out = {"planes": x.numpy(), "policy": policy.numpy(), "value": value.numpy()}
for k, v in model.state_dict().items():
    out["sd/" + k] = v.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resnet_2x16.npz"), **out)
print("wrote", len(planes), "positions;", sum(v.size for k, v in out.items() if k.startswith("sd/")), "parameters")
