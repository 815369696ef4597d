"""Generates tests/golden/resnet_ckpt_model1.npz from the reference's OWN trained checkpoint, weights/model_1.pt
(a `ResNet(2, 16)` state dict, 11 924 parameters), evaluated by the reference's own model/resnet.py in this container:
the checkpoint's tensors (a 48 KB data fixture, not source) + the fp32 outputs on positions from oracle games.
The test then requires this repo's model to LOAD that state dict strictly and reproduce the outputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/model")
import resnet as ref_resnet  # noqa: E402  (the reference's own file)
from oracle import oracle as orc  # noqa: E402

sd = torch.load("/root/reference/weights/model_1.pt", map_location="cpu", weights_only=False)
model = ref_resnet.ResNet(2, 16)
model.load_state_dict(sd, strict=True)
model.eval()
rng = np.random.default_rng(11)
planes = []
g = orc.Game()
ply = 0
while not g.is_terminal() and len(planes) < 16:
    lt = g.legal_tiles()
    if ply % 19 == 0:
        planes.append(g.board_state().astype(np.float32))
    g.apply(int(lt[rng.integers(len(lt))]))
    ply += 1
x = torch.from_numpy(np.stack(planes))
with torch.no_grad():
    policy, value = model(x)
out = {"planes": x.numpy(), "policy": policy.numpy(), "value": value.numpy()}
for k, v in sd.items():
    out["sd/" + k] = v.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resnet_ckpt_model1.npz"), **out)
print("wrote", len(planes), "positions; policy max", float(policy.max()), "value[0]", value[0].tolist())
