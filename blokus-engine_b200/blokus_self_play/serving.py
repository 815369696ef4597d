"""The human-play wire format of the reference's model server (model/model_server.py:34-57,
caller gui/src/app.rs:16-27,55-96) on this package's evaluators.

Request  : {"player": int, "data": [5][20][20] bool}   — Game::get_board_state of the position (mover frame)
Response : {"policy": [400] float, "values": [4] float, "status": 200}

The reference answers one request with one forward pass (`model(boards.unsqueeze(0))`, model_server.py:45-47) and
never reads `player`; `process_requests` is the batched form (one evaluator call for many requests), which is what
a B200 wants.  No web framework is imported here: `process_request` is the body of the POST /process_request
handler, so `app.post("/process_request")(lambda r: process_request(r.dict(), evaluator))` is the whole server.
The evaluator is any callable planes[B,5,20,20] float32 (CUDA tensor) -> (policy[B,400], value[B,4]), e.g.
`LeafEvaluator(model)` or `TensorCoreLeafEvaluator(model)`.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np

DIM = 20


def _planes(payload: Dict) -> np.ndarray:
    if "data" not in payload or "player" not in payload:
        raise ValueError("request needs 'player' and 'data' (model/model_server.py:34-36)")
    a = np.asarray(payload["data"])
    if a.shape != (5, DIM, DIM):
        raise ValueError(f"'data' must be [5][{DIM}][{DIM}], got {a.shape}")
    return a.astype(np.float32)


def process_requests(payloads: Sequence[Dict], evaluator: Callable, device=None) -> List[Dict]:
    """One evaluator call for a batch of /process_request bodies; answers in request order."""
    import torch
    if not payloads:
        return []
    batch = torch.from_numpy(np.stack([_planes(p) for p in payloads]))
    if device is None:
        device = torch.device("cuda", 0)
    with torch.no_grad():
        policy, values = evaluator(batch.to(device))
    policy, values = policy.float().cpu(), values.float().cpu()
    return [{"policy": policy[i].tolist(), "values": values[i].tolist(), "status": 200} for i in range(len(payloads))]


def process_request(payload: Dict, evaluator: Callable, device=None) -> Dict:
    """Body of POST /process_request (model/model_server.py:38-57)."""
    return process_requests([payload], evaluator, device)[0]


def request_from_game(game) -> Dict:
    """What the GUI sends for a position (gui/src/app.rs:55-96): the mover and Game::get_board_state as bools."""
    state = np.asarray(game.get_board_state())
    return {"player": int(game.current_player()), "data": state.astype(bool).tolist()}
