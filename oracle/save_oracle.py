"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's `save()`
(model/training.py:70-119): finished-game tuple -> (states, policies, values) training arrays.
The reference's module cannot be imported here (it needs torchrl / tensordict / the PyO3 wheel), so the
loop is restated line for line with numpy (np.rot90 == torch.rot90 for the same k and plane)."""
import numpy as np

DIM = 20


def save_arrays(game):
    history, policies, values = game
    num_moves = len(history)
    state_data = np.zeros((num_moves, 5, DIM, DIM), dtype=np.float32)          # training.py:78
    policy_data = np.zeros((num_moves, DIM * DIM), dtype=np.float32)           # :79
    value_data = np.tile(np.asarray(values, dtype=np.float32), (num_moves, 1)) # :80
    new_state = np.zeros((5, DIM, DIM), dtype=np.float32)                      # :84
    for i, (move, policy) in enumerate(zip(history, policies)):
        player, tile = move
        state_data[i] = np.concatenate((new_state[player:4], new_state[:player], new_state[4][None]), axis=0)  # :89
        for action, prob in policy:                                            # :92-98
            policy_data[i, action] = prob
            state_data[i, 4, action // DIM, action % DIM] = 1
        state_data[i] = np.rot90(state_data[i], k=player, axes=(1, 2))         # :101
        policy_data[i] = np.rot90(policy_data[i].reshape(DIM, DIM), k=player).reshape(-1)  # :102
        new_state[player, tile // DIM, tile % DIM] = 1                         # :105-106
    return state_data, policy_data, value_data
