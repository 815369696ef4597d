"""The N>1 path on CPU: two gloo ranks each play their shard (kernel sources on the CPU warp emulator),
rank 0 gathers the finished-game tuples; the result must equal one rank playing all ids."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from blokus_self_play.shard import shard_range
    for n_total in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n_total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == n_total
            for (a, na), (b, _) in zip(spans, spans[1:]):
                assert a + na == b
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    assert shard_range(65536, 3, 8) == (3 * 8192, 8192)        # config 5: 8192 games per GPU
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n_total, seed, emu_path, out_path):
    sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blokus_self_play import GameBatch, Lib, PLAYOUT_HASH
    from blokus_self_play.shard import shard_range, gather_finished, reduce_max_sum
    first, n = shard_range(n_total, rank, world)
    batch = GameBatch(n, lib=Lib(emu_path))
    res = batch.playout(seed=seed, first_game_id=first, flags=PLAYOUT_HASH)
    local = [(first + g, int(res["hash"][g]), int(res["steps"][g]), batch.scores()[g].tolist()) for g in range(n)]
    allg = gather_finished(local, dst=0)
    (tmax,), (steps,) = reduce_max_sum([float(rank + 1)], [float(res["total_steps"])])
    if rank == 0:
        np.save(out_path, np.array([[gid, h & 0xFFFFFFFF, h >> 32, st] + sc for gid, h, st, sc in allg], dtype=np.int64))
        assert tmax == float(world) and steps == sum(st for _, _, st, _ in allg)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_one(tmp_path, emu_lib, orc):
    n_total, seed = 3, 41
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, n_total, seed, emu_lib.path, out), nprocs=2, join=True)
    got = np.load(out)
    assert got[:, 0].tolist() == list(range(n_total))           # gathered in global-id order
    for row in got:
        ref = orc.playout(seed, int(row[0]), 0)
        h = int(row[1]) | (int(row[2]) << 32)
        assert h == ref["hash"] and int(row[3]) == ref["n_plies"] and row[4:].tolist() == list(ref["scores"])
