"""Multi-rank host logic on CPU: world_size-2 and -3 `gloo` runs of the video sharding plan
and the single gather step.  The per-video row tables come from the CPU
tracker oracle on fixture detections, so the check is the one the north star asks for:
the R-rank gathered output equals the 1-rank output byte for byte."""
import os
import socket

import numpy as np
import pytest

import helpers
from vbt_b200 import shard


def test_lpt_assign_balances_and_is_deterministic():
    counts = [len(helpers.fixture_detections(n)[2]) for n in sorted(helpers.golden_tables())]
    for world in (1, 2, 4, 8):
        plan = shard.lpt_assign(counts, world)
        assert sorted(i for p in plan for i in p) == list(range(len(counts)))
        loads = [sum(counts[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(counts)          # LPT bound
        assert plan == shard.lpt_assign(counts, world)
    assert shard.lpt_assign([5, 5, 5], 2) == [[0, 2], [1]]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, names, out_dir):
    import torch.distributed as dist
    from oracle import ocsort as oo
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        vids = [helpers.fixture_detections(n) for n in names]
        plan = shard.lpt_assign([len(v[2]) for v in vids], world)
        local = {i: oo.track_rows(vids[i][2], vids[i][0], frame_numbers=vids[i][1]) for i in plan[rank]}
        full = shard.gather_row_tables(local, len(names))
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), **{str(k): v for k, v in full.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gathered_tables_equal_single_rank(tmp_path, world):
    import torch.multiprocessing as mp
    from oracle import ocsort as oo
    names = sorted(helpers.golden_tables())[:7]
    mp.spawn(_worker, args=(world, _free_port(), names, str(tmp_path)), nprocs=world, join=True)
    vids = [helpers.fixture_detections(n) for n in names]
    want = [oo.track_rows(v[2], v[0], frame_numbers=v[1]) for v in vids]
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f'rank{rank}.npz'))
        assert sorted(z.files, key=int) == [str(i) for i in range(len(names))]
        for i, w in enumerate(want):
            assert z[str(i)].tobytes() == w.tobytes(), (rank, names[i])


def test_gather_without_process_group_is_identity():
    t = {0: np.arange(16.0).reshape(2, 8), 2: np.zeros((0, 8))}
    out = shard.gather_row_tables(t, 3)
    assert np.array_equal(out[0], t[0]) and out[2].shape == (0, 8)
