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


# ---- one long video, contiguous frame chunks per rank (second scheme) -----------

def test_chunk_bounds_cover_every_frame_once():
    for n in (0, 1, 7, 64, 1800):
        for world in (1, 2, 3, 8):
            b = shard.chunk_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _fixture_table(name, D=25):
    import torch
    fps, numbers, dets = helpers.fixture_detections(name)
    n = len(dets)
    tab = np.zeros((n, D, 6))
    cnt = np.zeros(n, np.int32)
    for i, d in enumerate(dets):
        tab[i, :len(d)] = d
        cnt[i] = len(d)
    return fps, torch.from_numpy(tab), torch.from_numpy(cnt), torch.from_numpy(np.asarray(numbers, np.int32))


def _chunk_worker(rank, world, port, name, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        fps, tab, cnt, nos = _fixture_table(name)
        lo, hi = shard.chunk_bounds(len(tab), world)[rank]
        full = shard.gather_detection_tables(tab[lo:hi].contiguous(), cnt[lo:hi].contiguous(), nos[lo:hi].contiguous())
        np.savez(os.path.join(out_dir, f'chunk{rank}.npz'), dets=full[0].numpy(), counts=full[1].numpy(),
                 numbers=full[2].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gathered_detection_chunks_equal_the_whole_table(tmp_path, world):
    """Each rank holds the detections of its contiguous frame chunk; after the one gather every rank
    holds the whole table in frame order, and the tracker oracle run on it gives the 1-rank rows."""
    import torch.multiprocessing as mp
    from oracle import ocsort as oo
    name = sorted(helpers.golden_tables())[2]
    mp.spawn(_chunk_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    fps, tab, cnt, nos = _fixture_table(name)
    want = oo.track_rows([tab[i, :cnt[i]].numpy() for i in range(len(tab))], fps, frame_numbers=nos.tolist())
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f'chunk{rank}.npz'))
        assert z['dets'].tobytes() == tab.numpy().tobytes()
        assert np.array_equal(z['counts'], cnt.numpy()) and np.array_equal(z['numbers'], nos.numpy())
        got = oo.track_rows([z['dets'][i, :z['counts'][i]] for i in range(len(z['dets']))], fps,
                            frame_numbers=z['numbers'].tolist())
        assert got.tobytes() == want.tobytes()
