"""Multi-GPU sharding of the hot path: whole videos per rank (or, for one long video, contiguous
frame chunks per rank), one gather at the end.

The reference processes its sources one after the other, each with a fresh interpreter
and a fresh tracker (track.py:88-101,157) and writes one pickle per video
(track.py:117-126): videos are independent units, so the path shards with no data-path
collective.  One process per GPU (`torch.distributed`, NCCL on the GPUs,
gloo in the CPU tests):

* ``lpt_assign``        -- greedy longest-processing-time deal of videos to ranks
                           (frame count is the cost: every frame costs the same);
* ``gather_row_tables`` -- the ONE exchange step: all_gather of per-video row counts,
                           then all_gather of each rank's row tables packed into one
                           padded [rows, 8] float64 buffer.  Never called per frame.
* ``track_video_chunks`` -- ONE long video over all ranks: detection (K1-K6, per-frame
                           independent, ~98 % of the work) on contiguous frame chunks, one gather
                           of the packed detection tables (1.2 kB per frame), then the sequential
                           tracker / velocity recurrence (K7, K8) over the whole table on rank 0.
                           No tracker state crosses ranks, so no hand-off is needed and the result
                           is the 1-rank result byte for byte (second scheme).
"""
from __future__ import annotations

import numpy as np


def lpt_assign(frame_counts, world_size):
    """Video indices per rank.  Longest first, each to the currently lightest rank; ties go
    to the lowest rank / lowest video index so every rank computes the same plan."""
    order = sorted(range(len(frame_counts)), key=lambda i: (-int(frame_counts[i]), i))
    load = [0] * world_size
    plan = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(frame_counts[i])
    return plan


def gather_row_tables(local_tables, n_videos, device=None, group=None):
    """local_tables: {video index: float64 [n,8] row table} of THIS rank's videos.
    Returns {video index: float64 [n,8]} holding every video, identical on every rank.

    Two collectives in total: counts (int64 [n_videos]) and one padded payload."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {k: np.asarray(v, dtype=np.float64).reshape(-1, 8) for k, v in local_tables.items()}
    world = dist.get_world_size(group)
    dev = device if device is not None else (
        torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu'))
    counts = torch.zeros(n_videos, dtype=torch.int64, device=dev)
    for k, v in local_tables.items():
        counts[k] = len(v)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    per_rank = [int(c.sum().item()) for c in all_counts]
    pad = max(max(per_rank), 1)
    payload = torch.zeros((pad, 8), dtype=torch.float64, device=dev)
    at = 0
    for k in sorted(local_tables):
        t = torch.as_tensor(np.asarray(local_tables[k], dtype=np.float64).reshape(-1, 8), device=dev)
        payload[at:at + len(t)] = t
        at += len(t)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    out = {}
    for r in range(world):
        c = all_counts[r].cpu().numpy()
        buf = gathered[r].cpu().numpy()
        at = 0
        for k in np.nonzero(c)[0]:
            out[int(k)] = buf[at:at + int(c[k])].copy()
            at += int(c[k])
    for k in range(n_videos):
        out.setdefault(k, np.zeros((0, 8)))
    return out


def track_videos(videos, detector, detection_threshold=0.5, frame_stride=1, rank=None, world=None,
                 group=None, gather=True, **pipe_kw):
    """Run the hot path over many videos, sharded by whole videos across the ranks.

    videos: list of dicts ``{'fps': float, 'frames': uint8 [N,H,W,3] CUDA or pinned-host tensor}``
    (BGR, as decoded).  Each rank processes the videos `lpt_assign` gives it with one
    `VideoPipeline` (tracker and velocity state reset per video, like the fresh interpreter and
    tracker per source of track.py:88-101,157), then ONE gather makes every row table visible on
    every rank.  Returns ``(tables, phases)``: tables = {video index: float64 [n,8]} for all
    videos, phases = {video index: {id: [Phase]}} for this rank's videos.  gather=False: no
    collective, `tables` holds this rank's videos only (a 1-rank reference run inside a larger job)."""
    import torch
    import torch.distributed as dist
    from .pipeline import VideoPipeline
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    plan = lpt_assign([int(v['frames'].shape[0]) // frame_stride for v in videos], world)
    pipe = None
    local, phases = {}, {}
    B = detector.max_batch
    handles = []                     # (video index, pending result): videos overlap, nothing drains between them
    prev = None
    for vi in plan[rank]:
        v = videos[vi]
        if pipe is None:
            pipe = VideoPipeline(detector, v['fps'], detection_threshold, **pipe_kw)
        else:
            handles.append((prev, pipe.next_video(v['fps'])))     # fresh tracker per source (track.py:157)
        prev = vi
        frames = v['frames']
        n = int(frames.shape[0])
        # 1-based frame_count of the kept frames (track.py:161,166)
        keep = torch.arange(frame_stride, n + 1, frame_stride, dtype=torch.int32)
        numbers = keep.to('cuda')
        for s in range(0, len(keep), B):
            idx = keep[s:s + B].long() - 1
            chunk = frames[idx[0]:idx[-1] + 1] if frame_stride == 1 else frames[idx.to(frames.device)]
            pipe.process(chunk.contiguous(), numbers[s:s + B], swap_rb=True)
    if pipe is not None:
        last = pipe.finish()
        for vi, h in handles:
            res = h.result()
            local[vi], phases[vi] = res['rows'], res['phases']
        local[prev], phases[prev] = last['rows'], last['phases']
    tables = gather_row_tables(local, len(videos), group=group) if gather else local
    return tables, phases


def chunk_bounds(n_items, world_size):
    """[(first, last_exclusive)] per rank: contiguous, balanced, earlier ranks take the remainder."""
    base, rem = divmod(int(n_items), world_size)
    out, at = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((at, at + n))
        at += n
    return out


def gather_detection_tables(dets, counts, numbers, group=None):
    """All ranks' detection tables concatenated in rank order (= frame order for contiguous
    chunks): dets f64 [n,D,6], counts i32 [n], frame numbers i32 [n], torch tensors on the
    collective's device.  Two collectives: the chunk lengths, then one padded payload per table."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dets, counts, numbers
    world = dist.get_world_size(group)
    dev = dets.device
    n = torch.tensor([dets.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x.item()) for x in ns]
    pad = max(max(ns), 1)
    D = dets.shape[1]
    # one float64 payload: [pad, D*6 + 2] = detections, count, frame number (exact in float64)
    payload = torch.zeros((pad, D * 6 + 2), dtype=torch.float64, device=dev)
    m = dets.shape[0]
    payload[:m, :D * 6] = dets.reshape(m, D * 6)
    payload[:m, D * 6] = counts.to(torch.float64)
    payload[:m, D * 6 + 1] = numbers.to(torch.float64)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    full = torch.cat([g[:k] for g, k in zip(gathered, ns)])
    return (full[:, :D * 6].reshape(-1, D, 6).contiguous(), full[:, D * 6].to(torch.int32),
            full[:, D * 6 + 1].to(torch.int32))


def track_video_chunks(video, detector, detection_threshold=0.5, frame_stride=1, rank=None, world=None,
                       group=None, result_on_all_ranks=False, **pipe_kw):
    """ONE video, its kept frames split into contiguous chunks across the ranks.

    video: ``{'fps': float, 'frames': uint8 [N,H,W,3] tensor}`` or ``{'fps', 'n_frames', 'load':
    callable(first, last_exclusive) -> uint8 [m,H,W,3]}`` (0-based source frame range), so that a
    rank only ever decodes / holds its own chunk.  Returns the `VideoPipeline.finish()` dict on
    rank 0 (on every rank with result_on_all_ranks), else None."""
    import torch
    import torch.distributed as dist
    from .pipeline import VideoPipeline
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    n_src = int(video['n_frames']) if 'load' in video else int(video['frames'].shape[0])
    keep = torch.arange(frame_stride, n_src + 1, frame_stride, dtype=torch.int32)   # 1-based (track.py:161,166)
    lo, hi = chunk_bounds(len(keep), world)[rank]
    pipe = VideoPipeline(detector, video['fps'], detection_threshold, **pipe_kw)
    # Rank 0 owns the first chunk: it tracks its own frames while it detects them (the tracker runs on
    # the side stream as in a one-pass run) and contributes an empty table to the gather; the other
    # ranks' tables then continue the same recurrence.  The recurrence itself stays sequential --
    # ~19 us per frame on one warp, the bound of any single-video run.
    own = rank == 0 and not result_on_all_ranks and world > 1
    dets, counts, numbers = detect_chunk(pipe, video, keep[lo:hi], frame_stride, track=own)
    dets, counts, numbers = gather_detection_tables(dets, counts, numbers, group=group)
    if rank != 0 and not result_on_all_ranks:
        return None
    pipe.track_table(dets, counts, numbers)
    return pipe.finish()


def detect_chunk(pipe, video, keep, frame_stride=1, track=False):
    """K1-K6 + pack over the kept frames `keep` (1-based frame numbers, int32 tensor) of one
    video -> this chunk's detection table on the device.  track=True: the chunk also goes through
    the tracker / velocity kernels at once and the returned table is empty."""
    B = pipe.det.max_batch
    numbers = keep.to('cuda')
    for s in range(0, len(keep), B):
        idx = keep[s:s + B].long() - 1
        if 'load' in video:
            chunk = video['load'](int(idx[0]), int(idx[-1]) + 1)
            chunk = chunk[::frame_stride] if frame_stride > 1 else chunk
        else:
            frames = video['frames']
            chunk = frames[idx[0]:idx[-1] + 1] if frame_stride == 1 else frames[idx.to(frames.device)]
        pipe.process(chunk.contiguous(), numbers[s:s + B], swap_rb=True, track=track)
    return pipe.detection_table()
