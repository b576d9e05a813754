#!/usr/bin/env python
"""torchrun check of shard.track_video_chunks over NCCL: one synthetic clip, its frames split in
contiguous chunks across the ranks, ONE gather of the detection tables, tracker + velocity on
rank 0 -- the rows must equal the single-GPU one-pass result byte for byte.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port 29533 scripts/chunk_shard_check.py [--frames 512]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=512)
    ap.add_argument('--batch', type=int, default=64)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        dist.init_process_group('nccl')
    from vbt_b200 import effdet, shard
    from vbt_b200.interpreter import Detector
    from vbt_b200.pipeline import VideoPipeline
    g = effdet.build_synthetic('lite0')
    det = Detector(g, max_batch=a.batch)
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, size=(270, 480, 3), dtype=np.uint8)
    n, fps, thr = a.frames, 30.0, 0.3

    def load(lo, hi):          # a rank only materialises the frames of its own batches
        return torch.as_tensor(np.stack([np.roll(base, 3 * i, axis=1) for i in range(lo, hi)]), device='cuda')

    video = {'fps': fps, 'n_frames': n, 'load': load}
    torch.cuda.synchronize()
    t0 = time.time()
    res = shard.track_video_chunks(video, det, detection_threshold=thr, row_cap=1 << 15)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        pipe = VideoPipeline(det, fps, thr, row_cap=1 << 15)
        for s in range(0, n, a.batch):
            e = min(n, s + a.batch)
            pipe.process(load(s, e), torch.arange(s + 1, e + 1, dtype=torch.int32, device='cuda'), swap_rb=True)
        ref = pipe.finish()
        same = res['rows'].tobytes() == ref['rows'].tobytes()
        ph = sorted(res['phases']) == sorted(ref['phases']) and all(
            [(p.time_start, p.time_end, p.rom, p.type) for p in res['phases'][k]] ==
            [(p.time_start, p.time_end, p.rom, p.type) for p in ref['phases'][k]] for k in res['phases'])
        print(f'world={world} frames={n} rows={len(ref["rows"])} rows_equal={same} phases_equal={ph} '
              f'chunked_wall_s={dt:.3f}')
        assert same and ph and len(ref['rows']) > 0
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
