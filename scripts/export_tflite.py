#!/usr/bin/env python
"""Export a synthetic post-training-quantised EfficientDet-Lite as a standard `.tflite` file.

usage: python scripts/export_tflite.py lite0 out.tflite [--seed 1234] [--vectors out.npz]

The file holds the very weights `synthetic:lite0` runs on the GPU.  With --vectors, the CPU oracle's
raw class / box outputs for four seeded frames are written next to it, so that someone who has
tflite_runtime can run the file through the real interpreter and compare."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('variant', choices=['lite0', 'lite1', 'lite2'])
    ap.add_argument('out')
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--vectors', default=None)
    a = ap.parse_args()
    from vbt_b200 import effdet, tflite_writer
    g = effdet.build_synthetic(a.variant, seed=a.seed)
    tflite_writer.save(g, a.out)
    print(f'{a.out}: {os.path.getsize(a.out)} bytes, input {g.S}x{g.S}, {g.n_anchors} anchors')
    if a.vectors:
        from oracle import effdet as OE
        from vbt_b200.synth import synthetic_model_inputs
        x = synthetic_model_inputs(4, g.S, seed=a.seed + 7)
        cls, box, _ = OE.run(g, x)
        np.savez_compressed(a.vectors, images=x, raw_scores_q=cls, raw_boxes_q=box, box_scale=g.box_scale,
                            box_zp=g.box_zp, anchors=g.anchors())
        print(f'{a.vectors}: images uint8 {x.shape}, int8 scores (scale 1/256, zp -128) {cls.shape}, int8 boxes {box.shape}')


if __name__ == '__main__':
    main()
