"""Detector on the device + an `Interpreter`-shaped facade.

``Detector`` is the batched production object (frames in HBM -> K1 preprocess -> network
-> K6 post-process, all through libvbt_b200.so).  ``Interpreter`` mirrors the
tflite_runtime surface the reference touches (SURVEY.md 8b "Detector facade"):

    interpreter = Interpreter(model_path=..., num_threads=4)      # track.py:93
    interpreter.allocate_tensors()                                # track.py:94
    interpreter.get_input_details()[0]['shape']                   # odt.py:86  -> (1,S,S,3)
    fn = interpreter.get_signature_runner()                       # odt.py:58
    out = fn(images=uint8[1,S,S,3])                               # odt.py:61
    out['output_0'] count f32[1], ['output_1'] scores f32[1,25],
    out['output_2'] classes f32[1,25], ['output_3'] boxes f32[1,25,4] (ymin,xmin,ymax,xmax)

Model files: ``*.vbtm`` (the layer-program blob of vbt_b200/effdet.py) or the spec
``synthetic:lite0|lite1|lite2[:seed]``.  The reference's own ``.tflite`` blobs are absent
from its checkout; `vbt_b200.tflite_reader` converts them when supplied.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib, effdet

MAX_DET = _lib.MAX_DETECTIONS


def load_model_bytes(model_path):
    """(blob bytes, display name) for a model path / spec."""
    if isinstance(model_path, effdet.Graph):
        return effdet.pack_blob(model_path), model_path.variant
    if model_path.startswith('synthetic:'):
        parts = model_path.split(':')
        seed = int(parts[2]) if len(parts) > 2 else 1234
        return effdet.pack_blob(effdet.build_synthetic(parts[1], seed=seed)), parts[1]
    if not os.path.isfile(model_path):
        raise ValueError(f'Could not open {model_path!r}.')
    if model_path.endswith('.tflite'):
        from .tflite_reader import tflite_to_graph
        return effdet.pack_blob(tflite_to_graph(model_path)), os.path.basename(model_path)
    with open(model_path, 'rb') as f:
        return f.read(), os.path.basename(model_path)


class Detector:
    """One model resident on the current device; buffers sized for `max_batch` frames."""

    def __init__(self, model, max_batch=64, iou_threshold=0.5, max_det=MAX_DET):
        self.torch = t = _lib.require_cuda()
        self.source = model
        blob, self.name = load_model_bytes(model)
        h = C.c_void_p()
        _lib.check(_lib.lib().vbt_model_create(blob, len(blob), C.byref(h)))
        self.handle = h
        info = (C.c_longlong * 8)()
        _lib.check(_lib.lib().vbt_model_info(h, info))
        self.S, self.N, self.ws_per_frame, self.n_ops = int(info[0]), int(info[1]), int(info[2]), int(info[3])
        self.Np = int(info[6])
        self.iou_threshold, self.max_det, self.max_batch = iou_threshold, max_det, max_batch
        B = max_batch
        self.workspace = t.empty(max(B * self.ws_per_frame, 256), dtype=t.uint8, device='cuda')
        self.resized = t.empty((B, self.S, self.S, 3), dtype=t.uint8, device='cuda')
        self.raw_cls = t.zeros((B, self.Np), dtype=t.int8, device='cuda')
        self.raw_box = t.zeros((B, self.Np, 4), dtype=t.int8, device='cuda')
        self.boxes = t.zeros((B, max_det, 4), dtype=t.float32, device='cuda')
        self.classes = t.zeros((B, max_det), dtype=t.float32, device='cuda')
        self.scores = t.zeros((B, max_det), dtype=t.float32, device='cuda')
        self.count = t.zeros(B, dtype=t.float32, device='cuda')
        self.index = t.zeros((B, max_det), dtype=t.int32, device='cuda')

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            try:
                _lib.lib().vbt_model_destroy(h)
            except Exception:      # interpreter shutdown: module globals are gone
                pass

    # -- per-op device timing (bench.py) -----------------------------------------------------
    def profile(self, enable=True):
        _lib.check(_lib.lib().vbt_model_profile(self.handle, int(enable)))

    def plan(self):
        """int32 [n_ops]: ops covered by the launch that starts at each op (0 = inside a group)."""
        g = np.ones(max(self.n_ops, 1), dtype=np.int32)
        _lib.check(_lib.lib().vbt_model_plan(self.handle, _lib.ptr(g)))
        return g[:self.n_ops]

    def plan_kinds(self):
        """int32 [n_ops]: 1 where the launch starting at the op is a whole MBConv block."""
        g = np.zeros(max(self.n_ops, 1), dtype=np.int32)
        _lib.check(_lib.lib().vbt_model_plan_kinds(self.handle, _lib.ptr(g)))
        return g[:self.n_ops]

    def op_times(self):
        """(ms per op accumulated [n_ops], number of vbt_detect calls covered)."""
        ms = np.zeros(max(self.n_ops, 1), dtype=np.float64)
        calls = C.c_longlong(0)
        _lib.check(_lib.lib().vbt_model_op_times(self.handle, _lib.ptr(ms), C.byref(calls)))
        return ms[:self.n_ops], int(calls.value)

    # -- stages ----------------------------------------------------------------------------
    def preprocess(self, frames, swap_rb=True, stream=None):
        """frames: uint8 [B,H,W,3], a CUDA tensor or a PINNED host tensor (read in place over
        PCIe: only the 2*S source rows the bilinear kernel touches cross the bus)
        -> self.resized[:B] (RGB, SxS)."""
        if not frames.is_cuda and not frames.is_pinned():
            raise ValueError('host frames must be in pinned memory (tensor.pin_memory())')
        B, H, W, _ = frames.shape
        assert B <= self.max_batch and frames.is_contiguous()
        _lib.check(_lib.lib().vbt_preprocess_u8(frames.data_ptr(), B, H, W, int(swap_rb),
                                                self.resized.data_ptr(), self.S,
                                                _lib.stream_ptr(stream)))
        return self.resized[:B]

    def preprocess_rows(self, ingest, table, n, swap_rb=True, stream=None):
        """K1 over a compacted row table produced by `ingest.RowSparseIngest.upload`."""
        assert n <= self.max_batch and ingest.S == self.S
        _lib.check(_lib.lib().vbt_preprocess_rows_u8(table.data_ptr(), n, ingest.H, ingest.W,
                                                     ingest.rows_per_frame, ingest.row_map.data_ptr(),
                                                     int(swap_rb), self.resized.data_ptr(), self.S,
                                                     _lib.stream_ptr(stream)))
        return self.resized[:n]

    def network(self, images, stream=None):
        """images: uint8 CUDA [B,S,S,3] RGB -> raw int8 class scores / box encodings."""
        B = images.shape[0]
        assert B <= self.max_batch and images.is_contiguous()
        _lib.check(_lib.lib().vbt_detect(self.handle, images.data_ptr(), B,
                                         self.workspace.data_ptr(), self.workspace.numel(),
                                         self.raw_cls.data_ptr(), self.raw_box.data_ptr(),
                                         _lib.stream_ptr(stream)))
        return self.raw_cls[:B], self.raw_box[:B]

    def postprocess(self, B, min_score_q=-128, raw_cls=None, raw_box=None, stream=None):
        cls = self.raw_cls if raw_cls is None else raw_cls
        box = self.raw_box if raw_box is None else raw_box
        _lib.check(_lib.lib().vbt_postprocess_q8(
            self.handle, cls.data_ptr(), box.data_ptr(), B, self.iou_threshold, self.max_det,
            int(min_score_q), self.boxes.data_ptr(), self.classes.data_ptr(),
            self.scores.data_ptr(), self.count.data_ptr(), self.index.data_ptr(),
            _lib.stream_ptr(stream)))
        return self.boxes[:B], self.classes[:B], self.scores[:B], self.count[:B], self.index[:B]

    def detect(self, frames, swap_rb=True, threshold=None, stream=None):
        """Full a2-a5 for a batch of device frames.  threshold: when given, candidates
        below it are skipped inside NMS (exact for `score >= threshold` consumers,
        SURVEY.md appendix D); None reproduces the op (all 25 outputs)."""
        B = frames.shape[0]
        if frames.shape[1] == self.S and frames.shape[2] == self.S and not swap_rb:
            images = frames
        else:
            images = self.preprocess(frames, swap_rb, stream)
        self.network(images, stream)
        return self.postprocess(B, score_to_q(threshold), stream=stream)


def score_to_q(threshold):
    """Smallest int8 LOGISTIC output whose dequantised score is >= threshold."""
    if threshold is None:
        return -128
    level = int(np.ceil(np.float32(threshold) * 256.0))
    return int(min(127, max(-128, level - 128)))


class Interpreter:
    """tflite_runtime.interpreter.Interpreter look-alike over `Detector` (batch 1)."""

    def __init__(self, model_path=None, num_threads=None, max_batch=1, **_ignored):
        """max_batch (not a TFLite argument): frames per call the detector's buffers are sized for; the
        track CLI passes its --batch here so that the model is loaded once, not once more inside track()."""
        self.model_path = model_path
        self.num_threads = num_threads           # accepted for drop-in; the GPU path ignores it
        self._det = Detector(model_path, max_batch=max(1, int(max_batch)))

    def allocate_tensors(self):
        return None                              # buffers are allocated with the model

    def get_input_details(self):
        S = self._det.S
        return [{'name': 'serving_default_images:0', 'index': 0,
                 'shape': np.array([1, S, S, 3], dtype=np.int32), 'dtype': np.uint8,
                 'quantization': (1.0 / 128.0, 127)}]

    def get_output_details(self):
        d = self._det.max_det
        return [{'name': n, 'shape': np.array(s, dtype=np.int32), 'dtype': np.float32}
                for n, s in (('output_0', [1]), ('output_1', [1, d]), ('output_2', [1, d]),
                             ('output_3', [1, d, 4]))]

    def get_signature_runner(self, signature_key=None):
        det = self._det
        t = det.torch

        def run(images=None, **kw):
            img = np.ascontiguousarray(np.asarray(images), dtype=np.uint8)
            if img.shape != (1, det.S, det.S, 3):
                raise ValueError(f'Cannot set tensor: Dimension mismatch. Got {img.shape} but '
                                 f'expected (1, {det.S}, {det.S}, 3) for input 0.')
            dev = t.as_tensor(img, device='cuda')
            det.network(dev)
            boxes, classes, scores, count, _ = det.postprocess(1)
            return {'output_0': count.cpu().numpy().reshape(1),
                    'output_1': scores.cpu().numpy(), 'output_2': classes.cpu().numpy(),
                    'output_3': boxes.cpu().numpy()}

        return run

    @property
    def detector(self):
        return self._det
