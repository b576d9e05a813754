"""The reference's detection helpers (odt.py) over the CUDA kernels.

Same names, argument meaning and return shapes as odt.py in the reference:
`preprocess_image` (odt.py:10-19) -> K1, `detect_objects` (odt.py:53-77) and `run_odt`
(odt.py:80-99) -> network + K6 through the interpreter facade,
`results_to_sorttracker_inputs` (odt.py:102-118), and the three geometry helpers
(odt.py:22-50), which stay host arithmetic on four Python scalars.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def preprocess_image(frame, input_size):
    """uint8 [H,W,3] RGB frame -> (uint8 [1,S,S,3] resized, the original frame).
    Bilinear, half-pixel centres, stretch (no aspect preservation), truncating cast --
    computed by vbt_preprocess_u8 on the device."""
    torch = _lib.require_cuda()
    original = np.ascontiguousarray(np.asarray(frame, dtype=np.uint8))
    h, w = int(input_size[0]), int(input_size[1])
    if h != w:
        raise ValueError('the EfficientDet-Lite inputs are square')
    dev = torch.as_tensor(original[None], device='cuda')
    out = torch.empty((1, h, h, 3), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.lib().vbt_preprocess_u8(dev.data_ptr(), 1, original.shape[0],
                                            original.shape[1], 0, out.data_ptr(), h,
                                            _lib.stream_ptr()))
    return out.cpu().numpy(), original


def calc_plate_width(bounding_box):
    _, xmin, _, xmax = bounding_box
    return abs(xmax - xmin)


def calc_plate_height(bounding_box):
    ymin, _, ymax, _ = bounding_box
    return abs(ymax - ymin)


def calc_bounding_box_center(bounding_box):
    ymin, xmin, ymax, xmax = bounding_box
    return ((xmin + xmax) / 2, (ymin + ymax) / 2)


def detect_objects(interpreter, image, threshold):
    """List of {'bounding_box': f32[4] (ymin,xmin,ymax,xmax), 'score': f32}, score
    descending, only scores >= threshold."""
    output = interpreter.get_signature_runner()(images=image)
    count = int(np.squeeze(output['output_0']))
    scores = np.squeeze(output['output_1'])
    boxes = np.squeeze(output['output_3'])
    return [{'bounding_box': boxes[i], 'score': scores[i]}
            for i in range(count) if scores[i] >= threshold]


def run_odt(frame, interpreter, threshold=0.5):
    _, input_height, input_width, _ = interpreter.get_input_details()[0]['shape']
    image, _ = preprocess_image(frame, (input_height, input_width))
    return detect_objects(interpreter, image, threshold=threshold)


def results_to_sorttracker_inputs(orig_results):
    """list[dict] -> float64 [N,6] = xmin,ymin,xmax,ymax,score,0 (np.empty((0,6)) if none)."""
    out = np.empty((len(orig_results), 6), dtype=np.float64)
    for i, res in enumerate(orig_results):
        ymin, xmin, ymax, xmax = res['bounding_box']
        out[i] = (xmin, ymin, xmax, ymax, res['score'], 0)
    return out
