"""K1 / network / K6 parity on a B200 (through the C ABI) vs the CPU oracle.
Bars: K1 bit-exact uint8; network raw int8 outputs bit-exact; K6 selected anchor indices,
scores and boxes bit-exact (the exp() the op needs is a 256-entry LUT shared by both)."""
import numpy as np
import pytest

from oracle import effdet as OE, postprocess as OP, resize as OR, ocsort as oo

pytestmark = pytest.mark.gpu

_models = {}


def model(variant):
    if variant not in _models:
        from vbt_b200 import effdet
        from vbt_b200.interpreter import Detector
        g = effdet.build_synthetic(variant)
        _models[variant] = (g, Detector(g, max_batch=8))
    return _models[variant]


# ---- K1 -------------------------------------------------------------------------------

@pytest.mark.parametrize('shape,S,swap', [((1080, 1920), 320, True), ((1080, 1920), 384, False),
                                          ((1080, 1920), 448, True), ((416, 416), 320, False),
                                          ((1920, 1080), 320, True), ((320, 320), 320, False),
                                          ((37, 53), 320, True)])
def test_preprocess_bit_exact(shape, S, swap):
    import torch
    from vbt_b200 import _lib
    rng = np.random.default_rng(hash((shape, S)) % 2**32)
    frames = rng.integers(0, 256, size=(3,) + shape + (3,), dtype=np.uint8)
    frames[1, : shape[0] // 2] = 255         # saturated and black regions hit the cast edges
    frames[2, :, : shape[1] // 2] = 0
    dev = torch.as_tensor(frames, device='cuda')
    out = torch.empty((3, S, S, 3), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.lib().vbt_preprocess_u8(dev.data_ptr(), 3, shape[0], shape[1], int(swap),
                                            out.data_ptr(), S, _lib.stream_ptr()))
    want = OR.preprocess_batch(frames, S, swap_rb=swap)
    assert np.array_equal(out.cpu().numpy(), want)


def test_preprocess_reads_pinned_host_frames_in_place():
    """Zero-copy ingest: K1 pulls the source rows it needs straight from pinned host memory."""
    import torch
    from vbt_b200 import _lib
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, size=(2, 1080, 1920, 3), dtype=np.uint8)
    host = torch.from_numpy(frames).pin_memory()
    out = torch.empty((2, 320, 320, 3), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.lib().vbt_preprocess_u8(host.data_ptr(), 2, 1080, 1920, 1, out.data_ptr(), 320,
                                            _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), OR.preprocess_batch(frames, 320, swap_rb=True))


@pytest.mark.parametrize('S', [320, 384, 448])
def test_row_sparse_ingest_equals_full_frames(S):
    """Host frames sent row-sparse (only the rows the resize reads, strided DMA) give the same
    K1 output as the full frames; the plan moves 2*S of the 1080 rows."""
    import torch
    from vbt_b200 import _lib
    from vbt_b200.ingest import RowSparseIngest, touched_rows
    rng = np.random.default_rng(S)
    frames = rng.integers(0, 256, size=(3, 1080, 1920, 3), dtype=np.uint8)
    host = torch.from_numpy(frames).pin_memory()
    ing = RowSparseIngest(4, 1080, 1920, S)
    assert ing.rows_per_frame == len(touched_rows(1080, S)) == 2 * S
    out = torch.empty((3, S, S, 3), dtype=torch.uint8, device='cuda')
    for _ in range(4):                      # cycles through the table ring
        table, n, slot, ready = ing.upload(host)
        torch.cuda.current_stream().wait_event(ready)
        _lib.check(_lib.lib().vbt_preprocess_rows_u8(table.data_ptr(), n, 1080, 1920, ing.rows_per_frame,
                                                     ing.row_map.data_ptr(), 1, out.data_ptr(), S,
                                                     _lib.stream_ptr()))
        ing.free[slot] = torch.cuda.Event()
        ing.free[slot].record()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), OR.preprocess_batch(frames, S, swap_rb=True))


def test_preprocess_image_helper_matches_oracle():
    from vbt_b200.odt import preprocess_image
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, size=(416, 416, 3), dtype=np.uint8)
    resized, orig = preprocess_image(frame, (320, 320))
    assert resized.shape == (1, 320, 320, 3) and resized.dtype == np.uint8
    assert np.array_equal(resized[0], OR.resize_bilinear_u8(frame, 320))
    assert np.array_equal(orig, frame)


# ---- K6 -------------------------------------------------------------------------------

def run_post(det, g, cls, box, min_q=-128):
    import torch
    B = cls.shape[0]
    pc = np.zeros((B, det.Np), np.int8); pc[:, :g.n_anchors] = cls
    pb = np.zeros((B, det.Np, 4), np.int8); pb[:, :g.n_anchors] = box
    r = det.postprocess(B, min_q, torch.as_tensor(pc, device='cuda'), torch.as_tensor(pb, device='cuda'))
    return [x.cpu().numpy() for x in r]


def check_post(det, g, cls, box, min_q=-128):
    boxes, classes, scores, count, index = run_post(det, g, cls, box, min_q)
    a = g.anchors()
    for b in range(cls.shape[0]):
        ob, oc, osc, cnt, oi = OP.detection_postprocess(cls[b], box[b], a, g.box_scale, g.box_zp,
                                                        min_score_q=min_q)
        assert count[b] == cnt
        assert np.array_equal(index[b], oi)                 # detection indices: exact
        assert np.array_equal(scores[b], osc)
        assert np.array_equal(boxes[b], ob)
        assert np.array_equal(classes[b], oc)


@pytest.mark.parametrize('variant', ['lite0', 'lite2'])
def test_postprocess_random_and_ties(variant):
    from vbt_b200 import effdet
    from vbt_b200.interpreter import Detector
    g = effdet.anchors_only(variant, box_scale=0.02, box_zp=3)
    det = Detector(g, max_batch=8)
    N = g.n_anchors
    rng = np.random.default_rng(0)
    cls = np.full((6, N), -120, np.int8)
    box = rng.integers(-60, 60, size=(6, N, 4)).astype(np.int8)
    cls[0] = rng.integers(-128, 128, N)                         # dense random scores
    cls[1, rng.choice(N, 40, replace=False)] = rng.integers(0, 127, 40)   # sparse peaks
    cls[2] = 5                                                  # ONE level holds every anchor
    cls[3] = rng.choice(np.array([-128, 90], np.int8), N)       # two giant tie groups
    cls[4] = -128                                               # nothing above -128 but all valid
    cls[5, :3000] = 100                                         # one level just above the stage cap
    check_post(det, g, cls, box)
    check_post(det, g, cls, box, min_q=0)                       # threshold shortcut (score >= 0.5)
    check_post(det, g, cls[4:5], box[4:5], min_q=-127)          # no candidate at all -> count 0


def test_postprocess_overlapping_clusters():
    """Small offsets: neighbouring anchors produce heavily overlapping boxes, so
    suppression (not just ordering) decides the output."""
    from vbt_b200 import effdet
    from vbt_b200.interpreter import Detector
    g = effdet.anchors_only('lite0', box_scale=0.004, box_zp=0)
    det = Detector(g, max_batch=4)
    rng = np.random.default_rng(4)
    N = g.n_anchors
    cls = rng.integers(-128, -60, size=(4, N)).astype(np.int8)
    for b in range(4):
        hot = rng.choice(N - 400, 3)
        for h in hot:
            cls[b, h:h + 300] = rng.integers(20, 127, 300)
    box = rng.integers(-20, 20, size=(4, N, 4)).astype(np.int8)
    check_post(det, g, cls, box)
    check_post(det, g, cls, box, min_q=10)


def test_pack_detections_matches_odt_helpers():
    import torch
    from vbt_b200 import _lib
    rng = np.random.default_rng(2)
    F, D = 9, 25
    boxes = rng.uniform(0, 1, (F, D, 4)).astype(np.float32)
    scores = np.sort(rng.integers(0, 256, (F, D)) / 256.0, axis=1)[:, ::-1].astype(np.float32)
    count = rng.integers(0, D + 1, F).astype(np.float32)
    count[0] = 0
    dets = torch.zeros((F, D, 6), dtype=torch.float64, device='cuda')
    n = torch.zeros(F, dtype=torch.int32, device='cuda')
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    d_boxes, d_scores, d_count = dev(boxes), dev(scores), dev(count)     # keep alive
    _lib.check(_lib.lib().vbt_pack_detections(d_boxes.data_ptr(), d_scores.data_ptr(),
                                              d_count.data_ptr(), F, D, 0.5, dets.data_ptr(),
                                              n.data_ptr(), _lib.stream_ptr()))
    dets, n = dets.cpu().numpy(), n.cpu().numpy()
    for f in range(F):
        want = OP.tracker_inputs(OP.detect_results(boxes[f], scores[f], count[f], 0.5))
        assert n[f] == len(want)
        assert np.array_equal(dets[f, :n[f]], want.reshape(-1, 6))


# ---- network --------------------------------------------------------------------------

@pytest.mark.parametrize('variant,frames', [('lite0', 3), ('lite1', 1), ('lite2', 1)])
def test_network_raw_outputs_bit_exact(variant, frames):
    import torch
    from vbt_b200.synth import synthetic_model_inputs
    g, det = model(variant)
    x = synthetic_model_inputs(frames, g.S, seed=21)
    want_cls, want_box, _ = OE.run(g, x)
    cls, box = det.network(torch.as_tensor(x, device='cuda'))
    cls, box = cls.cpu().numpy()[:, :g.n_anchors], box.cpu().numpy()[:, :g.n_anchors]
    assert np.array_equal(box, want_box)
    assert np.array_equal(cls, want_cls)


def test_interpreter_facade_and_run_odt():
    from vbt_b200.interpreter import Interpreter
    from vbt_b200.odt import run_odt, results_to_sorttracker_inputs
    from vbt_b200.synth import synthetic_model_inputs
    g, _ = model('lite0')
    interp = Interpreter(model_path=g, num_threads=4)
    interp.allocate_tensors()
    assert tuple(interp.get_input_details()[0]['shape']) == (1, 320, 320, 3)
    rng = np.random.default_rng(8)
    frame = rng.integers(0, 256, size=(416, 416, 3), dtype=np.uint8)
    img = OR.resize_bilinear_u8(frame, 320)[None]
    out = interp.get_signature_runner()(images=img)
    cls, box, _ = OE.run(g, img)
    ob, oc, osc, cnt, _ = OP.detection_postprocess(cls[0], box[0], g.anchors(), g.box_scale, g.box_zp)
    assert out['output_0'].shape == (1,) and out['output_0'][0] == cnt
    assert np.array_equal(out['output_1'][0], osc) and np.array_equal(out['output_3'][0], ob)
    assert out['output_2'].shape == (1, 25)
    res = run_odt(frame, interp, threshold=0.3)
    want = OP.detect_results(ob, osc, cnt, 0.3)
    assert len(res) == len(want)
    for r, w in zip(res, want):
        assert r['score'] == w['score'] and np.array_equal(r['bounding_box'], w['bounding_box'])
    assert np.array_equal(results_to_sorttracker_inputs(res), OP.tracker_inputs(want).reshape(-1, 6))
    with pytest.raises(ValueError):
        interp.get_signature_runner()(images=img[:, :100])


# ---- whole per-video pipeline -----------------------------------------------------------

def test_video_pipeline_equals_oracle_chain():
    """frames -> K1 -> network -> K6 -> threshold -> K7 -> rows, vs the same chain of
    CPU oracles; then K8 phases per id vs the velocity oracle on those rows."""
    import torch
    from oracle import velocity as ov
    from vbt_b200.pipeline import VideoPipeline
    from vbt_b200.interpreter import Detector
    g, _ = model('lite0')
    det = Detector(g, max_batch=4)
    rng = np.random.default_rng(3)
    n, H, W = 10, 270, 480
    base = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    frames = np.stack([np.roll(base, 3 * i, axis=0) for i in range(n)])
    thr = 0.3
    pipe = VideoPipeline(det, fps=30.0, detection_threshold=thr, row_cap=4096)
    for s in range(0, n, 4):
        e = min(s + 4, n)
        pipe.process(torch.as_tensor(frames[s:e], device='cuda'),
                     torch.arange(s + 1, e + 1, dtype=torch.int32, device='cuda'), swap_rb=True)
    res = pipe.finish()
    imgs = OR.preprocess_batch(frames, g.S, swap_rb=True)
    cls, box, _ = OE.run(g, imgs)
    a = g.anchors()
    dets = []
    for b in range(n):
        ob, oc, osc, cnt, _ = OP.detection_postprocess(cls[b], box[b], a, g.box_scale, g.box_zp)
        dets.append(OP.tracker_inputs(OP.detect_results(ob, osc, cnt, thr)).reshape(-1, 6))
    want = oo.track_rows(dets, 30.0, max_age=30, iou_threshold=0.1)
    assert sum(len(d) for d in dets) > 0
    assert res['rows'].shape == want.shape
    assert np.array_equal(res['rows'], want)
    for tid, phases in res['phases'].items():
        sel = want[want[:, 0] == tid][:, 1:]
        w = ov.analyze_series(sel, 0.45)
        got = np.array([[p.time_start, p.time_end, p.y_start, p.y_end, p.rom, p.type]
                        for p in phases]).reshape(-1, 6)
        assert np.array_equal(got, w)


def test_interpreter_loads_a_tflite_file(tmp_path):
    """track.py:93 `Interpreter(model_path=<x.tflite>)`: the exported synthetic Lite0 loaded back
    through vbt_b200.tflite_reader gives the detections of the in-memory graph."""
    from vbt_b200.interpreter import Interpreter
    from vbt_b200.odt import run_odt
    from vbt_b200.synth import synthetic_model_inputs
    from vbt_b200 import tflite_writer
    g, _ = model('lite0')
    path = str(tmp_path / 'efficientdet_lite0_synthetic.tflite')
    tflite_writer.save(g, path)
    a, b = Interpreter(model_path=g, num_threads=4), Interpreter(model_path=path, num_threads=4)
    a.allocate_tensors(); b.allocate_tensors()
    assert list(b.get_input_details()[0]['shape']) == [1, 320, 320, 3]
    x = synthetic_model_inputs(2, g.S, seed=33)
    for img in x:
        ra = a.get_signature_runner()(images=img[None])
        rb = b.get_signature_runner()(images=img[None])
        for k in ('output_0', 'output_1', 'output_2', 'output_3'):
            assert np.array_equal(ra[k], rb[k]), k
        la, lb = run_odt(img, a, 0.3), run_odt(img, b, 0.3)
        assert len(la) == len(lb)
        for p, q in zip(la, lb):
            assert np.array_equal(p['bounding_box'], q['bounding_box']) and p['score'] == q['score']
    with pytest.raises(ValueError):
        Interpreter(model_path=str(tmp_path / 'missing.tflite'))


@pytest.mark.gpu
def test_model_blob_with_out_of_range_offsets_is_refused():
    """vbt_model_create checks every weight / bias / multiplier region and every tensor extent of the blob
    (a truncated .vbtm or a bad conversion must fail at load time, not inside a kernel)."""
    import struct
    import ctypes
    from vbt_b200 import _lib, effdet
    g = effdet.build_synthetic('lite0')
    blob = bytearray(effdet.pack_blob(g))
    L = _lib.lib()

    def create(b):
        h = ctypes.c_void_p()
        rc = L.vbt_model_create(bytes(b), len(b), ctypes.byref(h))
        if rc == 0:
            L.vbt_model_destroy(h)
        return rc

    assert create(blob) == 0
    # first op record starts right after the 128-byte header; its w_off is the first int64 after 30 int32
    hdr = 128
    rec0 = bytes(blob[hdr:hdr + 224])
    w_off_pos = hdr + struct.calcsize('<i3ii i ii ii ii iiii 3i i ii ii')
    bad = bytearray(blob)
    struct.pack_into('<q', bad, w_off_pos, 1 << 40)
    assert create(bad) == _lib.EFORMAT
    # a tensor pushed past the per-frame workspace
    n_ops = struct.unpack_from('<i', blob, 24)[0]
    t0 = hdr + 224 * n_ops
    bad = bytearray(blob)
    struct.pack_into('<q', bad, t0, 1 << 40)
    assert create(bad) == _lib.EFORMAT
    # a truncated file
    assert create(blob[:len(blob) // 2]) == _lib.EFORMAT
    assert rec0 == bytes(blob[hdr:hdr + 224])
