"""vbt_b200.ingest.DecodeRing (CPU): the background decode thread delivers exactly the frames, in the
order and with the 1-based frame numbers of the reference loop (track.py:159-171), for any stride and
batch size -- checked against a plain cv2 read loop written like the reference's."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')


def _write_clip(path, n, h=48, w=64):
    rng = np.random.default_rng(3)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'MJPG'), 30.0, (w, h))
    assert wr.isOpened()
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for i in range(n):
        wr.write(np.roll(base, 3 * i, axis=1))
    wr.release()


def _reference_loop(path, stride):
    cap = cv2.VideoCapture(path)
    out, frame_count = [], 0
    while cap.isOpened():
        ret, frame = cap.read()
        frame_count += 1
        if not ret:
            break
        if frame_count % stride:
            continue
        out.append((frame_count, frame))
    cap.release()
    return out


@pytest.mark.parametrize('stride,batch', [(1, 5), (4, 3), (16, 64), (1, 64)])
def test_decode_ring_equals_the_reference_read_loop(tmp_path, stride, batch):
    from vbt_b200.ingest import DecodeRing
    path = str(tmp_path / 'clip.avi')
    _write_clip(path, 37)
    want = _reference_loop(path, stride)
    ring = DecodeRing(path, batch=batch, stride=stride, n_slots=2, pin=False)
    assert (ring.H, ring.W) == (48, 64) and ring.fps == 30.0
    got = []
    for frames, numbers, slot in ring:
        assert 0 < len(numbers) == frames.shape[0] <= batch
        got.extend((n, f.numpy().copy()) for n, f in zip(numbers, frames))
        ring.release(slot)
    assert [n for n, _ in got] == [n for n, _ in want]
    assert all(np.array_equal(a, b) for (_, a), (_, b) in zip(got, want))
    assert ring.frames_read == 37


def test_decode_ring_missing_file():
    from vbt_b200.ingest import DecodeRing
    with pytest.raises(FileNotFoundError):
        DecodeRing('/nonexistent/clip.mp4', pin=False)
