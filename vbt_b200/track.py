"""`track.py`-compatible CLI and per-video driver on the B200 path.

Same command line as the reference (track.py:65-72, including the misspelt
``--detection_treshold``; ``--threads`` is accepted and ignored), same return value of
``track()`` (track.py:129-260: dict of lists with keys id,time,x,y,dx,dy,
norm_plate_height,norm_plate_width) and the same pickled DataFrame (track.py:103-126).

Differences, all explicit:
* ``--frame_stride`` (default 16 = HEAD's ``if frame_count % 16: continue``, track.py:166;
  1 reproduces the bundled dfs_ocsort fixtures) and ``--batch`` (frames per GPU batch);
* frames are decoded on the host by cv2 exactly as in the reference, then moved to the
  GPU in batches; detection, tracking and row assembly run in CUDA;
* the annotated video export (track.py:152-154,241-242) is a second pass over the file once the
  track table exists: the frames the reference would have written (processed frames with at
  least one detection >= threshold) get the same overlay -- white box + "NN%, tracking_id: K"
  label (track.py:28-49), bar path of the last 120 centres + end marker (track.py:52-62) -- drawn
  with cv2 on the RGB-converted frame exactly like the reference (so the mp4's colours are
  swapped exactly like the reference's), and are written with the same mp4v VideoWriter;
* the live preview (cv2.imshow, track.py:237-239) is not built (headless).
"""
from __future__ import annotations

import os

import click
import numpy as np

from . import _lib
from .interpreter import Interpreter
from .pipeline import VideoPipeline, export_dataframe, rows_to_data

MAX_AGE = 30     # track.py:22
COLORS = [(115, 3, 252), (255, 255, 255)]     # track.py:23


def draw_bounding_box(image, tracking_id, bounding_box, score, color):
    """Box outline + label, in place (track.py:28-49): bounding_box = (ymin, xmin, ymax, xmax)
    normalised; pixel corners truncate; the label moves below the top edge near the image top."""
    import cv2
    h, w = image.shape[:2]
    ymin, xmin, ymax, xmax = bounding_box
    x0, x1, y0, y1 = int(xmin * w), int(xmax * w), int(ymin * h), int(ymax * h)
    cv2.rectangle(image, (x0, y0), (x1, y1), color, 2)
    ty = y0 - 15 if y0 - 15 > 15 else y0 + 15
    cv2.putText(image, '{:.0f}%, tracking_id: {}'.format(score * 100, tracking_id), (x0, ty),
                cv2.FONT_HERSHEY_DUPLEX, 1, color, 2)


def draw_bar_path(image, bar_path, color):
    """Last 120 bar positions as an open polyline + a filled end marker (track.py:52-62)."""
    import cv2
    pts = bar_path[-120:] if len(bar_path) > 120 else bar_path
    cv2.polylines(image, [pts], isClosed=False, color=color, thickness=2)
    cv2.circle(image, center=tuple(int(v) for v in pts[-1]), radius=10, color=color, thickness=-1)


def export_annotated_video(src, video_path, result, fps, frame_stride):
    """Writes the video track.py:152-154,241-242 writes: one frame per processed frame that had a
    detection >= threshold, annotated from the track table.  Returns the number of frames written."""
    import cv2
    rows, details = result['rows'], result['details']
    by_frame = {}
    for r, d in zip(rows, details):
        by_frame.setdefault(int(round(r[1] * fps)), []).append((r, d))
    wanted = set(result['frames_with_results'])
    cap = cv2.VideoCapture(src)
    size = (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
    writer = cv2.VideoWriter(video_path, cv2.VideoWriter_fourcc(*'mp4v'), fps, size)
    bar_paths, written, frame_count = {}, 0, 0
    while cap.isOpened():
        ret, frame = cap.read()
        frame_count += 1
        if not ret:
            break
        if frame_count % frame_stride or frame_count not in wanted:
            continue
        img = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)          # track.py:171; written as is (track.py:242)
        for r, d in by_frame.get(frame_count, []):
            tid = int(r[0])
            draw_bounding_box(img, tid, (d[1], d[0], d[3], d[2]), d[4], COLORS[1])
            centre = np.array([r[2] * img.shape[1], r[3] * img.shape[0]], dtype=np.int32)
            bar_paths[tid] = np.concatenate((bar_paths[tid], [centre]), dtype=np.int32) \
                if tid in bar_paths else np.array([centre], np.int32)
            draw_bar_path(img, bar_paths[tid], COLORS[1])
        writer.write(img)
        written += 1
    cap.release()
    writer.release()
    return written


def track(src, interpreter, detection_treshold, display_image_height=720, video_path=None,
          frame_stride=16, batch=64, return_pipeline_result=False):
    """Runs detection + tracking over one video file and returns the captured data."""
    torch = _lib.require_cuda()
    from .ingest import DecodeRing
    # decode thread -> ring of pinned batches (ingest.DecodeRing): frames the stride skips are only
    # demuxed, kept frames are decoded straight into page-locked memory while the GPU works
    ring = DecodeRing(src, batch=batch, stride=frame_stride)
    fps = ring.fps
    det = interpreter.detector if isinstance(interpreter, Interpreter) else interpreter
    if det.max_batch < batch:
        det = type(det)(interpreter.model_path, max_batch=batch)
    pipe = VideoPipeline(det, fps, detection_treshold, keep_details=video_path is not None,
                         tracker_kw=dict(max_age=MAX_AGE, iou_threshold=0.1))   # track.py:157
    pipe.use_row_sparse_ingest(ring.H, ring.W)             # host frames: send only the rows K1 reads
    for frames, numbers, slot in ring:
        nums = torch.as_tensor(np.asarray(numbers, dtype=np.int32), device='cuda')
        if pipe.ingest is not None:
            pipe.process(frames, nums, swap_rb=True)       # cv2 frames are BGR (track.py:171)
            ring.release(slot, pipe.input_consumed)
        else:                                              # odd row sizes: plain copy of the batch
            dev = frames.to('cuda', non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            pipe.process(dev, nums, swap_rb=True)
            ring.release(slot, done)
    result = pipe.finish()
    if video_path is not None:
        export_annotated_video(src, video_path, result, fps, frame_stride)
    data = rows_to_data(result['rows'])
    return (data, result) if return_pipeline_result else data


@click.command()
@click.argument('src', type=str, nargs=-1)
@click.option('--model', default='models/efficientdet_lite0_whole.tflite', type=str, show_default=True,
              help='Model: a .tflite / .vbtm file or synthetic:lite0|lite1|lite2.')
@click.option('--detection_treshold', default=0.5, type=float, show_default=True,
              help='Object detection threshold.')
@click.option('--display_image_height', default=720, type=int, show_default=True,
              help='Displayed image height in pixels (preview is not built; accepted for compatibility).')
@click.option('--df_dir', default=None, show_default=True,
              help="Directory for exporting the dataframes. If not set the dataframe won't be exported.")
@click.option('--video_dir', default=None, show_default=True,
              help='Directory for exporting the annotated video.')
@click.option('--threads', default=4, show_default=True,
              help='Accepted for compatibility; inference runs on the GPU.')
@click.option('--frame_stride', default=16, show_default=True,
              help='Process every n-th frame (reference HEAD: 16; fixtures: 1).')
@click.option('--batch', default=64, show_default=True, help='Frames per GPU batch.')
def main(src, model, detection_treshold, display_image_height, df_dir, video_dir, threads,
         frame_stride, batch):
    """Track weight plates in the given videos and export the per-frame dataframes."""
    if df_dir is not None:
        os.makedirs(df_dir, exist_ok=True)
    if video_dir is not None:
        os.makedirs(video_dir, exist_ok=True)
    for s in src:
        if not os.path.isfile(s):
            raise FileNotFoundError()
        interpreter = Interpreter(model_path=model, num_threads=threads, max_batch=batch)   # one model load per video
        interpreter.allocate_tensors()
        video_path = None
        if video_dir is not None:
            video_path = os.path.join(video_dir, f'{os.path.basename(s).split(".")[0]}.mp4')
        data = track(s, interpreter, detection_treshold, display_image_height, video_path,
                     frame_stride=frame_stride, batch=batch)
        if df_dir is not None:
            export_dataframe(data, s, model, df_dir)


if __name__ == '__main__':
    main()
