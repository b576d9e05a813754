"""Generate the committed golden vectors from the reference checkout.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

Writes
* tests/golden/dfs_ocsort.npz   -- the 34 per-frame DataFrames of dfs_ocsort/ (rows in
  their original append order = DataFrame index order, float64 [n,8]
  id,time,x,y,dx,dy,norm_plate_height,norm_plate_width; the index; fps; file name),
* tests/golden/velocity_phases.npz -- for every (file, id): the phases the LIVE reference
  classes (plot.analyze_df -> VelocityTracker/RunningAverage/Phase) produce on the
  plot.py-smoothed series, float64 [k,6] t_start,t_end,y_start,y_end,rom,type; plus the
  same for qualysis_dfs/ as extra known-answer inputs,
* tests/golden/eval_detections_stats.json -- shape statistics of dfs/eval_detections.pkl.gz
  (25 detections per image, 1/256 score grid, score-descending blocks),
* tests/golden/figs_ocsort_labels.json -- the ROM / ACV text labels (2 decimals,
  plot.py:178,186) pulled out of figs_ocsort/*.pdf by inflating the PDF streams.

Nothing here is reference source code; it is data the reference ships plus outputs of
the reference run on that data.
"""
import glob
import json
import os
import re
import sys
import zlib
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

REF = os.environ.get('VBT_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_plot():
    for m in ('seaborn', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import plot  # noqa: the reference's own module
    return plot


def smooth_like_plot(df, tid):
    d = df.query(f'id == {tid}').drop(columns=['id'])
    for col in ['x', 'y', 'dx', 'dy']:
        d[col] = d[col].rolling(window=5, center=False, min_periods=1).mean()
    for col in ['norm_plate_height', 'norm_plate_width']:
        d[col] = d[col].expanding(min_periods=1).mean()
    return d


def pdf_labels(path):
    data = open(path, 'rb').read()
    labels = []
    for m in re.finditer(rb'stream\r?\n(.*?)endstream', data, re.S):
        try:
            txt = zlib.decompress(m.group(1))
        except zlib.error:
            continue
        for s in re.findall(rb'\((.*?)\)', txt):
            if re.fullmatch(rb'\d\.\d\d', s):
                labels.append(s.decode())
    return sorted(labels)


def main():
    plot = load_reference_plot()
    tables, phases, labels = {}, {}, {}
    for sub in ('dfs_ocsort', 'qualysis_dfs'):
        for p in sorted(glob.glob(os.path.join(REF, sub, '*.pkl.gz'))):
            name = os.path.basename(p)
            m = plot.filename_regexp.match(name)
            if not m:
                continue
            df = pd.read_pickle(p)
            key = f'{sub}/{name[:-len(".pkl.gz")]}'
            if sub == 'dfs_ocsort':
                d = df.sort_index()
                tables[key + '|rows'] = d.to_numpy(dtype=np.float64)
                tables[key + '|index'] = d.index.to_numpy(dtype=np.int64)
            for tid in sorted(df['id'].unique()):
                sm = smooth_like_plot(df, tid)
                ph = plot.analyze_df(sm, 0.45)
                arr = np.array([[q.time_start, q.time_end, q.y_start, q.y_end, q.rom, q.type]
                                for q in ph], dtype=np.float64).reshape(-1, 6)
                phases[f'{key}|{tid}|phases'] = arr
                if sub != 'dfs_ocsort':     # raw series needed as kernel input
                    phases[f'{key}|{tid}|raw'] = df.query(f'id == {tid}').drop(
                        columns=['id']).to_numpy(dtype=np.float64)
    for p in sorted(glob.glob(os.path.join(REF, 'figs_ocsort', '*.pdf'))):
        labels[os.path.basename(p)[:-4]] = pdf_labels(p)
    # detector known-answer STATISTICS (the per-image detections need the absent .tflite
    # weights to reproduce; what the post-process oracle can be held to is their shape)
    ev = pd.read_pickle(os.path.join(REF, 'dfs', 'eval_detections.pkl.gz'))
    sc = ev['Score'].to_numpy().reshape(-1, 25)
    stats = {
        'rows': int(len(ev)), 'models': sorted(ev['Model'].unique().tolist()),
        'detections_per_image': 25, 'images_per_model': int(len(ev) // 25 // 6),
        'scores_on_1_256_grid': bool(np.all(sc * 256 == np.round(sc * 256))),
        'blocks_score_descending': bool(np.all(np.diff(sc, axis=1) <= 0)),
        'min_score': float(sc.min()), 'max_score': float(sc.max()),
    }
    with open(os.path.join(HERE, 'eval_detections_stats.json'), 'w') as f:
        json.dump(stats, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, 'dfs_ocsort.npz'), **tables)
    np.savez_compressed(os.path.join(HERE, 'velocity_phases.npz'), **phases)
    with open(os.path.join(HERE, 'figs_ocsort_labels.json'), 'w') as f:
        json.dump(labels, f, indent=0, sort_keys=True)
    print({k: os.path.getsize(os.path.join(HERE, k)) for k in
           ('dfs_ocsort.npz', 'velocity_phases.npz', 'figs_ocsort_labels.json')})


if __name__ == '__main__':
    main()
