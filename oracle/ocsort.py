"""CPU oracle for the tracker stage (rows a6-a8 of SURVEY.md section 8) -- TEST INFRASTRUCTURE.

Restates `OCSort(max_age=30, asso_func="diou", iou_threshold=0.1).update(dets, _)` as
the reference calls it (track.py:157,186-199) plus the row assembly of track.py:189-234.

The arithmetic lives in a third-party package that is ABSENT from /root/reference and
UNPINNED in requirements.txt: PyPI `ocsort` (the packaging of noahcao/OC_SORT with the
class/confidence columns, which is the variant whose output rows have the 7 columns
track.py:190 unpacks) on top of a filterpy-1.4.5-style Kalman filter
(requirements.txt:63).  This file restates that published algorithm [3P-MEM].

Parity status:
* Kalman filter + output gating + empty-frame skip: PINNED -- replaying the boxes
  rebuilt from dfs_ocsort/*.pkl.gz reproduces the stored dx,dy (= kf.x[4:6],
  track.py:199) bit-exactly on every track that is visible from its first hit
  (tests/test_oracle_ocsort.py).
* association internals (Hungarian tie-breaks, OCR round, `det_thresh` default, which
  column multiplies the velocity-direction cost): PARITY UNPINNED -- exercised only
  weakly by the fixtures (1-3 well separated plates).

Numerics: float64 everywhere, IEEE multiply and add kept separate and accumulated in
k-ascending order (what a BLAS dgemm does on these block-sparse 7x7 matrices: every
inner product has at most one inexact term).  The CUDA kernel
(`vbt_b200/csrc/tracker.cu`) uses the same order with FMA contraction disabled.
"""
from __future__ import annotations

import numpy as np

DIM_X, DIM_Z = 7, 4


def _mm(a, b):
    """Dense product accumulated k-ascending with separate multiply and add."""
    acc = np.zeros((a.shape[0], b.shape[1]))
    for k in range(a.shape[1]):
        acc = acc + a[:, k:k + 1] * b[k:k + 1, :]
    return acc


_F = np.eye(DIM_X)
_F[0, 4] = _F[1, 5] = _F[2, 6] = 1.0
_H = np.zeros((DIM_Z, DIM_X))
_H[0, 0] = _H[1, 1] = _H[2, 2] = _H[3, 3] = 1.0
_R = np.diag([1.0, 1.0, 10.0, 10.0])
_Q = np.diag([1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.0001])
_P0 = np.diag([10.0, 10.0, 10.0, 10.0, 10000.0, 10000.0, 10000.0])
_I = np.eye(DIM_X)


def box_to_z(b):
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, w / (h + 1e-6)])


def x_to_box(x):
    w = np.sqrt(x[2] * x[3])
    h = x[2] / w
    return np.array([x[0] - w / 2.0, x[1] - h / 2.0, x[0] + w / 2.0, x[1] + h / 2.0])


def unit_direction(b1, b2):
    """(dy, dx)/norm between two box centres (velocity of a track)."""
    cx1, cy1 = (b1[0] + b1[2]) / 2.0, (b1[1] + b1[3]) / 2.0
    cx2, cy2 = (b2[0] + b2[2]) / 2.0, (b2[1] + b2[3]) / 2.0
    dy, dx = cy2 - cy1, cx2 - cx1
    n = np.sqrt(dy ** 2 + dx ** 2) + 1e-6
    return np.array([dy / n, dx / n])


class Track:
    """One KalmanBoxTracker + its KalmanFilterNew, flattened."""

    def __init__(self, det5, cls, tid):
        self.x = np.zeros(DIM_X)
        self.x[:4] = box_to_z(det5)
        self.P = _P0.copy()
        self.id = tid
        self.tsu = 0                  # time_since_update
        self.hits = 0
        self.hit_streak = 0
        self.age = 0
        self.conf = det5[4]
        self.cls = cls
        self.last_obs = None          # placeholder [-1]*5 upstream
        self.obs = {}                 # age -> det5
        self.velocity = None
        self.observed = False         # KalmanFilterNew.observed
        self.frozen = None            # (x, P) at the first missed frame
        self.prev_z = None            # last real measurement fed to the filter
        self.missed = 0               # consecutive update(None) calls since prev_z

    # -- Kalman filter ---------------------------------------------------------------
    def kf_predict(self):
        self.x = _mm(_F, self.x[:, None])[:, 0]
        self.P = _mm(_mm(_F, self.P), _F.T) + _Q

    def kf_correct(self, z):
        y = z - _mm(_H, self.x[:, None])[:, 0]
        pht = _mm(self.P, _H.T)
        s = _mm(_H, pht) + _R
        si = np.zeros((DIM_Z, DIM_Z))
        for i in range(DIM_Z):        # S is diagonal for this model: inv == reciprocal
            si[i, i] = 1.0 / s[i, i]
        k = _mm(pht, si)
        self.x = self.x + _mm(k, y[:, None])[:, 0]
        ikh = _I - _mm(k, _H)
        self.P = _mm(_mm(ikh, self.P), ikh.T) + _mm(_mm(k, _R), k.T)

    def kf_update(self, z):
        if z is None:
            if self.observed:                       # first miss: freeze
                self.frozen = (self.x.copy(), self.P.copy())
            self.observed = False
            self.missed += 1
            return
        if (not self.observed) and self.frozen is not None:
            # observation-centric re-update: rewind to the frozen state and walk a
            # straight virtual trajectory from the previous to the new measurement
            self.x, self.P = self.frozen[0].copy(), self.frozen[1].copy()
            gap = self.missed + 1
            x1, y1, s1, r1 = self.prev_z
            x2, y2, s2, r2 = z
            w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
            w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
            dx, dy = (x2 - x1) / gap, (y2 - y1) / gap
            dw, dh = (w2 - w1) / gap, (h2 - h1) / gap
            virt = None
            for i in range(gap):
                w = w1 + (i + 1) * dw
                h = h1 + (i + 1) * dh
                virt = np.array([x1 + (i + 1) * dx, y1 + (i + 1) * dy, w * h, w / h])
                self.kf_correct(virt)
                if i != gap - 1:
                    self.kf_predict()
            # upstream rewinds its observation history together with the filter, so the
            # history ends with the last VIRTUAL box, not with z
            self.prev_z = virt
        else:
            self.prev_z = z.copy()
        self.observed = True
        self.missed = 0
        self.kf_correct(z)

    # -- KalmanBoxTracker ------------------------------------------------------------
    def predict(self):
        if self.x[6] + self.x[2] <= 0:
            self.x[6] *= 0.0
        self.kf_predict()
        self.age += 1
        if self.tsu > 0:
            self.hit_streak = 0
        self.tsu += 1
        return x_to_box(self.x)

    def update(self, det5, cls, delta_t):
        if det5 is None:
            self.kf_update(None)
            return
        self.conf = det5[4]
        self.cls = cls
        if self.last_obs is not None and self.last_obs.sum() >= 0:
            prev = None
            for i in range(delta_t):
                if self.age - (delta_t - i) in self.obs:
                    prev = self.obs[self.age - (delta_t - i)]
                    break
            if prev is None:
                prev = self.last_obs
            self.velocity = unit_direction(prev, det5)
        self.last_obs = det5
        self.obs[self.age] = det5
        self.tsu = 0
        self.hits += 1
        self.hit_streak += 1
        self.kf_update(box_to_z(det5))

    def k_previous(self, k):
        if not self.obs:
            return np.array([-1.0] * 5)
        for i in range(k):
            if self.age - (k - i) in self.obs:
                return self.obs[self.age - (k - i)]
        return self.obs[max(self.obs)]


def iou_matrix(a, b):
    """Plain IoU, [len(a), len(b)]."""
    a = a[:, None, :]
    b = b[None, :, :]
    w = np.maximum(0.0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]))
    h = np.maximum(0.0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]))
    wh = w * h
    with np.errstate(divide='ignore', invalid='ignore'):
        return wh / ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
                     + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - wh)


def diou_matrix(a, b):
    """(DIoU + 1) / 2, [len(a), len(b)]  (asso_func="diou", track.py:157)."""
    iou = iou_matrix(a, b)
    a = a[:, None, :]
    b = b[None, :, :]
    cxa, cya = (a[..., 0] + a[..., 2]) / 2.0, (a[..., 1] + a[..., 3]) / 2.0
    cxb, cyb = (b[..., 0] + b[..., 2]) / 2.0, (b[..., 1] + b[..., 3]) / 2.0
    inner = (cxa - cxb) ** 2 + (cya - cyb) ** 2
    ox1 = np.minimum(a[..., 0], b[..., 0])
    oy1 = np.minimum(a[..., 1], b[..., 1])
    ox2 = np.maximum(a[..., 2], b[..., 2])
    oy2 = np.maximum(a[..., 3], b[..., 3])
    outer = (ox2 - ox1) ** 2 + (oy2 - oy1) ** 2
    with np.errstate(divide='ignore', invalid='ignore'):
        return (iou - inner / outer + 1) / 2.0


def assign_min_cost(cost):
    """Rectangular linear assignment (shortest augmenting paths with potentials).

    Returns pairs (row, col) sorted by row for min(n, m) assignments.  The CUDA kernel
    runs the identical procedure (same scan order, same strict `<` tie rule), so ties
    resolve identically on both sides; against scipy/lap only the optimum is pinned.
    """
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    if n == 0 or m == 0:
        return []
    transposed = n > m
    if transposed:
        cost = cost.T
        n, m = m, n
    inf = np.inf
    u = np.zeros(n + 1)
    v = np.zeros(m + 1)
    p = np.zeros(m + 1, dtype=np.int64)       # p[j] = row matched to column j (1-based)
    way = np.zeros(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        p[0] = i
        j0 = 0
        minv = np.full(m + 1, inf)
        used = np.zeros(m + 1, dtype=bool)
        while True:
            used[j0] = True
            i0 = p[j0]
            delta = inf
            j1 = 0
            for j in range(1, m + 1):
                if not used[j]:
                    cur = cost[i0 - 1, j - 1] - u[i0] - v[j]
                    if cur < minv[j]:
                        minv[j] = cur
                        way[j] = j0
                    if minv[j] < delta:
                        delta = minv[j]
                        j1 = j
            if j1 == 0:
                raise ValueError('assignment cost matrix holds NaN/inf')
            for j in range(m + 1):
                if used[j]:
                    u[p[j]] += delta
                    v[j] -= delta
                else:
                    minv[j] -= delta
            j0 = j1
            if p[j0] == 0:
                break
        while True:
            j1 = way[j0]
            p[j0] = p[j1]
            j0 = j1
            if j0 == 0:
                break
    pairs = []
    for j in range(1, m + 1):
        if p[j] != 0:
            pairs.append((j - 1, p[j] - 1) if transposed else (p[j] - 1, j - 1))
    pairs.sort()
    return pairs


class OCSortOracle:
    def __init__(self, det_thresh=0.2, max_age=30, min_hits=3, iou_threshold=0.1,
                 delta_t=3, inertia=0.2, vdc_uses_class_column=True):
        self.det_thresh = det_thresh
        self.max_age = max_age
        self.min_hits = min_hits
        self.iou_threshold = iou_threshold
        self.delta_t = delta_t
        self.inertia = inertia
        self.vdc_uses_class_column = vdc_uses_class_column
        self.tracks = []
        self.frame_count = 0
        self.next_id = 0

    def _first_round(self, dets, trk_boxes):
        nd, nt = len(dets), len(self.tracks)
        if nt == 0:
            return [], list(range(nd)), []
        iou = iou_matrix(dets[:, :4], trk_boxes) if nd else np.zeros((0, nt))
        pairs = []
        if nd > 0:
            hit = iou > self.iou_threshold
            if hit.sum(1).max() == 1 and hit.sum(0).max() == 1:
                pairs = [(int(d), int(t)) for d, t in zip(*np.where(hit))]
            else:
                vel = np.array([t.velocity if t.velocity is not None else np.zeros(2)
                                for t in self.tracks])
                prev = np.array([t.k_previous(self.delta_t) for t in self.tracks])
                cxd = (dets[:, 0] + dets[:, 2]) / 2.0
                cyd = (dets[:, 1] + dets[:, 3]) / 2.0
                cxp = (prev[:, 0] + prev[:, 2]) / 2.0
                cyp = (prev[:, 1] + prev[:, 3]) / 2.0
                ddx = cxd[None, :] - cxp[:, None]            # [trk, det]
                ddy = cyd[None, :] - cyp[:, None]
                norm = np.sqrt(ddx ** 2 + ddy ** 2) + 1e-6
                ddx, ddy = ddx / norm, ddy / norm
                cosang = np.clip(vel[:, 1:2] * ddx + vel[:, 0:1] * ddy, -1, 1)
                ang = (np.pi / 2.0 - np.abs(np.arccos(cosang))) / np.pi
                valid = (prev[:, 4] >= 0).astype(np.float64)[:, None]
                mult = dets[:, 5] if self.vdc_uses_class_column else dets[:, 4]
                angle_cost = ((valid * ang) * self.inertia).T * mult[:, None]
                pairs = assign_min_cost(-(iou + angle_cost))
        md = {d for d, _ in pairs}
        mt = {t for _, t in pairs}
        un_d = [d for d in range(nd) if d not in md]
        un_t = [t for t in range(nt) if t not in mt]
        matches = []
        for d, t in pairs:
            if iou[d, t] < self.iou_threshold:
                un_d.append(d)
                un_t.append(t)
            else:
                matches.append((d, t))
        return matches, un_d, un_t

    def update(self, dets):
        """dets: float64 [N,6] = x1,y1,x2,y2,score,cls (odt.py:102-118).

        Returns float64 [M,9]: x1,y1,x2,y2,id+1,cls,conf,kf_dx,kf_dy -- the 7 upstream
        columns plus kf.x[4:6] of the emitting track (what track.py:194-199 reads)."""
        dets = np.asarray(dets, dtype=np.float64).reshape(-1, 6)
        self.frame_count += 1
        dets = dets[dets[:, 4] > self.det_thresh]
        keep, boxes = [], []
        for t in self.tracks:
            b = t.predict()
            if not np.any(np.isnan(b)):
                keep.append(t)
                boxes.append(b)
        self.tracks = keep
        trk_boxes = np.array(boxes).reshape(-1, 4)
        last_boxes = [t.last_obs if t.last_obs is not None else np.array([-1.0] * 5)
                      for t in self.tracks]

        matches, un_d, un_t = self._first_round(dets, trk_boxes)
        for d, t in matches:
            self.tracks[t].update(dets[d, :5].copy(), dets[d, 5], self.delta_t)

        if un_d and un_t:                                            # OCR round
            left = diou_matrix(dets[un_d][:, :4], np.array([last_boxes[t][:4] for t in un_t]))
            if left.max() > self.iou_threshold:
                rm_d, rm_t = [], []
                for a, b in assign_min_cost(-left):
                    if left[a, b] < self.iou_threshold:
                        continue
                    d, t = un_d[a], un_t[b]
                    self.tracks[t].update(dets[d, :5].copy(), dets[d, 5], self.delta_t)
                    rm_d.append(d)
                    rm_t.append(t)
                un_d = sorted(set(un_d) - set(rm_d))
                un_t = sorted(set(un_t) - set(rm_t))
        for t in un_t:
            self.tracks[t].update(None, None, self.delta_t)
        for d in un_d:
            self.tracks.append(Track(dets[d, :5].copy(), dets[d, 5], self.next_id))
            self.next_id += 1

        out = []
        survivors = []
        for t in reversed(self.tracks):
            if t.last_obs is None or t.last_obs.sum() < 0:
                box = x_to_box(t.x)
            else:
                box = t.last_obs[:4]
            if t.tsu < 1 and (t.hit_streak >= self.min_hits
                              or self.frame_count <= self.min_hits):
                out.append(np.concatenate((box, [t.id + 1, t.cls, t.conf, t.x[4], t.x[5]])))
            if not (t.tsu > self.max_age):
                survivors.append(t)
        self.tracks = survivors[::-1]
        return np.array(out).reshape(-1, 9)


def track_rows(frames_dets, fps, frame_numbers=None, **kw):
    """track.py:159-234 over pre-computed detections.

    frames_dets: list of [N,6] arrays (one per processed frame; empty frames are skipped
    and do NOT step the tracker, track.py:180-181).  frame_numbers: 1-based frame_count
    of each entry (track.py:161); default 1..n.  Returns float64 [rows, 8] in append
    order: id,time,x,y,dx,dy,norm_plate_height,norm_plate_width."""
    trk = OCSortOracle(**kw)
    rows = []
    for i, d in enumerate(frames_dets):
        fc = frame_numbers[i] if frame_numbers is not None else i + 1
        d = np.asarray(d, dtype=np.float64).reshape(-1, 6)
        if len(d) == 0:
            continue
        time = fc / fps
        for r in trk.update(d):
            xmin, ymin, xmax, ymax = r[:4]
            rows.append([int(r[4]), time, (xmin + xmax) / 2, (ymin + ymax) / 2, r[7], r[8],
                         abs(ymax - ymin), abs(xmax - xmin)])
    return np.array(rows, dtype=np.float64).reshape(-1, 8)
