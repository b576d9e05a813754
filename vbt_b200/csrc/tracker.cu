// K7 -- OC-SORT plate tracker, one warp per video, state resident on the device.
//
// replaces: ocsort.OCSort(max_age=30, asso_func="diou", iou_threshold=0.1) and
// tracker.update(dets, []) (track.py:157,186-187) plus the row assembly of
// track.py:189-234.  The package source is not part of the reference checkout; the
// algorithm restated here is the published OC-SORT (observation-centric re-update,
// velocity-direction cost, second association round on last observations) over a
// filterpy-style 7-state constant-velocity Kalman filter -- see oracle/ocsort.py for the
// CPU restatement this kernel is checked against and DESIGN.md for what is pinned.
//
// Layout: lanes of the warp own tracks (Kalman predict / correct are lane-parallel),
// lane 0 runs the association bookkeeping.  fp64, -fmad=false, dense 7x7 products
// accumulated k-ascending so results are bit-identical to the oracle.
// Latency-bound sequential recurrence over the frame axis: no roofline (us per frame).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kMaxT = 256;  // live tracks per video (compile-time ceiling)
constexpr int kMaxD = 32;   // detections per frame
constexpr int NX = 7, NZ = 4;
constexpr int kMaxF = 512;  // frames per launch whose counts / numbers are staged in shared memory
constexpr int NP = 13;      // stored covariance entries, see below

// The covariance of this filter never leaves a block structure: P0, Q and R are diagonal,
// F couples state i only with i+4 (position <- velocity) and H reads states 0..3, so the
// pairs (0,4), (1,5), (2,6) and the lone state 3 never mix.  Every other entry of the dense
// 7x7 matrices the oracle multiplies is an exact zero, and adding 0 * finite terms does not
// change an IEEE sum, so only the 13 in-block entries are stored and updated -- with the
// SAME operations in the SAME order as the dense k-ascending products (checked bit for bit
// against oracle/ocsort.py).  Layout: block b in 0..2 (a = b, v = b + 4):
//   P[4b+0] = P(a,a)  P[4b+1] = P(a,v)  P[4b+2] = P(v,a)  P[4b+3] = P(v,v);  P[12] = P(3,3)
struct Trk {
  double x[NX];
  double P[NP];
  double fx[NX];            // state frozen at the first missed frame
  double fP[NP];
  double prev_z[NZ];        // measurement the next re-update interpolates from
  double last_obs[5];
  double obs_box[4][5];     // observations of the last ages, slot = age & 3
  double vel[2];            // (dy, dx) unit direction
  double conf, cls;
  int obs_age[4];
  int id, tsu, hits, hit_streak, age;
  int has_last, has_vel, observed, has_frozen, missed, n_obs;
};

struct Video {
  int frame_count, next_id, n_tracks, status;
  int order[kMaxT];         // slot of the i-th tracker in list order
  unsigned char used[kMaxT];
};

struct Params {
  double det_thresh, iou_threshold, inertia;
  int max_age, min_hits, delta_t, vdc_cls;
};

// x = F x ; P = F P F' + Q.  Dense form: t = F P adds row v to row a, P' = t F' adds column
// v to column a, Q lands on the diagonal.
__device__ __forceinline__ void kf_predict(double* x, double* P) {
  x[0] = x[0] + x[4]; x[1] = x[1] + x[5]; x[2] = x[2] + x[6];
  const double qa[3] = {1.0, 1.0, 1.0}, qv[3] = {0.01, 0.01, 0.0001};
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const double t_aa = P[4 * b + 0] + P[4 * b + 2];
    const double t_av = P[4 * b + 1] + P[4 * b + 3];
    const double t_va = P[4 * b + 2];
    const double t_vv = P[4 * b + 3];
    P[4 * b + 0] = (t_aa + t_av) + qa[b];
    P[4 * b + 1] = t_av;
    P[4 * b + 2] = t_va + t_vv;
    P[4 * b + 3] = t_vv + qv[b];
  }
  P[12] = P[12] + 1.0;
}

// y = z - Hx ; S = HPH' + R (diagonal) ; K = PH' inv(S) as a multiply by the reciprocal ;
// x += K y ; P = (I-KH) P (I-KH)' + K R K'  (Joseph form), restricted to the blocks.
__device__ __forceinline__ void kf_correct(double* x, double* P, const double* z) {
  const double R[NZ] = {1.0, 1.0, 10.0, 10.0};
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const double Paa = P[4 * b + 0], Pav = P[4 * b + 1], Pva = P[4 * b + 2], Pvv = P[4 * b + 3];
    const double y = z[b] - x[b];
    const double si = 1.0 / (Paa + R[b]);
    const double Ka = Paa * si, Kv = Pva * si;
    x[b] = x[b] + Ka * y;
    x[b + 4] = x[b + 4] + Kv * y;
    const double ia = 1.0 - Ka, iv = 0.0 - Kv;          // (I - KH)(a,a), (I - KH)(v,a)
    const double t_aa = ia * Paa, t_av = ia * Pav;
    const double t_va = iv * Paa + Pva, t_vv = iv * Pav + Pvv;
    const double ra = Ka * R[b], rv = Kv * R[b];
    P[4 * b + 0] = t_aa * ia + ra * Ka;
    P[4 * b + 1] = (t_aa * iv + t_av) + ra * Kv;
    P[4 * b + 2] = t_va * ia + rv * Ka;
    P[4 * b + 3] = (t_va * iv + t_vv) + rv * Kv;
  }
  {
    const double P33 = P[12];
    const double y = z[3] - x[3];
    const double si = 1.0 / (P33 + R[3]);
    const double K3 = P33 * si;
    x[3] = x[3] + K3 * y;
    const double i3 = 1.0 - K3;
    const double t = i3 * P33;
    P[12] = t * i3 + (K3 * R[3]) * K3;
  }
}

__device__ __forceinline__ void box_to_z(const double* b, double* z) {
  double w = b[2] - b[0], h = b[3] - b[1];
  z[0] = b[0] + w / 2.0; z[1] = b[1] + h / 2.0; z[2] = w * h; z[3] = w / (h + 1e-6);
}

__device__ __forceinline__ void x_to_box(const double* x, double* b) {
  double w = sqrt(x[2] * x[3]);
  double h = x[2] / w;
  b[0] = x[0] - w / 2.0; b[1] = x[1] - h / 2.0; b[2] = x[0] + w / 2.0; b[3] = x[1] + h / 2.0;
}

__device__ void kf_update(Trk& t, const double* z) {   // z == nullptr: no observation
  if (!z) {
    if (t.observed) {
      for (int i = 0; i < NX; ++i) t.fx[i] = t.x[i];
      for (int i = 0; i < NP; ++i) t.fP[i] = t.P[i];
      t.has_frozen = 1;
    }
    t.observed = 0;
    t.missed += 1;
    return;
  }
  double x[NX], P[NP];
  if (!t.observed && t.has_frozen) {        // observation-centric re-update
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = t.fx[i];
#pragma unroll
    for (int i = 0; i < NP; ++i) P[i] = t.fP[i];
    int gap = t.missed + 1;
    double x1 = t.prev_z[0], y1 = t.prev_z[1], s1 = t.prev_z[2], r1 = t.prev_z[3];
    double w1 = sqrt(s1 * r1), h1 = sqrt(s1 / r1);
    double w2 = sqrt(z[2] * z[3]), h2 = sqrt(z[2] / z[3]);
    double dx = (z[0] - x1) / gap, dy = (z[1] - y1) / gap;
    double dw = (w2 - w1) / gap, dh = (h2 - h1) / gap;
    double v[NZ] = {0, 0, 0, 0};
    for (int i = 0; i < gap; ++i) {
      double w = w1 + (i + 1) * dw, h = h1 + (i + 1) * dh;
      v[0] = x1 + (i + 1) * dx; v[1] = y1 + (i + 1) * dy; v[2] = w * h; v[3] = w / h;
      kf_correct(x, P, v);
      if (i != gap - 1) kf_predict(x, P);
    }
    for (int i = 0; i < NZ; ++i) t.prev_z[i] = v[i];   // history ends with the virtual box
  } else {
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = t.x[i];
#pragma unroll
    for (int i = 0; i < NP; ++i) P[i] = t.P[i];
    for (int i = 0; i < NZ; ++i) t.prev_z[i] = z[i];
  }
  t.observed = 1;
  t.missed = 0;
  kf_correct(x, P, z);
#pragma unroll
  for (int i = 0; i < NX; ++i) t.x[i] = x[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) t.P[i] = P[i];
}

__device__ void trk_predict(Trk& t, double* box) {
  if (t.x[6] + t.x[2] <= 0) t.x[6] *= 0.0;
  double x[NX], P[NP];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = t.x[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = t.P[i];
  kf_predict(x, P);
#pragma unroll
  for (int i = 0; i < NX; ++i) t.x[i] = x[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) t.P[i] = P[i];
  t.age += 1;
  if (t.tsu > 0) t.hit_streak = 0;
  t.tsu += 1;
  x_to_box(t.x, box);
}

__device__ __forceinline__ const double* obs_at(const Trk& t, int age) {
  if (age < 0) return nullptr;
  return (t.obs_age[age & 3] == age) ? t.obs_box[age & 3] : nullptr;
}

// observation delta_t frames back (or the nearest newer one), else the latest, else null
__device__ const double* k_previous(const Trk& t, int k) {
  if (t.n_obs == 0) return nullptr;
  for (int i = 0; i < k; ++i) {
    const double* o = obs_at(t, t.age - (k - i));
    if (o) return o;
  }
  return t.last_obs;
}

__device__ void trk_update(Trk& t, const double* det, int delta_t) {  // det: 6 doubles or null
  if (!det) { kf_update(t, nullptr); return; }
  t.conf = det[4];
  t.cls = det[5];
  if (t.has_last) {
    double sum = t.last_obs[0] + t.last_obs[1] + t.last_obs[2] + t.last_obs[3] + t.last_obs[4];
    if (sum >= 0) {
      const double* prev = nullptr;
      for (int i = 0; i < delta_t && !prev; ++i) prev = obs_at(t, t.age - (delta_t - i));
      if (!prev) prev = t.last_obs;
      double cx1 = (prev[0] + prev[2]) / 2.0, cy1 = (prev[1] + prev[3]) / 2.0;
      double cx2 = (det[0] + det[2]) / 2.0, cy2 = (det[1] + det[3]) / 2.0;
      double dy = cy2 - cy1, dx = cx2 - cx1;
      double n = sqrt(dy * dy + dx * dx) + 1e-6;
      t.vel[0] = dy / n; t.vel[1] = dx / n;
      t.has_vel = 1;
    }
  }
  for (int i = 0; i < 5; ++i) { t.last_obs[i] = det[i]; t.obs_box[t.age & 3][i] = det[i]; }
  t.obs_age[t.age & 3] = t.age;
  t.has_last = 1;
  t.n_obs += 1;
  t.tsu = 0;
  t.hits += 1;
  t.hit_streak += 1;
  double z[NZ];
  box_to_z(det, z);
  kf_update(t, z);
}

__device__ void trk_init(Trk& t, const double* det, int id) {
  for (int i = 0; i < NX; ++i) t.x[i] = 0.0;
  double z[NZ];
  box_to_z(det, z);
  for (int i = 0; i < NZ; ++i) t.x[i] = z[i];
  for (int i = 0; i < NP; ++i) t.P[i] = 0.0;
  for (int b = 0; b < 3; ++b) { t.P[4 * b + 0] = 10.0; t.P[4 * b + 3] = 10000.0; }
  t.P[12] = 10.0;
  t.id = id; t.tsu = 0; t.hits = 0; t.hit_streak = 0; t.age = 0;
  t.conf = det[4]; t.cls = det[5];
  t.has_last = 0; t.has_vel = 0; t.observed = 0; t.has_frozen = 0; t.missed = 0; t.n_obs = 0;
  for (int i = 0; i < 5; ++i) t.last_obs[i] = -1.0;
  for (int i = 0; i < 4; ++i) t.obs_age[i] = -1;
  t.vel[0] = t.vel[1] = 0.0;
  for (int i = 0; i < NZ; ++i) t.prev_z[i] = 0.0;
}

__device__ __forceinline__ double iou_of(const double* a, const double* b) {
  double w = fmax(0.0, fmin(a[2], b[2]) - fmax(a[0], b[0]));
  double h = fmax(0.0, fmin(a[3], b[3]) - fmax(a[1], b[1]));
  double wh = w * h;
  return wh / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - wh);
}

__device__ __forceinline__ double diou_of(const double* a, const double* b) {
  double iou = iou_of(a, b);
  double cxa = (a[0] + a[2]) / 2.0, cya = (a[1] + a[3]) / 2.0;
  double cxb = (b[0] + b[2]) / 2.0, cyb = (b[1] + b[3]) / 2.0;
  double ex = cxa - cxb, ey = cya - cyb;
  double inner = ex * ex + ey * ey;
  double ox = fmax(a[2], b[2]) - fmin(a[0], b[0]);
  double oy = fmax(a[3], b[3]) - fmin(a[1], b[1]);
  double outer = ox * ox + oy * oy;
  return (iou - inner / outer + 1) / 2.0;
}

// Rectangular min-cost assignment, same procedure and tie rule as
// oracle/ocsort.py:assign_min_cost.  cost is [n][ld] (row-major, m used columns).
// Writes row_of_col[j] (-1 = free) for the ORIGINAL orientation via out_row/out_col pairs.
__device__ int assign_min_cost(const double* cost, int ld, int n, int m, int* pair_r,
                               int* pair_c) {
  if (n == 0 || m == 0) return 0;
  const bool tr = n > m;
  const int N = tr ? m : n, M = tr ? n : m;
  double u[kMaxT + 1], v[kMaxT + 1], minv[kMaxT + 1];
  int p[kMaxT + 1], way[kMaxT + 1];
  bool used[kMaxT + 1];
  for (int j = 0; j <= M; ++j) { v[j] = 0.0; p[j] = 0; way[j] = 0; }
  for (int i = 0; i <= N; ++i) u[i] = 0.0;
  for (int i = 1; i <= N; ++i) {
    p[0] = i;
    int j0 = 0;
    for (int j = 0; j <= M; ++j) { minv[j] = INFINITY; used[j] = false; }
    while (true) {
      used[j0] = true;
      int i0 = p[j0], j1 = 0;
      double delta = INFINITY;
      for (int j = 1; j <= M; ++j) {
        if (used[j]) continue;
        double cij = tr ? cost[(j - 1) * ld + (i0 - 1)] : cost[(i0 - 1) * ld + (j - 1)];
        double cur = cij - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      if (j1 == 0) return -1;          // NaN / inf costs: no augmenting column exists
      for (int j = 0; j <= M; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
      if (p[j0] == 0) break;
    }
    while (true) {
      int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
      if (j0 == 0) break;
    }
  }
  // emit pairs sorted by row (row = first index of the original orientation)
  int cnt = 0;
  if (!tr) {
    for (int i = 1; i <= N; ++i)
      for (int j = 1; j <= M; ++j)
        if (p[j] == i) { pair_r[cnt] = i - 1; pair_c[cnt] = j - 1; ++cnt; }
  } else {
    for (int j = 1; j <= M; ++j)
      if (p[j] != 0) { pair_r[cnt] = j - 1; pair_c[cnt] = p[j] - 1; ++cnt; }
  }
  return cnt;
}

// Warp-parallel form of assign_min_cost for min(n,m) <= max(n,m) <= 31: lane j owns column
// j of the (possibly transposed) problem -- v[j], minv[j], way[j], used[j], p[j] live in
// registers, u[] in shared memory; the column scan of the serial procedure becomes one
// step per lane plus a warp arg-min that breaks ties towards the lowest column, which is
// what the serial strict `<` scan does.  Same floating-point expressions, same result.
// All 32 lanes must call it; returns the pair count (-1: no augmenting column).
__device__ int assign_min_cost_warp(const double* cost, int ld, int n, int m, int* pair_r, int* pair_c,
                                    double* scratch /* [32] shared */, int lane) {
  if (n == 0 || m == 0) return 0;
  const bool tr = n > m;
  const int N = tr ? m : n, M = tr ? n : m;
  const unsigned full = 0xffffffffu;
  // lane j: column j (v, minv, way, used, p); lane r: row r (its potential u, whether it is on the alternating tree)
  double v = 0.0, minv = INFINITY, u = 0.0;
  int p = 0, way = 0;
  bool used = false;
  const bool col = lane >= 1 && lane <= M;
  for (int i = 1; i <= N; ++i) {
    if (lane == 0) p = i;
    int j0 = 0;
    minv = INFINITY;
    used = false;
    bool in_tree = false;
    while (true) {
      if (lane == j0) used = true;
      const int i0 = __shfl_sync(full, p, j0);
      if (lane == i0) in_tree = true;   // the rows whose potential moves: p[j] of the used columns (the serial u[p[j]] += delta)
      const double u_i0 = __shfl_sync(full, u, i0);
      double key = INFINITY;
      if (col && !used) {
        const double cij = tr ? cost[(lane - 1) * ld + (i0 - 1)] : cost[(i0 - 1) * ld + (lane - 1)];
        const double cur = cij - u_i0 - v;
        if (cur < minv) { minv = cur; way = j0; }
        key = minv;
      }
      // arg-min over the candidate columns (key below +inf), ties to the lowest column -- the serial scan's strict
      // `<`: the keys as order-preserving 64-bit integers (-0.0 folded onto +0.0 first: equal as doubles), high
      // words through one warp min-reduction, low words of the survivors through a second, the lowest lane of
      // what is left by ballot.  Five dependent warp operations instead of three shuffles per halving step.
      const bool cand = key < INFINITY;
      const unsigned long long kb = (unsigned long long)__double_as_longlong(key + 0.0);
      const unsigned long long ko = !cand ? ~0ull : ((kb >> 63) ? ~kb : (kb | 0x8000000000000000ull));
      const unsigned hi = (unsigned)(ko >> 32), lo = (unsigned)ko;
      const unsigned min_hi = __reduce_min_sync(full, hi);
      const unsigned min_lo = __reduce_min_sync(full, hi == min_hi ? lo : 0xffffffffu);
      const unsigned win = __ballot_sync(full, cand && hi == min_hi && lo == min_lo);
      if (win == 0) return -1;
      const int j1 = __ffs(win) - 1;
      const double delta = __shfl_sync(full, key, j1);
      if (in_tree) u += delta;
      if (used) v -= delta;             // lanes 0..M that are on the alternating tree
      else if (col) minv -= delta;
      j0 = j1;
      if (__shfl_sync(full, p, j0) == 0) break;
    }
    while (true) {                      // augment along the way[] chain
      const int j1 = __shfl_sync(full, way, j0);
      const int pj1 = __shfl_sync(full, p, j1);
      if (lane == j0) p = pj1;
      j0 = j1;
      if (j0 == 0) break;
    }
  }
  // pairs sorted by row of the ORIGINAL orientation
  int cnt;
  if (!tr) {
    // row r (1..N) is matched to the column whose p == r: the columns post themselves at their row's slot
    int* col_of_row = reinterpret_cast<int*>(scratch);
    col_of_row[lane] = -1;
    __syncwarp();
    if (col && p != 0) col_of_row[p] = lane;
    __syncwarp();
    const int c = (lane >= 1 && lane <= N) ? col_of_row[lane] : -1;
    const unsigned mask = __ballot_sync(full, c >= 0);
    if (c >= 0) { const int pos = __popc(mask & ((1u << lane) - 1u)); pair_r[pos] = lane - 1; pair_c[pos] = c - 1; }
    cnt = __popc(mask);
  } else {
    const bool has = col && p != 0;     // original row = this column
    const unsigned mask = __ballot_sync(full, has);
    if (has) { const int pos = __popc(mask & ((1u << lane) - 1u)); pair_r[pos] = lane - 1; pair_c[pos] = p - 1; }
    cnt = __popc(mask);
  }
  __syncwarp();
  return cnt;
}

constexpr int kCache = 32;  // track slots 0..kCache-1 are staged in shared memory for the launch

struct Shared {                 // carved from dynamic shared memory, ld = max_tracks
  double (*dets)[6];            // [kMaxD][6]
  double (*tbox)[4];            // [ld][4] predicted boxes, list order
  double* iou;                  // [kMaxD][ld]
  double* cost;                 // [kMaxD][ld]
  int *pair_d, *pair_t;         // [kMaxD]
  int *un_d, *un_t;             // [kMaxD], [ld]
  int *pr, *pc;                 // [kMaxD] second-round assignment (positions in un_d / un_t)
  unsigned char* flag_t;        // [ld]  per-track scratch flags
  unsigned char* flag_d;        // [kMaxD]
  Trk* cache;                   // [kCache] low slots of this video's track table
  Video* vid;                   // this video's list state
  double u[32];                 // row potentials of the warp-parallel assignment
  int cnt[kMaxF], fno[kMaxF];   // detection count / frame number of every frame of this launch
  int n_pairs, n_un_d, n_un_t;
};

__host__ __device__ inline size_t shared_bytes(int ld) {
  return sizeof(double) * (kMaxD * 6 + (size_t)ld * 4 + 2 * (size_t)kMaxD * ld) +
         sizeof(int) * (5 * kMaxD + (size_t)ld) + ((size_t)ld + 15) / 16 * 16 + kMaxD + 16 +
         sizeof(Trk) * kCache + sizeof(Video);
}

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// One warp steps one video through its next frames.  Every list operation of the
// reference's Python (filtering detections, dropping tracks, building the matched /
// unmatched index lists, emitting rows in reverse list order, deleting old tracks) is an
// order-preserving compaction, done here with warp ballots over 32-wide chunks so that no
// lane ever walks a list alone; only births (rare, order-dependent slot allocation) and
// the > 31-wide assignment fallback are serial.
// VBT_TRK_DBG=1: cycles per phase of video 0 (lane 0), summed over the frames of every launch
__device__ long long g_trk_dbg[16];
__device__ int g_trk_dbg_on;
#define TRK_TICK(i)                                                                          \
  do {                                                                                       \
    if (DBG && dbg) { const long long n__ = clock64(); g_trk_dbg[i] += n__ - dbg_t; dbg_t = n__; }   \
  } while (0)

template <bool DBG>
__global__ void __maxnreg__(255) tracker_update_kernel(
    Video* videos, Trk* tracks, Params prm, const double* dets, const int32_t* det_count,
    const int32_t* frame_no, const double* fps, const int32_t* n_frames, int F, int max_det, int max_tracks,
    double* rows, int32_t* row_count, int row_cap, double* last_out, int32_t* last_out_count,
    double* row_details) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ Shared sh;
  const unsigned full = 0xffffffffu;
  const int v = blockIdx.x, lane = threadIdx.x;
  const bool dbg = DBG && g_trk_dbg_on && v == 0 && lane == 0;
  long long dbg_t = dbg ? clock64() : 0;
  if (lane == 0) {
    unsigned char* p = dyn_smem;
    sh.dets = reinterpret_cast<double(*)[6]>(p); p += sizeof(double) * kMaxD * 6;
    sh.tbox = reinterpret_cast<double(*)[4]>(p); p += sizeof(double) * 4 * (size_t)max_tracks;
    sh.iou = reinterpret_cast<double*>(p); p += sizeof(double) * (size_t)kMaxD * max_tracks;
    sh.cost = reinterpret_cast<double*>(p); p += sizeof(double) * (size_t)kMaxD * max_tracks;
    sh.pair_d = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.pair_t = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.un_d = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.pr = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.pc = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.un_t = reinterpret_cast<int*>(p); p += sizeof(int) * (size_t)max_tracks;
    sh.flag_t = p; p += ((size_t)max_tracks + 15) / 16 * 16;
    sh.flag_d = p; p += kMaxD;
    p = reinterpret_cast<unsigned char*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    sh.cache = reinterpret_cast<Trk*>(p); p += sizeof(Trk) * kCache;
    sh.vid = reinterpret_cast<Video*>(p);
  }
  __syncwarp();
  const int ld = max_tracks;
  // The recurrence is a chain of dependent accesses to a few KB of state: stage the list
  // state and the low track slots (births take the lowest free slot) in shared memory
  // for the whole launch, write them back at the end.
  static_assert(sizeof(Trk) % 8 == 0 && sizeof(Video) % 4 == 0, "word copies");
  Video& vid = *sh.vid;
  Trk* gtrk = tracks + (size_t)v * max_tracks;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(videos + v);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sh.vid);
    for (int i = lane; i < (int)(sizeof(Video) / 4); i += 32) dst[i] = src[i];
    __syncwarp();
    constexpr int W = sizeof(Trk) / 8;
    for (int s = 0; s < kCache && s < max_tracks; ++s) {
      if (!vid.used[s]) continue;
      const uint64_t* a = reinterpret_cast<const uint64_t*>(gtrk + s);
      uint64_t* b = reinterpret_cast<uint64_t*>(sh.cache + s);
      for (int i = lane; i < W; i += 32) b[i] = a[i];
    }
    __syncwarp();
  }
  Trk* const cache = sh.cache;
  auto T = [&](int slot) -> Trk& { return slot < kCache ? cache[slot] : gtrk[slot]; };
  const int nf = min(n_frames[v], F);
  const double vfps = fps[v];
  double* vrows = rows + (size_t)v * row_cap * VBT_ROW_COLS;
  int rc = row_count[v];
  int frame_count = vid.frame_count;
  const double thr = prm.iou_threshold;

  // detections of the next frame are fetched while the current one is processed
  int nxt_n = 0;
  double nxt_d[6] = {0, 0, 0, 0, 0, 0};
  // a dependent global load per frame (count -> which lanes load a detection; frame number -> time stamp) would
  // sit on the recurrence's critical path: both tables are staged once per launch
  const bool staged = nf <= kMaxF;
  if (staged) {
    for (int i = lane; i < nf; i += 32) { sh.cnt[i] = det_count[(size_t)v * F + i]; sh.fno[i] = frame_no[(size_t)v * F + i]; }
    __syncwarp();
  }
  auto fetch = [&](int f) {
    nxt_n = 0;
    if (f < nf) {
      nxt_n = min(staged ? sh.cnt[f] : det_count[(size_t)v * F + f], max_det);
      if (lane < nxt_n) {
        const double* fd = dets + (((size_t)v * F + f) * max_det + lane) * 6;
#pragma unroll
        for (int j = 0; j < 6; ++j) nxt_d[j] = fd[j];
      }
    }
  };
  fetch(0);
  TRK_TICK(0);                                          // launch prologue: state -> shared memory

  for (int f = 0; f < nf; ++f) {
    const int nd0 = nxt_n;
    double d6[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) d6[j] = nxt_d[j];
    fetch(f + 1);
    if (nd0 <= 0) continue;                         // track.py:180-181
    frame_count += 1;
    // ---- detections above det_thresh, order kept ---------------------------------------
    int nd;
    {
      const bool keep = lane < nd0 && d6[4] > prm.det_thresh;
      const unsigned m = __ballot_sync(full, keep);
      nd = __popc(m);
      if (keep) {
        const int pos = __popc(m & lanemask_lt(lane));
#pragma unroll
        for (int j = 0; j < 6; ++j) sh.dets[pos][j] = d6[j];
      }
    }
    TRK_TICK(1);
    // ---- predict every track, drop the ones whose box went NaN (list order kept) ---------
    int nt = vid.n_tracks;
    {
      int kept = 0;
      for (int base = 0; base < nt; base += 32) {
        const int t = base + lane;
        const bool valid = t < nt;
        const int slot = valid ? vid.order[t] : 0;
        double b[4] = {0, 0, 0, 0};
        bool bad = false;
        if (valid) {
          trk_predict(T(slot), b);
          bad = isnan(b[0]) || isnan(b[1]) || isnan(b[2]) || isnan(b[3]);
        }
        const unsigned m = __ballot_sync(full, valid && !bad);
        if (valid && bad) vid.used[slot] = 0;
        if (valid && !bad) {
          const int p = kept + __popc(m & lanemask_lt(lane));
          vid.order[p] = slot;                      // p <= t: never overwrites an unread entry
#pragma unroll
          for (int j = 0; j < 4; ++j) sh.tbox[p][j] = b[j];
        }
        kept += __popc(m);
        __syncwarp();
      }
      nt = kept;
    }
    __syncwarp();
    TRK_TICK(2);
    // ---- first association round ----------------------------------------------------
    // overlaps above the threshold are counted per detection (un_d) and per track (un_t) while the matrix is
    // written -- the index lists these arrays hold later are not built yet; pr[d] = a track d overlaps
    const int n_pairs_all = nd * nt;
    for (int t = lane; t < nt; t += 32) sh.un_t[t] = 0;
    if (lane < kMaxD) { sh.un_d[lane] = 0; sh.pr[lane] = -1; }
    __syncwarp();
    for (int i = lane; i < n_pairs_all; i += 32) {
      const int d = i / nt, t = i - d * nt;
      const double v = iou_of(sh.dets[d], sh.tbox[t]);
      sh.iou[d * ld + t] = v;
      if (v > thr) { atomicAdd(&sh.un_d[d], 1); atomicAdd(&sh.un_t[t], 1); sh.pr[d] = t; }
    }
    __syncwarp();
    TRK_TICK(12);
    int rmax = lane < nd ? sh.un_d[lane] : 0, cmax = 0;   // hits per detection / per track
    const int hit_t = lane < nd ? sh.pr[lane] : -1;       // the track, when there is exactly one
    for (int t = lane; t < nt; t += 32) cmax = max(cmax, sh.un_t[t]);
    const int my_hits = rmax;
    rmax = __reduce_max_sync(full, rmax);
    cmax = __reduce_max_sync(full, cmax);
    TRK_TICK(13);
    int n_pairs = 0;
    if (nt > 0 && nd > 0) {
      if (rmax == 1 && cmax == 1) {                 // every overlap is unambiguous
        const bool has = my_hits == 1;
        const unsigned m = __ballot_sync(full, has);
        n_pairs = __popc(m);
        if (has) {
          const int pos = __popc(m & lanemask_lt(lane));
          sh.pair_d[pos] = lane; sh.pair_t[pos] = hit_t;
        }
      } else {
        const double kPi = 3.141592653589793;
        auto pair_cost = [&](int d, int t) -> double {
          const Trk& tk = T(vid.order[t]);
          const double* prev = k_previous(tk, prm.delta_t);
          double pb[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
          if (prev) for (int j = 0; j < 5; ++j) pb[j] = prev[j];
          double cxd = (sh.dets[d][0] + sh.dets[d][2]) / 2.0, cyd = (sh.dets[d][1] + sh.dets[d][3]) / 2.0;
          double cxp = (pb[0] + pb[2]) / 2.0, cyp = (pb[1] + pb[3]) / 2.0;
          double ddx = cxd - cxp, ddy = cyd - cyp;
          double norm = sqrt(ddx * ddx + ddy * ddy) + 1e-6;
          ddx = ddx / norm; ddy = ddy / norm;
          double vy = tk.has_vel ? tk.vel[0] : 0.0, vx = tk.has_vel ? tk.vel[1] : 0.0;
          double c = vx * ddx + vy * ddy;
          c = fmin(fmax(c, -1.0), 1.0);
          double ang = (kPi / 2.0 - fabs(acos(c))) / kPi;
          double valid = (pb[4] >= 0) ? 1.0 : 0.0;
          double mult = prm.vdc_cls ? sh.dets[d][5] : sh.dets[d][4];
          double angle_cost = ((valid * ang) * prm.inertia) * mult;
          return -(sh.iou[d * ld + t] + angle_cost);
        };
        for (int i = lane; i < n_pairs_all; i += 32) {
          const int d = i / nt, t = i - d * nt;
          sh.cost[d * ld + t] = pair_cost(d, t);
        }
        __syncwarp();
        TRK_TICK(14);
        int np = 0;
        if (max(nd, nt) <= 31) {
          np = assign_min_cost_warp(sh.cost, ld, nd, nt, sh.pair_d, sh.pair_t, sh.u, lane);
        } else {
          if (lane == 0) np = assign_min_cost(sh.cost, ld, nd, nt, sh.pair_d, sh.pair_t);
          np = __shfl_sync(full, np, 0);
        }
        if (np < 0) { np = 0; if (lane == 0) vid.status = VBT_EINVAL; }
        n_pairs = np;
        TRK_TICK(15);
      }
    }
    __syncwarp();
    TRK_TICK(3);
    // ---- unmatched lists (ascending), then pairs below the IoU threshold are undone and
    //      appended to both lists in pair order (upstream behaviour) ------------------------
    int n_un_d = 0, n_un_t = 0;
    {
      for (int t = lane; t < nt; t += 32) sh.flag_t[t] = 0;
      if (lane < kMaxD) sh.flag_d[lane] = 0;
      __syncwarp();
      int pd = 0, pt = 0;
      bool rej = false;
      if (lane < n_pairs) {
        pd = sh.pair_d[lane]; pt = sh.pair_t[lane];
        sh.flag_d[pd] = 1; sh.flag_t[pt] = 1;
        rej = sh.iou[pd * ld + pt] < thr;
      }
      __syncwarp();
      {
        const bool un = lane < nd && !sh.flag_d[lane];
        const unsigned m = __ballot_sync(full, un);
        if (un) sh.un_d[__popc(m & lanemask_lt(lane))] = lane;
        n_un_d = __popc(m);
      }
      for (int base = 0; base < nt; base += 32) {
        const int t = base + lane;
        const bool un = t < nt && !sh.flag_t[t];
        const unsigned m = __ballot_sync(full, un);
        if (un) sh.un_t[n_un_t + __popc(m & lanemask_lt(lane))] = t;
        n_un_t += __popc(m);
      }
      const unsigned mr = __ballot_sync(full, rej);
      const unsigned mk = __ballot_sync(full, lane < n_pairs && !rej);
      __syncwarp();
      if (rej) {
        const int pos = __popc(mr & lanemask_lt(lane));
        sh.un_d[n_un_d + pos] = pd; sh.un_t[n_un_t + pos] = pt;
      } else if (lane < n_pairs) {
        const int pos = __popc(mk & lanemask_lt(lane));
        sh.pair_d[pos] = pd; sh.pair_t[pos] = pt;
      }
      n_un_d += __popc(mr); n_un_t += __popc(mr);
      n_pairs = __popc(mk);
    }
    __syncwarp();
    TRK_TICK(4);
    if (lane < n_pairs) trk_update(T(vid.order[sh.pair_t[lane]]), sh.dets[sh.pair_d[lane]], prm.delta_t);
    __syncwarp();
    TRK_TICK(5);
    // ---- second round: unmatched detections vs last observations (DIoU) --------------
    if (n_un_d > 0 && n_un_t > 0) {
      const int a_n = n_un_d, b_n = n_un_t;
      double mx = -INFINITY;
      bool any_nan = false;
      for (int i = lane; i < a_n * b_n; i += 32) {
        const int a = i / b_n, b = i - a * b_n;
        const Trk& tk = T(vid.order[sh.un_t[b]]);
        const double q = diou_of(sh.dets[sh.un_d[a]], tk.last_obs);   // [-1]*5 when unseen
        sh.cost[a * ld + b] = q;
        sh.iou[a * ld + b] = -q;
        if (isnan(q)) any_nan = true; else mx = fmax(mx, q);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(full, mx, o));
      any_nan = __any_sync(full, any_nan);
      __syncwarp();
      if (!any_nan && mx > thr) {
        int np = 0;
        if (max(a_n, b_n) <= 31) {
          np = assign_min_cost_warp(sh.iou, ld, a_n, b_n, sh.pr, sh.pc, sh.u, lane);
        } else {
          if (lane == 0) np = assign_min_cost(sh.iou, ld, a_n, b_n, sh.pr, sh.pc);
          np = __shfl_sync(full, np, 0);
        }
        if (np < 0) { np = 0; if (lane == 0) vid.status = VBT_EINVAL; }
        __syncwarp();
        // accepted pairs in assignment order; the rest stays unmatched, as sorted sets
        for (int t = lane; t < nt; t += 32) sh.flag_t[t] = 0;
        if (lane < kMaxD) sh.flag_d[lane] = 0;
        __syncwarp();
        bool acc = false;
        int ad = 0, at = 0;
        if (lane < np) {
          const int pa = sh.pr[lane], pb = sh.pc[lane];
          acc = !(sh.cost[pa * ld + pb] < thr);
          ad = sh.un_d[pa]; at = sh.un_t[pb];
        }
        const unsigned ma = __ballot_sync(full, acc);
        if (acc) {
          const int pos = __popc(ma & lanemask_lt(lane));
          sh.pair_d[pos] = ad; sh.pair_t[pos] = at;
        }
        n_pairs = __popc(ma);
        // keep flags, indexed by detection / track: unmatched entries that were not accepted
        __syncwarp();
        if (lane < a_n) sh.flag_d[sh.un_d[lane]] = 1;
        for (int b = lane; b < b_n; b += 32) sh.flag_t[sh.un_t[b]] = 1;
        __syncwarp();
        if (acc) { sh.flag_d[ad] = 0; sh.flag_t[at] = 0; }
        __syncwarp();
        {
          const bool un = lane < nd && sh.flag_d[lane];
          const unsigned m = __ballot_sync(full, un);
          __syncwarp();
          if (un) sh.un_d[__popc(m & lanemask_lt(lane))] = lane;
          n_un_d = __popc(m);
        }
        n_un_t = 0;
        for (int base = 0; base < nt; base += 32) {
          const int t = base + lane;
          const bool un = t < nt && sh.flag_t[t];
          const unsigned m = __ballot_sync(full, un);
          if (un) sh.un_t[n_un_t + __popc(m & lanemask_lt(lane))] = t;   // position <= t
          n_un_t += __popc(m);
        }
        __syncwarp();
        if (lane < n_pairs) trk_update(T(vid.order[sh.pair_t[lane]]), sh.dets[sh.pair_d[lane]], prm.delta_t);
        __syncwarp();
      }
    }
    TRK_TICK(6);
    for (int i = lane; i < n_un_t; i += 32)
      trk_update(T(vid.order[sh.un_t[i]]), nullptr, prm.delta_t);
    __syncwarp();
    TRK_TICK(7);
    // ---- births (slot allocation is order dependent: one lane) -----------------------------
    if (n_un_d > 0) {
      if (lane == 0) {
        for (int i = 0; i < n_un_d; ++i) {
          int slot = -1;
          for (int s = 0; s < max_tracks; ++s) if (!vid.used[s]) { slot = s; break; }
          if (slot < 0 || nt >= max_tracks) { vid.status = VBT_ECAPACITY; break; }
          vid.used[slot] = 1;
          trk_init(T(slot), sh.dets[sh.un_d[i]], vid.next_id++);
          vid.order[nt++] = slot;
        }
      }
      nt = __shfl_sync(full, nt, 0);
      __syncwarp();
    }
    TRK_TICK(8);
    // ---- output rows in reverse list order, then deaths -------------------------------------
    {
      const double time = (double)(staged ? sh.fno[f] : frame_no[(size_t)v * F + f]) / vfps;     // track.py:169
      const bool last_frame = (f == nf - 1);
      int n_out = 0;
      for (int top = nt - 1; top >= 0; top -= 32) {
        const int t = top - lane;
        bool emit = false;
        double box[4] = {0, 0, 0, 0}, dxy[2] = {0, 0}, cc[2] = {0, 0};
        int id = 0;
        if (t >= 0) {
          const Trk& tk = T(vid.order[t]);
          emit = tk.tsu < 1 && (tk.hit_streak >= prm.min_hits || frame_count <= prm.min_hits);
          if (emit) {
            const double sum = tk.last_obs[0] + tk.last_obs[1] + tk.last_obs[2] + tk.last_obs[3] + tk.last_obs[4];
            if (!tk.has_last || sum < 0) x_to_box(tk.x, box);
            else for (int j = 0; j < 4; ++j) box[j] = tk.last_obs[j];
            dxy[0] = tk.x[4]; dxy[1] = tk.x[5]; cc[0] = tk.cls; cc[1] = tk.conf; id = tk.id + 1;
          }
        }
        const unsigned m = __ballot_sync(full, emit);
        if (emit) {
          const int pos = __popc(m & lanemask_lt(lane));
          if (rc + pos >= row_cap) {
            vid.status = VBT_ECAPACITY;
          } else {
            double* r = vrows + (size_t)(rc + pos) * VBT_ROW_COLS;
            r[0] = (double)id;
            r[1] = time;
            r[2] = (box[0] + box[2]) / 2;          // odt.py:43-50
            r[3] = (box[1] + box[3]) / 2;
            r[4] = dxy[0];                         // track.py:199
            r[5] = dxy[1];
            r[6] = fabs(box[3] - box[1]);          // odt.py:32-40
            r[7] = fabs(box[2] - box[0]);          // odt.py:22-29
            if (row_details) {                     // what track.py:190 unpacks for the overlay
              double* d = row_details + ((size_t)v * row_cap + rc + pos) * VBT_ROW_DETAIL_COLS;
              d[0] = box[0]; d[1] = box[1]; d[2] = box[2]; d[3] = box[3]; d[4] = cc[1];
            }
          }
          if (last_out && last_frame && n_out + pos < max_det) {
            double* o = last_out + ((size_t)v * max_det + n_out + pos) * 9;
            o[0] = box[0]; o[1] = box[1]; o[2] = box[2]; o[3] = box[3];
            o[4] = (double)id; o[5] = cc[0]; o[6] = cc[1]; o[7] = dxy[0]; o[8] = dxy[1];
          }
        }
        const int c = __popc(m);
        rc = min(rc + c, row_cap);
        n_out += c;
      }
      if (last_out_count && last_frame && lane == 0) last_out_count[v] = min(n_out, max_det);
      int kept = 0;
      for (int base = 0; base < nt; base += 32) {
        const int t = base + lane;
        const bool valid = t < nt;
        const int slot = valid ? vid.order[t] : 0;
        const bool dead = valid && T(slot).tsu > prm.max_age;
        const unsigned m = __ballot_sync(full, valid && !dead);
        if (dead) vid.used[slot] = 0;
        if (valid && !dead) vid.order[kept + __popc(m & lanemask_lt(lane))] = slot;
        kept += __popc(m);
        __syncwarp();
      }
      if (lane == 0) vid.n_tracks = kept;
    }
    __syncwarp();
    TRK_TICK(9);
    if (DBG && dbg) g_trk_dbg[11] += 1;
  }
  __threadfence();                                  // rows before their count: the velocity kernel may run beside the next launch
  if (lane == 0) { vid.frame_count = frame_count; row_count[v] = rc; }
  {                                                 // write the staged state back
    __syncwarp();
    constexpr int W = sizeof(Trk) / 8;
    for (int s = 0; s < kCache && s < max_tracks; ++s) {
      if (!vid.used[s]) continue;
      const uint64_t* a = reinterpret_cast<const uint64_t*>(sh.cache + s);
      uint64_t* b = reinterpret_cast<uint64_t*>(gtrk + s);
      for (int i = lane; i < W; i += 32) b[i] = a[i];
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sh.vid);
    uint32_t* dst = reinterpret_cast<uint32_t*>(videos + v);
    for (int i = lane; i < (int)(sizeof(Video) / 4); i += 32) dst[i] = src[i];
  }
  // a call whose last frame was empty still reports "nothing emitted"
  if (lane == 0 && last_out_count && nf > 0 && det_count[(size_t)v * F + nf - 1] <= 0)
    last_out_count[v] = 0;
}

__global__ void tracker_reset_kernel(Video* videos, int V) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  Video& vid = videos[v];
  vid.frame_count = vid.next_id = vid.n_tracks = vid.status = 0;
  for (int i = 0; i < kMaxT; ++i) { vid.order[i] = 0; vid.used[i] = 0; }
}

__global__ void tracker_peek_kernel(const Video* videos, const Trk* tracks, int v, int max_tracks,
                                    double* out, int32_t* out_n) {
  if (blockIdx.x || threadIdx.x) return;
  const Video& vid = videos[v];
  for (int t = 0; t < vid.n_tracks; ++t) {
    const Trk& tk = tracks[(size_t)v * max_tracks + vid.order[t]];
    for (int j = 0; j < NX; ++j) out[t * 9 + j] = tk.x[j];
    out[t * 9 + 7] = (double)tk.id;
    out[t * 9 + 8] = (double)tk.tsu;
  }
  *out_n = vid.n_tracks;
}

__global__ void tracker_status_kernel(const Video* videos, int V, int32_t* out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) out[v] = videos[v].status;
}

}  // namespace

struct vbt_tracker {
  int V, max_tracks;
  Params prm;
  Video* videos;
  Trk* tracks;
  double* peek;      // [kMaxT*9 + 1]
  double* row_details = nullptr;   // optional f64 [V,row_cap,5] next to the rows (vbt_tracker_row_details)
  int32_t* scratch;  // [max(V,1)]
};

extern "C" {

int vbt_tracker_create(int V, int max_tracks, const vbt_tracker_params* p, vbt_tracker** out) {
  VBT_REQUIRE(out && p && V > 0, "vbt_tracker_create: bad arguments");
  VBT_REQUIRE(max_tracks > 0 && max_tracks <= kMaxT, "vbt_tracker_create: max_tracks must be 1..%d",
              kMaxT);
  VBT_REQUIRE(p->delta_t >= 1 && p->delta_t <= 3, "vbt_tracker_create: delta_t must be 1..3");
  if (int rc = vbt::ensure_device()) return rc;
  vbt_tracker* t = new vbt_tracker();
  t->V = V; t->max_tracks = max_tracks;
  t->prm.det_thresh = p->det_thresh; t->prm.iou_threshold = p->iou_threshold;
  t->prm.inertia = p->inertia; t->prm.max_age = p->max_age; t->prm.min_hits = p->min_hits;
  t->prm.delta_t = p->delta_t; t->prm.vdc_cls = p->vdc_uses_class_column;
  VBT_CHECK_CUDA(cudaMalloc(&t->videos, sizeof(Video) * (size_t)V));
  VBT_CHECK_CUDA(cudaMalloc(&t->tracks, sizeof(Trk) * (size_t)V * max_tracks));
  VBT_CHECK_CUDA(cudaFuncSetAttribute(tracker_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)shared_bytes(kMaxT)));
  VBT_CHECK_CUDA(cudaFuncSetAttribute(tracker_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)shared_bytes(kMaxT)));
  VBT_CHECK_CUDA(cudaMalloc(&t->peek, sizeof(double) * (kMaxT * 9)));
  VBT_CHECK_CUDA(cudaMalloc(&t->scratch, sizeof(int32_t) * (size_t)(V + 1)));
  *out = t;
  return vbt_tracker_reset(t, nullptr);
}

void vbt_tracker_destroy(vbt_tracker* t) {
  if (!t) return;
  cudaFree(t->videos); cudaFree(t->tracks); cudaFree(t->peek); cudaFree(t->scratch);
  delete t;
}

int vbt_tracker_reset(vbt_tracker* t, void* stream) {
  VBT_REQUIRE(t, "vbt_tracker_reset: null handle");
  tracker_reset_kernel<<<vbt::ceil_div(t->V, 64), 64, 0, (cudaStream_t)stream>>>(t->videos, t->V);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

int vbt_tracker_update(vbt_tracker* t, const double* dev_dets, const int32_t* dev_det_count,
                       const int32_t* dev_frame_no, const double* dev_fps,
                       const int32_t* dev_n_frames, int F, int max_det, double* dev_rows,
                       int32_t* dev_row_count, int row_cap, double* dev_last_out,
                       int32_t* dev_last_out_count, void* stream) {
  VBT_REQUIRE(t && dev_dets && dev_det_count && dev_frame_no && dev_fps && dev_n_frames &&
                  dev_rows && dev_row_count, "vbt_tracker_update: null pointer");
  VBT_REQUIRE(F > 0 && max_det > 0 && max_det <= kMaxD && row_cap > 0,
              "vbt_tracker_update: F=%d max_det=%d (<=%d) row_cap=%d", F, max_det, kMaxD, row_cap);
  static const bool dbg = [] {
    const char* e = getenv("VBT_TRK_DBG");
    const int on = e && e[0] == '1';
    if (on) cudaMemcpyToSymbol(g_trk_dbg_on, &on, sizeof(int));
    return on != 0;
  }();
  auto kern = dbg ? tracker_update_kernel<true> : tracker_update_kernel<false>;   // phase counters: their own build
  kern<<<t->V, 32, shared_bytes(t->max_tracks), (cudaStream_t)stream>>>(
      t->videos, t->tracks, t->prm, dev_dets, dev_det_count, dev_frame_no, dev_fps, dev_n_frames,
      F, max_det, t->max_tracks, dev_rows, dev_row_count, row_cap, dev_last_out, dev_last_out_count,
      t->row_details);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

int vbt_tracker_row_details(vbt_tracker* t, double* dev_row_details) {
  VBT_REQUIRE(t, "vbt_tracker_row_details: null handle");
  t->row_details = dev_row_details;
  return VBT_OK;
}

int vbt_tracker_status(vbt_tracker* t, int32_t* host_status, void* stream) {
  VBT_REQUIRE(t && host_status, "vbt_tracker_status: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  tracker_status_kernel<<<vbt::ceil_div(t->V, 64), 64, 0, st>>>(t->videos, t->V, t->scratch);
  VBT_LAUNCHED(1);
  VBT_CHECK_CUDA(cudaMemcpyAsync(host_status, t->scratch, sizeof(int32_t) * (size_t)t->V,
                                 cudaMemcpyDeviceToHost, st));
  VBT_CHECK_CUDA(cudaStreamSynchronize(st));
  if (getenv("VBT_TRK_DBG")) {
    long long h[16];
    cudaMemcpyFromSymbol(h, g_trk_dbg, sizeof(h));
    static const char* nm[10] = {"prologue", "compact dets", "predict", "assoc 1", "unmatched lists", "update matched", "OCR round",
                                 "update unmatched", "births", "output+deaths"};
    const double fr = h[11] > 0 ? (double)h[11] : 1.0;
    fprintf(stderr, "[tracker video 0: %lld frames stepped]\n", h[11]);
    for (int i = 0; i < 10; ++i) fprintf(stderr, "   %-18s %10.0f cycles/frame\n", nm[i], (double)h[i] / fr);
    static const char* nm2[4] = {"  assoc 1: IoU", "  assoc 1: counts", "  assoc 1: costs", "  assoc 1: solver"};
    for (int i = 0; i < 4; ++i) fprintf(stderr, "   %-18s %10.0f cycles/frame\n", nm2[i], (double)h[12 + i] / fr);
  }
  for (int v = 0; v < t->V; ++v)
    if (host_status[v] != 0) {
      if (host_status[v] == VBT_EINVAL) {
        vbt::set_error("tracker video %d: association cost matrix held NaN/inf (degenerate boxes)", v);
        return VBT_EINVAL;
      }
      vbt::set_error("tracker video %d overflowed (more than %d live tracks, %d detections per "
                     "frame, or the row table)", v, t->max_tracks, kMaxD);
      return VBT_ECAPACITY;
    }
  return VBT_OK;
}

int vbt_tracker_peek(vbt_tracker* t, int v, double* host_tracks, void* stream) {
  VBT_REQUIRE(t && host_tracks && v >= 0 && v < t->V, "vbt_tracker_peek: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  tracker_peek_kernel<<<1, 1, 0, st>>>(t->videos, t->tracks, v, t->max_tracks, t->peek,
                                       t->scratch + t->V);
  VBT_LAUNCHED(1);
  int32_t n = 0;
  VBT_CHECK_CUDA(cudaMemcpyAsync(&n, t->scratch + t->V, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  VBT_CHECK_CUDA(cudaStreamSynchronize(st));
  if (n > 0)
    VBT_CHECK_CUDA(cudaMemcpy(host_tracks, t->peek, sizeof(double) * 9 * (size_t)n,
                              cudaMemcpyDeviceToHost));
  return n;
}

}  // extern "C"
