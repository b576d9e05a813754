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

struct Trk {
  double x[NX];
  double P[NX * NX];
  double fx[NX];            // state frozen at the first missed frame
  double fP[NX * NX];
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

// The oracle multiplies dense 7x7 matrices k-ascending.  F = I + shift and H = [I4 0] are
// 0/1 matrices, so every dense inner product below collapses to the few terms written
// out here IN THE SAME ORDER; the skipped terms are exact zeros (0 * finite), which do
// not change an IEEE sum.  Everything is unrolled over constant indices so x and P live
// in registers.
__device__ __forceinline__ void kf_predict(double* x, double* P) {
  // x = F x
  x[0] = x[0] + x[4]; x[1] = x[1] + x[5]; x[2] = x[2] + x[6];
  // t = F P : rows 0..2 pick up rows 4..6
  double t[NX * NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) {
#pragma unroll
    for (int i = 0; i < NX; ++i)
      t[i * NX + j] = (i < 3) ? P[i * NX + j] + P[(i + 4) * NX + j] : P[i * NX + j];
  }
  // P = t F^T + Q : columns 0..2 pick up columns 4..6
  const double q[NX] = {1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.0001};
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double v = (j < 3) ? t[i * NX + j] + t[i * NX + j + 4] : t[i * NX + j];
      P[i * NX + j] = (i == j) ? v + q[i] : v;
    }
  }
}

__device__ __forceinline__ void kf_correct(double* x, double* P, const double* z) {
  const double R[NZ] = {1.0, 1.0, 10.0, 10.0};
  double y[NZ], si[NZ], K[NX * NZ], ikh[NX * NZ];
#pragma unroll
  for (int j = 0; j < NZ; ++j) {
    y[j] = z[j] - x[j];                                   // z - Hx
    si[j] = 1.0 / (P[j * NX + j] + R[j]);                 // S = HPH' + R is diagonal
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int j = 0; j < NZ; ++j) K[i * NZ + j] = P[i * NX + j] * si[j];   // K = PH' inv(S)
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {                          // x += K y
    double acc = K[i * NZ + 0] * y[0];
    acc = acc + K[i * NZ + 1] * y[1];
    acc = acc + K[i * NZ + 2] * y[2];
    acc = acc + K[i * NZ + 3] * y[3];
    x[i] = x[i] + acc;
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {                          // (I - KH)[:, :4]; columns 4..6 are I
#pragma unroll
    for (int k = 0; k < NZ; ++k) ikh[i * NZ + k] = ((i == k) ? 1.0 : 0.0) - K[i * NZ + k];
  }
  double t1[NX * NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {                          // t1 = (I-KH) P
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double acc = ikh[i * NZ + 0] * P[0 * NX + j];
      acc = acc + ikh[i * NZ + 1] * P[1 * NX + j];
      acc = acc + ikh[i * NZ + 2] * P[2 * NX + j];
      acc = acc + ikh[i * NZ + 3] * P[3 * NX + j];
      if (i >= NZ) acc = acc + P[i * NX + j];
      t1[i * NX + j] = acc;
    }
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {                          // P = t1 (I-KH)' + (K R) K'
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double acc = t1[i * NX + 0] * ikh[j * NZ + 0];
      acc = acc + t1[i * NX + 1] * ikh[j * NZ + 1];
      acc = acc + t1[i * NX + 2] * ikh[j * NZ + 2];
      acc = acc + t1[i * NX + 3] * ikh[j * NZ + 3];
      if (j >= NZ) acc = acc + t1[i * NX + j];
      double krk = (K[i * NZ + 0] * R[0]) * K[j * NZ + 0];
      krk = krk + (K[i * NZ + 1] * R[1]) * K[j * NZ + 1];
      krk = krk + (K[i * NZ + 2] * R[2]) * K[j * NZ + 2];
      krk = krk + (K[i * NZ + 3] * R[3]) * K[j * NZ + 3];
      P[i * NX + j] = acc + krk;
    }
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
      for (int i = 0; i < NX * NX; ++i) t.fP[i] = t.P[i];
      t.has_frozen = 1;
    }
    t.observed = 0;
    t.missed += 1;
    return;
  }
  double x[NX], P[NX * NX];
  if (!t.observed && t.has_frozen) {        // observation-centric re-update
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = t.fx[i];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) P[i] = t.fP[i];
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
    for (int i = 0; i < NX * NX; ++i) P[i] = t.P[i];
    for (int i = 0; i < NZ; ++i) t.prev_z[i] = z[i];
  }
  t.observed = 1;
  t.missed = 0;
  kf_correct(x, P, z);
#pragma unroll
  for (int i = 0; i < NX; ++i) t.x[i] = x[i];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) t.P[i] = P[i];
}

__device__ void trk_predict(Trk& t, double* box) {
  if (t.x[6] + t.x[2] <= 0) t.x[6] *= 0.0;
  double x[NX], P[NX * NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = t.x[i];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) P[i] = t.P[i];
  kf_predict(x, P);
#pragma unroll
  for (int i = 0; i < NX; ++i) t.x[i] = x[i];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) t.P[i] = P[i];
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
  const double p0[NX] = {10.0, 10.0, 10.0, 10.0, 10000.0, 10000.0, 10000.0};
  for (int i = 0; i < NX * NX; ++i) t.P[i] = 0.0;
  for (int i = 0; i < NX; ++i) t.P[i * NX + i] = p0[i];
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
                                    double* u_s /* [32] shared */, int lane) {
  if (n == 0 || m == 0) return 0;
  const bool tr = n > m;
  const int N = tr ? m : n, M = tr ? n : m;
  const unsigned full = 0xffffffffu;
  double v = 0.0, minv = INFINITY;
  int p = 0, way = 0;
  bool used = false;
  u_s[lane] = 0.0;
  __syncwarp();
  const bool col = lane >= 1 && lane <= M;
  for (int i = 1; i <= N; ++i) {
    if (lane == 0) p = i;
    int j0 = 0;
    minv = INFINITY;
    used = false;
    while (true) {
      if (lane == j0) used = true;
      const int i0 = __shfl_sync(full, p, j0);
      double key = INFINITY;
      if (col && !used) {
        const double cij = tr ? cost[(lane - 1) * ld + (i0 - 1)] : cost[(i0 - 1) * ld + (lane - 1)];
        const double cur = cij - u_s[i0] - v;
        if (cur < minv) { minv = cur; way = j0; }
        key = minv;
      }
      int j1 = (key < INFINITY) ? lane : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double k2 = __shfl_xor_sync(full, key, o);
        const int jj = __shfl_xor_sync(full, j1, o);
        // candidates carry j1 == 0 when their key is not finite-below-inf
        const bool take = (jj != 0) && (j1 == 0 || k2 < key || (k2 == key && jj < j1));
        if (take) { key = k2; j1 = jj; }
      }
      if (j1 == 0) return -1;
      const double delta = key;
      __syncwarp();
      if (used) {                       // lanes 0..M that are on the alternating tree
        u_s[p] += delta;                // distinct rows: p is injective over used columns
        v -= delta;
      } else if (col) {
        minv -= delta;
      }
      __syncwarp();
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
  int cnt = 0;
  if (!tr) {
    // row i-1 (i = 1..N) is matched to the column whose p == i
    for (int i = 1; i <= N; ++i) {
      const unsigned mask = __ballot_sync(full, col && p == i);
      if (mask) {
        if (lane == 0) { pair_r[cnt] = i - 1; pair_c[cnt] = __ffs(mask) - 2; }
        ++cnt;
      }
    }
  } else {
    for (int j = 1; j <= M; ++j) {
      const int pj = __shfl_sync(full, p, j);
      if (pj != 0) {
        if (lane == 0) { pair_r[cnt] = j - 1; pair_c[cnt] = pj - 1; }
        ++cnt;
      }
    }
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
  unsigned char* nanflag;       // [ld]
  double u[32];                 // row potentials of the warp-parallel assignment
  Trk* cache;                   // [kCache] low slots of this video's track table
  Video* vid;                   // this video's list state
  int ld;
  int n_pairs, n_un_d, n_un_t, nd, nt, go;
};

__host__ __device__ inline size_t shared_bytes(int ld) {
  return sizeof(double) * (kMaxD * 6 + (size_t)ld * 4 + 2 * (size_t)kMaxD * ld) +
         sizeof(int) * (3 * kMaxD + (size_t)ld) + ((size_t)ld + 15) / 16 * 16 + 16 +
         sizeof(Trk) * kCache + sizeof(Video);
}

__global__ void __launch_bounds__(32) tracker_update_kernel(
    Video* videos, Trk* tracks, Params prm, const double* dets, const int32_t* det_count,
    const int32_t* frame_no, const double* fps, const int32_t* n_frames, int F, int max_det, int max_tracks,
    double* rows, int32_t* row_count, int row_cap, double* last_out, int32_t* last_out_count) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ Shared sh;
  const int v = blockIdx.x, lane = threadIdx.x;
  if (lane == 0) {
    unsigned char* p = dyn_smem;
    sh.ld = max_tracks;
    sh.dets = reinterpret_cast<double(*)[6]>(p); p += sizeof(double) * kMaxD * 6;
    sh.tbox = reinterpret_cast<double(*)[4]>(p); p += sizeof(double) * 4 * (size_t)max_tracks;
    sh.iou = reinterpret_cast<double*>(p); p += sizeof(double) * (size_t)kMaxD * max_tracks;
    sh.cost = reinterpret_cast<double*>(p); p += sizeof(double) * (size_t)kMaxD * max_tracks;
    sh.pair_d = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.pair_t = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.un_d = reinterpret_cast<int*>(p); p += sizeof(int) * kMaxD;
    sh.un_t = reinterpret_cast<int*>(p); p += sizeof(int) * (size_t)max_tracks;
    sh.nanflag = p; p += ((size_t)max_tracks + 15) / 16 * 16;
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

  for (int f = 0; f < nf; ++f) {
    const int nd0 = det_count[(size_t)v * F + f];
    if (nd0 <= 0) continue;                         // track.py:180-181
    const double* fd = dets + ((size_t)v * F + f) * max_det * 6;
    if (lane == 0) {
      vid.frame_count += 1;
      int nd = 0;
      for (int i = 0; i < nd0 && i < max_det; ++i) {
        if (fd[i * 6 + 4] > prm.det_thresh) {
          if (nd >= kMaxD) { vid.status = VBT_ECAPACITY; break; }
          for (int j = 0; j < 6; ++j) sh.dets[nd][j] = fd[i * 6 + j];
          ++nd;
        }
      }
      sh.nd = nd;
      sh.nt = vid.n_tracks;
    }
    __syncwarp();
    // ---- predict every track (lane-parallel) ---------------------------------------
    for (int t = lane; t < sh.nt; t += 32) {
      double b[4];
      trk_predict(T(vid.order[t]), b);
      for (int j = 0; j < 4; ++j) sh.tbox[t][j] = b[j];
      sh.nanflag[t] = (isnan(b[0]) || isnan(b[1]) || isnan(b[2]) || isnan(b[3])) ? 1 : 0;
    }
    __syncwarp();
    if (lane == 0) {                                // drop NaN tracks, keep list order
      int k = 0;
      for (int t = 0; t < sh.nt; ++t) {
        int slot = vid.order[t];
        if (sh.nanflag[t]) { vid.used[slot] = 0; continue; }
        vid.order[k] = slot;
        for (int j = 0; j < 4; ++j) sh.tbox[k][j] = sh.tbox[t][j];
        ++k;
      }
      sh.nt = vid.n_tracks = k;
    }
    __syncwarp();
    const int nd = sh.nd, nt = sh.nt;
    // ---- first association round ----------------------------------------------------
    for (int i = lane; i < nd * nt; i += 32) {
      int d = i / nt, t = i % nt;
      sh.iou[d * ld + t] = iou_of(sh.dets[d], sh.tbox[t]);
    }
    __syncwarp();
    int rmax = 0, cmax = 0;                         // max hits per detection / per track
    for (int d = lane; d < nd; d += 32) {
      int c = 0;
      for (int t = 0; t < nt; ++t) c += sh.iou[d * ld + t] > prm.iou_threshold;
      rmax = max(rmax, c);
    }
    for (int t = lane; t < nt; t += 32) {
      int c = 0;
      for (int d = 0; d < nd; ++d) c += sh.iou[d * ld + t] > prm.iou_threshold;
      cmax = max(cmax, c);
    }
    rmax = __reduce_max_sync(0xffffffffu, rmax);
    cmax = __reduce_max_sync(0xffffffffu, cmax);
    if (lane == 0) {
      sh.n_pairs = 0;
      sh.go = 0;                                    // 1: Hungarian needed
      if (nt > 0 && nd > 0) {
        if (rmax == 1 && cmax == 1) {
          for (int d = 0; d < nd; ++d)
            for (int t = 0; t < nt; ++t)
              if (sh.iou[d * ld + t] > prm.iou_threshold) {
                sh.pair_d[sh.n_pairs] = d; sh.pair_t[sh.n_pairs] = t; ++sh.n_pairs;
              }
        } else {
          sh.go = 1;
        }
      }
    }
    __syncwarp();
    if (sh.go) {
      const double kPi = 3.141592653589793;
      for (int i = lane; i < nd * nt; i += 32) {
        int d = i / nt, t = i % nt;
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
        sh.cost[d * ld + t] = -(sh.iou[d * ld + t] + angle_cost);
      }
      __syncwarp();
      int np = 0;
      if (max(nd, nt) <= 31) {
        np = assign_min_cost_warp(sh.cost, ld, nd, nt, sh.pair_d, sh.pair_t, sh.u, lane);
      } else if (lane == 0) {
        np = assign_min_cost(sh.cost, ld, nd, nt, sh.pair_d, sh.pair_t);
      }
      if (lane == 0) {
        if (np < 0) { np = 0; vid.status = VBT_EINVAL; }
        sh.n_pairs = np;
      }
      __syncwarp();
    }
    if (lane == 0) {       // unmatched lists + low-IoU rejection, upstream order
      bool md[kMaxD], mt[kMaxT];
      for (int d = 0; d < nd; ++d) md[d] = false;
      for (int t = 0; t < nt; ++t) mt[t] = false;
      for (int i = 0; i < sh.n_pairs; ++i) { md[sh.pair_d[i]] = true; mt[sh.pair_t[i]] = true; }
      int nud = 0, nut = 0, k = 0;
      for (int d = 0; d < nd; ++d) if (!md[d]) sh.un_d[nud++] = d;
      for (int t = 0; t < nt; ++t) if (!mt[t]) sh.un_t[nut++] = t;
      for (int i = 0; i < sh.n_pairs; ++i) {
        int d = sh.pair_d[i], t = sh.pair_t[i];
        if (sh.iou[d * ld + t] < prm.iou_threshold) { sh.un_d[nud++] = d; sh.un_t[nut++] = t; }
        else { sh.pair_d[k] = d; sh.pair_t[k] = t; ++k; }
      }
      sh.n_pairs = k; sh.n_un_d = nud; sh.n_un_t = nut;
    }
    __syncwarp();
    for (int i = lane; i < sh.n_pairs; i += 32)
      trk_update(T(vid.order[sh.pair_t[i]]), sh.dets[sh.pair_d[i]], prm.delta_t);
    __syncwarp();
    // ---- second round: unmatched detections vs last observations (DIoU) --------------
    if (sh.n_un_d > 0 && sh.n_un_t > 0) {
      const int a_n = sh.n_un_d, b_n = sh.n_un_t;
      for (int i = lane; i < a_n * b_n; i += 32) {
        int a = i / b_n, b = i % b_n;
        const Trk& tk = T(vid.order[sh.un_t[b]]);
        sh.cost[a * ld + b] = diou_of(sh.dets[sh.un_d[a]], tk.last_obs);   // [-1]*5 when unseen
      }
      __syncwarp();
      if (lane == 0) {
        double mx = -INFINITY;
        bool any_nan = false;
        for (int a = 0; a < a_n; ++a)
          for (int b = 0; b < b_n; ++b) {
            double q = sh.cost[a * ld + b];
            if (isnan(q)) any_nan = true; else mx = fmax(mx, q);
          }
        sh.n_pairs = 0;
        sh.go = (!any_nan && mx > prm.iou_threshold) ? 1 : 0;
      }
      __syncwarp();
      if (sh.go) {
        for (int i = lane; i < a_n * b_n; i += 32) {
          int a = i / b_n, b = i % b_n;
          sh.iou[a * ld + b] = -sh.cost[a * ld + b];
        }
        __syncwarp();
        int pr[kMaxD], pc[kMaxD];
        int np = 0;
        if (max(a_n, b_n) <= 31) np = assign_min_cost_warp(sh.iou, ld, a_n, b_n, pr, pc, sh.u, lane);
        else if (lane == 0) np = assign_min_cost(sh.iou, ld, a_n, b_n, pr, pc);
        if (lane == 0) {
          if (np < 0) { np = 0; vid.status = VBT_EINVAL; }
          bool rm_d[kMaxD], rm_t[kMaxT];
          for (int a = 0; a < a_n; ++a) rm_d[a] = false;
          for (int b = 0; b < b_n; ++b) rm_t[b] = false;
          int k = 0;
          for (int i = 0; i < np; ++i) {
            if (sh.cost[pr[i] * ld + pc[i]] < prm.iou_threshold) continue;
            sh.pair_d[k] = sh.un_d[pr[i]]; sh.pair_t[k] = sh.un_t[pc[i]]; ++k;
            rm_d[pr[i]] = true; rm_t[pc[i]] = true;
          }
          sh.n_pairs = k;
          // np.setdiff1d: sorted unique remainder
          int nud = 0, nut = 0;
          bool keep_d[kMaxD], keep_t[kMaxT];
          for (int d = 0; d < nd; ++d) keep_d[d] = false;
          for (int t = 0; t < nt; ++t) keep_t[t] = false;
          for (int a = 0; a < a_n; ++a) if (!rm_d[a]) keep_d[sh.un_d[a]] = true;
          for (int b = 0; b < b_n; ++b) if (!rm_t[b]) keep_t[sh.un_t[b]] = true;
          for (int d = 0; d < nd; ++d) if (keep_d[d]) sh.un_d[nud++] = d;
          for (int t = 0; t < nt; ++t) if (keep_t[t]) sh.un_t[nut++] = t;
          sh.n_un_d = nud; sh.n_un_t = nut;
        }
      }
      __syncwarp();
      for (int i = lane; i < sh.n_pairs; i += 32)
        trk_update(T(vid.order[sh.pair_t[i]]), sh.dets[sh.pair_d[i]], prm.delta_t);
      __syncwarp();
    }
    for (int i = lane; i < sh.n_un_t; i += 32)
      trk_update(T(vid.order[sh.un_t[i]]), nullptr, prm.delta_t);
    __syncwarp();
    // ---- births, output rows, deaths (lane 0, list order matters) ---------------------
    if (lane == 0) {
      for (int i = 0; i < sh.n_un_d; ++i) {
        int slot = -1;
        for (int s = 0; s < max_tracks; ++s) if (!vid.used[s]) { slot = s; break; }
        if (slot < 0 || vid.n_tracks >= max_tracks) { vid.status = VBT_ECAPACITY; break; }
        vid.used[slot] = 1;
        trk_init(T(slot), sh.dets[sh.un_d[i]], vid.next_id++);
        vid.order[vid.n_tracks++] = slot;
      }
      const double time = (double)frame_no[(size_t)v * F + f] / vfps;     // track.py:169
      const bool last_frame = (f == nf - 1);
      int n_out = 0;
      int rc = row_count[v];
      for (int t = vid.n_tracks - 1; t >= 0; --t) {
        Trk& tk = T(vid.order[t]);
        double box[4];
        double sum = tk.last_obs[0] + tk.last_obs[1] + tk.last_obs[2] + tk.last_obs[3] + tk.last_obs[4];
        if (!tk.has_last || sum < 0) x_to_box(tk.x, box);
        else for (int j = 0; j < 4; ++j) box[j] = tk.last_obs[j];
        if (tk.tsu < 1 && (tk.hit_streak >= prm.min_hits || vid.frame_count <= prm.min_hits)) {
          if (rc >= row_cap) {
            vid.status = VBT_ECAPACITY;
          } else {
            double* r = vrows + (size_t)rc * VBT_ROW_COLS;
            r[0] = (double)(tk.id + 1);
            r[1] = time;
            r[2] = (box[0] + box[2]) / 2;          // odt.py:43-50
            r[3] = (box[1] + box[3]) / 2;
            r[4] = tk.x[4];                        // track.py:199
            r[5] = tk.x[5];
            r[6] = fabs(box[3] - box[1]);          // odt.py:32-40
            r[7] = fabs(box[2] - box[0]);          // odt.py:22-29
            ++rc;
          }
          if (last_out && last_frame && n_out < max_det) {
            double* o = last_out + ((size_t)v * max_det + n_out) * 9;
            o[0] = box[0]; o[1] = box[1]; o[2] = box[2]; o[3] = box[3];
            o[4] = (double)(tk.id + 1); o[5] = tk.cls; o[6] = tk.conf; o[7] = tk.x[4]; o[8] = tk.x[5];
          }
          ++n_out;
        }
      }
      row_count[v] = rc;
      if (last_out_count && last_frame) last_out_count[v] = min(n_out, max_det);
      int k = 0;
      for (int t = 0; t < vid.n_tracks; ++t) {
        int slot = vid.order[t];
        if (T(slot).tsu > prm.max_age) { vid.used[slot] = 0; continue; }
        vid.order[k++] = slot;
      }
      vid.n_tracks = k;
    }
    __syncwarp();
  }
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
  VBT_CHECK_CUDA(cudaFuncSetAttribute(tracker_update_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
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
  tracker_update_kernel<<<t->V, 32, shared_bytes(t->max_tracks), (cudaStream_t)stream>>>(
      t->videos, t->tracks, t->prm, dev_dets, dev_det_count, dev_frame_no, dev_fps, dev_n_frames,
      F, max_det, t->max_tracks, dev_rows, dev_row_count, row_cap, dev_last_out, dev_last_out_count);
  VBT_LAUNCHED(1);
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
