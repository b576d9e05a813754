// K8 -- velocity smoothing + concentric/eccentric phase segmentation, one thread per
// (video, id) lane, streaming: a lane's state persists on the device between calls so a
// video can be fed in frame batches exactly like the reference feeds
// VelocityTracker.process_measurements one sample at a time.
//
// replaces: plot.py:87-95 (rolling(5)/expanding means -- pandas' Kahan add/remove
// window mean), plot.analyze_df (plot.py:33-47), VelocityTracker.py:92-230,
// RunningAverage.py:16-27, Phase.py:6-40, and the per-id cumulative path of
// track.py:109-113.
//
// Latency-bound fp64 recurrence: no roofline, reported as microseconds per video.
// Built with -fmad=false: every multiply/add rounds separately, like CPython floats.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kConcentric = 0, kEccentric = 1, kHold = 2;   // Phase.py:12-14
constexpr int kStartThreshold = 3, kEndThreshold = 1;       // VelocityTracker.py:11-12
constexpr int kRaWindow = 30;                               // VelocityTracker.py:44
constexpr int kRoll = 5;                                    // plot.py:90-92

struct Lane {
  // pandas window-mean state for columns x,y,dx,dy (rolling 5) and h,w (expanding)
  double sum[6], comp_add[6], comp_rem[6], prev[6];
  int neg[6], same[6];
  double ring[4][kRoll];
  long long seen;            // rows consumed
  // VelocityTracker
  int phase, have_max, have_prev, neg_cnt, pos_cnt;
  double max_y_diff, y_prev;
  // RunningAverage(30), shared by width and height (VelocityTracker.py:44-45,98-99)
  double ra_ring[kRaWindow];
  double ra_total;
  int ra_count, ra_head;
  // path + phases live in side arrays
  int path_len, n_phases, status;
  // track.py:109-113 cumulative Euclidean path of the RAW centres
  double cum_path, raw_x_prev, raw_y_prev;
};

struct Ctx {
  Lane* s;
  double* px; double* py; double* pw; double* ph; double* pt;   // path columns
  double* phases;                                                // [phase_cap][6]
  int path_cap, phase_cap;
  double plate_diameter, diff_threshold, min_distance;
};

__device__ __forceinline__ void lane_init(Lane& s) {
  for (int c = 0; c < 6; ++c) {
    s.sum[c] = s.comp_add[c] = s.comp_rem[c] = s.prev[c] = 0.0;
    s.neg[c] = s.same[c] = 0;
  }
  for (int c = 0; c < 4; ++c)
    for (int i = 0; i < kRoll; ++i) s.ring[c][i] = 0.0;
  s.seen = 0;
  s.phase = kHold;
  s.have_max = s.have_prev = s.neg_cnt = s.pos_cnt = 0;
  s.max_y_diff = s.y_prev = 0.0;
  for (int i = 0; i < kRaWindow; ++i) s.ra_ring[i] = 0.0;
  s.ra_total = 0.0;
  s.ra_count = s.ra_head = 0;
  s.path_len = s.n_phases = s.status = 0;
  s.cum_path = s.raw_x_prev = s.raw_y_prev = 0.0;
}

// pandas roll_mean: one value enters; for the rolling columns the value that left the
// window (row seen-5) is removed first.  `window` 0 = expanding.
__device__ __forceinline__ double window_mean(Lane& s, int c, double val, int window) {
  long long i = s.seen;
  long long nobs = window ? (i < window ? i : window) : i;   // before this row
  if (window && i >= window) {
    double old = s.ring[c][i % window];
    nobs -= 1;
    double y = -old - s.comp_rem[c];
    double t = s.sum[c] + y;
    s.comp_rem[c] = t - s.sum[c] - y;
    s.sum[c] = t;
    if (signbit(old)) s.neg[c] -= 1;
  }
  if (window) s.ring[c][i % window] = val;
  nobs += 1;
  double y = val - s.comp_add[c];
  double t = s.sum[c] + y;
  s.comp_add[c] = t - s.sum[c] - y;
  s.sum[c] = t;
  if (signbit(val)) s.neg[c] += 1;
  if (i == 0) s.prev[c] = val;
  s.same[c] = (val == s.prev[c]) ? s.same[c] + 1 : 1;
  s.prev[c] = val;
  double r = s.sum[c] / (double)nobs;
  if (s.same[c] >= nobs) r = s.prev[c];
  else if (s.neg[c] == 0 && r < 0) r = 0.0;
  else if (s.neg[c] == nobs && r > 0) r = 0.0;
  return r;
}

// RunningAverage.update (RunningAverage.py:16-27)
__device__ __forceinline__ double ra_update(Lane& s, double val) {
  int tail = (s.ra_head + s.ra_count) % kRaWindow;
  s.ra_ring[tail] = val;
  s.ra_total += val;
  s.ra_count += 1;
  if (s.ra_count >= kRaWindow) {
    double avg = s.ra_total / (double)kRaWindow;
    s.ra_total -= s.ra_ring[s.ra_head];
    s.ra_head = (s.ra_head + 1) % kRaWindow;
    s.ra_count -= 1;
    return avg;
  }
  return s.ra_total / (double)s.ra_count;
}

__device__ __forceinline__ void path_reset(Ctx& c) { c.s->path_len = 0; }

__device__ __forceinline__ void path_append(Ctx& c, double x, double y, double w, double h,
                                            double t) {
  int n = c.s->path_len;
  if (n >= c.path_cap) { c.s->status = VBT_ECAPACITY; return; }
  c.px[n] = x; c.py[n] = y; c.pw[n] = w; c.ph[n] = h; c.pt[n] = t;
  c.s->path_len = n + 1;
}

// VelocityTracker._filter_phases (VelocityTracker.py:50-67)
__device__ void drop_small(Ctx& c) {
  double lim = c.s->max_y_diff / 2;
  int k = 0;
  for (int i = 0; i < c.s->n_phases; ++i) {
    double* p = c.phases + (size_t)i * 6;
    if (fabs(p[2] - p[3]) < lim) continue;
    if (k != i) for (int j = 0; j < 6; ++j) c.phases[(size_t)k * 6 + j] = p[j];
    ++k;
  }
  c.s->n_phases = k;
}

// VelocityTracker._end_phase (VelocityTracker.py:171-222)
__device__ void close_phase(Ctx& c) {
  Lane& s = *c.s;
  int n = s.path_len;
  if (n > 0) {
    int hi = 0, lo = 0;
    for (int i = 1; i < n; ++i) {           // first occurrence wins, like np.argmax/argmin
      if (c.py[i] > c.py[hi]) hi = i;
      if (c.py[i] < c.py[lo]) lo = i;
    }
    int a = (s.phase == kConcentric) ? hi : lo;
    int b = (s.phase == kConcentric) ? lo : hi;
    double y_diff = fabs(c.py[a] - c.py[b]);
    if (!s.have_max || y_diff > s.max_y_diff) {
      s.have_max = 1;
      s.max_y_diff = y_diff;
      drop_small(c);
    }
    if (y_diff > s.max_y_diff * c.diff_threshold) {
      double dist = 0.0;
      for (int i = a + 1; i <= b; ++i) {
        double ddx = fabs(c.px[i] - c.px[i - 1]) / ((c.pw[i] + c.pw[i - 1]) / 2) * c.plate_diameter;
        double ddy = fabs(c.py[i] - c.py[i - 1]) / ((c.ph[i] + c.ph[i - 1]) / 2) * c.plate_diameter;
        dist += ddx + ddy;
      }
      if (!(dist < c.min_distance)) {
        if (s.n_phases >= c.phase_cap) {
          s.status = VBT_ECAPACITY;
        } else {
          double* p = c.phases + (size_t)s.n_phases * 6;
          p[0] = c.pt[a]; p[1] = c.pt[b]; p[2] = c.py[a]; p[3] = c.py[b]; p[4] = dist;
          p[5] = (double)s.phase;
          s.n_phases += 1;
          drop_small(c);
        }
      }
    }
  }
  s.phase = kHold;
  s.neg_cnt = s.pos_cnt = 0;
}

// VelocityTracker.process_measurements (VelocityTracker.py:92-158)
__device__ void step(Ctx& c, double time, double x, double y, double dy, double h, double w) {
  Lane& s = *c.s;
  double width = ra_update(s, w);    // :98
  double height = ra_update(s, h);   // :99  same running average on purpose
  if (s.have_prev) dy = y - s.y_prev;
  if (s.phase != kHold) path_append(c, x, y, width, height, time);
  if (s.phase == kConcentric) {
    if (dy > 0) {
      s.pos_cnt += 1; s.neg_cnt = 0;
      if (s.pos_cnt >= kEndThreshold) close_phase(c);
    } else {
      s.pos_cnt = 0;
    }
  }
  if (s.phase == kEccentric) {
    if (dy < 0) {
      s.neg_cnt += 1; s.pos_cnt = 0;
      if (s.neg_cnt >= kEndThreshold) close_phase(c);
    } else {
      s.neg_cnt = 0; s.pos_cnt += 1;
    }
  }
  if (dy < 0 && s.phase == kHold) {
    s.neg_cnt += 1; s.pos_cnt = 0;
    if (s.neg_cnt == 1) path_reset(c); else path_append(c, x, y, width, height, time);
    if (s.neg_cnt >= kStartThreshold) { s.phase = kConcentric; s.neg_cnt = s.pos_cnt = 0; }
  }
  if (dy > 0 && s.phase == kHold) {
    s.pos_cnt += 1; s.neg_cnt = 0;
    if (s.pos_cnt == 1) path_reset(c); else path_append(c, x, y, width, height, time);
    if (s.pos_cnt >= kStartThreshold) { s.phase = kEccentric; s.neg_cnt = s.pos_cnt = 0; }
  }
  s.have_prev = 1;
  s.y_prev = y;
}

__global__ void velocity_reset_kernel(Lane* lanes, int L) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L) lane_init(lanes[l]);
}

__global__ void velocity_update_kernel(Lane* lanes, double* path, double* phases,
                                       int path_cap, int phase_cap, const double* rows,
                                       const int32_t* row_count, int row_cap,
                                       const int32_t* lane_table, const int32_t* lane_id,
                                       int32_t* lane_begin, int L, double plate_diameter,
                                       double diff_threshold, double min_distance,
                                       int smooth, int finish) {
  // one lane per CTA (thread 0): lanes are independent serial recurrences that diverge
  // from one another, so they must not share a warp
  const int l = blockIdx.x;
  if (l >= L || threadIdx.x != 0) return;
  Lane s = lanes[l];
  Ctx c;
  c.s = &s;
  double* pbase = path + (size_t)l * 5 * path_cap;
  c.px = pbase; c.py = pbase + path_cap; c.pw = pbase + 2 * (size_t)path_cap;
  c.ph = pbase + 3 * (size_t)path_cap; c.pt = pbase + 4 * (size_t)path_cap;
  c.phases = phases + (size_t)l * phase_cap * 6;
  c.path_cap = path_cap; c.phase_cap = phase_cap;
  c.plate_diameter = plate_diameter; c.diff_threshold = diff_threshold;
  c.min_distance = min_distance;

  int table = lane_table[l];
  int want = lane_id[l];
  int end = row_count[table];
  if (end > row_cap) end = row_cap;
  const double* tab = rows + (size_t)table * row_cap * VBT_ROW_COLS;
  for (int r = lane_begin[l]; r < end; ++r) {
    const double* row = tab + (size_t)r * VBT_ROW_COLS;
    if (want >= 0 && (int)row[0] != want) continue;
    double t = row[1], x = row[2], y = row[3], dx = row[4], dy = row[5], h = row[6], w = row[7];
    if (s.seen > 0) {   // track.py:109-113: sqrt(dx^2 + dy^2) of consecutive raw centres
      double ex = x - s.raw_x_prev, ey = y - s.raw_y_prev;
      s.cum_path += sqrt(ex * ex + ey * ey);
    }
    s.raw_x_prev = x; s.raw_y_prev = y;
    if (smooth) {       // plot.py:90-95
      x = window_mean(s, 0, x, kRoll);
      y = window_mean(s, 1, y, kRoll);
      dx = window_mean(s, 2, dx, kRoll);
      dy = window_mean(s, 3, dy, kRoll);
      h = window_mean(s, 4, h, 0);
      w = window_mean(s, 5, w, 0);
    }
    (void)dx;           // VelocityTracker never reads dx (VelocityTracker.py:92-158)
    s.seen += 1;
    step(c, t, x, y, dy, h, w);
  }
  lane_begin[l] = end;
  if (finish && s.phase != kHold) close_phase(c);   // VelocityTracker.py:224-230
  lanes[l] = s;
}

__global__ void velocity_export_kernel(const Lane* lanes, int L, int32_t* phase_count,
                                       double* lane_state) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const Lane& s = lanes[l];
  phase_count[l] = s.n_phases;
  double* o = lane_state + (size_t)l * 8;
  o[0] = (double)s.phase;
  o[1] = s.have_max ? s.max_y_diff : nan("");
  o[2] = (double)s.seen;
  o[3] = s.cum_path;
  o[4] = (double)s.status;
  o[5] = s.have_prev ? s.y_prev : nan("");
  o[6] = (double)s.neg_cnt;
  o[7] = (double)s.pos_cnt;
}

__global__ void running_average_kernel(double* state, int window, const double* values, int n,
                                       double* out) {
  if (blockIdx.x || threadIdx.x) return;
  double total = state[window];
  int count = (int)state[window + 1];
  int head = (int)state[window + 2];
  for (int i = 0; i < n; ++i) {
    double val = values[i];
    state[(head + count) % window] = val;
    total += val;
    count += 1;
    if (count >= window) {
      out[i] = total / (double)window;
      total -= state[head];
      head = (head + 1) % window;
      count -= 1;
    } else {
      out[i] = total / (double)count;
    }
  }
  state[window] = total;
  state[window + 1] = (double)count;
  state[window + 2] = (double)head;
}

}  // namespace

struct vbt_velocity {
  int L, path_cap, phase_cap;
  Lane* lanes;
  double* path;
  double* phases;
  int32_t* phase_count;
  double* lane_state;
};

extern "C" {

int vbt_velocity_create(int L, int path_cap, int phase_cap, vbt_velocity** out) {
  VBT_REQUIRE(out && L > 0 && path_cap > 0 && phase_cap > 0, "vbt_velocity_create: bad sizes");
  if (int rc = vbt::ensure_device()) return rc;
  vbt_velocity* v = new vbt_velocity();
  v->L = L; v->path_cap = path_cap; v->phase_cap = phase_cap;
  VBT_CHECK_CUDA(cudaMalloc(&v->lanes, sizeof(Lane) * (size_t)L));
  VBT_CHECK_CUDA(cudaMalloc(&v->path, sizeof(double) * 5 * (size_t)path_cap * L));
  VBT_CHECK_CUDA(cudaMalloc(&v->phases, sizeof(double) * 6 * (size_t)phase_cap * L));
  VBT_CHECK_CUDA(cudaMalloc(&v->phase_count, sizeof(int32_t) * (size_t)L));
  VBT_CHECK_CUDA(cudaMalloc(&v->lane_state, sizeof(double) * 8 * (size_t)L));
  *out = v;
  return vbt_velocity_reset(v, nullptr);
}

void vbt_velocity_destroy(vbt_velocity* v) {
  if (!v) return;
  cudaFree(v->lanes); cudaFree(v->path); cudaFree(v->phases); cudaFree(v->phase_count);
  cudaFree(v->lane_state);
  delete v;
}

int vbt_velocity_reset(vbt_velocity* v, void* stream) {
  VBT_REQUIRE(v, "vbt_velocity_reset: null handle");
  velocity_reset_kernel<<<vbt::ceil_div(v->L, 64), 64, 0, (cudaStream_t)stream>>>(v->lanes, v->L);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

int vbt_velocity_update(vbt_velocity* v, const double* dev_rows, const int32_t* dev_row_count,
                        int row_cap, const int32_t* dev_lane_table, const int32_t* dev_lane_id,
                        int32_t* dev_lane_begin, int L, double plate_diameter,
                        double diff_threshold, double min_distance, int smooth, int finish,
                        void* stream) {
  VBT_REQUIRE(v && dev_rows && dev_row_count && dev_lane_table && dev_lane_id && dev_lane_begin,
              "vbt_velocity_update: null pointer");
  VBT_REQUIRE(L > 0 && L <= v->L && row_cap > 0, "vbt_velocity_update: L=%d exceeds lanes=%d", L,
              v->L);
  velocity_update_kernel<<<L, 32, 0, (cudaStream_t)stream>>>(
      v->lanes, v->path, v->phases, v->path_cap, v->phase_cap, dev_rows, dev_row_count, row_cap,
      dev_lane_table, dev_lane_id, dev_lane_begin, L, plate_diameter, diff_threshold,
      min_distance, smooth, finish);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

int vbt_velocity_read(vbt_velocity* v, double* host_phases, int32_t* host_phase_count,
                      double* host_lane_state, void* stream) {
  VBT_REQUIRE(v && host_phases && host_phase_count && host_lane_state,
              "vbt_velocity_read: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  velocity_export_kernel<<<vbt::ceil_div(v->L, 64), 64, 0, st>>>(v->lanes, v->L, v->phase_count,
                                                                  v->lane_state);
  VBT_LAUNCHED(1);
  VBT_CHECK_CUDA(cudaMemcpyAsync(host_phases, v->phases,
                                 sizeof(double) * 6 * (size_t)v->phase_cap * v->L,
                                 cudaMemcpyDeviceToHost, st));
  VBT_CHECK_CUDA(cudaMemcpyAsync(host_phase_count, v->phase_count, sizeof(int32_t) * (size_t)v->L,
                                 cudaMemcpyDeviceToHost, st));
  VBT_CHECK_CUDA(cudaMemcpyAsync(host_lane_state, v->lane_state, sizeof(double) * 8 * (size_t)v->L,
                                 cudaMemcpyDeviceToHost, st));
  VBT_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int l = 0; l < v->L; ++l) {
    if ((int)host_lane_state[(size_t)l * 8 + 4] != 0) {
      vbt::set_error("velocity lane %d overflowed its path (%d) or phase (%d) table", l,
                     v->path_cap, v->phase_cap);
      return VBT_ECAPACITY;
    }
  }
  return VBT_OK;
}

int vbt_running_average(double* dev_state, int window, const double* dev_values, int n,
                        double* dev_out, void* stream) {
  VBT_REQUIRE(dev_state && dev_values && dev_out && window > 0 && n >= 0,
              "vbt_running_average: bad arguments");
  if (int rc = vbt::ensure_device()) return rc;
  if (n == 0) return VBT_OK;
  running_average_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_state, window, dev_values, n,
                                                           dev_out);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

}  // extern "C"
