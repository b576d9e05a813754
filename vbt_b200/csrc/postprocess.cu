// K6 -- detection post-processing, one CTA per frame.
//
// replaces: the TFLite_Detection_PostProcess custom op that closes the exported
// EfficientDet-Lite graph (invoked through signature_fn at odt.py:61), in the mode the
// exporter configures: max_detections 25, one class per detection, fast (class-agnostic)
// NMS, nms_score_threshold -inf, IoU threshold 0.5, x/y/h/w scales 1, no clipping.
// Also vbt_pack_detections = odt.py:68-75 + odt.py:102-118.
//
// Order of candidates = score descending, ties by ascending anchor index (stable sort).
// Scores are 8-bit (LOGISTIC output, scale 1/256), so the order is produced by a
// 256-bin histogram, a range compaction of the best levels into shared memory and a
// bitonic sort of (255-level, anchor) keys; greedy suppression then walks the sorted
// keys 32 at a time in one warp.  fp32 arithmetic in the op's own operation order,
// -fmad=false.  Warp-level, latency-bound (north_star item 3).
#include <math.h>

#include "model.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCap = 2048;          // candidate keys staged per round
constexpr int kMaxDet = 32;

struct Box { float ymin, xmin, ymax, xmax; };

__device__ __forceinline__ Box decode(const float* __restrict__ anchors,
                                      const int8_t* __restrict__ enc,
                                      const float* __restrict__ exp_lut, float scale, int zp,
                                      int idx) {
  const float4 a = *reinterpret_cast<const float4*>(anchors + (size_t)idx * 4);  // y,x,h,w
  const char4 q = *reinterpret_cast<const char4*>(enc + (size_t)idx * 4);        // ty,tx,th,tw
  const float ty = scale * (float)((int)q.x - zp);
  const float tx = scale * (float)((int)q.y - zp);
  const float ycenter = ty * a.z + a.x;
  const float xcenter = tx * a.w + a.y;
  const float half_h = 0.5f * exp_lut[(int)q.z + 128] * a.z;
  const float half_w = 0.5f * exp_lut[(int)q.w + 128] * a.w;
  Box b;
  b.ymin = ycenter - half_h; b.xmin = xcenter - half_w;
  b.ymax = ycenter + half_h; b.xmax = xcenter + half_w;
  return b;
}

__device__ __forceinline__ float iou(const Box& i, const Box& j) {
  const float area_i = (i.ymax - i.ymin) * (i.xmax - i.xmin);
  const float area_j = (j.ymax - j.ymin) * (j.xmax - j.xmin);
  if (area_i <= 0 || area_j <= 0) return 0.0f;
  const float iymin = fmaxf(i.ymin, j.ymin), ixmin = fmaxf(i.xmin, j.xmin);
  const float iymax = fminf(i.ymax, j.ymax), ixmax = fminf(i.xmax, j.xmax);
  const float inter = fmaxf(iymax - iymin, 0.0f) * fmaxf(ixmax - ixmin, 0.0f);
  return inter / (area_i + area_j - inter);
}

struct Shared {
  int hist[256];
  unsigned int keys[kCap];
  Box sel_box[kMaxDet];
  int sel_idx[kMaxDet];
  int sel_q[kMaxDet];
  int n_sel, n_cand, lo, hi, mode_b, scan_pos, warp_tot[kThreads / 32];
};

// greedy suppression over keys[0..n) (sorted), executed by warp 0
__device__ void greedy(Shared& sh, int n, const float* anchors, const int8_t* enc,
                       const float* exp_lut, float scale, int zp, float iou_thr, int max_det) {
  const int lane = threadIdx.x;
  int n_sel = sh.n_sel;
  for (int base = 0; base < n && n_sel < max_det; base += 32) {
    const int i = base + lane;
    bool alive = i < n;
    Box b = {0.f, 0.f, 0.f, 0.f};
    int idx = 0, level = 0;
    if (alive) {
      const unsigned int key = sh.keys[i];
      idx = key & 0xffff;
      level = 255 - (int)(key >> 16);
      b = decode(anchors, enc, exp_lut, scale, zp, idx);
      for (int s = 0; s < n_sel; ++s)
        if (iou(sh.sel_box[s], b) > iou_thr) { alive = false; break; }
    }
    while (n_sel < max_det) {
      const unsigned int m = __ballot_sync(0xffffffffu, alive);
      if (!m) break;
      const int leader = __ffs(m) - 1;
      Box lb;
      lb.ymin = __shfl_sync(0xffffffffu, b.ymin, leader);
      lb.xmin = __shfl_sync(0xffffffffu, b.xmin, leader);
      lb.ymax = __shfl_sync(0xffffffffu, b.ymax, leader);
      lb.xmax = __shfl_sync(0xffffffffu, b.xmax, leader);
      if (lane == leader) {
        sh.sel_box[n_sel] = b; sh.sel_idx[n_sel] = idx; sh.sel_q[n_sel] = level;
        alive = false;
      } else if (alive && lane > leader) {
        if (iou(lb, b) > iou_thr) alive = false;
      }
      ++n_sel;
    }
    __syncwarp();
  }
  if (lane == 0) sh.n_sel = n_sel;
}

__global__ void __launch_bounds__(kThreads) postprocess_kernel(
    const int8_t* __restrict__ cls, const int8_t* __restrict__ box, const float* __restrict__ anchors,
    const float* __restrict__ exp_lut, int N, int Np, float box_scale, int box_zp, float iou_thr,
    int max_det, int min_q, float* __restrict__ out_boxes, float* __restrict__ out_classes,
    float* __restrict__ out_scores, float* __restrict__ out_count, int32_t* __restrict__ out_index) {
  extern __shared__ __align__(16) unsigned char dyn[];
  int8_t* sc = reinterpret_cast<int8_t*>(dyn);             // [Np] scores of this frame
  __shared__ Shared sh;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int8_t* fcls = cls + (size_t)b * Np;
  const int8_t* fbox = box + (size_t)b * Np * 4;

  for (int i = tid; i < 256; i += kThreads) sh.hist[i] = 0;
  if (tid == 0) { sh.n_sel = 0; }
  __syncthreads();
  // stage the scores (128-bit loads; Np is a multiple of 16) and histogram them
  for (int i = tid * 16; i < Np; i += kThreads * 16) {
    const int4 v = *reinterpret_cast<const int4*>(fcls + i);
    *reinterpret_cast<int4*>(sc + i) = v;
  }
  __syncthreads();
  for (int i = tid; i < N; i += kThreads) {
    const int q = (int)sc[i] + 128;
    // warp-aggregated increment: lanes holding the same level elect one adder
    const unsigned int active = __activemask();
    const unsigned int peers = __match_any_sync(active, q);
    if ((__ffs(peers) - 1) == (tid & 31)) atomicAdd(&sh.hist[q], __popc(peers));
  }
  __syncthreads();

  int hi = 255;                                   // level = q + 128
  const int min_level = min_q + 128;
  while (true) {
    if (tid == 0) {
      int h = hi;
      while (h >= min_level && sh.hist[h] == 0) --h;
      sh.hi = h; sh.mode_b = 0; sh.n_cand = 0; sh.scan_pos = 0;
      if (h >= min_level) {
        if (sh.hist[h] > kCap) { sh.mode_b = 1; sh.lo = h; }
        else {
          int tot = sh.hist[h], l = h;
          while (l - 1 >= min_level && tot + sh.hist[l - 1] <= kCap) { --l; tot += sh.hist[l]; }
          sh.lo = l;
        }
      }
    }
    __syncthreads();
    hi = sh.hi;
    if (hi < min_level || sh.n_sel >= max_det) break;
    const int lo = sh.lo;
    if (!sh.mode_b) {
      for (int i = tid; i < N; i += kThreads) {
        const int level = (int)sc[i] + 128;
        if (level >= lo && level <= hi) {
          const int pos = atomicAdd(&sh.n_cand, 1);
          sh.keys[pos] = ((unsigned int)(255 - level) << 16) | (unsigned int)i;
        }
      }
      __syncthreads();
      const int n = sh.n_cand;
      int n2 = 32;
      while (n2 < n) n2 <<= 1;
      for (int i = n + tid; i < n2; i += kThreads) sh.keys[i] = 0xffffffffu;
      __syncthreads();
      for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < n2; i += kThreads) {
            const int p = i ^ j;
            if (p > i) {
              const unsigned int a = sh.keys[i], c = sh.keys[p];
              const bool up = (i & k) == 0;
              if ((a > c) == up) { sh.keys[i] = c; sh.keys[p] = a; }
            }
          }
          __syncthreads();
        }
      if (tid < 32) greedy(sh, n, anchors, fbox, exp_lut, box_scale, box_zp, iou_thr, max_det);
      __syncthreads();
    } else {
      // one level holds more than kCap anchors: walk it in anchor order, kCap at a time
      int pos = 0;
      while (pos < N && sh.n_sel < max_det) {
        if (tid == 0) sh.n_cand = 0;
        __syncthreads();
        while (pos < N && sh.n_cand + kThreads <= kCap) {
          const int i = pos + tid;
          const bool hit = i < N && ((int)sc[i] + 128) == hi;
          const unsigned int m = __ballot_sync(0xffffffffu, hit);
          if ((tid & 31) == 0) sh.warp_tot[tid >> 5] = __popc(m);
          __syncthreads();
          int off = sh.n_cand;
          for (int w = 0; w < (tid >> 5); ++w) off += sh.warp_tot[w];
          if (hit) sh.keys[off + __popc(m & ((1u << (tid & 31)) - 1))] =
                       ((unsigned int)(255 - hi) << 16) | (unsigned int)i;
          __syncthreads();
          if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kThreads / 32; ++w) tot += sh.warp_tot[w];
            sh.n_cand += tot;
          }
          pos += kThreads;
          __syncthreads();
        }
        if (tid < 32) greedy(sh, sh.n_cand, anchors, fbox, exp_lut, box_scale, box_zp, iou_thr, max_det);
        __syncthreads();
      }
    }
    hi = lo - 1;
  }
  __syncthreads();
  // write the four output tensors, zero padded (odt.py:64-66 reads count, scores, boxes)
  const int n_sel = sh.n_sel;
  for (int i = tid; i < max_det; i += kThreads) {
    float* ob = out_boxes + ((size_t)b * max_det + i) * 4;
    if (i < n_sel) {
      ob[0] = sh.sel_box[i].ymin; ob[1] = sh.sel_box[i].xmin;
      ob[2] = sh.sel_box[i].ymax; ob[3] = sh.sel_box[i].xmax;
      out_scores[(size_t)b * max_det + i] = 0.00390625f * (float)sh.sel_q[i];
      out_index[(size_t)b * max_det + i] = sh.sel_idx[i];
    } else {
      ob[0] = ob[1] = ob[2] = ob[3] = 0.0f;
      out_scores[(size_t)b * max_det + i] = 0.0f;
      out_index[(size_t)b * max_det + i] = -1;
    }
    out_classes[(size_t)b * max_det + i] = 0.0f;
  }
  if (tid == 0) out_count[b] = (float)n_sel;
}

__global__ void pack_detections_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                       const float* __restrict__ count, int F, int max_det,
                                       float threshold, double* __restrict__ dets,
                                       int32_t* __restrict__ det_count) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int n = (int)count[f];
  int k = 0;
  for (int i = 0; i < n && i < max_det; ++i) {
    const float s = scores[(size_t)f * max_det + i];
    if (s >= threshold) {                                   // odt.py:71
      const float* bb = boxes + ((size_t)f * max_det + i) * 4;   // ymin,xmin,ymax,xmax
      double* d = dets + ((size_t)f * max_det + k) * 6;
      d[0] = (double)bb[1]; d[1] = (double)bb[0]; d[2] = (double)bb[3]; d[3] = (double)bb[2];
      d[4] = (double)s; d[5] = 0.0;                         // odt.py:116
      ++k;
    }
  }
  det_count[f] = k;
}

}  // namespace

extern "C" {

int vbt_postprocess_q8(const vbt_model* m, const int8_t* dev_cls, const int8_t* dev_box, int B,
                       float iou_threshold, int max_det, int min_score_q, float* dev_boxes,
                       float* dev_classes, float* dev_scores, float* dev_count,
                       int32_t* dev_index, void* stream) {
  VBT_REQUIRE(m && dev_cls && dev_box && dev_boxes && dev_classes && dev_scores && dev_count &&
                  dev_index, "vbt_postprocess_q8: null pointer");
  VBT_REQUIRE(B > 0 && max_det > 0 && max_det <= kMaxDet, "vbt_postprocess_q8: B=%d max_det=%d", B,
              max_det);
  VBT_REQUIRE(min_score_q >= -128 && min_score_q <= 127, "vbt_postprocess_q8: min_score_q range");
  VBT_REQUIRE(m->hdr.n_anchors < 65536, "vbt_postprocess_q8: anchor index must fit 16 bits");
  const int Np = m->hdr.n_anchors_pad;
  static thread_local int smem_set = 0;
  if (smem_set < Np) {
    VBT_CHECK_CUDA(cudaFuncSetAttribute(postprocess_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, Np));
    smem_set = Np;
  }
  postprocess_kernel<<<B, kThreads, Np, (cudaStream_t)stream>>>(
      dev_cls, dev_box, m->dev_anchors, m->dev_exp_lut, m->hdr.n_anchors, Np, m->hdr.box_scale,
      m->hdr.box_zp, iou_threshold, max_det, min_score_q, dev_boxes, dev_classes, dev_scores,
      dev_count, dev_index);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

int vbt_pack_detections(const float* dev_boxes, const float* dev_scores, const float* dev_count,
                        int F, int max_det, float threshold, double* dev_dets,
                        int32_t* dev_det_count, void* stream) {
  VBT_REQUIRE(dev_boxes && dev_scores && dev_count && dev_dets && dev_det_count && F > 0 &&
                  max_det > 0, "vbt_pack_detections: bad arguments");
  if (int rc = vbt::ensure_device()) return rc;
  pack_detections_kernel<<<vbt::ceil_div(F, 128), 128, 0, (cudaStream_t)stream>>>(
      dev_boxes, dev_scores, dev_count, F, max_det, threshold, dev_dets, dev_det_count);
  VBT_LAUNCHED(1);
  return VBT_OK;
}

}  // extern "C"
