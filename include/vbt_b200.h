/* vbt_b200 -- C ABI of the B200-native vbt hot path (libvbt_b200.so).
 *
 * The reference (simonkosina/vbt) has no FFI of its own: its hot path is Python calling
 * into third-party wheels (tensorflow, tflite_runtime, ocsort, pandas).  Every entry
 * point below replaces one of those call sites; the "replaces" line cites it as
 * file:line in the reference checkout.  The Python side (the vbt_b200 package) binds these
 * with ctypes and mirrors the reference's own call surfaces (Interpreter, OCSort,
 * VelocityTracker, track.py CLI); INTEGRATION.md shows the stubs a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; every `dev` pointer is device memory owned by the
 *    caller; nothing is allocated after a *_create call returns;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *    stream-ordered, no entry point synchronises unless its comment says so;
 *  - return 0 on success, a negative VBT_E* code otherwise; vbt_last_error() returns the
 *    message of the last failure on the calling thread;
 *  - there is no CPU path: on a machine without a usable sm_100 device every compute
 *    entry point fails with VBT_ECUDA.
 */
#ifndef VBT_B200_H
#define VBT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBT_OK 0
#define VBT_EINVAL (-1)    /* bad argument */
#define VBT_ECUDA (-2)     /* CUDA runtime / driver error, or no device */
#define VBT_ECAPACITY (-3) /* a caller-sized table overflowed (rows, tracks, phases) */
#define VBT_EFORMAT (-4)   /* malformed model blob */

#define VBT_MAX_DETECTIONS 25 /* tflite_max_detections of the exported models */
#define VBT_ROW_COLS 8        /* id,time,x,y,dx,dy,norm_plate_height,norm_plate_width */
#define VBT_PHASE_COLS 6      /* time_start,time_end,y_start,y_end,rom,type */

int vbt_abi_version(void);
const char* vbt_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's
 * gpu_launches); graph replays count the kernels they contain */
long long vbt_launch_count(void);

/* ---- K1 frame preprocessing -------------------------------------------------------
 * replaces: cv2.cvtColor BGR->RGB (track.py:171) + odt.preprocess_image (odt.py:10-19,
 * tf.image.resize bilinear, half-pixel centres, no antialias, stretch) + tf.cast uint8.
 * frames: u8 [B,H,W,3]; out: u8 [B,S,S,3] RGB.  swap_rb != 0 when `frames` is BGR. */
int vbt_preprocess_u8(const uint8_t* dev_frames, int B, int H, int W, int swap_rb,
                      uint8_t* dev_out, int S, void* stream);

/* Row-sparse ingest of host frames (replaces the host->device half of the same call sites:
 * the reference resizes on the host, so only the resized image ever moves; here only the
 * source rows the bilinear kernel touches cross PCIe -- 2*S of H, e.g. 16 of every 27 rows
 * for 1080 -> 320).
 * vbt_copy_rows_h2d: host_frames u8 [B,H,W,3] in PINNED host memory; the touched rows repeat
 * with `period` = H / gcd(H,S) source rows; host_rows[n_rows] are the touched row indices
 * inside one period.  Issues n_rows strided DMA copies (cudaMemcpy2DAsync) on `stream` into
 * dev_rows u8 [B * H/period, n_rows, W*3].
 * vbt_preprocess_rows_u8: K1 over that table; dev_row_map i32 [H] maps a source row to its
 * table row within the frame (rows_per_frame = H/period * n_rows). */
int vbt_copy_rows_h2d(const uint8_t* host_frames, int B, int H, int W, int period,
                      const int32_t* host_rows, int n_rows, uint8_t* dev_rows, void* stream);
int vbt_preprocess_rows_u8(const uint8_t* dev_rows, int B, int H, int W, int rows_per_frame,
                           const int32_t* dev_row_map, int swap_rb, uint8_t* dev_out, int S,
                           void* stream);

/* ---- K2-K5 the EfficientDet-Lite network ------------------------------------------
 * replaces: tflite_runtime Interpreter(model_path) / allocate_tensors (track.py:93-94)
 * and the graph part of signature_fn(images=...) (odt.py:58-61).
 * The blob is the layer program written by vbt_b200/effdet.py (host memory). */
typedef struct vbt_model vbt_model;
int vbt_model_create(const void* blob, size_t blob_bytes, vbt_model** out);
void vbt_model_destroy(vbt_model* m);
/* info[0]=input size S, [1]=anchors N, [2]=workspace bytes per frame, [3]=ops,
 * [4]=classes, [5]=kernels launched per vbt_detect call, [6]=Np (N rounded up to 16) */
int vbt_model_info(const vbt_model* m, long long info[8]);
/* launch plan: group_len i32 [ops] on the host; group_len[i] = number of consecutive ops the
 * kernel launched at op i covers (fused [ADD ->] DW3x3 -> PW groups), 0 for ops inside a group */
int vbt_model_plan(const vbt_model* m, int32_t* host_group_len);
/* kind i32 [ops] of the launch that starts at op i: 0 = a single op or a fused BiFPN node / head
 * stage, 1 = a whole MBConv block ([1x1 expand ->] depthwise -> 1x1 project [+ residual]) */
int vbt_model_plan_kinds(const vbt_model* m, int32_t* host_kind);
/* in: u8 [B,S,S,3] RGB; out_cls: i8 [B,Np] post-LOGISTIC scores (scale 1/256, zp -128);
 * out_box: i8 [B,Np,4] (ty,tx,th,tw) with the model's box quantisation; Np = info[6] =
 * N rounded up to 16 (row stride; the pad entries are never read).
 * workspace: >= B * info[2] bytes, 256-byte aligned. */
int vbt_detect(vbt_model* m, const uint8_t* dev_in, int B, void* dev_workspace,
               size_t workspace_bytes, int8_t* dev_out_cls, int8_t* dev_out_box,
               void* stream);

/* Per-op device timing for bench.py's roofline: while enabled, vbt_detect brackets every
 * op with CUDA events on its stream; vbt_model_op_times synchronises and returns the
 * accumulated milliseconds per op (host f64 [ops]) and the number of calls covered. */
int vbt_model_profile(vbt_model* m, int enable);
int vbt_model_op_times(vbt_model* m, double* host_ms, long long* calls);

/* ---- K6 detection post-processing --------------------------------------------------
 * replaces: the TFLite_Detection_PostProcess custom op inside signature_fn (odt.py:61)
 * in fast-NMS mode (class-agnostic greedy NMS, stable score-descending order, IoU > thr
 * suppresses, no clipping).  Anchors (ycentre,xcentre,h,w) and the box dequantisation
 * come from the model.  min_score_q: candidates with int8 score below it never reach
 * NMS (-128 = the op's own behaviour, nms_score_threshold = -inf).
 * outputs per frame, zero padded: boxes f32 [B,max_det,4] (ymin,xmin,ymax,xmax),
 * classes f32 [B,max_det], scores f32 [B,max_det], count f32 [B] (the four tensors
 * odt.py:64-66 reads as output_3/2/1/0) and the selected anchor index i32 [B,max_det]. */
int vbt_postprocess_q8(const vbt_model* m, const int8_t* dev_cls, const int8_t* dev_box,
                       int B, float iou_threshold, int max_det, int min_score_q,
                       float* dev_boxes, float* dev_classes, float* dev_scores,
                       float* dev_count, int32_t* dev_index, void* stream);

/* ---- a5/a6 threshold filter + tracker input packing --------------------------------
 * replaces: the `scores[i] >= threshold` loop of odt.detect_objects (odt.py:68-75) and
 * odt.results_to_sorttracker_inputs (odt.py:102-118).
 * dets: f64 [F,max_det,6] = xmin,ymin,xmax,ymax,score,0 ; det_count: i32 [F]. */
int vbt_pack_detections(const float* dev_boxes, const float* dev_scores,
                        const float* dev_count, int F, int max_det, float threshold,
                        double* dev_dets, int32_t* dev_det_count, void* stream);

/* ---- K7 tracker ---------------------------------------------------------------------
 * replaces: OCSort(max_age, asso_func="diou", iou_threshold) (track.py:157),
 * tracker.update(dets, []) (track.py:186-187) and the per-track row assembly
 * (track.py:189-234: kf.x[4:6], box centre, plate height/width).
 * One handle holds V independent videos (one warp each); state lives on the device. */
typedef struct vbt_tracker vbt_tracker;
typedef struct vbt_tracker_params {
  double det_thresh;    /* package default, see DESIGN.md (0.2) */
  double iou_threshold; /* 0.1 at track.py:157 */
  double inertia;       /* 0.2 */
  int max_age;          /* 30 at track.py:22,157 */
  int min_hits;         /* 3 */
  int delta_t;          /* 3 */
  int vdc_uses_class_column; /* 1: velocity-direction cost is multiplied by the class
                                column (0) as the 6-column package does */
} vbt_tracker_params;
int vbt_tracker_create(int V, int max_tracks, const vbt_tracker_params* p,
                       vbt_tracker** out);
void vbt_tracker_destroy(vbt_tracker* t);
int vbt_tracker_reset(vbt_tracker* t, void* stream); /* all videos back to frame 0 */
/* Steps every video v through its next n_frames[v] (<= F) frames.
 * dets: f64 [V,F,max_det,6]; det_count: i32 [V,F]; frame_no: i32 [V,F] the 1-based
 * frame_count of track.py:161; fps: f64 [V]; n_frames: i32 [V].
 * Frames with det_count == 0 do not step the tracker (track.py:180-181).
 * rows: f64 [V,row_cap,8] appended in the reference's append order; row_count: i32 [V]
 * is read-modify-written (caller zeroes it when a video starts).
 * emit_all != 0 additionally writes tracker.update()'s own return value for the LAST
 * stepped frame of each video to last_out f64 [V,max_det,9] (x1,y1,x2,y2,id,cls,conf,
 * kf_dx,kf_dy) and last_out_count i32 [V] (both may be NULL). */
int vbt_tracker_update(vbt_tracker* t, const double* dev_dets, const int32_t* dev_det_count,
                       const int32_t* dev_frame_no, const double* dev_fps,
                       const int32_t* dev_n_frames, int F, int max_det, double* dev_rows,
                       int32_t* dev_row_count, int row_cap, double* dev_last_out,
                       int32_t* dev_last_out_count, void* stream);
/* Optional: from now on vbt_tracker_update also writes, for every appended row, the tracker
 * output box and confidence the reference unpacks at track.py:190 for its overlay
 * (draw_bounding_box, track.py:28-49): f64 [V,row_cap,VBT_ROW_DETAIL_COLS] = (xmin,ymin,xmax,ymax,
 * score), same row index as `rows`.  NULL switches it off. */
#define VBT_ROW_DETAIL_COLS 5
int vbt_tracker_row_details(vbt_tracker* t, double* dev_row_details);
/* status i32 [V]: 0 ok, VBT_ECAPACITY if tracks/rows overflowed.  Synchronises. */
int vbt_tracker_status(vbt_tracker* t, int32_t* host_status, void* stream);
/* copies track.kf.x (7 doubles) + id + time_since_update of every live track of video
 * v, list order, to host: f64 [max_tracks,9]; returns the live count.  Synchronises.
 * (the reference reads tracker.trackers[i].id / .kf.x at track.py:194-199) */
int vbt_tracker_peek(vbt_tracker* t, int v, double* host_tracks, void* stream);

/* ---- K8 velocity --------------------------------------------------------------------
 * replaces: plot.py:87-95 (select id, rolling(5)/expanding means) and plot.analyze_df
 * (plot.py:33-47) -> VelocityTracker.process_measurements / end_processing
 * (VelocityTracker.py:92-230), RunningAverage.update (RunningAverage.py:16-27),
 * Phase (Phase.py:6-40); also the per-id cumulative path of track.py:109-113.
 * L lanes, one thread each; a lane consumes the rows of ONE id of ONE row table. */
typedef struct vbt_velocity vbt_velocity;
int vbt_velocity_create(int L, int path_cap, int phase_cap, vbt_velocity** out);
void vbt_velocity_destroy(vbt_velocity* v);
int vbt_velocity_reset(vbt_velocity* v, void* stream);
/* Per lane l: table = lane_table[l] (index into [T,row_cap,8] `rows`), id filter
 * lane_id[l] (<0 accepts every row), consumes rows [lane_begin[l], row_count[table]).
 * smooth != 0 applies plot.py:90-95 before the state machine.  finish != 0 also runs
 * end_processing().  lane_begin is advanced by the kernel (streaming). */
int vbt_velocity_update(vbt_velocity* v, const double* dev_rows, const int32_t* dev_row_count,
                        int row_cap, const int32_t* dev_lane_table, const int32_t* dev_lane_id,
                        int32_t* dev_lane_begin, int L, double plate_diameter,
                        double diff_threshold, double min_distance, int smooth, int finish,
                        void* stream);
/* Results, device -> host, synchronises: phases f64 [L,phase_cap,6], phase_count i32 [L],
 * lane_state f64 [L,8] = current_phase, max_y_diff (NaN when unset), rows consumed,
 * cumulative path length (track.py:109-113), status, y_prev, neg_cnt, pos_cnt. */
int vbt_velocity_read(vbt_velocity* v, double* host_phases, int32_t* host_phase_count,
                      double* host_lane_state, void* stream);

/* RunningAverage.update over a sequence (RunningAverage.py:16-27): state f64
 * [window+3] on the device (ring..., total, count, head), values/out f64 [n]. */
int vbt_running_average(double* dev_state, int window, const double* dev_values, int n,
                        double* dev_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBT_B200_H */
