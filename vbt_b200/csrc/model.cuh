// Device-side description of one EfficientDet-Lite model: the layer program written by
// vbt_b200/effdet.py, its weights, anchors and output quantisation.
#pragma once
#include <map>
#include <set>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace vbt {

// Blob layout (little endian):
//   BlobHeader | OpRecord[n_ops] | data section (256-byte aligned offsets)
constexpr uint32_t kBlobMagic = 0x4d544256u;  // "VBTM"
constexpr int kBlobVersion = 11;

struct BlobHeader {
  uint32_t magic;
  int32_t version;
  int32_t input_size;      // S
  int32_t n_anchors;       // N
  int32_t n_anchors_pad;   // N rounded up to 16: row stride of the raw outputs
  int32_t n_classes;
  int32_t n_ops;
  int32_t n_tensors;
  int64_t ws_bytes_per_frame;   // activation workspace per frame
  int64_t data_offset;          // byte offset of the data section from blob start
  int64_t data_bytes;
  int64_t anchors_off;          // f32 [N,4] (ycentre,xcentre,h,w), offsets into data
  int64_t exp_lut_off;          // f32 [256]  exp(dequant(q)) for q = -128..127
  float box_scale;              // dequantisation of the raw box output
  int32_t box_zp;
  int32_t in_zp;                // uint8 input zero point (127)
  int32_t reserved[11];
};
static_assert(sizeof(BlobHeader) == 128, "blob header layout");

enum OpType : int32_t {
  OP_STEM = 1,      // 3x3 s2 conv on the uint8 input, ReLU6
  OP_PW = 2,        // 1x1 conv (+ optional fused quantised residual add)
  OP_DW = 3,        // depthwise kxk
  OP_ADD = 4,       // quantised n-ary add with resampling (BiFPN fusion), ReLU6
  OP_MAXPOOL = 5,   // 3x3 s2 SAME
  OP_LOGISTIC = 6,  // int8 LUT applied in place on the class output
};

enum Resample : int32_t { RS_NONE = 0, RS_UP_NEAREST = 1, RS_DOWN_MAXPOOL = 2 };

struct OpRecord {
  int32_t type;
  int32_t in[3];           // tensor ids (-1 unused); PW: in[1] = residual
  int32_t out;
  int32_t n_in;
  int32_t k, stride;
  int32_t cin, cout;       // logical channels
  int32_t cin_p, cout_p;   // physical (padded to 16) channels
  int32_t h_in, w_in, h_out, w_out;
  int32_t zp_in[3];
  int32_t zp_out;
  int32_t act_lo, act_hi;  // clamp of the int8 result
  int32_t pad_top, pad_left;
  int64_t w_off, bias_off, scale_off, lut_off;   // data-section byte offsets (-1 none); DW: lut_off =
                                                 // block-diagonal tensor-core weights (dw_umma.cu)
  int64_t out_elem_offset; // within-frame element offset (level offset for the heads)
  // OP_ADD / fused residual: integer rescale  out = clamp(((sum_i (x_i - zp_i)*mult_i)
  //                                             + round) >> shift) + zp_out)
  int32_t add_mult[3];
  int32_t add_shift;
  int32_t resample[3];
  int32_t in_h[3], in_w[3];
  // output placement: element offset of pixel p of frame b =
  //   out_off(b) + p * out_pix_stride ; out_off(b) = tensor base + b * out_batch_stride
  int32_t out_kind;        // 0 workspace tensor, 1 raw class output, 2 raw box output
  int32_t out_pix_stride;
  int32_t branch;          // 0: trunk (program order); k > 0: head chain k, independent of the others
  int32_t requant_fast;    // 1: |acc * mult| < 2^15 - 256 for every possible input (Requant::pack4)
  int32_t pw_dtype;        // OP_PW inside a fused head stage: 0 int8 (kind::i8), 1 bf16 (kind::f16)
  // OP_DW inside an MBConv block the fused kernel can take (csrc/mbconv_umma.cu): weight-image offset
  // in the data section / 256 + 1 (0: none), image stride in bytes, number of 32-channel chunks,
  // index of the run's first op relative to this one (-1: expand conv in front, 0: none)
  int32_t mb[4];
};
static_assert(sizeof(OpRecord) == 224, "op record layout");

struct TensorRecord {
  int64_t ws_offset;   // per-frame byte offset inside the workspace
  int32_t h, w, c, c_p;
};
static_assert(sizeof(TensorRecord) == 24, "tensor record layout");

}  // namespace vbt

struct vbt_model {
  vbt::BlobHeader hdr;
  std::vector<vbt::OpRecord> ops;
  std::vector<vbt::TensorRecord> tensors;
  uint8_t* dev_data = nullptr;     // the blob's data section on the device
  const float* dev_anchors = nullptr;
  const float* dev_exp_lut = nullptr;
  int kernels_per_detect = 0;
  int device = -1;
  // side streams for the independent head chains (fork after the trunk, join at the end)
  static constexpr int kMaxBranches = 16;
  cudaStream_t branch_stream[kMaxBranches] = {};
  cudaEvent_t fork_event[kMaxBranches] = {}, join_event[kMaxBranches] = {};
  int fork_after[kMaxBranches] = {};   // index of the trunk op that produces branch k's input
  int n_branches = 0;      // highest branch id used by the program
  // launch plan: fuse[i] = number of consecutive ops the launch starting at op i covers
  // ([ADD ->] DW3x3 -> PW on small maps run as one node_umma kernel), 0 for ops inside a group
  std::vector<int> fuse;
  std::vector<int> fuse_kind;   // per launch start: 0 single op / fused node (node_umma.cu), 1 MBConv block (mbconv_umma.cu)
  // work counters of the persistent kernels' dynamic tile schedulers: 16 ints per (workspace, op),
  // all zero between launches (the last CTA of a launch puts them back).  One block of n_ops * 16
  // per workspace the model has been run on, so concurrent lanes never share a counter.
  static constexpr int kCounterSlots = 8;
  int32_t* dev_counters = nullptr;            // [kCounterSlots][n_ops][16]
  std::vector<const void*> counter_owner;     // workspace pointer of each slot in use
  mutable int32_t* cur_counters = nullptr;    // the current op's 16 counters (set by run_ops)
  // CUDA graphs of the layer program, one per (buffers, batch) a caller has used twice
  struct GraphKey {
    const void *in, *ws, *cls, *box; int B;
    bool operator<(const GraphKey& o) const {
      return std::tie(in, ws, cls, box, B) < std::tie(o.in, o.ws, o.cls, o.box, o.B);
    }
  };
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::map<GraphKey, int> graph_kernels;   // kernel nodes of each captured graph
  int kernels_last_run = 0;                // kernels the last run_ops enqueued (fused runs count once)
  std::set<GraphKey> graph_seen;
  // optional per-op timing (bench.py): a ring of event sets, harvested lazily
  bool profile = false;
  static constexpr int kProfRing = 64;
  std::vector<cudaEvent_t> prof_events;   // [kProfRing][n_ops + 1]
  std::vector<double> prof_ms;            // [n_ops] accumulated
  int prof_head = 0, prof_pending = 0;    // sets recorded and not yet harvested
  long long prof_calls = 0;
};
