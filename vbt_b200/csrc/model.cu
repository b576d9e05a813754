// Model handle: parses the layer-program blob (vbt_b200/effdet.py writes it), uploads the
// weight / anchor / LUT data section once and keeps the op table on the host.
// replaces: tflite_runtime.Interpreter(model_path) + allocate_tensors() (track.py:93-94).
#include <string.h>

#include "model.cuh"

using namespace vbt;

static int harvest_one(vbt_model* m) {
  const int n = (int)m->ops.size();
  const int set = (m->prof_head - m->prof_pending + vbt_model::kProfRing * 2) % vbt_model::kProfRing;
  cudaEvent_t* ev = m->prof_events.data() + (size_t)set * (n + 1);
  VBT_CHECK_CUDA(cudaEventSynchronize(ev[n]));
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    VBT_CHECK_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
    m->prof_ms[i] += ms;
  }
  m->prof_pending -= 1;
  m->prof_calls += 1;
  return VBT_OK;
}

namespace vbt {
// called by vbt_detect: returns the event set to record into (nullptr = profiling off)
cudaEvent_t* profile_begin(vbt_model* m) {
  if (!m->profile) return nullptr;
  if (m->prof_pending == vbt_model::kProfRing && harvest_one(m) != VBT_OK) return nullptr;
  cudaEvent_t* ev = m->prof_events.data() + (size_t)m->prof_head * (m->ops.size() + 1);
  m->prof_head = (m->prof_head + 1) % vbt_model::kProfRing;
  m->prof_pending += 1;
  return ev;
}
}  // namespace vbt

extern "C" {

int vbt_model_create(const void* blob, size_t blob_bytes, vbt_model** out) {
  VBT_REQUIRE(blob && out, "vbt_model_create: null pointer");
  if (int rc = ensure_device()) return rc;
  if (blob_bytes < sizeof(BlobHeader)) {
    set_error("vbt_model_create: blob of %zu bytes is shorter than its header", blob_bytes);
    return VBT_EFORMAT;
  }
  BlobHeader hdr;
  memcpy(&hdr, blob, sizeof(hdr));
  if (hdr.magic != kBlobMagic || hdr.version != kBlobVersion) {
    set_error("vbt_model_create: bad magic/version (%08x, %d); expected (%08x, %d)", hdr.magic,
              hdr.version, kBlobMagic, kBlobVersion);
    return VBT_EFORMAT;
  }
  const size_t table = sizeof(BlobHeader) + sizeof(OpRecord) * (size_t)hdr.n_ops +
                       sizeof(TensorRecord) * (size_t)hdr.n_tensors;
  if (hdr.n_ops < 0 || hdr.n_tensors < 0 || hdr.data_bytes < 0 || hdr.ws_bytes_per_frame < 0 || hdr.data_offset < (int64_t)table ||
      (size_t)(hdr.data_offset + hdr.data_bytes) > blob_bytes || hdr.n_anchors <= 0 ||
      hdr.n_anchors_pad < hdr.n_anchors || hdr.n_anchors_pad % 16 != 0) {
    set_error("vbt_model_create: inconsistent blob header");
    return VBT_EFORMAT;
  }
  vbt_model* m = new vbt_model();
  m->hdr = hdr;
  const uint8_t* p = static_cast<const uint8_t*>(blob) + sizeof(BlobHeader);
  m->ops.resize(hdr.n_ops);
  if (hdr.n_ops) memcpy(m->ops.data(), p, sizeof(OpRecord) * (size_t)hdr.n_ops);
  p += sizeof(OpRecord) * (size_t)hdr.n_ops;
  m->tensors.resize(hdr.n_tensors);
  if (hdr.n_tensors) memcpy(m->tensors.data(), p, sizeof(TensorRecord) * (size_t)hdr.n_tensors);
  auto in_data = [&](int64_t off, size_t bytes) {
    return off >= 0 && (size_t)off + bytes <= (size_t)hdr.data_bytes && off % 16 == 0;
  };
  if (!in_data(hdr.anchors_off, sizeof(float) * 4 * (size_t)hdr.n_anchors) ||
      !in_data(hdr.exp_lut_off, sizeof(float) * 256)) {
    delete m;
    set_error("vbt_model_create: anchor / LUT offsets outside the data section");
    return VBT_EFORMAT;
  }
  for (const OpRecord& op : m->ops) {
    bool ok = ((op.out >= 0 && op.out < hdr.n_tensors) || (op.out == -1 && op.out_kind != 0)) &&
              op.n_in >= 1 && op.n_in <= 3;
    for (int i = 0; i < op.n_in && ok; ++i) ok = op.in[i] >= -1 && op.in[i] < hdr.n_tensors;
    if (!ok) {
      delete m;
      set_error("vbt_model_create: op references a tensor outside the table");
      return VBT_EFORMAT;
    }
  }
  // every data region an op names lies inside the data section, every tensor inside the per-frame workspace: a
  // truncated or malformed .vbtm / converted .tflite is refused here instead of sending a kernel out of bounds
  for (const OpRecord& op : m->ops) {
    bool ok = true;
    size_t wbytes = 0;
    if (op.type == OP_STEM) wbytes = (size_t)9 * op.cout_p * 4;                 // [9 taps][cout_p] words
    else if (op.type == OP_PW) wbytes = (size_t)op.cout_p * op.cin_p;
    else if (op.type == OP_DW) wbytes = (size_t)op.k * op.k * op.cout_p * 4;    // [k * k][c_p] words
    if (wbytes) {
      ok = op.cout_p > 0 && op.cout_p % 16 == 0 && op.k >= 1 && op.k <= 7 && in_data(op.w_off, wbytes) && in_data(op.bias_off, sizeof(int32_t) * (size_t)op.cout_p) &&
           in_data(op.scale_off, sizeof(float) * (size_t)op.cout_p);
      if (ok && op.type == OP_PW && op.lut_off >= 0) ok = in_data(op.lut_off, 256);
      if (ok && op.type == OP_DW && op.lut_off >= 0)
        ok = in_data(op.lut_off, (size_t)((op.cout_p / 16 + 1) / 2) * (size_t)(op.k * op.k) * 1024);
      if (ok && op.type == OP_DW && op.mb[0] > 0)
        ok = op.mb[1] > 0 && op.mb[2] > 0 && in_data((int64_t)(op.mb[0] - 1) * 256, (size_t)op.mb[1] * (size_t)op.mb[2]);
    }
    if (!ok) {
      delete m;
      set_error("vbt_model_create: op '%d' names weights / bias / multipliers outside the %lld-byte data section", (int)op.type,
                (long long)hdr.data_bytes);
      return VBT_EFORMAT;
    }
  }
  for (const TensorRecord& t : m->tensors) {
    const long long bytes = (long long)t.h * t.w * t.c_p;
    if (t.h < 0 || t.w < 0 || t.c_p < 0 || t.c_p % 16 != 0 || (t.ws_offset >= 0 && t.ws_offset + bytes > hdr.ws_bytes_per_frame)) {
      delete m;
      set_error("vbt_model_create: a %dx%dx%d tensor at workspace offset %lld does not fit the %lld bytes per frame", t.h, t.w, t.c_p,
                (long long)t.ws_offset, (long long)hdr.ws_bytes_per_frame);
      return VBT_EFORMAT;
    }
  }
  cudaGetDevice(&m->device);
  cudaError_t e = cudaMalloc(&m->dev_data, (size_t)hdr.data_bytes);
  if (e == cudaSuccess)
    e = cudaMemcpy(m->dev_data, static_cast<const uint8_t*>(blob) + hdr.data_offset,
                   (size_t)hdr.data_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("vbt_model_create: uploading %lld bytes failed: %s", (long long)hdr.data_bytes,
              cudaGetErrorString(e));
    if (m->dev_data) cudaFree(m->dev_data);
    delete m;
    return VBT_ECUDA;
  }
  if (hdr.n_ops > 0) {
    const size_t cbytes = sizeof(int32_t) * 16 * (size_t)hdr.n_ops * vbt_model::kCounterSlots;
    if (cudaMalloc(&m->dev_counters, cbytes) != cudaSuccess || cudaMemset(m->dev_counters, 0, cbytes) != cudaSuccess) {
      set_error("vbt_model_create: allocating the scheduler counters failed");
      cudaFree(m->dev_data);
      delete m;
      return VBT_ECUDA;
    }
  }
  m->dev_anchors = reinterpret_cast<const float*>(m->dev_data + hdr.anchors_off);
  m->dev_exp_lut = reinterpret_cast<const float*>(m->dev_data + hdr.exp_lut_off);
  m->kernels_per_detect = hdr.n_ops;
  for (const OpRecord& op : m->ops) {
    if (op.branch < 0 || op.branch > vbt_model::kMaxBranches) {
      vbt_model_destroy(m);
      set_error("vbt_model_create: op branch %d outside 0..%d", op.branch, vbt_model::kMaxBranches);
      return VBT_EFORMAT;
    }
    if (op.branch > m->n_branches) m->n_branches = op.branch;
  }
  for (int k = 0; k < m->n_branches; ++k) {
    VBT_CHECK_CUDA(cudaStreamCreateWithFlags(&m->branch_stream[k], cudaStreamNonBlocking));
    VBT_CHECK_CUDA(cudaEventCreateWithFlags(&m->join_event[k], cudaEventDisableTiming));
  }
  // ---- launch plan: which op runs of the program are served by one fused kernel -----------
  {
    const int n = (int)m->ops.size();
    std::vector<int> readers(hdr.n_tensors, 0);
    for (const OpRecord& op : m->ops)
      for (int i = 0; i < op.n_in; ++i)
        if (op.in[i] >= 0) readers[op.in[i]] += 1;
    static const int max_hw = [] { const char* e = getenv("VBT_FUSE_HW"); return e ? atoi(e) : 64; }();
    auto dw_ok = [&](const OpRecord& d) {
      return d.type == OP_DW && d.k == 3 && d.stride == 1 && d.cout_p <= 128 && d.lut_off >= 0 &&
             d.h_in <= max_hw && d.w_in <= max_hw && d.out >= 0 && readers[d.out] == 1;
    };
    auto pw_ok = [&](const OpRecord& p, const OpRecord& d) {
      return p.type == OP_PW && p.n_in == 1 && p.in[0] == d.out && p.branch == d.branch && p.cout_p <= 128 &&
             p.cout_p % 16 == 0 && p.cin_p == d.cout_p;
    };
    // one kernel's CTAs write the run's output while others still read its inputs: the two
    // must not share workspace memory (effdet.plan_workspace keeps them apart)
    auto overlaps = [&](int tin, const OpRecord& last) {
      if (last.out < 0 || tin < 0) return false;
      const TensorRecord& to = m->tensors[last.out];
      const TensorRecord& ti = m->tensors[tin];
      if (ti.ws_offset < 0) return false;               // the model input: caller's buffer
      const int64_t o0 = to.ws_offset, o1 = o0 + (int64_t)to.h * to.w * to.c_p;
      const int64_t i0 = ti.ws_offset, i1 = i0 + (int64_t)ti.h * ti.w * ti.c_p;
      return o0 < i1 && i0 < o1;
    };
    auto disjoint = [&](int first, int last) {           // no external input of ops[first..last) aliases the output
      for (int j = first; j < last; ++j)
        for (int i = 0; i < m->ops[j].n_in; ++i)
          if (overlaps(m->ops[j].in[i], m->ops[last])) return false;
      return true;
    };
    auto add_dw_pw = [&](int i) {                        // ops[i] = ADD feeding DW3x3 -> PW
      const OpRecord& o = m->ops[i];
      return o.type == OP_ADD && i + 2 < n && o.out >= 0 && readers[o.out] == 1 &&
             m->ops[i + 1].in[0] == o.out && m->ops[i + 1].branch == o.branch && dw_ok(m->ops[i + 1]) &&
             pw_ok(m->ops[i + 2], m->ops[i + 1]);
    };
    m->fuse.assign(n, 1);
    m->fuse_kind.assign(n, 0);
    // MBConv blocks: [expand PW ->] DW -> project PW as one mbconv_umma kernel
    static const bool mb_on = [] { const char* e = getenv("VBT_MBCONV"); return !(e && e[0] == '0'); }();
    std::vector<char> in_mb(n, 0);
    for (int i = 0; mb_on && i < n; ++i) {
      const OpRecord& d = m->ops[i];
      if (d.type != OP_DW || d.mb[0] <= 0) continue;
      const int first = i + d.mb[3], last = i + 1;
      if (first < 0 || first > i || last >= n || d.mb[3] < -1) continue;
      const OpRecord& p = m->ops[last];
      bool ok = p.type == OP_PW && p.in[0] == d.out && p.out_kind == 0 && d.out >= 0 && readers[d.out] == 1 &&
                p.branch == d.branch && p.cin_p == d.cout_p &&
                (size_t)(d.mb[0] - 1) * 256 + (size_t)d.mb[1] * d.mb[2] <= (size_t)hdr.data_bytes;
      if (ok && first < i) {
        const OpRecord& e = m->ops[first];
        // the expand stage is a 1x1 conv -- or the network's stem (3x3 s2 on the uint8 input, run as a
        // K = 27 GEMM over im2col rows)
        const bool pw_e = e.type == OP_PW && e.n_in == 1 && e.out_kind == 0 && (p.n_in == 1 || p.in[1] == e.in[0]);
        const bool stem_e = e.type == OP_STEM && e.k == 3 && e.stride == 2 && e.cout_p <= 32 && p.n_in == 1 &&
                            e.in[0] >= 0 && m->tensors[e.in[0]].ws_offset < 0;
        ok = (pw_e || stem_e) && e.out >= 0 && d.in[0] == e.out && readers[e.out] == 1 && e.cout_p == d.cout_p;
      } else if (ok) {
        ok = p.n_in == 1;
      }
      if (!ok || !disjoint(first, last)) continue;
      m->fuse[first] = last - first + 1;
      m->fuse_kind[first] = 1;
      for (int j = first + 1; j <= last; ++j) m->fuse[j] = 0;
      for (int j = first; j <= last; ++j) in_mb[j] = 1;
    }
    for (int i = 0; i < n;) {
      const OpRecord& o = m->ops[i];
      if (in_mb[i]) { i += 1; continue; }
      if (max_hw > 0 && o.type == OP_ADD && o.n_in == 2 && o.out >= 0 && readers[o.out] == 1 && i + 3 < n &&
          m->ops[i + 1].type == OP_ADD && m->ops[i + 1].n_in == 2 && m->ops[i + 1].branch == o.branch &&
          (m->ops[i + 1].in[0] == o.out || m->ops[i + 1].in[1] == o.out) && add_dw_pw(i + 1) &&
          disjoint(i, i + 3)) {
        m->fuse[i] = 4; m->fuse[i + 1] = m->fuse[i + 2] = m->fuse[i + 3] = 0;
        i += 4;
      } else if (max_hw > 0 && add_dw_pw(i) && disjoint(i, i + 2)) {
        m->fuse[i] = 3; m->fuse[i + 1] = m->fuse[i + 2] = 0;
        i += 3;
      } else if (max_hw > 0 && i + 1 < n && dw_ok(o) && pw_ok(m->ops[i + 1], o) && disjoint(i, i + 1)) {
        m->fuse[i] = 2; m->fuse[i + 1] = 0;
        i += 2;
      } else {
        i += 1;
      }
    }
    m->kernels_per_detect = 0;
    for (int f : m->fuse) m->kernels_per_detect += f > 0;
  }
  // a branch may start as soon as the trunk op that writes its input tensor is enqueued
  for (int k = 0; k < m->n_branches; ++k) {
    VBT_CHECK_CUDA(cudaEventCreateWithFlags(&m->fork_event[k], cudaEventDisableTiming));
    int first = -1, last_trunk = -1;
    for (int i = 0; i < (int)m->ops.size(); ++i) {
      if (m->ops[i].branch == 0) last_trunk = i;
      if (m->ops[i].branch == k + 1 && first < 0) first = i;
    }
    m->fork_after[k] = last_trunk;
    if (first >= 0)
      for (int i = 0; i < (int)m->ops.size(); ++i)
        if (m->ops[i].branch == 0 && m->ops[i].out == m->ops[first].in[0]) m->fork_after[k] = i;
  }
  *out = m;
  return VBT_OK;
}

void vbt_model_destroy(vbt_model* m) {
  if (!m) return;
  for (cudaEvent_t e : m->prof_events) cudaEventDestroy(e);
  for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);
  for (int k = 0; k < vbt_model::kMaxBranches; ++k) {
    if (m->branch_stream[k]) cudaStreamDestroy(m->branch_stream[k]);
    if (m->join_event[k]) cudaEventDestroy(m->join_event[k]);
  }
  for (int k = 0; k < vbt_model::kMaxBranches; ++k)
    if (m->fork_event[k]) cudaEventDestroy(m->fork_event[k]);
  if (m->dev_data) cudaFree(m->dev_data);
  if (m->dev_counters) cudaFree(m->dev_counters);
  delete m;
}

int vbt_model_info(const vbt_model* m, long long info[8]) {
  VBT_REQUIRE(m && info, "vbt_model_info: null pointer");
  info[0] = m->hdr.input_size;
  info[1] = m->hdr.n_anchors;
  info[2] = m->hdr.ws_bytes_per_frame;
  info[3] = m->hdr.n_ops;
  info[4] = m->hdr.n_classes;
  info[5] = m->kernels_per_detect;
  info[6] = m->hdr.n_anchors_pad;
  info[7] = 0;
  return VBT_OK;
}

int vbt_model_plan(const vbt_model* m, int32_t* host_group_len) {
  VBT_REQUIRE(m && host_group_len, "vbt_model_plan: null pointer");
  for (size_t i = 0; i < m->ops.size(); ++i) host_group_len[i] = m->fuse[i];
  return VBT_OK;
}

int vbt_model_plan_kinds(const vbt_model* m, int32_t* host_kind) {
  VBT_REQUIRE(m && host_kind, "vbt_model_plan_kinds: null pointer");
  for (size_t i = 0; i < m->ops.size(); ++i) host_kind[i] = m->fuse_kind[i];
  return VBT_OK;
}

int vbt_model_profile(vbt_model* m, int enable) {
  VBT_REQUIRE(m, "vbt_model_profile: null handle");
  if (enable && m->prof_events.empty()) {
    const size_t n = (size_t)vbt_model::kProfRing * (m->ops.size() + 1);
    m->prof_events.resize(n);
    for (size_t i = 0; i < n; ++i) VBT_CHECK_CUDA(cudaEventCreate(&m->prof_events[i]));
    m->prof_ms.assign(m->ops.size(), 0.0);
  }
  if (enable && !m->profile) {
    m->prof_ms.assign(m->ops.size(), 0.0);
    m->prof_calls = 0; m->prof_pending = 0; m->prof_head = 0;
  }
  m->profile = enable != 0;
  return VBT_OK;
}

int vbt_model_op_times(vbt_model* m, double* host_ms, long long* calls) {
  VBT_REQUIRE(m && host_ms && calls, "vbt_model_op_times: null pointer");
  while (m->prof_pending > 0)
    if (int rc = harvest_one(m)) return rc;
  for (size_t i = 0; i < m->ops.size(); ++i) host_ms[i] = i < m->prof_ms.size() ? m->prof_ms[i] : 0.0;
  *calls = m->prof_calls;
  return VBT_OK;
}

}  // extern "C"
