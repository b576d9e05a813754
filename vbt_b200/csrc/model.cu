// Model handle: parses the layer-program blob (vbt_b200/effdet.py writes it), uploads the
// weight / anchor / LUT data section once and keeps the op table on the host.
// replaces: tflite_runtime.Interpreter(model_path) + allocate_tensors() (track.py:93-94).
#include <string.h>

#include "model.cuh"

using namespace vbt;

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
  if (hdr.n_ops < 0 || hdr.n_tensors < 0 || hdr.data_offset < (int64_t)table ||
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
  m->dev_anchors = reinterpret_cast<const float*>(m->dev_data + hdr.anchors_off);
  m->dev_exp_lut = reinterpret_cast<const float*>(m->dev_data + hdr.exp_lut_off);
  m->kernels_per_detect = hdr.n_ops;
  *out = m;
  return VBT_OK;
}

void vbt_model_destroy(vbt_model* m) {
  if (!m) return;
  if (m->dev_data) cudaFree(m->dev_data);
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

}  // extern "C"
