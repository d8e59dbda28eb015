// C-ABI plumbing: error reporting, weight blobs, device query.
#include <stdarg.h>

#include <string>

#include "common.cuh"

namespace t2p {
static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
}  // namespace t2p

extern "C" {

int t2p_version(void) { return 100; }

const char* t2p_last_error(void) { return t2p::g_last_error.c_str(); }

int t2p_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  T2P_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  T2P_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  T2P_REQUIRE(prop.major == 10, T2P_ERR_UNSUPPORTED, "device %s has compute capability %d.%d; this library is sm_100a only",
              prop.name, prop.major, prop.minor);
  return T2P_OK;
}

int t2p_weights_create(const float* h_blob, size_t n_floats, t2p_weights** out) {
  T2P_REQUIRE(h_blob != nullptr && out != nullptr && n_floats > 0, T2P_ERR_INVALID, "weights_create: null/empty blob");
  t2p_weights* w = new t2p_weights();
  w->n_floats = n_floats;
  w->d_blob = nullptr;
  cudaError_t e = cudaMalloc(&w->d_blob, n_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(w->d_blob, h_blob, n_floats * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    t2p::set_error("weights_create: %s", cudaGetErrorString(e));
    if (w->d_blob) cudaFree(w->d_blob);
    delete w;
    return T2P_ERR_CUDA;
  }
  *out = w;
  return T2P_OK;
}

int t2p_weights_destroy(t2p_weights* w) {
  if (w == nullptr) return T2P_OK;
  if (w->d_blob) cudaFree(w->d_blob);
  delete w;
  return T2P_OK;
}

const float* t2p_weights_device_ptr(const t2p_weights* w) { return w ? w->d_blob : nullptr; }

}  // extern "C"
