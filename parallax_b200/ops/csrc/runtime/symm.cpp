// Symmetric heap: device allocations that every rank can address.
//
// Replaces the reference's transport layer — NCCL communicators bootstrapped
// over MPI (horovod/common/ops/nccl_operations.cc:111-153), gRPC RecvTensor
// with host staging (tensorflow/core/distributed_runtime/rpc/
// grpc_worker_service.cc:427-500) and the verbs/GDR transports
// (tensorflow/contrib/{verbs,gdr}) — with plain peer-addressable memory:
// each rank cudaMalloc's its segment, exports a CUDA IPC handle, and maps
// every peer's segment (peer access enabled lazily over NVLink).  Handles are
// exchanged by the Python control plane (torch.distributed object gather).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace {
std::mutex g_mu;
std::unordered_map<void*, size_t> g_allocs;   // local segments
std::unordered_map<void*, int> g_mapped;      // peer mappings (ptr -> refcount)
thread_local char g_err[256];
int fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
}  // namespace

extern "C" {

const char* px_last_error() { return g_err; }

int px_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int px_set_device(int dev) {
  cudaError_t e = cudaSetDevice(dev);
  return e == cudaSuccess ? 0 : fail(e, "cudaSetDevice");
}

// Allocate a zero-filled segment on the current device.
int px_symm_alloc(size_t bytes, void** out) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(e, "cudaMalloc");
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) { cudaFree(p); return fail(e, "cudaMemset"); }
  std::lock_guard<std::mutex> lk(g_mu);
  g_allocs[p] = bytes;
  *out = p;
  return 0;
}

int px_symm_free(void* p) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return -1;
    g_allocs.erase(it);
  }
  cudaError_t e = cudaFree(p);
  return e == cudaSuccess ? 0 : fail(e, "cudaFree");
}

// 64-byte opaque handle for a segment allocated with px_symm_alloc.
int px_ipc_export(void* p, unsigned char* out64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return fail(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(h) == 64, "handle size");
  memcpy(out64, &h, 64);
  return 0;
}

// Map a peer's segment into this process (enables peer access lazily).
int px_ipc_import(const unsigned char* in64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(e, "cudaIpcOpenMemHandle");
  std::lock_guard<std::mutex> lk(g_mu);
  g_mapped[p] += 1;
  *out = p;
  return 0;
}

int px_ipc_close(void* p) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_mapped.find(p);
    if (it == g_mapped.end()) return -1;
    if (--it->second > 0) return 0;
    g_mapped.erase(it);
  }
  cudaError_t e = cudaIpcCloseMemHandle(p);
  return e == cudaSuccess ? 0 : fail(e, "cudaIpcCloseMemHandle");
}

// Same-process multi-device worlds (tests, single-process launch): direct
// peer access instead of IPC.
int px_enable_peer(int peer_dev) {
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur == peer_dev) return 0;
  int can = 0;
  cudaDeviceCanAccessPeer(&can, cur, peer_dev);
  if (!can) { snprintf(g_err, sizeof(g_err), "no P2P %d->%d", cur, peer_dev); return -1; }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  return e == cudaSuccess ? 0 : fail(e, "cudaDeviceEnablePeerAccess");
}

int px_memcpy_h2d_async(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
  return e == cudaSuccess ? 0 : fail(e, "cudaMemcpyAsync");
}

int px_memset_async(void* dst, int value, size_t bytes, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(dst, value, bytes, s);
  return e == cudaSuccess ? 0 : fail(e, "cudaMemsetAsync");
}

size_t px_symm_live_bytes() {
  std::lock_guard<std::mutex> lk(g_mu);
  size_t t = 0;
  for (auto& kv : g_allocs) t += kv.second;
  return t;
}

}  // extern "C"
