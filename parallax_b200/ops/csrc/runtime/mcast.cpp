// NVLS multicast segments: a symmetric allocation that is additionally bound
// to an NVSwitch multicast object, so `multimem.ld_reduce` / `multimem.st` on
// the multicast address reduce / broadcast inside the switch.
//
// Built on the CUDA VMM driver API (entry points resolved at run time, no
// libcuda link): cuMemCreate (POSIX-fd shareable) → cuMulticastCreate on
// rank 0 → fd exchange over abstract Unix sockets (SCM_RIGHTS) →
// cuMulticastAddDevice / BindMem → map unicast + multicast views.
// The reference has no counterpart (NCCL 2.2 ring over PCIe/IB,
// horovod/common/ops/nccl_operations.cc:60-109); this is the "reduce inside the
// switch" path SURVEY §5.8 calls for on large buckets.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>
#include <vector>

namespace {

thread_local char g_err[256];
#define FAIL(...) do { snprintf(g_err, sizeof(g_err), __VA_ARGS__); return -1; } while (0)

template <typename F> F drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<F>(p);
}
#define DRV(name) static auto p_##name = drv<decltype(&name)>(#name); \
  if (!p_##name) FAIL("driver entry point %s unavailable", #name)
#define CK(call, what) do { CUresult _r = (call); if (_r != CUDA_SUCCESS) \
  FAIL("%s failed: CUresult %d", what, (int)_r); } while (0)

struct Seg {
  CUmemGenericAllocationHandle local = 0, mc = 0;
  CUdeviceptr uc_va = 0, mc_va = 0;
  size_t size = 0;
  int dev = 0;
};

int listen_fd = -1;
std::string sock_name(const char* job, int rank) {
  return std::string("parallax_mc_") + job + "_" + std::to_string(rank);
}
void fill_addr(sockaddr_un& a, socklen_t& len, const std::string& name) {
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  a.sun_path[0] = '\0';                                   // abstract namespace
  strncpy(a.sun_path + 1, name.c_str(), sizeof(a.sun_path) - 2);
  len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

}  // namespace

extern "C" {

const char* px_mc_last_error() { return g_err; }

// 1 if the current device supports NVSwitch multicast
int px_mc_supported() {
  DRV(cuDeviceGet); DRV(cuDeviceGetAttribute);
  int dev_ord = 0;
  cudaGetDevice(&dev_ord);
  CUdevice dev;
  if (p_cuDeviceGet(&dev, dev_ord) != CUDA_SUCCESS) return 0;
  int v = 0;
  if (p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS)
    return 0;
  return v;
}

// ---- fd exchange ----------------------------------------------------------------
int px_fd_listen(const char* job, int rank) {
  if (listen_fd >= 0) close(listen_fd);
  listen_fd = socket(AF_UNIX, SOCK_DGRAM, 0);
  if (listen_fd < 0) FAIL("socket: %s", strerror(errno));
  sockaddr_un a; socklen_t len;
  fill_addr(a, len, sock_name(job, rank));
  if (bind(listen_fd, (sockaddr*)&a, len) < 0) FAIL("bind: %s", strerror(errno));
  timeval tv{60, 0};                       // never block a rank forever on a lost fd
  setsockopt(listen_fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  return 0;
}

int px_fd_send(const char* job, int dst_rank, int fd, int tag) {
  int s = socket(AF_UNIX, SOCK_DGRAM, 0);
  if (s < 0) FAIL("socket: %s", strerror(errno));
  sockaddr_un a; socklen_t len;
  fill_addr(a, len, sock_name(job, dst_rank));
  msghdr msg{}; iovec io{};
  int payload = tag;
  io.iov_base = &payload; io.iov_len = sizeof(payload);
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  msg.msg_name = &a; msg.msg_namelen = len;
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  int rc = -1;
  for (int tries = 0; tries < 200 && rc < 0; ++tries) {     // receiver may not be bound yet
    rc = (int)sendmsg(s, &msg, 0);
    if (rc < 0) usleep(20000);
  }
  close(s);
  if (rc < 0) FAIL("sendmsg: %s", strerror(errno));
  return 0;
}

int px_fd_recv(int* fd_out, int* tag_out) {
  if (listen_fd < 0) FAIL("px_fd_listen not called");
  msghdr msg{}; iovec io{};
  int payload = 0;
  io.iov_base = &payload; io.iov_len = sizeof(payload);
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(listen_fd, &msg, 0) < 0) FAIL("recvmsg: %s", strerror(errno));
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_type != SCM_RIGHTS) FAIL("no fd in message");
  memcpy(fd_out, CMSG_DATA(c), sizeof(int));
  *tag_out = payload;
  return 0;
}

void px_fd_close_listener() { if (listen_fd >= 0) { close(listen_fd); listen_fd = -1; } }
void px_fd_close(int fd) { close(fd); }

// ---- segment construction (each step is collective; Python sequences them) ----------
// granularity-rounded size for a multicast segment of `bytes` over `world` devices
long long px_mc_round_size(size_t bytes, int world) {
  DRV(cuMulticastGetGranularity);
  CUmulticastObjectProp prop{};
  prop.numDevices = (unsigned)world;
  prop.size = bytes;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g = 0;
  CK(p_cuMulticastGetGranularity(&g, &prop, CU_MULTICAST_GRANULARITY_MINIMUM), "granularity");
  if (g < (2u << 20)) g = 2u << 20;                       // VMM allocation granularity
  return (long long)((bytes + g - 1) / g * g);
}

// local physical allocation + unicast mapping; returns opaque segment handle
int px_mc_seg_create(size_t size, void** seg_out, void** uc_ptr_out) {
  DRV(cuMemCreate); DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemSetAccess);
  int dev = 0;
  cudaGetDevice(&dev);
  Seg* s = new Seg();
  s->size = size; s->dev = dev;
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CK(p_cuMemCreate(&s->local, size, &prop, 0), "cuMemCreate");
  CK(p_cuMemAddressReserve(&s->uc_va, size, 0, 0, 0), "cuMemAddressReserve");
  CK(p_cuMemMap(s->uc_va, size, 0, s->local, 0), "cuMemMap");
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CK(p_cuMemSetAccess(s->uc_va, size, &acc, 1), "cuMemSetAccess");
  cudaMemset((void*)s->uc_va, 0, size);
  *seg_out = s; *uc_ptr_out = (void*)s->uc_va;
  return 0;
}

// rank 0: create the multicast object and export it as an fd
int px_mc_create_export(void* seg, int world, int* fd_out) {
  DRV(cuMulticastCreate); DRV(cuMemExportToShareableHandle);
  Seg* s = (Seg*)seg;
  CUmulticastObjectProp prop{};
  prop.numDevices = (unsigned)world; prop.size = s->size;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CK(p_cuMulticastCreate(&s->mc, &prop), "cuMulticastCreate");
  int fd = -1;
  CK(p_cuMemExportToShareableHandle(&fd, s->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
     "export multicast handle");
  *fd_out = fd;
  return 0;
}

// other ranks: import the multicast object from the received fd
int px_mc_import(void* seg, int fd) {
  DRV(cuMemImportFromShareableHandle);
  Seg* s = (Seg*)seg;
  CK(p_cuMemImportFromShareableHandle(&s->mc, (void*)(uintptr_t)fd,
                                      CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
     "import multicast handle");
  return 0;
}

int px_mc_add_device(void* seg) {
  DRV(cuMulticastAddDevice); DRV(cuDeviceGet);
  Seg* s = (Seg*)seg;
  CUdevice dev;
  CK(p_cuDeviceGet(&dev, s->dev), "cuDeviceGet");
  CK(p_cuMulticastAddDevice(s->mc, dev), "cuMulticastAddDevice");
  return 0;
}

// after ALL ranks added their device: bind local memory and map the multicast view
int px_mc_bind_map(void* seg, void** mc_ptr_out) {
  DRV(cuMulticastBindMem); DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemSetAccess);
  Seg* s = (Seg*)seg;
  CK(p_cuMulticastBindMem(s->mc, 0, s->local, 0, s->size, 0), "cuMulticastBindMem");
  CK(p_cuMemAddressReserve(&s->mc_va, s->size, 0, 0, 0), "reserve mc va");
  CK(p_cuMemMap(s->mc_va, s->size, 0, s->mc, 0), "map mc");
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = s->dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CK(p_cuMemSetAccess(s->mc_va, s->size, &acc, 1), "set access mc");
  *mc_ptr_out = (void*)s->mc_va;
  return 0;
}

int px_mc_seg_destroy(void* seg) {
  DRV(cuMemUnmap); DRV(cuMemRelease); DRV(cuMemAddressFree);
  Seg* s = (Seg*)seg;
  if (s->mc_va) { p_cuMemUnmap(s->mc_va, s->size); p_cuMemAddressFree(s->mc_va, s->size); }
  if (s->mc) p_cuMemRelease(s->mc);
  if (s->uc_va) { p_cuMemUnmap(s->uc_va, s->size); p_cuMemAddressFree(s->uc_va, s->size); }
  if (s->local) p_cuMemRelease(s->local);
  delete s;
  return 0;
}

}  // extern "C"
