// Symmetric NVLink peer-memory heap (the replacement for the reference's host
// numpy staging buffers + MPI).
//
// Every rank allocates the SAME number of bytes with the CUDA VMM driver API
// (cuMemCreate, POSIX-fd shareable), exchanges the file descriptors with its
// peers over abstract Unix-domain datagram sockets (SCM_RIGHTS), maps every
// peer's allocation into its own address space (-> plain ld/st/atomics travel
// over NVLink 5 / NVSwitch), and — when the platform supports NVLS — binds all
// allocations to one multicast object so that a single `multimem.st` is
// replicated by the switch to every GPU and `multimem.ld_reduce` returns the
// in-switch sum.  A cudaIpc (legacy) mode is provided as a fallback for
// platforms without fd-exportable VMM allocations (no multicast there).
//
// The driver API is resolved at run time through cudaGetDriverEntryPoint so
// that this extension imports on machines without libcuda (the CPU build box).
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace atomo {

#define DRV_FN(name) static decltype(&name) p_##name = nullptr
DRV_FN(cuMemCreate);
DRV_FN(cuMemRelease);
DRV_FN(cuMemAddressReserve);
DRV_FN(cuMemAddressFree);
DRV_FN(cuMemMap);
DRV_FN(cuMemUnmap);
DRV_FN(cuMemSetAccess);
DRV_FN(cuMemGetAllocationGranularity);
DRV_FN(cuMemExportToShareableHandle);
DRV_FN(cuMemImportFromShareableHandle);
DRV_FN(cuMulticastCreate);
DRV_FN(cuMulticastAddDevice);
DRV_FN(cuMulticastBindMem);
DRV_FN(cuMulticastGetGranularity);
DRV_FN(cuMulticastUnbind);
DRV_FN(cuDeviceGetAttribute);
DRV_FN(cuGetErrorString);

static std::string g_err;

static bool load_driver() {
  static int state = 0;  // 0 unknown, 1 ok, -1 failed
  if (state != 0) return state > 0;
#define LOAD(name)                                                                                   \
  do {                                                                                               \
    void* fn = nullptr;                                                                              \
    cudaDriverEntryPointQueryResult qr;                                                              \
    if (cudaGetDriverEntryPoint(#name, &fn, cudaEnableDefault, &qr) != cudaSuccess || fn == nullptr) { \
      g_err = std::string("driver entry point not found: ") + #name;                                 \
      cudaGetLastError();                                                                            \
      state = -1;                                                                                    \
      return false;                                                                                  \
    }                                                                                                \
    p_##name = reinterpret_cast<decltype(p_##name)>(fn);                                             \
  } while (0)
  LOAD(cuMemCreate); LOAD(cuMemRelease); LOAD(cuMemAddressReserve); LOAD(cuMemAddressFree); LOAD(cuMemMap);
  LOAD(cuMemUnmap); LOAD(cuMemSetAccess); LOAD(cuMemGetAllocationGranularity);
  LOAD(cuMemExportToShareableHandle); LOAD(cuMemImportFromShareableHandle); LOAD(cuMulticastCreate);
  LOAD(cuMulticastAddDevice); LOAD(cuMulticastBindMem); LOAD(cuMulticastGetGranularity);
  LOAD(cuDeviceGetAttribute); LOAD(cuGetErrorString);
#undef LOAD
  {  // optional (only used at teardown): never let its absence disable the VMM/NVLS path
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMulticastUnbind", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn)
      p_cuMulticastUnbind = reinterpret_cast<decltype(p_cuMulticastUnbind)>(fn);
    else
      cudaGetLastError();
  }
  state = 1;
  return true;
}

static bool drv_ok(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return true;
  const char* s = nullptr;
  if (p_cuGetErrorString) p_cuGetErrorString(r, &s);
  g_err = std::string(what) + " failed: " + (s ? s : "unknown") + " (" + std::to_string((int)r) + ")";
  return false;
}

static size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

// ------------------------------------------------------------------------------------------------
// fd passing over abstract unix datagram sockets
// ------------------------------------------------------------------------------------------------
static void make_addr(const std::string& job, int rank, sockaddr_un* addr, socklen_t* len) {
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  std::string name = "atomo_b200." + job + "." + std::to_string(rank);
  // abstract namespace: leading NUL, no filesystem entry to clean up
  size_t n = name.size() < sizeof(addr->sun_path) - 2 ? name.size() : sizeof(addr->sun_path) - 2;
  memcpy(addr->sun_path + 1, name.data(), n);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

struct FdMsg {
  int kind;  // 0 = memory handle, 1 = multicast handle
  int src;
};

static bool send_fd(int sock, const std::string& job, int dst, int fd, FdMsg msg, double timeout_s) {
  sockaddr_un addr; socklen_t alen;
  make_addr(job, dst, &addr, &alen);
  struct iovec iov; iov.iov_base = &msg; iov.iov_len = sizeof(msg);
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  struct msghdr mh; memset(&mh, 0, sizeof(mh));
  mh.msg_name = &addr; mh.msg_namelen = alen; mh.msg_iov = &iov; mh.msg_iovlen = 1;
  mh.msg_control = ctrl; mh.msg_controllen = sizeof(ctrl);
  struct cmsghdr* cm = CMSG_FIRSTHDR(&mh);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
  for (;;) {
    if (sendmsg(sock, &mh, 0) >= 0) return true;
    if (errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN && errno != ENOBUFS) {
      g_err = std::string("sendmsg: ") + strerror(errno);
      return false;
    }
    struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
    if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > timeout_s) {
      g_err = "send_fd: peer socket never appeared";
      return false;
    }
    usleep(2000);  // the peer has not bound its socket yet
  }
}

static bool recv_fd(int sock, int* fd, FdMsg* msg, double timeout_s) {
  struct timeval tv; tv.tv_sec = (long)timeout_s; tv.tv_usec = 0;
  setsockopt(sock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  struct iovec iov; iov.iov_base = msg; iov.iov_len = sizeof(*msg);
  char ctrl[CMSG_SPACE(sizeof(int))];
  struct msghdr mh; memset(&mh, 0, sizeof(mh));
  mh.msg_iov = &iov; mh.msg_iovlen = 1; mh.msg_control = ctrl; mh.msg_controllen = sizeof(ctrl);
  ssize_t n = recvmsg(sock, &mh, 0);
  if (n < (ssize_t)sizeof(*msg)) { g_err = std::string("recvmsg: ") + strerror(errno); return false; }
  struct cmsghdr* cm = CMSG_FIRSTHDR(&mh);
  if (!cm || cm->cmsg_type != SCM_RIGHTS) { g_err = "recvmsg: no fd attached"; return false; }
  memcpy(fd, CMSG_DATA(cm), sizeof(int));
  return true;
}

// ------------------------------------------------------------------------------------------------
struct SymmHeap {
  int rank = 0, world = 1, device = 0;
  size_t bytes = 0;      // rounded allocation size
  std::string job;
  std::string mode;      // "vmm" | "ipc" | "local"
  int sock = -1;
  CUmemGenericAllocationHandle handle = 0;
  std::vector<CUmemGenericAllocationHandle> peer_handles;
  std::vector<uint64_t> ptrs;  // [world] device pointers (ptrs[rank] = local)
  uint64_t mc_ptr = 0;
  CUmemGenericAllocationHandle mc_handle = 0;
  bool mc_added = false;
  size_t gran = 0, mc_gran = 0;
  void* ipc_base = nullptr;
  std::vector<int> pending_mc_fd;
};

static bool map_handle(CUmemGenericAllocationHandle h, size_t bytes, size_t gran, int device, uint64_t* out) {
  CUdeviceptr p = 0;
  if (!drv_ok(p_cuMemAddressReserve(&p, bytes, gran, 0, 0), "cuMemAddressReserve")) return false;
  if (!drv_ok(p_cuMemMap(p, bytes, 0, h, 0), "cuMemMap")) return false;
  CUmemAccessDesc acc; memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (!drv_ok(p_cuMemSetAccess(p, bytes, &acc, 1), "cuMemSetAccess")) return false;
  *out = (uint64_t)p;
  return true;
}

extern "C" {

const char* atomo_heap_last_error() { return g_err.c_str(); }

// returns 1 when NVLS multicast objects are supported on `device`
int atomo_heap_multicast_supported(int device) {
  if (!load_driver()) return 0;
  int v = 0;
  if (p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS) return 0;
  return v;
}
int atomo_heap_posix_fd_supported(int device) {
  if (!load_driver()) return 0;
  int v = 0;
  if (p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, device) !=
      CUDA_SUCCESS)
    return 0;
  return v;
}

// phase 1 (vmm mode): allocate, exchange fds with every peer, map peers.  Collective.
void* atomo_heap_create_vmm(int rank, int world, int device, size_t bytes, const char* job, int want_mc,
                            double timeout_s) {
  g_err.clear();
  if (!load_driver()) return nullptr;
  cudaSetDevice(device);
  cudaFree(0);  // make sure the primary context exists
  SymmHeap* h = new SymmHeap();
  h->rank = rank; h->world = world; h->device = device; h->job = job; h->mode = "vmm";

  CUmemAllocationProp prop; memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  if (!drv_ok(p_cuMemGetAllocationGranularity(&h->gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
              "cuMemGetAllocationGranularity")) { delete h; return nullptr; }
  size_t g = h->gran;
  if (want_mc && world > 1) {
    CUmulticastObjectProp mp; memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)world; mp.size = round_up(bytes, g); mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (p_cuMulticastGetGranularity(&h->mc_gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
        h->mc_gran > g)
      g = h->mc_gran;
  }
  h->bytes = round_up(bytes, g);
  if (!drv_ok(p_cuMemCreate(&h->handle, h->bytes, &prop, 0), "cuMemCreate")) { delete h; return nullptr; }
  h->ptrs.assign(world, 0);
  h->peer_handles.assign(world, 0);
  if (!map_handle(h->handle, h->bytes, g, device, &h->ptrs[rank])) { delete h; return nullptr; }
  cudaMemset((void*)h->ptrs[rank], 0, h->bytes);
  cudaDeviceSynchronize();
  if (world == 1) return h;

  // ---- fd exchange --------------------------------------------------------------------------
  h->sock = socket(AF_UNIX, SOCK_DGRAM, 0);
  sockaddr_un addr; socklen_t alen;
  make_addr(h->job, rank, &addr, &alen);
  if (h->sock < 0 || bind(h->sock, (sockaddr*)&addr, alen) != 0) {
    g_err = std::string("bind unix socket: ") + strerror(errno);
    delete h; return nullptr;
  }
  int fd = -1;
  if (!drv_ok(p_cuMemExportToShareableHandle(&fd, h->handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
              "cuMemExportToShareableHandle")) { delete h; return nullptr; }
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    if (!send_fd(h->sock, h->job, p, fd, FdMsg{0, rank}, timeout_s)) { delete h; return nullptr; }
  }
  close(fd);
  int got = 0;
  while (got < world - 1) {
    int pfd = -1; FdMsg m;
    if (!recv_fd(h->sock, &pfd, &m, timeout_s)) { delete h; return nullptr; }
    if (m.kind == 1) { h->pending_mc_fd.push_back(pfd); continue; }  // early multicast fd from rank 0
    CUmemGenericAllocationHandle ph;
    if (!drv_ok(p_cuMemImportFromShareableHandle(&ph, (void*)(uintptr_t)pfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
                "cuMemImportFromShareableHandle")) { delete h; return nullptr; }
    close(pfd);
    h->peer_handles[m.src] = ph;
    if (!map_handle(ph, h->bytes, g, device, &h->ptrs[m.src])) { delete h; return nullptr; }
    ++got;
  }
  return h;
}

// phase 2a: rank 0 creates the multicast object and ships its fd; every rank adds its device.
// Returns 1 on success, 0 on (recoverable) failure -> caller runs without multicast.
int atomo_heap_mc_phase_a(void* hp, double timeout_s) {
  SymmHeap* h = (SymmHeap*)hp;
  if (h->world == 1 || h->mode != "vmm") return 0;
  if (!atomo_heap_multicast_supported(h->device)) { g_err = "multicast not supported on this device"; return 0; }
  if (h->rank == 0) {
    CUmulticastObjectProp mp; memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)h->world; mp.size = h->bytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (!drv_ok(p_cuMulticastCreate(&h->mc_handle, &mp), "cuMulticastCreate")) return 0;
    int fd = -1;
    if (!drv_ok(p_cuMemExportToShareableHandle(&fd, h->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
                "export multicast handle")) return 0;
    for (int p = 1; p < h->world; ++p)
      if (!send_fd(h->sock, h->job, p, fd, FdMsg{1, 0}, timeout_s)) return 0;
    close(fd);
  } else {
    int fd = -1;
    if (!h->pending_mc_fd.empty()) { fd = h->pending_mc_fd.back(); h->pending_mc_fd.pop_back(); }
    else {
      FdMsg m;
      if (!recv_fd(h->sock, &fd, &m, timeout_s) || m.kind != 1) return 0;
    }
    if (!drv_ok(p_cuMemImportFromShareableHandle(&h->mc_handle, (void*)(uintptr_t)fd,
                                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
                "import multicast handle")) return 0;
    close(fd);
  }
  if (!drv_ok(p_cuMulticastAddDevice(h->mc_handle, h->device), "cuMulticastAddDevice")) return 0;
  h->mc_added = true;
  return 1;
}

// phase 2b (after a barrier: every device has been added): bind memory, map the multicast VA.
int atomo_heap_mc_phase_b(void* hp) {
  SymmHeap* h = (SymmHeap*)hp;
  if (!h->mc_added) return 0;
  if (!drv_ok(p_cuMulticastBindMem(h->mc_handle, 0, h->handle, 0, h->bytes, 0), "cuMulticastBindMem")) return 0;
  size_t g = h->mc_gran > h->gran ? h->mc_gran : h->gran;
  uint64_t p = 0;
  if (!map_handle(h->mc_handle, h->bytes, g, h->device, &p)) return 0;
  h->mc_ptr = p;
  return 1;
}

// ---- cudaIpc fallback -----------------------------------------------------------------------------
void* atomo_heap_create_ipc(int rank, int world, int device, size_t bytes, unsigned char* handle_out64) {
  g_err.clear();
  cudaSetDevice(device);
  SymmHeap* h = new SymmHeap();
  h->rank = rank; h->world = world; h->device = device; h->mode = world == 1 ? "local" : "ipc";
  h->bytes = round_up(bytes, 2u << 20);
  if (cudaMalloc(&h->ipc_base, h->bytes) != cudaSuccess) {
    g_err = std::string("cudaMalloc: ") + cudaGetErrorString(cudaGetLastError());
    delete h; return nullptr;
  }
  cudaMemset(h->ipc_base, 0, h->bytes);
  cudaDeviceSynchronize();
  h->ptrs.assign(world, 0);
  h->ptrs[rank] = (uint64_t)h->ipc_base;
  if (world > 1) {
    cudaIpcMemHandle_t ih;
    if (cudaIpcGetMemHandle(&ih, h->ipc_base) != cudaSuccess) {
      g_err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
      delete h; return nullptr;
    }
    memcpy(handle_out64, &ih, sizeof(ih));
  }
  return h;
}

int atomo_heap_open_ipc(void* hp, const unsigned char* all_handles /* world x 64 */) {
  SymmHeap* h = (SymmHeap*)hp;
  for (int p = 0; p < h->world; ++p) {
    if (p == h->rank) continue;
    cudaIpcMemHandle_t ih;
    memcpy(&ih, all_handles + 64 * p, sizeof(ih));
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      g_err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError());
      return 0;
    }
    h->ptrs[p] = (uint64_t)ptr;
  }
  return 1;
}

// ---- accessors --------------------------------------------------------------------------------------
uint64_t atomo_heap_ptr(void* hp, int rank) { return ((SymmHeap*)hp)->ptrs[rank]; }
uint64_t atomo_heap_mc_ptr(void* hp) { return ((SymmHeap*)hp)->mc_ptr; }
uint64_t atomo_heap_bytes(void* hp) { return ((SymmHeap*)hp)->bytes; }
const char* atomo_heap_mode(void* hp) { return ((SymmHeap*)hp)->mode.c_str(); }

void atomo_heap_destroy(void* hp) {
  SymmHeap* h = (SymmHeap*)hp;
  if (!h) return;
  cudaDeviceSynchronize();
  if (h->mode == "vmm") {
    if (h->mc_ptr) {
      p_cuMemUnmap((CUdeviceptr)h->mc_ptr, h->bytes);
      p_cuMemAddressFree((CUdeviceptr)h->mc_ptr, h->bytes);
      if (p_cuMulticastUnbind) p_cuMulticastUnbind(h->mc_handle, h->device, 0, h->bytes);  // release the switch binding
    }
    for (int p = 0; p < h->world; ++p) {
      if (h->ptrs[p]) { p_cuMemUnmap((CUdeviceptr)h->ptrs[p], h->bytes); p_cuMemAddressFree((CUdeviceptr)h->ptrs[p], h->bytes); }
      if (p != h->rank && h->peer_handles[p]) p_cuMemRelease(h->peer_handles[p]);
    }
    if (h->mc_handle) p_cuMemRelease(h->mc_handle);
    if (h->handle) p_cuMemRelease(h->handle);
    if (h->sock >= 0) close(h->sock);
  } else {
    for (int p = 0; p < h->world; ++p)
      if (p != h->rank && h->ptrs[p]) cudaIpcCloseMemHandle((void*)h->ptrs[p]);
    if (h->ipc_base) cudaFree(h->ipc_base);
  }
  delete h;
}

}  // extern "C"
}  // namespace atomo
