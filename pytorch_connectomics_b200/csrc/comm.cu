// pcb200 — the exchange steps of the path behind the C ABI (pcb_comm_init, pcb_grad_allreduce,
// pcb_sw_exchange_overlap), so that a host that is not Python/torch.distributed can run the data-parallel step and the
// z-slab sharded tile loop.  Reference: the gradient mean of Lightning's DDPStrategy (connectomics/training/lightning/
// trainer.py:231-256) and the accumulator reduction of the lazy path (connectomics/inference/lazy_distributed.py:78-169, here
// replaced by the neighbour exchange of inference/sharded.py).
//
// NCCL is bound at RUN time (dlopen + dlsym), not at link time: the process usually already holds the copy torch bundles
// (same SONAME libnccl.so.2 — RTLD_NOLOAD finds it), and a second, different NCCL linked into this library would shadow
// its symbols.  Nothing here runs unless pcb_comm_* is called; without an NCCL in the process or on the loader path the calls
// fail loudly (PCB_ERR_UNSUPPORTED + message), they never fall back to a host path.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <nccl.h>

#include <mutex>

#include "../../include/pcb200.h"
#include "pcb_common.cuh"

namespace pcb {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  bool ok = false;
  char why[256] = "";
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
  NcclApi& n = g_nccl;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {                       // the copy already in the process (torch's) wins
    n.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
    if (n.handle) break;
  }
  for (int i = 0; !n.handle && i < 2; ++i) n.handle = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
  if (!n.handle) {
    snprintf(n.why, sizeof(n.why), "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    return;
  }
#define PCB_NCCL_SYM(field, sym)                                                          \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, sym));                    \
  if (!n.field) { snprintf(n.why, sizeof(n.why), "NCCL symbol %s missing", sym); return; }
  PCB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  PCB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  PCB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  PCB_NCCL_SYM(AllReduce, "ncclAllReduce")
  PCB_NCCL_SYM(Send, "ncclSend")
  PCB_NCCL_SYM(Recv, "ncclRecv")
  PCB_NCCL_SYM(GroupStart, "ncclGroupStart")
  PCB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  PCB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  PCB_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef PCB_NCCL_SYM
  n.ok = true;
}

static const NcclApi* nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) { set_error("%s", g_nccl.why); return nullptr; }
  return &g_nccl;
}

#define PCB_NCCL_CALL(api, expr, what)                                          \
  do {                                                                          \
    ncclResult_t r__ = (expr);                                                  \
    if (r__ != ncclSuccess) {                                                   \
      pcb::set_error("%s: %s", what, (api)->GetErrorString(r__));               \
      return PCB_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

static bool nccl_type(int dtype, ncclDataType_t* t, size_t* bytes) {
  switch (dtype) {
    case PCB_F32: *t = ncclFloat32; *bytes = 4; return true;
    case PCB_F16: *t = ncclFloat16; *bytes = 2; return true;
    case PCB_BF16: *t = ncclBfloat16; *bytes = 2; return true;
    default: return false;
  }
}

// x *= s, 128-bit accesses on the aligned body (the gradient arena is 256-byte aligned), scalar tail.
template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* __restrict__ x, int64_t n, float s) {
  constexpr int PER = 16 / sizeof(T);
  const int64_t nv = n / PER;
  uint4* xv = reinterpret_cast<uint4*>(x);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 raw = xv[i];
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int k = 0; k < PER; ++k) e[k] = (T)((float)e[k] * s);
    xv[i] = raw;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = nv * PER; i < n; ++i) x[i] = (T)((float)x[i] * s);
}

template <typename T>
static int launch_scale(void* x, int64_t n, float s, cudaStream_t st) {
  if (n <= 0) return PCB_OK;
  const int64_t want = (n / (16 / (int64_t)sizeof(T)) + 255) / 256;
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  scale_kernel<T><<<grid, 256, 0, st>>>(reinterpret_cast<T*>(x), n, s);
  PCB_CHECK_LAUNCH("scale_kernel");
  return PCB_OK;
}

}  // namespace pcb

struct pcb_comm {
  ncclComm_t comm;
  int rank, world, device;
};

extern "C" int pcb_comm_unique_id(void* id_out) {
  PCB_CHECK_ARG(id_out != nullptr, "pcb_comm_unique_id: id_out is NULL");
  static_assert(sizeof(ncclUniqueId) == PCB_COMM_ID_BYTES, "PCB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  const pcb::NcclApi* n = pcb::nccl();
  if (!n) return PCB_ERR_UNSUPPORTED;
  PCB_NCCL_CALL(n, n->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id_out)), "ncclGetUniqueId");
  return PCB_OK;
}

extern "C" int pcb_comm_init(const void* unique_id, int rank, int world, pcb_comm** out) {
  using namespace pcb;
  PCB_CHECK_ARG(unique_id != nullptr && out != nullptr, "pcb_comm_init: NULL argument");
  PCB_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "pcb_comm_init: rank %d is not in [0, %d)", rank, world);
  *out = nullptr;
  const NcclApi* n = nccl();
  if (!n) return PCB_ERR_UNSUPPORTED;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("pcb_comm_init: no CUDA device (there is no host path for the exchange)");
    return PCB_ERR_CUDA;
  }
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  PCB_NCCL_CALL(n, n->CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  pcb_comm* c = new pcb_comm{comm, rank, world, dev};
  *out = c;
  return PCB_OK;
}

extern "C" void pcb_comm_destroy(pcb_comm* c) {
  if (!c) return;
  const pcb::NcclApi* n = pcb::nccl();
  if (n && c->comm) n->CommDestroy(c->comm);
  delete c;
}

extern "C" int pcb_comm_rank(const pcb_comm* c) { return c ? c->rank : -1; }
extern "C" int pcb_comm_world(const pcb_comm* c) { return c ? c->world : -1; }

extern "C" int pcb_comm_nccl_version(void) {
  const pcb::NcclApi* n = pcb::nccl();
  int v = 0;
  if (!n || n->GetVersion(&v) != ncclSuccess) return -1;
  return v;
}

extern "C" int pcb_grad_allreduce(pcb_comm* c, void* arena, int64_t numel, int dtype, float scale, void* stream) {
  using namespace pcb;
  PCB_CHECK_ARG(c != nullptr && c->comm != nullptr, "pcb_grad_allreduce: communicator is NULL");
  PCB_CHECK_ARG(numel >= 0 && (arena != nullptr || numel == 0), "pcb_grad_allreduce: bad buffer");
  ncclDataType_t t;
  size_t eb;
  PCB_CHECK_ARG(nccl_type(dtype, &t, &eb), "pcb_grad_allreduce: unsupported dtype %d", dtype);
  PCB_CHECK_ARG((reinterpret_cast<uintptr_t>(arena) & 15) == 0, "pcb_grad_allreduce: the arena must be 16-byte aligned");
  const NcclApi* n = nccl();
  if (!n) return PCB_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (numel == 0) return PCB_OK;
  if (c->world > 1) PCB_NCCL_CALL(n, n->AllReduce(arena, arena, (size_t)numel, t, ncclSum, c->comm, st), "ncclAllReduce");
  if (scale != 1.0f) {
    if (dtype == PCB_F32) return launch_scale<float>(arena, numel, scale, st);
    if (dtype == PCB_F16) return launch_scale<__half>(arena, numel, scale, st);
    return launch_scale<__nv_bfloat16>(arena, numel, scale, st);
  }
  return PCB_OK;
}

extern "C" int pcb_sw_exchange_overlap(pcb_comm* c, int nsend, const void* const* send_bufs, const int64_t* send_numel,
                                       const int* send_peer, int nrecv, void* const* recv_bufs, const int64_t* recv_numel,
                                       const int* recv_peer, int dtype, void* stream) {
  using namespace pcb;
  PCB_CHECK_ARG(c != nullptr && c->comm != nullptr, "pcb_sw_exchange_overlap: communicator is NULL");
  PCB_CHECK_ARG(nsend >= 0 && nrecv >= 0, "pcb_sw_exchange_overlap: negative message count");
  PCB_CHECK_ARG(nsend == 0 || (send_bufs && send_numel && send_peer), "pcb_sw_exchange_overlap: NULL send list");
  PCB_CHECK_ARG(nrecv == 0 || (recv_bufs && recv_numel && recv_peer), "pcb_sw_exchange_overlap: NULL receive list");
  ncclDataType_t t;
  size_t eb;
  PCB_CHECK_ARG(nccl_type(dtype, &t, &eb), "pcb_sw_exchange_overlap: unsupported dtype %d", dtype);
  for (int i = 0; i < nsend; ++i)
    PCB_CHECK_ARG(send_peer[i] >= 0 && send_peer[i] < c->world && send_peer[i] != c->rank && send_numel[i] > 0 && send_bufs[i],
                  "pcb_sw_exchange_overlap: bad send %d (peer %d, %lld elements)", i, send_peer[i], (long long)send_numel[i]);
  for (int i = 0; i < nrecv; ++i)
    PCB_CHECK_ARG(recv_peer[i] >= 0 && recv_peer[i] < c->world && recv_peer[i] != c->rank && recv_numel[i] > 0 && recv_bufs[i],
                  "pcb_sw_exchange_overlap: bad receive %d (peer %d, %lld elements)", i, recv_peer[i], (long long)recv_numel[i]);
  if (nsend + nrecv == 0) return PCB_OK;
  const NcclApi* n = nccl();
  if (!n) return PCB_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // one group: every send/recv of this rank is posted before any of them has to complete (two faces per interior rank),
  // so neighbours cannot deadlock on message order
  PCB_NCCL_CALL(n, n->GroupStart(), "ncclGroupStart");
  for (int i = 0; i < nsend; ++i) {
    ncclResult_t r = n->Send(send_bufs[i], (size_t)send_numel[i], t, send_peer[i], c->comm, st);
    if (r != ncclSuccess) { n->GroupEnd(); set_error("ncclSend: %s", n->GetErrorString(r)); return PCB_ERR_CUDA; }
  }
  for (int i = 0; i < nrecv; ++i) {
    ncclResult_t r = n->Recv(recv_bufs[i], (size_t)recv_numel[i], t, recv_peer[i], c->comm, st);
    if (r != ncclSuccess) { n->GroupEnd(); set_error("ncclRecv: %s", n->GetErrorString(r)); return PCB_ERR_CUDA; }
  }
  PCB_NCCL_CALL(n, n->GroupEnd(), "ncclGroupEnd");
  return PCB_OK;
}
