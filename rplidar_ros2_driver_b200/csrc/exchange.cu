// exchange.cu -- the one exchange of the multi-GPU PointCloud2 path (BASELINE.json configs[4]):
// every rank contributes its fused per-GPU cloud, every rank ends up with all of them, in rank order.
//
// Host side in C++ behind the C-ABI (include/rpl_b200.h, rpl_exchange_*), one process per GPU.  The only thing
// the embedding process supplies is the 128-byte NCCL unique id (rank 0 creates it, any out-of-band channel
// carries it: torch.distributed in bench.py, a ROS parameter or a file in a deployment).  NCCL is loaded at run
// time (dlopen "libnccl.so.2"): a single-GPU user of this library does not need it.
//
// One step (rpl_exchange_allgather), buffer b = step & 1:
//   caller's stream S:   offsets + pack kernel: the per-scan clouds -> this rank's slot of gather buffer b
//                        (header = point count, then the dense points), then an event
//   exchange stream X:   (highest priority, waits for that event)
//       RPL_EXCHANGE_NCCL  one in-place ncclAllGather of the slot over NVLink / NVSwitch
//       RPL_EXCHANGE_COPY  world-1 peer copies of the slot into the peers' buffers, which are mapped once
//                          through CUDA IPC -- the COPY ENGINES move the cloud over NVLink, no SM is involved --
//                          in a rotated order (rank+1, rank+2, ...) so that every receiver is written by one
//                          sender at a time; then a 4-byte all-reduce as the barrier "everyone has delivered"
//                        then the `done` event of buffer b
// S never waits for X: the next batch's scan kernels run while the cloud of this batch is in flight.  A consumer
// orders itself behind the transfer with rpl_exchange_wait and hands the buffer back with rpl_exchange_release
// (two buffers alternate; the exchange that is about to reuse a buffer waits for its release).
#include <dlfcn.h>
#include <nccl.h>  // declarations only: the library is not linked, see load_nccl()

#include <cstring>
#include <mutex>
#include <new>

#include "rpl_ctx.h"

namespace {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  void* handle = nullptr;
  bool ok = false;
  std::string err;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // an already loaded libnccl.so.2 (e.g. the one PyTorch brought) is reused: dlopen matches the soname
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) {
      api.err = std::string("dlopen libnccl.so.2: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(api.handle, name);
      if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.ok = api.err.empty();
  });
  return api;
}

constexpr uint32_t kSlotHeader = 16;  // bytes in front of a slot's points: [count][pad x3]

}  // namespace

struct rpl_exchange {
  rpl_ctx* ctx = nullptr;
  uint32_t world = 0, rank = 0, slot_points = 0, flags = 0;
  size_t slot_bytes = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t xstream = nullptr;
  unsigned char* buf[2] = {nullptr, nullptr};                 // [world][slot_bytes], own allocation
  unsigned char* peer[2][RPL_MAX_PEERS] = {};                 // peers' buffers (CUDA IPC mappings); [rank] = own
  uint32_t* offsets = nullptr;                                // [max_scans] scratch of the pack
  uint32_t offsets_cap = 0;
  float* flag = nullptr;                                      // barrier all-reduce operand
  cudaEvent_t packed[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, released[2] = {nullptr, nullptr};
  bool release_pending[2] = {false, false};
  uint64_t step = 0;
  bool peers_mapped = false;
};

namespace {

bool nccl_ok(rpl_ctx* c, ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return true;
  c->err = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  return false;
}
#define RPL_NCCL(c, call, code)                          \
  do {                                                   \
    if (!nccl_ok((c), (call), #call)) return (code);     \
  } while (0)

}  // namespace

extern "C" {

rpl_result rpl_exchange_unique_id(uint8_t* id_out) {
  if (!id_out) return RPL_RESULT_INVALID_DATA;
  static_assert(sizeof(ncclUniqueId) == RPL_EXCHANGE_ID_BYTES, "NCCL unique id size");
  NcclApi& n = nccl();
  if (!n.ok) {
    std::fprintf(stderr, "[rpl_b200] NCCL unavailable: %s\n", n.err.c_str());
    return RPL_RESULT_OPERATION_NOT_SUPPORT;
  }
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != ncclSuccess) return RPL_RESULT_OPERATION_FAIL;
  std::memcpy(id_out, &id, sizeof(id));
  return RPL_RESULT_OK;
}

void rpl_exchange_destroy(rpl_exchange* ex) {
  if (!ex) return;
  rpl_ctx* c = ex->ctx;
  cudaSetDevice(c->device);
  if (ex->xstream) cudaStreamSynchronize(ex->xstream);
  cudaDeviceSynchronize();
  // nobody unmaps or frees while a peer may still be copying into this rank: one last barrier
  if (ex->comm && ex->flag && ex->world > 1 && ex->xstream) {
    nccl().AllReduce(ex->flag, ex->flag, 1, ncclFloat32, ncclSum, ex->comm, ex->xstream);
    cudaStreamSynchronize(ex->xstream);
  }
  for (int b = 0; b < 2; ++b) {
    for (uint32_t p = 0; p < ex->world && p < RPL_MAX_PEERS; ++p)
      if (p != ex->rank && ex->peer[b][p]) cudaIpcCloseMemHandle(ex->peer[b][p]);
  }
  if (ex->comm) nccl().CommDestroy(ex->comm);
  for (int b = 0; b < 2; ++b) {
    cudaFree(ex->buf[b]);
    if (ex->packed[b]) cudaEventDestroy(ex->packed[b]);
    if (ex->done[b]) cudaEventDestroy(ex->done[b]);
    if (ex->released[b]) cudaEventDestroy(ex->released[b]);
  }
  cudaFree(ex->offsets);
  cudaFree(ex->flag);
  if (ex->xstream) cudaStreamDestroy(ex->xstream);
  delete ex;
}

rpl_result rpl_exchange_create(rpl_ctx* c, const uint8_t* id, uint32_t world, uint32_t rank, uint32_t slot_points,
                               uint32_t flags, rpl_exchange** out) {
  if (!c || !out || world == 0 || world > RPL_MAX_PEERS || rank >= world || slot_points == 0) return RPL_RESULT_INVALID_DATA;
  *out = nullptr;
  if (world > 1 && !id) {
    c->err = "a NCCL unique id is needed for world > 1";
    return RPL_RESULT_INVALID_DATA;
  }
  NcclApi& n = nccl();
  if (world > 1 && !n.ok) {
    c->err = "NCCL unavailable: " + n.err;
    return RPL_RESULT_OPERATION_NOT_SUPPORT;
  }
  RPL_CUDA(c, cudaSetDevice(c->device), RPL_RESULT_OPERATION_FAIL);
  rpl_exchange* ex = new (std::nothrow) rpl_exchange();
  if (!ex) return RPL_RESULT_INSUFFICIENT_MEMORY;
  ex->ctx = c;
  ex->world = world;
  ex->rank = rank;
  ex->slot_points = slot_points;
  ex->flags = flags;
  ex->slot_bytes = kSlotHeader + (size_t)slot_points * 16;
  auto fail = [&](rpl_result r) {
    const std::string why = c->err;
    rpl_exchange_destroy(ex);
    c->err = why;
    return r;
  };
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  if (!cuda_ok(c, cudaStreamCreateWithPriority(&ex->xstream, cudaStreamNonBlocking, hi), "cudaStreamCreateWithPriority"))
    return fail(RPL_RESULT_OPERATION_FAIL);
  for (int b = 0; b < 2; ++b) {
    if (!cuda_ok(c, cudaMalloc(reinterpret_cast<void**>(&ex->buf[b]), ex->slot_bytes * world), "cudaMalloc gather buffer") ||
        !cuda_ok(c, cudaMemset(ex->buf[b], 0, ex->slot_bytes * world), "cudaMemset") ||
        !cuda_ok(c, cudaEventCreateWithFlags(&ex->packed[b], cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(c, cudaEventCreateWithFlags(&ex->done[b], cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(c, cudaEventCreateWithFlags(&ex->released[b], cudaEventDisableTiming), "cudaEventCreate"))
      return fail(RPL_RESULT_INSUFFICIENT_MEMORY);
    ex->peer[b][rank] = ex->buf[b];
  }
  if (!cuda_ok(c, cudaMalloc(reinterpret_cast<void**>(&ex->flag), 256), "cudaMalloc") ||
      !cuda_ok(c, cudaMemset(ex->flag, 0, 256), "cudaMemset"))
    return fail(RPL_RESULT_INSUFFICIENT_MEMORY);
  if (world > 1) {
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    if (!nccl_ok(c, n.CommInitRank(&ex->comm, (int)world, uid, (int)rank), "ncclCommInitRank")) return fail(RPL_RESULT_OPERATION_FAIL);
    if ((flags & RPL_EXCHANGE_NO_PEER_MAP) == 0) {
      // the peers' gather buffers, mapped once: IPC handles travel through the communicator itself
      constexpr size_t H = sizeof(cudaIpcMemHandle_t);
      static_assert(H == 64, "IPC handle size");
      unsigned char* d_handles = nullptr;
      if (!cuda_ok(c, cudaMalloc(reinterpret_cast<void**>(&d_handles), 2 * H * world), "cudaMalloc")) return fail(RPL_RESULT_INSUFFICIENT_MEMORY);
      std::vector<unsigned char> h(2 * H * world, 0);
      bool good = true;
      for (int b = 0; b < 2 && good; ++b) {
        cudaIpcMemHandle_t mine;
        good = cuda_ok(c, cudaIpcGetMemHandle(&mine, ex->buf[b]), "cudaIpcGetMemHandle") &&
               cuda_ok(c, cudaMemcpyAsync(d_handles + ((size_t)b * world + rank) * H, &mine, H, cudaMemcpyHostToDevice, ex->xstream), "H2D") &&
               nccl_ok(c, n.AllGather(d_handles + ((size_t)b * world + rank) * H, d_handles + (size_t)b * world * H, H, ncclUint8, ex->comm, ex->xstream), "ncclAllGather(handles)");
      }
      good = good && cuda_ok(c, cudaMemcpyAsync(h.data(), d_handles, h.size(), cudaMemcpyDeviceToHost, ex->xstream), "D2H") &&
             cuda_ok(c, cudaStreamSynchronize(ex->xstream), "sync");
      cudaFree(d_handles);
      if (!good) return fail(RPL_RESULT_OPERATION_FAIL);
      for (int b = 0; b < 2; ++b)
        for (uint32_t p = 0; p < world; ++p) {
          if (p == rank) continue;
          cudaIpcMemHandle_t hp;
          std::memcpy(&hp, h.data() + ((size_t)b * world + p) * H, H);
          void* mapped = nullptr;
          if (!cuda_ok(c, cudaIpcOpenMemHandle(&mapped, hp, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return fail(RPL_RESULT_OPERATION_FAIL);
          ex->peer[b][p] = static_cast<unsigned char*>(mapped);
        }
      ex->peers_mapped = true;
    }
  }
  *out = ex;
  return RPL_RESULT_OK;
}

rpl_result rpl_exchange_allgather(rpl_exchange* ex, const float* xyzi, const uint32_t* point_counts, uint32_t n_scans,
                                  uint32_t stride, uint32_t mode, void* stream, uint32_t* buffer_index) {
  if (!ex || !xyzi || !point_counts) return RPL_RESULT_INVALID_DATA;
  rpl_ctx* c = ex->ctx;
  if (mode != RPL_EXCHANGE_NCCL && mode != RPL_EXCHANGE_COPY) {
    c->err = "mode must be RPL_EXCHANGE_NCCL or RPL_EXCHANGE_COPY";
    return RPL_RESULT_INVALID_DATA;
  }
  if (mode == RPL_EXCHANGE_COPY && ex->world > 1 && !ex->peers_mapped) {
    c->err = "RPL_EXCHANGE_COPY needs the peer mappings (created without RPL_EXCHANGE_NO_PEER_MAP)";
    return RPL_RESULT_INVALID_DATA;
  }
  RPL_CUDA(c, cudaSetDevice(c->device), RPL_RESULT_OPERATION_FAIL);
  cudaStream_t S = stream ? static_cast<cudaStream_t>(stream) : c->lane[0].stream;
  const uint32_t b = (uint32_t)(ex->step & 1u);
  ex->step++;
  if (n_scans > ex->offsets_cap) {
    RPL_CUDA(c, cudaStreamSynchronize(S), RPL_RESULT_OPERATION_FAIL);
    cudaFree(ex->offsets);
    ex->offsets = nullptr;
    RPL_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&ex->offsets), (size_t)n_scans * 4), RPL_RESULT_INSUFFICIENT_MEMORY);
    ex->offsets_cap = n_scans;
  }
  // buffer b was last filled two steps ago: its consumer must have handed it back, and the transfer that filled
  // it must be over before the pack overwrites this rank's slot
  if (ex->release_pending[b]) {
    RPL_CUDA(c, cudaStreamWaitEvent(S, ex->released[b], 0), RPL_RESULT_OPERATION_FAIL);
    ex->release_pending[b] = false;
  }
  RPL_CUDA(c, cudaStreamWaitEvent(S, ex->done[b], 0), RPL_RESULT_OPERATION_FAIL);
  unsigned char* slot = ex->buf[b] + (size_t)ex->rank * ex->slot_bytes;
  int launched = 0;
  RPL_CUDA(c, rpl::launch_cloud_fuse(reinterpret_cast<const float4*>(xyzi), point_counts, n_scans, stride,
                                     reinterpret_cast<float4*>(slot + kSlotHeader), ex->slot_points, ex->offsets,
                                     reinterpret_cast<uint32_t*>(slot), S, &launched),
           RPL_RESULT_OPERATION_FAIL);
  c->launches += launched;
  RPL_CUDA(c, cudaEventRecord(ex->packed[b], S), RPL_RESULT_OPERATION_FAIL);
  cudaStream_t X = ex->xstream;
  RPL_CUDA(c, cudaStreamWaitEvent(X, ex->packed[b], 0), RPL_RESULT_OPERATION_FAIL);
  if (ex->world > 1) {
    NcclApi& n = nccl();
    if (mode == RPL_EXCHANGE_NCCL) {
      RPL_NCCL(c, n.AllGather(slot, ex->buf[b], ex->slot_bytes, ncclUint8, ex->comm, X), RPL_RESULT_OPERATION_FAIL);
    } else {
      // peers must not still be reading what this step overwrites in THEIR buffer b: that is their release,
      // ordered by the barrier of the previous step (a rank passes it only after its own waits)
      for (uint32_t j = 1; j < ex->world; ++j) {
        const uint32_t p = (ex->rank + j) % ex->world;
        RPL_CUDA(c, cudaMemcpyAsync(ex->peer[b][p] + (size_t)ex->rank * ex->slot_bytes, slot, ex->slot_bytes,
                                    cudaMemcpyDeviceToDevice, X),
                 RPL_RESULT_OPERATION_FAIL);
      }
      // The barrier: once every rank has passed it, every rank's copies of this step have landed.  It is also
      // what lets the peers overwrite this rank's OTHER buffer in their next step, so it waits for the consumer
      // of that buffer (if one was registered with rpl_exchange_release).
      if (ex->release_pending[b ^ 1u])
        RPL_CUDA(c, cudaStreamWaitEvent(X, ex->released[b ^ 1u], 0), RPL_RESULT_OPERATION_FAIL);
      RPL_NCCL(c, n.AllReduce(ex->flag, ex->flag, 1, ncclFloat32, ncclSum, ex->comm, X), RPL_RESULT_OPERATION_FAIL);
    }
  }
  RPL_CUDA(c, cudaEventRecord(ex->done[b], X), RPL_RESULT_OPERATION_FAIL);
  if (buffer_index) *buffer_index = b;
  return RPL_RESULT_OK;
}

rpl_result rpl_exchange_wait(rpl_exchange* ex, uint32_t index, void* stream) {
  if (!ex || index > 1) return RPL_RESULT_INVALID_DATA;
  rpl_ctx* c = ex->ctx;
  RPL_CUDA(c, cudaSetDevice(c->device), RPL_RESULT_OPERATION_FAIL);
  cudaStream_t S = stream ? static_cast<cudaStream_t>(stream) : c->lane[0].stream;
  RPL_CUDA(c, cudaStreamWaitEvent(S, ex->done[index], 0), RPL_RESULT_OPERATION_FAIL);
  return RPL_RESULT_OK;
}

rpl_result rpl_exchange_release(rpl_exchange* ex, uint32_t index, void* stream) {
  if (!ex || index > 1) return RPL_RESULT_INVALID_DATA;
  rpl_ctx* c = ex->ctx;
  RPL_CUDA(c, cudaSetDevice(c->device), RPL_RESULT_OPERATION_FAIL);
  cudaStream_t S = stream ? static_cast<cudaStream_t>(stream) : c->lane[0].stream;
  RPL_CUDA(c, cudaEventRecord(ex->released[index], S), RPL_RESULT_OPERATION_FAIL);
  ex->release_pending[index] = true;
  return RPL_RESULT_OK;
}

rpl_result rpl_exchange_slot(rpl_exchange* ex, uint32_t index, uint32_t rank, const float** points, const uint32_t** count) {
  if (!ex || index > 1 || rank >= ex->world) return RPL_RESULT_INVALID_DATA;
  const unsigned char* slot = ex->buf[index] + (size_t)rank * ex->slot_bytes;
  if (count) *count = reinterpret_cast<const uint32_t*>(slot);
  if (points) *points = reinterpret_cast<const float*>(slot + kSlotHeader);
  return RPL_RESULT_OK;
}

rpl_result rpl_exchange_synchronize(rpl_exchange* ex) {
  if (!ex) return RPL_RESULT_INVALID_DATA;
  rpl_ctx* c = ex->ctx;
  RPL_CUDA(c, cudaSetDevice(c->device), RPL_RESULT_OPERATION_FAIL);
  RPL_CUDA(c, cudaStreamSynchronize(ex->xstream), RPL_RESULT_OPERATION_FAIL);
  return RPL_RESULT_OK;
}

}  // extern "C"
