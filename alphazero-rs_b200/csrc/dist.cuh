// The library's own communicator and multi-GPU fan-out (north star: "no PyTorch").
//
// The reference is single-process (rayon threads, crossbeam channels): it has no collective.  Here independent games shard
// over GPUs with NO collective on the self-play path; the one place that needs one is the optional data-parallel training
// step of Coach::learn (gradient all-reduce) and the three arena counters.  Round 1 left every NCCL call to the caller
// (torch.distributed in the ctypes mirror); now the library can do it itself:
//
//   * azb_dist_unique_id / azb_dist_init / azb_dist_destroy: an NCCL communicator per process (one process per GPU), built
//     from a 128-byte unique id that the host passes to its ranks by whatever means it has (a file, MPI, a socket);
//   * azb_dist_make: an azb_dist whose two all-reduce callbacks run on that communicator (ncclAllReduce in place on the
//     device gradient vector over NVLink / NVSwitch; the u64 counters staged through a small device buffer), so that
//     azb_coach_learn_dist needs no torch;
//   * azb_coach_self_play_multi: ONE call, one host thread per listed GPU, each playing its shard of the games on its own
//     coach (weights replicated); returns when every shard is done.
//
// NCCL is loaded with dlopen at first use (AZB200_NCCL_LIB, default libnccl.so.2): libazb200.so itself links only the CUDA
// runtime, and a host process that already carries an NCCL (torch does) shares it instead of loading a second copy.
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: no NCCL symbol is linked

#include <mutex>

namespace {

struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* name = std::getenv("AZB200_NCCL_LIB") ? std::getenv("AZB200_NCCL_LIB") : "libnccl.so.2";
    api.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (!api.so) {
      api.why = std::string("cannot load ") + name + ": " + dlerror();
      return;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(api.so, s);
      if (!p && api.why.empty()) api.why = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return api;
}

int nccl_ready() {
  NcclApi& a = nccl_api();
  if (!a.so || !a.why.empty()) return fail(AZB_ERR_UNSUPPORTED, "NCCL is not available: " + a.why);
  return AZB_OK;
}

#define AZB_NCCL(expr)                                                                                           \
  do {                                                                                                           \
    ncclResult_t _r = (expr);                                                                                    \
    if (_r != ncclSuccess) return fail(AZB_ERR_CUDA, std::string(#expr) + ": " + nccl_api().GetErrorString(_r)); \
  } while (0)

}  // namespace

struct azb_comm {
  ncclComm_t comm = nullptr;
  uint32_t rank = 0, world = 1;
  int device = 0;
  cudaStream_t stream = nullptr;
  DevBuf scratch;  // u64 counters staged for the host-side all-reduce
};

namespace {
// the two callbacks of azb_dist, backed by an azb_comm (`user` = the communicator)
int comm_allreduce_f32_device(void* device_ptr, uint64_t count, void* user) {
  azb_comm* c = static_cast<azb_comm*>(user);
  if (cudaSetDevice(c->device) != cudaSuccess) return 1;
  // the library's kernels run on the legacy default stream; c->stream is a blocking stream, so the reduction is ordered
  // after the gradient kernels and before whatever the caller launches next
  if (nccl_api().AllReduce(device_ptr, device_ptr, count, ncclFloat32, ncclSum, c->comm, c->stream) != ncclSuccess) return 1;
  return cudaStreamSynchronize(c->stream) == cudaSuccess ? 0 : 1;
}
int comm_allreduce_u64_host(uint64_t* host_ptr, uint64_t count, void* user) {
  azb_comm* c = static_cast<azb_comm*>(user);
  if (cudaSetDevice(c->device) != cudaSuccess) return 1;
  if (c->scratch.ensure(count * 8) != cudaSuccess) return 1;
  if (cudaMemcpyAsync(c->scratch.p, host_ptr, count * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return 1;
  if (nccl_api().AllReduce(c->scratch.p, c->scratch.p, count, ncclUint64, ncclSum, c->comm, c->stream) != ncclSuccess) return 1;
  if (cudaMemcpyAsync(host_ptr, c->scratch.p, count * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return 1;
  return cudaStreamSynchronize(c->stream) == cudaSuccess ? 0 : 1;
}
}  // namespace

extern "C" {

int azb_dist_unique_id(uint8_t out[AZB_DIST_ID_BYTES]) {
  if (!out) return fail(AZB_ERR_INVALID, "NULL argument");
  static_assert(AZB_DIST_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "the unique id travels as raw bytes");
  int rc = nccl_ready();
  if (rc) return rc;
  ncclUniqueId id;
  AZB_NCCL(nccl_api().GetUniqueId(&id));
  std::memcpy(out, id.internal, AZB_DIST_ID_BYTES);
  return AZB_OK;
}

int azb_dist_init(const uint8_t id_bytes[AZB_DIST_ID_BYTES], uint32_t rank, uint32_t world, int32_t device, azb_comm** out) {
  if (!id_bytes || !out) return fail(AZB_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (world == 0 || rank >= world) return fail(AZB_ERR_INVALID, "rank / world out of range");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  int rc = nccl_ready();
  if (rc) return rc;
  AZB_CUDA(cudaSetDevice(device));
  auto c = std::make_unique<azb_comm>();
  c->rank = rank;
  c->world = world;
  c->device = device;
  ncclUniqueId id;
  std::memcpy(id.internal, id_bytes, AZB_DIST_ID_BYTES);
  AZB_NCCL(nccl_api().CommInitRank(&c->comm, static_cast<int>(world), id, static_cast<int>(rank)));
  AZB_CUDA(cudaStreamCreate(&c->stream));
  *out = c.release();
  return AZB_OK;
}

int azb_dist_destroy(azb_comm* c) {
  if (!c) return AZB_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->comm) nccl_api().CommDestroy(c->comm);
  delete c;
  return AZB_OK;
}

int azb_dist_make(azb_comm* c, azb_dist* out) {
  if (!c || !out) return fail(AZB_ERR_INVALID, "NULL argument");
  out->rank = c->rank;
  out->world = c->world;
  out->allreduce_sum_f32_device = comm_allreduce_f32_device;
  out->allreduce_sum_u64_host = comm_allreduce_u64_host;
  out->user = c;
  return AZB_OK;
}

int azb_dist_allreduce_f64(azb_comm* c, double* values, uint64_t count, int32_t op) {
  if (!c || !values) return fail(AZB_ERR_INVALID, "NULL argument");
  if (op != AZB_DIST_SUM && op != AZB_DIST_MAX && op != AZB_DIST_MIN) return fail(AZB_ERR_INVALID, "op");
  AZB_CUDA(cudaSetDevice(c->device));
  AZB_CUDA(c->scratch.ensure(count * 8));
  AZB_CUDA(cudaMemcpyAsync(c->scratch.p, values, count * 8, cudaMemcpyHostToDevice, c->stream));
  const ncclRedOp_t rop = op == AZB_DIST_SUM ? ncclSum : (op == AZB_DIST_MAX ? ncclMax : ncclMin);
  AZB_NCCL(nccl_api().AllReduce(c->scratch.p, c->scratch.p, count, ncclFloat64, rop, c->comm, c->stream));
  AZB_CUDA(cudaMemcpyAsync(values, c->scratch.p, count * 8, cudaMemcpyDeviceToHost, c->stream));
  AZB_CUDA(cudaStreamSynchronize(c->stream));
  return AZB_OK;
}

// One call, every listed GPU: a host thread per device sets up its own coach (and its own copy of the network when the
// evaluator is AZB_EVAL_NNET: weights replicated from the same seed), plays games [first_game_id + d * games_per_device,
// ...) and writes stats[d].  No collective: the shards never talk to each other.
int azb_coach_self_play_multi(const azb_config* cfg, const azb_nnet_config* net_cfg, const int32_t* devices, uint32_t n_devices,
                              uint64_t games_per_device, uint64_t first_game_id, azb_selfplay_stats* stats, double* wall_ms) {
  if (!cfg || !devices || n_devices == 0) return fail(AZB_ERR_INVALID, "NULL argument");
  if (cfg->evaluator >= AZB_EVAL_NNET && !net_cfg) return fail(AZB_ERR_INVALID, "evaluator NNET needs a network configuration");
  std::vector<int> rcs(n_devices, AZB_OK);
  std::vector<std::string> errs(n_devices);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (uint32_t d = 0; d < n_devices; ++d)
    th.emplace_back([&, d] {
      azb_config c = *cfg;
      c.device = devices[d];
      c.checkpoint_directory = nullptr;  // self-play only: no history to resume
      azb_coach* coach = nullptr;
      azb_nnet* net = nullptr;
      int rc = azb_coach_setup(&c, &coach);
      if (!rc && cfg->evaluator >= AZB_EVAL_NNET) {
        azb_nnet_config nc = *net_cfg;
        nc.device = devices[d];
        rc = azb_nnet_create(&nc, &net);
        if (!rc) rc = azb_coach_set_nnet(coach, net);
      }
      azb_selfplay_stats st{};
      if (!rc) rc = azb_coach_self_play(coach, games_per_device, first_game_id + static_cast<uint64_t>(d) * games_per_device, &st);
      if (rc) errs[d] = azb_last_error();  // (thread-local message of this worker)
      if (stats) stats[d] = st;
      if (coach) azb_coach_destroy(coach);
      if (net) azb_nnet_destroy(net);
      rcs[d] = rc;
    });
  for (auto& t : th) t.join();
  if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  for (uint32_t d = 0; d < n_devices; ++d)
    if (rcs[d]) return fail(rcs[d], "device " + std::to_string(devices[d]) + ": " + errs[d]);
  return AZB_OK;
}

}  // extern "C"
