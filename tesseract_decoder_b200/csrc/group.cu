// Multi-GPU face of the C ABI (include/tesseract_b200.h, "multi-GPU" section).
//
// Shots are independent (tesseract.cc:665-671): the reference shards them over CPU threads (utils.h:75-115,
// tesseract_main.cc:559-616) and sums three counters.  Here each GPU holds a full copy of the DEM tables (a few tens of
// MB) and decodes its own shots; the ONLY exchange is one ncclAllReduce(sum) of {shots, logical errors, low
// confidence} (+ the optional error-use histogram of --dem-out, tesseract_main.cc:618-623) over NVLink / NVSwitch.
// Nothing is fused into a kernel because no kernel of this path is followed by a collective on its data.
//
// Two deployment shapes:
//   * one process per GPU (torchrun, MPI, ...): tsr_comm_unique_id on rank 0, shipped to the other ranks by whatever
//     the launcher offers, then tsr_comm_init_rank on every rank's decoder and tsr_comm_allreduce_sum_u64 per batch;
//   * one process driving several GPUs (the CLI's --gpus N): tsr_group_create / tsr_group_decode_b8 — one decoder, host
//     thread and communicator (ncclCommInitAll) per device, shots dealt in interleaved chunks because per-shot cost is
//     heavy-tailed.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) rather than at link time: a process that never asks for a
// communicator does not load it, and one that already carries a copy (e.g. the one bundled with PyTorch) shares it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "handle.h"

using namespace tsr_impl;

namespace {

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;

  static Nccl& get() {
    static Nccl api;
    static std::once_flag once;
    std::call_once(once, [] {
      for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
      }
      if (!api.lib) return;
      auto sym = [&](auto& fn, const char* name) { fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.lib, name)); };
      sym(api.GetUniqueId, "ncclGetUniqueId");
      sym(api.CommInitRank, "ncclCommInitRank");
      sym(api.CommInitAll, "ncclCommInitAll");
      sym(api.CommDestroy, "ncclCommDestroy");
      sym(api.AllReduce, "ncclAllReduce");
      sym(api.GroupStart, "ncclGroupStart");
      sym(api.GroupEnd, "ncclGroupEnd");
      sym(api.GetErrorString, "ncclGetErrorString");
      sym(api.GetVersion, "ncclGetVersion");
    });
    if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllReduce ||
        !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
      throw CudaError("NCCL: libnccl.so.2 could not be loaded (needed for multi-GPU decoding only)");
    return api;
  }
};

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw CudaError(std::string("NCCL: ") + what + ": " + Nccl::get().GetErrorString(r));
}
#define NCCL_OK(expr) nccl_check((expr), #expr)

}  // namespace

struct tsr_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf scratch;  // device staging of the reduced counters
};

void tsr_impl::comm_release(tsr_decoder& h) {
  if (!h.comm) return;
  if (h.comm->comm) {
    cudaSetDevice(h.device);
    Nccl::get().CommDestroy(h.comm->comm);
  }
  delete h.comm;
  h.comm = nullptr;
}

namespace {

// In-place sum over the communicator of `n` u64 counters held on the host: H2D, ncclAllReduce on the decoder's stream, D2H.
void allreduce_u64(tsr_decoder& h, uint64_t* values, uint64_t n) {
  if (!h.comm) throw std::invalid_argument("this decoder has no communicator: call tsr_comm_init_rank or use a tsr_group");
  if (n == 0) return;
  CUDA_OK(cudaSetDevice(h.device));
  h.comm->scratch.ensure(n * sizeof(uint64_t));
  void* d = h.comm->scratch.p;
  CUDA_OK(cudaMemcpyAsync(d, values, n * sizeof(uint64_t), cudaMemcpyHostToDevice, h.stream));
  NCCL_OK(Nccl::get().AllReduce(d, d, n, ncclUint64, ncclSum, h.comm->comm, h.stream));
  CUDA_OK(cudaMemcpyAsync(values, d, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, h.stream));
  CUDA_OK(cudaStreamSynchronize(h.stream));
}

}  // namespace

struct tsr_group {
  std::vector<tsr_decoder*> members;
  std::vector<DevBuf> d_dets, d_obs, d_cost, d_low;  // per member: its shots, contiguous
  uint64_t last_shots = 0, last_chunk = 0;           // shape of the last tsr_group_decode_b8 call
  bool last_has_errors = false;
  ~tsr_group() {
    for (size_t i = 0; i < members.size(); ++i) {
      if (!members[i]) continue;
      cudaSetDevice(members[i]->device);
      d_dets[i].release();
      d_obs[i].release();
      d_cost[i].release();
      d_low[i].release();
      tsr_destroy(members[i]);
    }
  }
};

extern "C" {

int tsr_comm_unique_id(uint8_t out[TSR_COMM_UNIQUE_ID_BYTES]) {
  return guarded([&] {
    static_assert(sizeof(ncclUniqueId) == TSR_COMM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_OK(Nccl::get().GetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
  });
}

int tsr_comm_init_rank(tsr_decoder* h, const uint8_t id_bytes[TSR_COMM_UNIQUE_ID_BYTES], int32_t rank, int32_t world) {
  return guarded([&] {
    require_device(*h);
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("need 0 <= rank < world");
    comm_release(*h);
    CUDA_OK(cudaSetDevice(h->device));
    auto c = std::make_unique<tsr_comm>();
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    NCCL_OK(Nccl::get().CommInitRank(&c->comm, world, id, rank));
    h->comm = c.release();
  });
}

int32_t tsr_comm_world(const tsr_decoder* h) {
  return h->comm ? h->comm->world : 0;
}

int tsr_comm_allreduce_sum_u64(tsr_decoder* h, uint64_t* values, uint64_t n) {
  return guarded([&] {
    require_device(*h);
    allreduce_u64(*h, values, n);
  });
}

int tsr_nccl_version(void) {
  int v = 0;
  guarded([&] {
    if (Nccl::get().GetVersion) Nccl::get().GetVersion(&v);
  });
  return v;
}

int tsr_group_create(const char* dem_text, const tsr_config* cfg, const int32_t* devices, uint32_t n_devices, tsr_group** out) {
  *out = nullptr;
  return guarded([&] {
    if (n_devices == 0) throw std::invalid_argument("a group needs at least one device");
    int have = 0;
    CUDA_OK(cudaGetDeviceCount(&have));
    std::vector<int> devs(n_devices);
    for (uint32_t i = 0; i < n_devices; ++i) {
      devs[i] = devices ? devices[i] : (int)i;
      if (devs[i] < 0 || devs[i] >= have)
        throw std::invalid_argument("device " + std::to_string(devs[i]) + " does not exist (" + std::to_string(have) + " visible)");
      for (uint32_t j = 0; j < i; ++j)
        if (devs[j] == devs[i]) throw std::invalid_argument("a device may appear in a group only once");
    }
    auto g = std::make_unique<tsr_group>();
    g->members.assign(n_devices, nullptr);
    g->d_dets = std::vector<DevBuf>(n_devices);
    g->d_obs = std::vector<DevBuf>(n_devices);
    g->d_cost = std::vector<DevBuf>(n_devices);
    g->d_low = std::vector<DevBuf>(n_devices);
    for (uint32_t i = 0; i < n_devices; ++i) {
      tsr_config c = *cfg;
      c.device = devs[i];
      const int rc = tsr_create(dem_text, &c, 0, &g->members[i]);
      if (rc == TSR_E_INVALID_ARGUMENT) throw std::invalid_argument(last_error());
      if (rc == TSR_E_CUDA) throw CudaError(last_error());
      if (rc != TSR_OK) throw std::runtime_error(last_error());
    }
    std::vector<ncclComm_t> comms(n_devices);
    NCCL_OK(Nccl::get().CommInitAll(comms.data(), (int)n_devices, devs.data()));
    for (uint32_t i = 0; i < n_devices; ++i) {
      auto c = new tsr_comm;
      c->comm = comms[i];
      c->rank = (int)i;
      c->world = (int)n_devices;
      g->members[i]->comm = c;
    }
    *out = g.release();
  });
}

void tsr_group_destroy(tsr_group* g) {
  delete g;
}

uint32_t tsr_group_size(const tsr_group* g) {
  return (uint32_t)g->members.size();
}

tsr_decoder* tsr_group_member(tsr_group* g, uint32_t i) {
  return i < g->members.size() ? g->members[i] : nullptr;
}

int tsr_group_decode_b8(tsr_group* g, const uint8_t* dets, uint64_t stride, uint64_t shots, uint64_t chunk_shots,
                        const uint8_t* obs_true, uint8_t* obs_out, double* cost_out, uint8_t* low_conf_out,
                        uint64_t tally_out[3], uint64_t* error_use_out) {
  return guarded([&] {
    const uint32_t n = (uint32_t)g->members.size();
    const uint64_t chunk = chunk_shots ? chunk_shots : 16384;
    const uint64_t n_chunks = (shots + chunk - 1) / chunk;
    const uint64_t O = tsr_num_observables(g->members[0]), ob = (O + 7) / 8;
    const uint64_t n_use = error_use_out ? tsr_num_dem_errors(g->members[0]) : 0;
    std::vector<std::string> failure(n);
    std::vector<int> failure_code(n, TSR_OK);
    std::vector<std::vector<uint64_t>> reduced(n, std::vector<uint64_t>(3 + n_use, 0));
    auto work = [&](uint32_t i) {
      tsr_decoder& h = *g->members[i];
      std::vector<uint64_t>& acc = reduced[i];
      try {
        require_device(h);
        CUDA_OK(cudaSetDevice(h.device));
        // This member's chunks i, i+n, ... laid out back to back on its device.
        uint64_t mine = 0;
        for (uint64_t k = i; k < n_chunks; k += n) mine += std::min(chunk, shots - k * chunk);
        if (mine) {
          g->d_dets[i].ensure(mine * stride);
          g->d_obs[i].ensure(std::max<uint64_t>(mine * ob, 16));
          g->d_cost[i].ensure(mine * 8);
          g->d_low[i].ensure(mine);
          uint64_t at = 0;
          for (uint64_t k = i; k < n_chunks; k += n) {
            const uint64_t len = std::min(chunk, shots - k * chunk);
            CUDA_OK(cudaMemcpyAsync(g->d_dets[i].as<uint8_t>() + at * stride, dets + k * chunk * stride, len * stride,
                                    cudaMemcpyHostToDevice, h.stream));
            at += len;
          }
          const bool keep = h.collect_errors;
          h.collect_errors = error_use_out != nullptr;
          decode_batch(h, g->d_dets[i].as<uint8_t>(), stride, mine, true, h.ensemble, g->d_obs[i].as<uint8_t>(),
                       g->d_cost[i].as<double>(), g->d_low[i].as<uint8_t>());
          h.collect_errors = keep;
          std::vector<uint8_t> obs_local(mine * ob + 1), low_local(mine);
          if (ob) CUDA_OK(cudaMemcpyAsync(obs_local.data(), g->d_obs[i].p, mine * ob, cudaMemcpyDeviceToHost, h.stream));
          CUDA_OK(cudaMemcpyAsync(low_local.data(), g->d_low[i].p, mine, cudaMemcpyDeviceToHost, h.stream));
          at = 0;
          for (uint64_t k = i; k < n_chunks; k += n) {
            const uint64_t len = std::min(chunk, shots - k * chunk);
            if (cost_out)
              CUDA_OK(cudaMemcpyAsync(cost_out + k * chunk, g->d_cost[i].as<double>() + at, len * 8, cudaMemcpyDeviceToHost, h.stream));
            at += len;
          }
          CUDA_OK(cudaStreamSynchronize(h.stream));
          // Scatter to the shots' own positions and count with the CLI's accounting (tesseract_main.cc:580-601): a
          // low-confidence shot is never a logical error; error_use only counts correctly predicted shots.
          at = 0;
          for (uint64_t k = i; k < n_chunks; k += n) {
            const uint64_t len = std::min(chunk, shots - k * chunk), base = k * chunk;
            if (obs_out && ob) std::memcpy(obs_out + base * ob, obs_local.data() + at * ob, len * ob);
            if (low_conf_out) std::memcpy(low_conf_out + base, low_local.data() + at, len);
            for (uint64_t j = 0; j < len; ++j) {
              const bool low = low_local[at + j] != 0;
              const bool match = !obs_true || std::memcmp(obs_local.data() + (at + j) * ob, obs_true + (base + j) * ob, ob) == 0;
              acc[1] += (!low && !match) ? 1 : 0;
              acc[2] += low ? 1 : 0;
              if (n_use && match && !low)
                for (uint64_t q = h.last_off[at + j]; q < h.last_off[at + j + 1]; ++q) ++acc[3 + h.last_idx[q]];
            }
            at += len;
          }
          acc[0] = mine;
        }
      } catch (const std::invalid_argument& e) {
        failure[i] = e.what();
        failure_code[i] = TSR_E_INVALID_ARGUMENT;
      } catch (const DeviceMemoryError& e) {
        failure[i] = e.what();
        failure_code[i] = TSR_E_DEVICE_MEMORY;
      } catch (const CudaError& e) {
        failure[i] = e.what();
        failure_code[i] = TSR_E_CUDA;
      } catch (const std::exception& e) {
        failure[i] = e.what();
        failure_code[i] = TSR_E_RUNTIME;
      }
      // Every member takes part in the collective, failed or not (a missing rank would hang the others).
      try {
        if (failure_code[i] != TSR_OK) std::fill(acc.begin(), acc.end(), 0);
        allreduce_u64(h, acc.data(), acc.size());
      } catch (const std::exception& e) {
        if (failure_code[i] == TSR_OK) {
          failure[i] = e.what();
          failure_code[i] = TSR_E_CUDA;
        }
      }
    };
    std::vector<std::thread> threads;
    for (uint32_t i = 1; i < n; ++i) threads.emplace_back(work, i);
    work(0);
    for (auto& t : threads) t.join();
    for (uint32_t i = 0; i < n; ++i) {
      if (failure_code[i] == TSR_OK) continue;
      const std::string msg = "device " + std::to_string(g->members[i]->device) + ": " + failure[i];
      switch (failure_code[i]) {
        case TSR_E_INVALID_ARGUMENT: throw std::invalid_argument(msg);
        case TSR_E_DEVICE_MEMORY: throw DeviceMemoryError(msg);
        case TSR_E_CUDA: throw CudaError(msg);
        default: throw std::runtime_error(msg);
      }
    }
    g->last_shots = shots;
    g->last_chunk = chunk;
    g->last_has_errors = error_use_out != nullptr;
    if (tally_out) std::copy(reduced[0].begin(), reduced[0].begin() + 3, tally_out);
    if (error_use_out) std::copy(reduced[0].begin() + 3, reduced[0].end(), error_use_out);
  });
}

uint64_t tsr_group_last_batch_nnz(const tsr_group* g) {
  uint64_t n = 0;
  if (g->last_has_errors)
    for (const tsr_decoder* m : g->members) n += m->last_idx.size();
  return n;
}

int tsr_group_last_batch_errors(const tsr_group* g, uint64_t* offsets, uint64_t* idx) {
  return guarded([&] {
    if (!g->last_has_errors) throw std::invalid_argument("the last tsr_group_decode_b8 call did not collect error lists (error_use_out was NULL)");
    const uint64_t n = g->members.size(), chunk = g->last_chunk, shots = g->last_shots;
    std::vector<uint64_t> local(n, 0);  // next local shot of each member
    uint64_t at = 0;
    for (uint64_t s = 0; s < shots; ++s) {
      const tsr_decoder& m = *g->members[(s / chunk) % n];
      const uint64_t j = local[(s / chunk) % n]++;
      offsets[s] = at;
      for (uint64_t q = m.last_off[j]; q < m.last_off[j + 1]; ++q) idx[at++] = m.last_idx[q];
    }
    offsets[shots] = at;
  });
}

}  // extern "C"
