// comm.cu — peer-memory fabric for the data-parallel training path (BASELINE.json north_star: "training is data-parallel
// over the batch with the gradient allreduce over NCCL/NVLink"; the reference itself is single-device).
//
// One process per GPU.  Each rank cudaMallocs a HEAP, exports it with cudaIpcGetMemHandle and maps every peer's heap
// (cudaIpcOpenMemHandle, peer access over NVLink 5 / NVSwitch).  The collectives are ordinary kernels over those
// mappings (comm_dev.cuh), so they are captured in the step graph and overlap with the convolutions:
//   * ssr_comm_adam_step     gradient reduce-scatter + Keras-Adam + parameter all-gather in ONE kernel: rank r pulls its
//                            1/world shard of every peer's gradient (fixed rank order: deterministic and identical on all
//                            ranks), updates m / v / param for that shard only, and stores the new parameters into every
//                            rank's buffer.  Replaces all-reduce -> full Adam on every rank.
//   * ssr_comm_allreduce_f32 one-shot all-reduce of small vectors (loss metrics)
//   * sync-BatchNorm / RaGAN variants live next to their single-device kernels (disc_kernels.cu) and use the same barrier.
// torch.distributed is only used by the host code to exchange the 64-byte IPC handles (plumbing).
#include <cuda_runtime.h>

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>

#include "comm_dev.cuh"
#include "internal.h"

using namespace ssr;

namespace ssr { struct LocalGroup; }

struct ssr_comm {
  ssr::LocalGroup* group = nullptr;   // "ranks" of one process on one device (tests): collectives rendezvous on the host
  int device = 0, rank = 0, world = 1;
  size_t heap_bytes = 0;
  uint8_t* local = nullptr;
  uint8_t* peer[kCommMaxWorld] = {};
  bool ipc_open[kCommMaxWorld] = {};
  uint32_t* counters = nullptr;
  unsigned long long* status = nullptr;
  bool opened = false;
  CommDev dev;
};

namespace ssr {

const CommDev* comm_dev(const ssr_comm* c) { return (c && c->opened) ? &c->dev : nullptr; }
size_t comm_heap_bytes(const ssr_comm* c) { return c ? c->heap_bytes : 0; }
bool comm_is_group(const ssr_comm* c) { return c && c->group != nullptr; }

struct OptState {    // device-resident optimizer clock (ssr_opt_prepare), read by the Adam kernels
  long long step;    // updates applied so far + 1 during the current step (Keras `iterations` + 1)
  float lr;          // learning rate of the current step (schedule evaluated at `iterations`)
  float lr_t;        // lr * sqrt(1 - b2^t) / (1 - b1^t)
  float pad_[12];
};
static_assert(sizeof(OptState) == 64, "OptState is 64 bytes");

// step += 1; lr from the piecewise-constant schedule (tf.keras PiecewiseConstantDecay: values[0] while
// iterations <= boundaries[0], values[i] while boundaries[i-1] < iterations <= boundaries[i], examples/training/
// example_without_yaml.py:287-297), or base_lr without one; Keras Adam's bias-corrected step size (sr_model.py:121-131).
__global__ void opt_prepare_kernel(OptState* st, float base_lr, float b1, float b2, const long long* bounds,
                                   const float* values, int nb) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long it = st->step;  // iterations before this update
  float lr = base_lr;
  if (nb > 0) {
    int i = 0;
    while (i < nb && it > bounds[i]) ++i;
    lr = values[i];
  }
  const double t = static_cast<double>(it + 1);
  st->step = it + 1;
  st->lr = lr;
  st->lr_t = static_cast<float>(lr * sqrt(1.0 - pow(static_cast<double>(b2), t)) / (1.0 - pow(static_cast<double>(b1), t)));
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr_t, float b1, float b2, float eps) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p = p - lr_t * m / (sqrtf(v) + eps);
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t count, const OptState* __restrict__ st, float b1, float b2,
                                float eps, float grad_scale) {
  const float lr_t = st->lr_t;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam1(pi, g[i] * grad_scale, mi, vi, lr_t, b1, b2, eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

constexpr int kCommAdamBlocks = 64;
constexpr int kCommAdamThreads = 512;

struct AdamArgs {
  int slot0;
  size_t grad_off, param_off;
  float *m, *v;
  int64_t lo, hi;
  const OptState* st;
  float b1, b2, eps;
};
// Elements [lo, hi) of the flat gradient / parameter buffers (both at the same offset in every rank's heap).
__device__ __forceinline__ void comm_adam_body(const CommDev& c, const AdamArgs& a) {
  const int slot = a.slot0 + blockIdx.x;
  comm_barrier(c, slot);  // every rank's gradients of this range are final (their backward kernels precede this one)
  const float lr_t = a.st->lr_t;
  const float inv_world = 1.f / static_cast<float>(c.world);
  const int64_t n4 = (a.hi - a.lo + 3) >> 2;  // buffers are padded to a multiple of 4 floats
  const int64_t per = (n4 + c.world - 1) / c.world;
  const int64_t i0 = c.rank * per, i1 = (i0 + per < n4) ? i0 + per : n4;
  for (int64_t i = i0 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < i1;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = a.lo + 4 * i;
    float4 g[kCommMaxWorld];
#pragma unroll
    for (int p = 0; p < kCommMaxWorld; ++p)
      if (p < c.world) g[p] = ld_peer_f4(reinterpret_cast<const float*>(c.heap[p] + a.grad_off) + e);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kCommMaxWorld; ++p) {  // fixed order: the same sum on whichever rank owns the shard
      if (p < c.world) {
        s.x += g[p].x;
        s.y += g[p].y;
        s.z += g[p].z;
        s.w += g[p].w;
      }
    }
    float* pl = reinterpret_cast<float*>(c.heap[c.rank] + a.param_off) + e;
    float4 pv = *reinterpret_cast<const float4*>(pl);
    float4 mv = *reinterpret_cast<const float4*>(a.m + e), vv = *reinterpret_cast<const float4*>(a.v + e);
    adam1(pv.x, s.x * inv_world, mv.x, vv.x, lr_t, a.b1, a.b2, a.eps);
    adam1(pv.y, s.y * inv_world, mv.y, vv.y, lr_t, a.b1, a.b2, a.eps);
    adam1(pv.z, s.z * inv_world, mv.z, vv.z, lr_t, a.b1, a.b2, a.eps);
    adam1(pv.w, s.w * inv_world, mv.w, vv.w, lr_t, a.b1, a.b2, a.eps);
    *reinterpret_cast<float4*>(a.m + e) = mv;
    *reinterpret_cast<float4*>(a.v + e) = vv;
#pragma unroll
    for (int p = 0; p < kCommMaxWorld; ++p)  // all-gather: the owner stores the new values into every rank's buffer
      if (p < c.world) *reinterpret_cast<float4*>(reinterpret_cast<float*>(c.heap[p] + a.param_off) + e) = pv;
  }
  __threadfence_system();
  comm_barrier(c, slot);  // all shards of all ranks have landed everywhere: the re-pack may read the parameters
}
__global__ void __launch_bounds__(kCommAdamThreads) comm_adam_kernel(const CommDev c, const AdamArgs a) { comm_adam_body(c, a); }

// out[i] = scale * sum_p in_p[i], count small.  The input is first copied into a staging region of the heap that
// alternates with the barrier epoch, so back-to-back calls never overwrite what a slower peer still reads.
struct SmallArgs {
  int slot;
  size_t stage_off;
  const float* in;
  float* out;
  int count;
  float scale;
};
__device__ __forceinline__ void comm_small_body(const CommDev& c, const SmallArgs& a) {
  const uint32_t e = c.counters[a.slot] + 1;
  const size_t off = a.stage_off + static_cast<size_t>(e & 1) * a.count * sizeof(float);
  float* mine = reinterpret_cast<float*>(c.heap[c.rank] + off);
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) mine[i] = a.in[i];
  __threadfence_system();
  comm_barrier(c, a.slot);
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < c.world; ++p) s += ld_peer_f(reinterpret_cast<const float*>(c.heap[p] + off) + i);
    a.out[i] = s * a.scale;
  }
}
__global__ void __launch_bounds__(256) comm_allreduce_small_kernel(const CommDev c, const SmallArgs a) { comm_small_body(c, a); }

struct BarrierArgs {
  int slot;
};
__global__ void comm_barrier_kernel(const CommDev c, const BarrierArgs a) { comm_barrier(c, a.slot); }

// ---- emulated ranks (ssr_comm_open_local): the kernels of all ranks run as ONE cooperative launch, blockIdx.y = rank -
// kernels that wait for each other must never be separate launches on one GPU (nothing guarantees they run together)
template <typename Args>
struct MultiParams {
  CommDev dev[kCommMaxLocal];
  Args arg[kCommMaxLocal];
};
__global__ void __launch_bounds__(kCommAdamThreads) comm_adam_multi_kernel(const __grid_constant__ MultiParams<AdamArgs> p) {
  comm_adam_body(p.dev[blockIdx.y], p.arg[blockIdx.y]);
}
__global__ void __launch_bounds__(256) comm_small_multi_kernel(const __grid_constant__ MultiParams<SmallArgs> p) {
  comm_small_body(p.dev[blockIdx.y], p.arg[blockIdx.y]);
}
__global__ void comm_barrier_multi_kernel(const __grid_constant__ MultiParams<BarrierArgs> p) {
  comm_barrier(p.dev[blockIdx.y], p.arg[blockIdx.y].slot);
}

// Host rendezvous of the emulated ranks (one host thread per rank, eager launches): every rank records what it has
// queued so far and parks its arguments; the last one to arrive makes the group stream wait for all ranks, launches the
// multi-rank kernel cooperatively, and makes every rank's stream wait for it.  No GPU-side waiting across launches.
struct LocalGroup {
  int world = 0, refs = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long generation = 0;
  int last_rc = SSR_OK;
  cudaStream_t gstream = nullptr;
  cudaEvent_t ready[kCommMaxLocal] = {}, done = nullptr;
  cudaStream_t stream[kCommMaxLocal] = {};
  const void* tag[kCommMaxLocal] = {};          // which collective each rank asked for (must agree)
  alignas(16) unsigned char blob[kCommMaxLocal][kCommBlobBytes];
  CommDev dev[kCommMaxLocal];
};

int comm_group_collective(ssr_comm* c, cudaStream_t stream, const void* tag, const void* args, size_t arg_bytes,
                          GroupLauncher launch) {
  LocalGroup* g = c->group;
  if (arg_bytes > kCommBlobBytes) return set_error(SSR_ERR_INVALID, "comm group: argument blob too large");
  std::unique_lock<std::mutex> lk(g->mu);
  const int r = c->rank;
  memcpy(g->blob[r], args, arg_bytes);
  g->stream[r] = stream;
  g->tag[r] = tag;
  g->dev[r] = c->dev;
  cudaError_t e = cudaEventRecord(g->ready[r], stream);
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "comm group: cudaEventRecord: %s", cudaGetErrorString(e));
  const unsigned long long gen = g->generation;
  if (++g->arrived == g->world) {
    int rc = SSR_OK;
    for (int p = 0; p < g->world && rc == SSR_OK; ++p) {
      if (g->tag[p] != tag) rc = set_error(SSR_ERR_INVALID, "comm group: the ranks called different collectives");
      else if (cudaStreamWaitEvent(g->gstream, g->ready[p], 0) != cudaSuccess) rc = set_error(SSR_ERR_CUDA, "comm group: wait");
    }
    if (rc == SSR_OK) {
      e = launch(g->dev, &g->blob[0][0], kCommBlobBytes, g->world, g->gstream);
      if (e != cudaSuccess) rc = set_error(SSR_ERR_CUDA, "comm group: cooperative launch: %s", cudaGetErrorString(e));
    }
    if (rc == SSR_OK) {
      cudaEventRecord(g->done, g->gstream);
      for (int p = 0; p < g->world; ++p) cudaStreamWaitEvent(g->stream[p], g->done, 0);
    }
    g->last_rc = rc;
    g->arrived = 0;
    ++g->generation;
    g->cv.notify_all();
    return rc;
  }
  if (!g->cv.wait_for(lk, std::chrono::seconds(60), [&] { return g->generation != gen; })) {
    --g->arrived;
    return set_error(SSR_ERR_INVALID, "comm group: a rank did not reach the collective within 60 s");
  }
  return g->last_rc == SSR_OK ? SSR_OK : set_error(g->last_rc, "comm group: the collective failed on the launching rank");
}

template <typename Args, typename Kernel>
static cudaError_t launch_multi(Kernel kernel, const CommDev* dev, const unsigned char* blobs, size_t stride, int world,
                                dim3 grid, dim3 block, cudaStream_t st) {
  MultiParams<Args> p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < world; ++r) {
    p.dev[r] = dev[r];
    memcpy(&p.arg[r], blobs + r * stride, sizeof(Args));
  }
  void* params[] = {&p};
  grid.y = world;
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), grid, block, params, 0, st);
}
static cudaError_t launch_adam_multi(const CommDev* dev, const unsigned char* blobs, size_t stride, int world, cudaStream_t st) {
  return launch_multi<AdamArgs>(comm_adam_multi_kernel, dev, blobs, stride, world, dim3(16), dim3(kCommAdamThreads), st);
}
static cudaError_t launch_small_multi(const CommDev* dev, const unsigned char* blobs, size_t stride, int world, cudaStream_t st) {
  return launch_multi<SmallArgs>(comm_small_multi_kernel, dev, blobs, stride, world, dim3(1), dim3(256), st);
}
static cudaError_t launch_barrier_multi(const CommDev* dev, const unsigned char* blobs, size_t stride, int world, cudaStream_t st) {
  return launch_multi<BarrierArgs>(comm_barrier_multi_kernel, dev, blobs, stride, world, dim3(1), dim3(32), st);
}

}  // namespace ssr

#define SSR_CUDA(call, what)                                                                 \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, what ": %s", cudaGetErrorString(e__)); \
  } while (0)
#define SSR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
  } while (0)

extern "C" size_t ssr_comm_data_offset(void) { return kCommDataOffset; }
extern "C" int ssr_comm_max_slots(void) { return kCommMaxSlots; }
extern "C" int ssr_comm_adam_slots(void) { return kCommAdamBlocks; }

extern "C" int ssr_comm_create(int device, int rank, int world, size_t heap_bytes, ssr_comm** out) {
  if (!out) return set_error(SSR_ERR_INVALID, "comm_create: out is NULL");
  *out = nullptr;
  if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world)
    return set_error(SSR_ERR_INVALID, "comm_create: need 1 <= world <= %d and 0 <= rank < world", kCommMaxWorld);
  if (heap_bytes < kCommDataOffset) return set_error(SSR_ERR_INVALID, "comm_create: heap smaller than its flag area");
  SSR_CUDA(cudaSetDevice(device), "comm_create: cudaSetDevice");
  ssr_comm* c = new ssr_comm();
  c->device = device;
  c->rank = rank;
  c->world = world;
  c->heap_bytes = heap_bytes;
  cudaError_t e = cudaMalloc(&c->local, heap_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, heap_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->counters, kCommMaxSlots * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(c->counters, 0, kCommMaxSlots * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&c->status, 64);
  if (e == cudaSuccess) e = cudaMemset(c->status, 0, 64);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(c->local);
    cudaFree(c->counters);
    cudaFree(c->status);
    delete c;
    return set_error(SSR_ERR_NOMEM, "comm_create: %s", cudaGetErrorString(e));
  }
  memset(&c->dev, 0, sizeof(c->dev));
  c->dev.rank = rank;
  c->dev.world = world;
  c->dev.counters = c->counters;
  c->dev.status = c->status;
  c->dev.spin_limit = 40000000000ll;  // ~20 s at 2 GHz: ranks may enter their first step seconds apart
  c->peer[rank] = c->local;
  if (world == 1) {
    c->dev.heap[0] = c->local;
    c->opened = true;
  }
  *out = c;
  return SSR_OK;
}

extern "C" int ssr_comm_destroy(ssr_comm* c) {
  if (!c) return SSR_OK;
  cudaSetDevice(c->device);
  if (c->group != nullptr) {
    LocalGroup* g = c->group;
    bool last;
    {
      std::lock_guard<std::mutex> lk(g->mu);
      last = --g->refs == 0;
    }
    if (last) {
      cudaStreamSynchronize(g->gstream);
      cudaStreamDestroy(g->gstream);
      cudaEventDestroy(g->done);
      for (int p = 0; p < g->world; ++p) cudaEventDestroy(g->ready[p]);
      delete g;
    }
  }
  for (int p = 0; p < c->world; ++p)
    if (c->ipc_open[p]) cudaIpcCloseMemHandle(c->peer[p]);
  cudaFree(c->local);
  cudaFree(c->counters);
  cudaFree(c->status);
  delete c;
  return SSR_OK;
}

extern "C" void* ssr_comm_heap(ssr_comm* c) { return c ? c->local : nullptr; }
extern "C" size_t ssr_comm_heap_size(ssr_comm* c) { return c ? c->heap_bytes : 0; }

extern "C" int ssr_comm_ipc_handle(ssr_comm* c, void* handle_out_64) {
  if (!c || !handle_out_64) return set_error(SSR_ERR_INVALID, "comm_ipc_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SSR_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  SSR_CUDA(cudaIpcGetMemHandle(&h, c->local), "cudaIpcGetMemHandle");
  memcpy(handle_out_64, &h, sizeof(h));
  return SSR_OK;
}

static int comm_finish_open(ssr_comm* c) {
  for (int p = 0; p < c->world; ++p) c->dev.heap[p] = c->peer[p];
  c->opened = true;
  return SSR_OK;
}

extern "C" int ssr_comm_open_ipc(ssr_comm* c, const void* handles) {
  if (!c || !handles) return set_error(SSR_ERR_INVALID, "comm_open_ipc: NULL argument");
  SSR_CUDA(cudaSetDevice(c->device), "comm_open_ipc: cudaSetDevice");
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + static_cast<size_t>(p) * SSR_COMM_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return set_error(SSR_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (peer access over NVLink is required for "
                                     "data-parallel training)", p, cudaGetErrorString(e));
    c->peer[p] = static_cast<uint8_t*>(ptr);
    c->ipc_open[p] = true;
  }
  return comm_finish_open(c);
}

// "ranks" living in one process on one device (tests: several trainers on one GPU exercise the same kernels)
extern "C" int ssr_comm_open_local(ssr_comm* const* comms, int world) {
  if (!comms || world < 2 || world > kCommMaxLocal)
    return set_error(SSR_ERR_INVALID, "comm_open_local: 2 <= world <= %d emulated ranks", kCommMaxLocal);
  for (int p = 0; p < world; ++p)
    if (!comms[p] || comms[p]->world != world || comms[p]->rank != p || comms[p]->device != comms[0]->device ||
        comms[p]->opened)
      return set_error(SSR_ERR_INVALID, "comm_open_local: comms must be the unopened ranks 0..world-1 of one device");
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, comms[0]->device);
  if (!coop) return set_error(SSR_ERR_UNSUPPORTED, "comm_open_local: the device has no cooperative launch");
  LocalGroup* g = new LocalGroup();
  g->world = world;
  g->refs = world;
  cudaError_t e = cudaStreamCreateWithFlags(&g->gstream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->done, cudaEventDisableTiming);
  for (int p = 0; p < world && e == cudaSuccess; ++p) e = cudaEventCreateWithFlags(&g->ready[p], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete g;
    return set_error(SSR_ERR_CUDA, "comm_open_local: %s", cudaGetErrorString(e));
  }
  for (int r = 0; r < world; ++r) {
    for (int p = 0; p < world; ++p) comms[r]->peer[p] = comms[p]->local;
    comms[r]->group = g;
    comm_finish_open(comms[r]);
  }
  return SSR_OK;
}

extern "C" int ssr_comm_set_spin_limit(ssr_comm* c, double seconds) {
  if (!c || seconds <= 0) return set_error(SSR_ERR_INVALID, "comm_set_spin_limit: bad argument");
  c->dev.spin_limit = static_cast<long long>(seconds * 2.0e9);
  return SSR_OK;
}

extern "C" int ssr_comm_status(ssr_comm* c, unsigned long long* host_timeouts) {
  if (!c || !host_timeouts) return set_error(SSR_ERR_INVALID, "comm_status: NULL argument");
  SSR_CUDA(cudaMemcpy(host_timeouts, c->status, sizeof(unsigned long long), cudaMemcpyDeviceToHost), "comm_status");
  return SSR_OK;
}

static int comm_check(const ssr_comm* c, const char* what) {
  if (!c) return set_error(SSR_ERR_INVALID, "%s: comm is NULL", what);
  if (!c->opened) return set_error(SSR_ERR_INVALID, "%s: peers not opened (ssr_comm_open_ipc / ssr_comm_open_local)", what);
  return SSR_OK;
}

extern "C" int ssr_comm_barrier(ssr_comm* c, int slot, void* stream) {
  if (int rc = comm_check(c, "comm_barrier")) return rc;
  if (slot < 0 || slot >= kCommMaxSlots) return set_error(SSR_ERR_INVALID, "comm_barrier: slot out of range");
  BarrierArgs a{slot};
  if (c->group) return comm_group_collective(c, static_cast<cudaStream_t>(stream), reinterpret_cast<const void*>(&launch_barrier_multi), &a, sizeof(a), launch_barrier_multi);
  comm_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(c->dev, a);
  SSR_CHECK_LAUNCH("comm_barrier");
  return SSR_OK;
}

extern "C" int ssr_comm_allreduce_f32(ssr_comm* c, int slot, size_t stage_off, const float* in, float* out, int count,
                                      float scale, void* stream) {
  if (int rc = comm_check(c, "comm_allreduce_f32")) return rc;
  if (!in || !out || count <= 0 || count > 65536 || slot < 0 || slot >= kCommMaxSlots || stage_off < kCommDataOffset ||
      stage_off % 16 || stage_off + 2 * static_cast<size_t>(count) * 4 > c->heap_bytes)
    return set_error(SSR_ERR_INVALID, "comm_allreduce_f32: bad argument (count <= 65536, staging of 2 * count floats "
                                      "inside the heap)");
  SmallArgs a{slot, stage_off, in, out, count, scale};
  if (c->group) return comm_group_collective(c, static_cast<cudaStream_t>(stream), reinterpret_cast<const void*>(&launch_small_multi), &a, sizeof(a), launch_small_multi);
  comm_allreduce_small_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(c->dev, a);
  SSR_CHECK_LAUNCH("comm_allreduce_f32");
  return SSR_OK;
}

extern "C" size_t ssr_opt_state_bytes(void) { return sizeof(OptState); }

extern "C" int ssr_opt_state_set(void* state, int64_t iterations, void* stream) {
  if (!state || iterations < 0) return set_error(SSR_ERR_INVALID, "opt_state_set: bad argument");
  OptState h;
  memset(&h, 0, sizeof(h));
  h.step = iterations;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SSR_CUDA(cudaMemcpyAsync(state, &h, sizeof(h), cudaMemcpyHostToDevice, st), "opt_state_set");
  SSR_CUDA(cudaStreamSynchronize(st), "opt_state_set: sync");  // `h` dies with this call
  return SSR_OK;
}

extern "C" int ssr_opt_state_get(const void* state, int64_t* host_iterations, float* host_lr, void* stream) {
  if (!state) return set_error(SSR_ERR_INVALID, "opt_state_get: NULL");
  OptState h;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SSR_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st), "opt_state_get");
  SSR_CUDA(cudaStreamSynchronize(st), "opt_state_get: sync");
  if (host_iterations) *host_iterations = h.step;
  if (host_lr) *host_lr = h.lr;
  return SSR_OK;
}

extern "C" int ssr_opt_prepare(void* state, float base_lr, float beta1, float beta2, const int64_t* boundaries_dev,
                               const float* values_dev, int n_boundaries, void* stream) {
  if (!state || n_boundaries < 0 || (n_boundaries > 0 && (!boundaries_dev || !values_dev)))
    return set_error(SSR_ERR_INVALID, "opt_prepare: bad argument");
  opt_prepare_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<OptState*>(state), base_lr, beta1, beta2, reinterpret_cast<const long long*>(boundaries_dev), values_dev,
      n_boundaries);
  SSR_CHECK_LAUNCH("opt_prepare");
  return SSR_OK;
}

extern "C" int ssr_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t count, const void* opt_state,
                                 float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!param || !grad || !m || !v || !opt_state || count < 0) return set_error(SSR_ERR_INVALID, "adam_step_dev: bad argument");
  if (count == 0) return SSR_OK;
  int64_t g = (count + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  adam_dev_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, m, v, count, static_cast<const OptState*>(opt_state), beta1, beta2, eps, grad_scale);
  SSR_CHECK_LAUNCH("adam_step_dev");
  return SSR_OK;
}

extern "C" int ssr_comm_adam_step(ssr_comm* c, int slot0, size_t grad_off, size_t param_off, float* m, float* v,
                                  int64_t lo, int64_t hi, const void* opt_state, float beta1, float beta2, float eps,
                                  void* stream) {
  if (int rc = comm_check(c, "comm_adam_step")) return rc;
  if (!m || !v || !opt_state || lo < 0 || hi < lo || (lo & 3))
    return set_error(SSR_ERR_INVALID, "comm_adam_step: bad range (lo must be a multiple of 4)");
  const int64_t hi4 = (hi + 3) & ~static_cast<int64_t>(3);
  if (slot0 < 0 || slot0 + kCommAdamBlocks > kCommMaxSlots || grad_off < kCommDataOffset || param_off < kCommDataOffset ||
      (grad_off & 15) || (param_off & 15) || grad_off + static_cast<size_t>(hi4) * 4 > c->heap_bytes ||
      param_off + static_cast<size_t>(hi4) * 4 > c->heap_bytes)
    return set_error(SSR_ERR_INVALID, "comm_adam_step: buffers must lie inside the heap, 16-byte aligned and padded to a "
                                      "multiple of 4 floats");
  if (hi == lo) return SSR_OK;
  AdamArgs a{slot0, grad_off, param_off, m, v, lo, hi, static_cast<const OptState*>(opt_state), beta1, beta2, eps};
  if (c->group) return comm_group_collective(c, static_cast<cudaStream_t>(stream), reinterpret_cast<const void*>(&launch_adam_multi), &a, sizeof(a), launch_adam_multi);
  comm_adam_kernel<<<kCommAdamBlocks, kCommAdamThreads, 0, static_cast<cudaStream_t>(stream)>>>(c->dev, a);
  SSR_CHECK_LAUNCH("comm_adam_step");
  return SSR_OK;
}
