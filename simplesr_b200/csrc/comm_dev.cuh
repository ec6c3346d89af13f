// comm_dev.cuh — device side of the peer-memory fabric (one process per GPU, buffers mapped into each other's address
// space with CUDA IPC over NVLink / NVSwitch).  Every collective of the data-parallel training path is a plain kernel
// that loads / stores peer memory directly and synchronises the ranks with flags in each other's heap, so it can sit
// inside the captured step graph next to the convolutions (new work: the reference has no parallelism).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ssr {

constexpr int kCommMaxWorld = 8;
constexpr int kCommMaxLocal = 4;       // emulated ranks of one process (tests): one cooperative launch covers them all
constexpr int kCommBlobBytes = 160;    // per-rank argument record of an emulated collective
constexpr int kCommMaxSlots = 16384;                                   // barrier slots per heap
constexpr size_t kCommSignalBytes = static_cast<size_t>(kCommMaxSlots) * kCommMaxWorld * 4;  // 512 KB of flags
constexpr size_t kCommDataOffset = 1u << 20;                           // data starts 1 MB into the heap

struct CommDev {
  int rank, world;
  uint8_t* heap[kCommMaxWorld];   // heap[p] = base of rank p's heap as mapped in THIS process (heap[rank] = local)
  uint32_t* counters;             // local, private: epoch counter per barrier slot
  unsigned long long* status;     // local: [0] = number of barrier waits that timed out
  long long spin_limit;           // clock64 ticks a barrier wait may take before it gives up (never hang the GPU)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data is read exactly once per kernel, after the barrier's acquire: relaxed system-scope loads (never a stale L1 line)
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Barrier across the ranks for ONE block: block b of a kernel on rank r meets block b of the same kernel on every other
// rank (per-block slots, no grid-wide dependency, so partially resident grids cannot deadlock).  The slot's epoch only
// grows: a rank that is one barrier ahead leaves a larger value, which still satisfies the slower rank's wait.
// Writes made by this block before the call are visible to the peers' blocks after it (release / acquire at system scope,
// cumulative over the block through bar.sync); callers that wrote peer memory add __threadfence_system() before calling.
__device__ __forceinline__ void comm_barrier(const CommDev& c, int slot) {
  __shared__ uint32_t s_epoch;
  __syncthreads();
  if (threadIdx.x == 0) {
    s_epoch = c.counters[slot] + 1;
    c.counters[slot] = s_epoch;
  }
  __syncthreads();
  const uint32_t e = s_epoch;
  if (threadIdx.x < c.world) {
    const int peer = threadIdx.x;
    uint32_t* theirs = reinterpret_cast<uint32_t*>(c.heap[peer]) + static_cast<size_t>(slot) * kCommMaxWorld + c.rank;
    const uint32_t* mine =
        reinterpret_cast<const uint32_t*>(c.heap[c.rank]) + static_cast<size_t>(slot) * kCommMaxWorld + peer;
    st_release_sys_u32(theirs, e);
    const long long t0 = clock64();
    // fail fast: once a wait has timed out on this rank every later barrier only posts its flag (results are invalid
    // anyway, and a chain of thousands of barriers must not add up to minutes of spinning)
    const bool dead = *reinterpret_cast<volatile unsigned long long*>(c.status) != 0ull;
    while (!dead && static_cast<int32_t>(ld_acquire_sys_u32(mine) - e) < 0) {
      if (clock64() - t0 > c.spin_limit) {
        atomicAdd(c.status, 1ull);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

}  // namespace ssr

// ---- emulated ranks: host-side rendezvous (comm.cu) -------------------------------------------------------------------
struct ssr_comm;
namespace ssr {
// launches the multi-rank kernel of one collective: dev[world], blobs = world argument records `stride` bytes apart
using GroupLauncher = cudaError_t (*)(const CommDev* dev, const unsigned char* blobs, size_t stride, int world,
                                      cudaStream_t st);
bool comm_is_group(const ssr_comm* c);
int comm_group_collective(ssr_comm* c, cudaStream_t stream, const void* tag, const void* args, size_t arg_bytes,
                          GroupLauncher launch);
}  // namespace ssr
