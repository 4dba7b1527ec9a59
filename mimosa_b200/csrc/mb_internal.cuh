// Internal definitions shared by the translation units of libmimosa_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mimosa_b200.h"
#include "mb_math.cuh"

namespace mb {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define MB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::mb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MB_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)
#define MB_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t _e = (expr);                                                                      \
    if (_e != ncclSuccess) {                                                                       \
      ::mb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_e)); \
      return MB_ERR_NCCL;                                                                          \
    }                                                                                              \
  } while (0)
#define MB_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != MB_OK) return _s; \
  } while (0)
#define MB_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) {                            \
      ::mb::set_error("%s: %s", __func__, msg); \
      return MB_ERR_INVALID_ARG;              \
    }                                         \
  } while (0)

}  // namespace mb

namespace mb {
// ---- peer-memory exchange of the per-iteration packet (single node, NVLink / NVSwitch) -----------------
// Every rank owns one small cudaMalloc'd block, exported to the other ranks with cudaIpcGetMemHandle:
//   ll     u64 [2 parity][kMaxRanks source][2 * kXchgDoubles]   the packets of all ranks for one exchange
//   xseq   u64                                                  number of exchanges this rank has completed
// A packet is kXchgDoubles doubles: the 48-double reduction packet of a linearisation, the six component
// localizabilities of the last one (entries 48..53), one word for the ranks' device-side barrier (entry 55).  Every
// double travels as two 8-byte words {32 data bits, 32-bit exchange number}, each written by ONE 8-byte store — a
// word is either absent or complete, so the receiver needs neither a fence on the sender's side nor a separate flag
// (the protocol of NCCL's LL transport).  Block 0 of k_icp_loop STORES its packet straight into every rank's block
// (peer stores over NVLink), waits for all ranks' words in its own block (local memory), adds the packets in rank
// order — bit-identical on every rank — and publishes the sum to the other blocks of its GPU.  No collective call,
// no extra kernel.  Two parities: a fast rank may already write exchange n + 1 while a slow one still reads
// exchange n; it cannot reach n + 2 before every rank has consumed n.
constexpr int kMaxRanks = 8;
constexpr int kXchgDoubles = 56;
struct PeerTable {
  int world, rank;
  unsigned long long* ll[kMaxRanks];
  unsigned long long* xseq;
};
constexpr size_t kXchgLlBytes = 2 * kMaxRanks * 2 * kXchgDoubles * sizeof(unsigned long long);
constexpr size_t kXchgBlockBytes = kXchgLlBytes + 256;
}  // namespace mb

// ---- handles ----------------------------------------------------------------------------------------
struct mb_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint64_t launches = 0;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  void* xchg_block = nullptr;           // this rank's mailbox / flags / counter (cudaMalloc, IPC-exported)
  void* xchg_peer[mb::kMaxRanks] = {};  // the other ranks' blocks, opened with cudaIpcOpenMemHandle
  mb::PeerTable* d_peer = nullptr;      // device copy of the table the kernels read; nullptr = use NCCL
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  void* pinned = nullptr;  // page-locked staging for host <-> device copies of scans
  size_t pinned_bytes = 0;
  void* pin_small = nullptr;  // 4 KiB page-locked (device-mapped) block for poses in / normal equations out
  unsigned host_seq = 0;      // sequence number of the last host-polled completion (mb_factor_linearize)
  // Resident linearisation kernel (mb_factor.cu, k_icp_loop in serve mode): after a host-facing linearisation the
  // kernel stays on the device for a short window and takes the next pose from the mapped block below instead of a
  // new launch.  srv_req numbers every host-facing request of this context; srv_live / srv_factor: a kernel that may
  // still be resident and the factor it serves.
  unsigned long long srv_req = 0;
  bool srv_live = false;
  struct mb_factor* srv_factor = nullptr;
  unsigned srv_window_us = 30;  // MB_RESIDENT_US (0: every call is a launch)
  // Size-keyed cache of released device blocks: a factor is created and destroyed for every scan with the
  // same sizes, and cudaMalloc/cudaFree would otherwise dominate the per-scan host cost.
  struct Block {
    void* p;
    size_t bytes;
  };
  std::vector<Block> pool;
  size_t pool_bytes = 0;
  std::unordered_map<void*, size_t> tracked;  // sizes of blocks handed out by dev_alloc_t
};

namespace mb {
// Words of the context's mapped block the resident kernel and the host talk through (byte offsets into pin_small;
// all monotonic request numbers, so nothing ever has to be reset):
constexpr size_t kSrvRec = 3072;   // the request record (host -> device), 4 lines of 64 bytes = 32 words of 8 bytes:
                                   //   lines 0..2: 7 payload words each — the 16 doubles [pose 12 | gravity 3 | lambda]
                                   //   at word i + i / 7 — and the REQUEST NUMBER in word 7 / 15 / 23, written last;
                                   //   word 24: the kernel serving requests <= this number must leave
constexpr size_t kSrvOut = 256;    // the result (device -> host): kSrvOutDoubles doubles [mb_linearization | component
                                   //   localizabilities 6 | pose 12] as flag-in-data words — each double is two 8-byte
                                   //   words {32 data bits, low 32 bits of the REQUEST NUMBER}, every word written by one
                                   //   store: no fence and no separate flag on the device, the host accepts the result
                                   //   when every word shows its request's number
constexpr size_t kSrvOutDoubles = sizeof(mb_linearization) / 8 + 6 + 12;
constexpr size_t kSrvExit = 2560;  // u64: the first request number a departed kernel did NOT serve (device -> host)
static_assert(kSrvOut + 2 * 8 * kSrvOutDoubles <= kSrvExit, "result words overlap the control words");
// The result of request n, if all of it has arrived.
inline bool read_result(const mb_ctx* c, unsigned long long n, unsigned long long* res /* kSrvOutDoubles */) {
  const unsigned long long* w = (const unsigned long long*)((const char*)c->pin_small + kSrvOut);
  const unsigned flag = (unsigned)n;
  if ((unsigned)(__atomic_load_n(w + 2 * kSrvOutDoubles - 1, __ATOMIC_ACQUIRE) >> 32) != flag) return false;
  for (size_t d = 0; d < kSrvOutDoubles; ++d) {
    const unsigned long long lo = __atomic_load_n(w + 2 * d, __ATOMIC_RELAXED), hi = __atomic_load_n(w + 2 * d + 1, __ATOMIC_RELAXED);
    if ((unsigned)(lo >> 32) != flag || (unsigned)(hi >> 32) != flag) return false;
    res[d] = (hi << 32) | (lo & 0xffffffffull);
  }
  return true;
}
// Ask a resident kernel to leave (it would on its own once its window closes): everything enqueued on the context's
// stream afterwards then starts without that delay.
inline void server_stop(mb_ctx* c) {
  if (c->srv_live) {
    __atomic_store_n((unsigned long long*)((char*)c->pin_small + kSrvRec) + 24, c->srv_req, __ATOMIC_RELEASE);
    c->srv_live = false;
  }
}
// Grow-only page-locked staging buffer of the context (synchronises the stream when it has to grow).
int pinned_reserve(mb_ctx* c, size_t bytes);
// Pooled device allocations (exact-size reuse).  dev_free never synchronises: the caller guarantees that no
// work touching the block is still pending on the context's stream.
int dev_alloc(mb_ctx* c, void** p, size_t bytes);
void dev_free(mb_ctx* c, void* p, size_t bytes);
void dev_pool_release(mb_ctx* c);
// true when `p` points into page-locked host memory the device can read by DMA (cudaHostAlloc / cudaHostRegister)
inline bool host_is_page_locked(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
// Same, with the size remembered per pointer (for owners that do not keep it).
template <typename T>
inline int dev_alloc_t(mb_ctx* c, T** p, size_t bytes) {
  void* v = nullptr;
  const int rc = dev_alloc(c, &v, bytes);
  *p = (T*)v;
  if (rc == MB_OK) c->tracked[v] = bytes;
  return rc;
}
inline void dev_free_t(mb_ctx* c, void* p) {
  if (!p) return;
  auto it = c->tracked.find(p);
  if (it == c->tracked.end()) {
    cudaFree(p);
    return;
  }
  const size_t bytes = it->second;
  c->tracked.erase(it);
  dev_free(c, p, bytes);
}
}

#include "mb_search.cuh"

namespace mb {

#if defined(__CUDACC__)
// Returns the packed (slot << 5 | count) word of the voxel at (x,y,z), or kEmpty.
__device__ __forceinline__ uint32_t table_find(const int4* __restrict__ table, uint32_t mask, int x, int y, int z) {
  uint32_t h = hash_coord(x, y, z) & mask;
  while (true) {
    const int4 e = __ldg(table + h);
    if ((uint32_t)e.w == kEmpty) return kEmpty;
    if (e.x == x && e.y == y && e.z == z) return (uint32_t)e.w;
    h = (h + 1) & mask;
  }
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may start while its predecessor
// on the stream is still running; pdl_wait() blocks until the predecessor has completed and its writes are
// visible, pdl_launch_dependents() lets the successor's blocks be scheduled early.  Both are no-ops for a
// plain launch.  This hides the ~3.6 us launch-to-launch gap between the small dependent kernels of one ICP
// iteration.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif  // __CUDACC__

}  // namespace mb
