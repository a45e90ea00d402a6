// Device-initiated all-gather over NVLink peer mappings (pf_allgather_views): the one collective of the view-sharded
// denoise step — the projected K|V of the local views before every EPPA block, and the eps outputs at the
// end of the step — as ONE kernel that is capturable in the step's CUDA graph.
//
// Replaces the host-issued ncclAllGather between graph segments of round 1. The reference has no counterpart (Lightning
// DDP over prompts, main.py:63; `predict` uses no collective): this is the exchange the view partition of
// models/pano/modules.py:44-48 needs (panorama queries attend to the keys / values of ALL views).
//
// Every rank owns, per call site, a receive buffer [nranks][slice_bytes] and nranks flag words, both mapped into every
// peer (CUDA IPC). The kernel (a) copies the local slice into slot `rank` of every peer's buffer with 16-byte stores
// over NVLink, (b) after all CTAs are done, publishes epoch e in flag `rank` of every peer (release, system scope),
// (c) waits until all nranks local flags have reached e (acquire, system scope). The epoch lives in device memory and is
// advanced by the kernel itself, so a replayed graph keeps working. A buffer is only rewritten one whole step later, and
// a rank can never be more than one collective ahead of a peer (each collective is also a barrier), so per-site buffers
// need no further flow control.
#include <string.h>

#include "pf_common.cuh"

namespace pf {

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct AllGatherParams {
  const uint4* local;          // slice_bytes, 16-byte aligned
  long long slice_vecs;        // slice_bytes / 16
  void* const* peer_data;      // [nranks] device array: base of every rank's receive buffer (own included)
  uint32_t* const* peer_flags; // [nranks] device array: base of every rank's flag words
  uint32_t* my_flags;          // [nranks] == peer_flags[rank]
  uint32_t* state;             // [0] epoch of the last completed call, [1] CTA arrival counter (zero between calls)
  int rank, nranks;
  long long timeout_cycles;
};

__global__ void __launch_bounds__(256) allgather_push_wait_kernel(const AllGatherParams p) {
  // (a) push: CTA c copies vectors [c*per, (c+1)*per) of the local slice to every peer
  const long long per = (p.slice_vecs + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * per;
  const long long v1 = min(p.slice_vecs, v0 + per);
  for (int pr = 0; pr < p.nranks; ++pr) {
    const int dst_rank = (p.rank + pr) % p.nranks;  // start with the own buffer, stagger the peers
    uint4* dst = static_cast<uint4*>(p.peer_data[dst_rank]) + (long long)p.rank * p.slice_vecs;
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) dst[v] = __ldg(p.local + v);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.state[1], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  // (b) the last CTA of this rank publishes the new epoch everywhere, (c) then waits for every rank's flag
  __threadfence_system();
  const uint32_t epoch = p.state[0] + 1u;
  if (threadIdx.x < p.nranks) st_release_sys_u32(p.peer_flags[threadIdx.x] + p.rank, epoch);
  if (threadIdx.x < p.nranks) {
    const long long t0 = clock64();
    // flags only grow; a peer that is one collective ahead on ANOTHER site cannot touch this site's words
    while ((int)(ld_acquire_sys_u32(p.my_flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(100);
      if (clock64() - t0 > p.timeout_cycles) {
        printf("pf_allgather_views: rank %d timed out waiting for rank %d (epoch %u, flag %u)\n", p.rank,
               (int)threadIdx.x, epoch, ld_acquire_sys_u32(p.my_flags + threadIdx.x));
        __trap();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.state[0] = epoch;
    p.state[1] = 0u;
  }
}

}  // namespace pf

// ---- receive-buffer plumbing: plain cudaMalloc memory exported / imported with CUDA IPC -------------------------------------
// (the caller's PyTorch allocator is not involved: an IPC handle names a whole cudaMalloc allocation, and the importing side
// must open it on ITS device with lazy peer access, which is how NCCL's P2P transport maps peer buffers too)
extern "C" int pf_comm_alloc(long long bytes, void** ptr) {
  using namespace pf;
  PF_CHECK_ARG(bytes > 0 && ptr, "pf_comm_alloc: bad arguments");
  void* p = nullptr;
  if (int rc = check_cuda(cudaMalloc(&p, (size_t)bytes), "cudaMalloc(comm buffer)")) return rc;
  if (int rc = check_cuda(cudaMemset(p, 0, (size_t)bytes), "cudaMemset(comm buffer)")) return rc;
  if (int rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return rc;
  *ptr = p;
  return PF_OK;
}

extern "C" int pf_comm_free(void* ptr) {
  return pf::check_cuda(cudaFree(ptr), "cudaFree(comm buffer)");
}

extern "C" int pf_ipc_export(const void* ptr, unsigned char handle[64]) {
  using namespace pf;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  PF_CHECK_ARG(ptr && handle, "pf_ipc_export: null pointer");
  cudaIpcMemHandle_t h;
  if (int rc = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)), "cudaIpcGetMemHandle")) return rc;
  memcpy(handle, &h, 64);
  return PF_OK;
}

extern "C" int pf_ipc_open(const unsigned char handle[64], void** ptr) {
  using namespace pf;
  PF_CHECK_ARG(ptr && handle, "pf_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  if (int rc = check_cuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return rc;
  *ptr = p;
  return PF_OK;
}

extern "C" int pf_ipc_close(void* ptr) {
  return pf::check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
}

extern "C" int pf_enable_peer_access(int peer_device) {
  using namespace pf;
  int dev = -1;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (peer_device == dev) return PF_OK;
  int can = 0;
  if (int rc = check_cuda(cudaDeviceCanAccessPeer(&can, dev, peer_device), "cudaDeviceCanAccessPeer")) return rc;
  if (!can) {
    set_error("pf_enable_peer_access: device %d cannot access device %d (no NVLink / PCIe peer path)", dev, peer_device);
    return PF_ERR_UNSUPPORTED;
  }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky "already enabled" status
    return PF_OK;
  }
  return check_cuda(e, "cudaDeviceEnablePeerAccess");
}

extern "C" int pf_allgather_views(const void* local, long long slice_bytes, void* const* peer_data,
                                  unsigned int* const* peer_flags, unsigned int* my_flags, unsigned int* state, int rank,
                                  int nranks, void* stream) {
  using namespace pf;
  PF_CHECK_ARG(local && peer_data && peer_flags && my_flags && state, "pf_allgather_views: null pointer");
  PF_CHECK_ARG(nranks >= 1 && nranks <= 64 && rank >= 0 && rank < nranks, "pf_allgather_views: bad rank %d of %d", rank,
               nranks);
  PF_CHECK_ARG(slice_bytes > 0 && slice_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(local) & 15) == 0,
               "pf_allgather_views: the slice must be a non-empty multiple of 16 bytes, 16-byte aligned");
  AllGatherParams p;
  p.local = static_cast<const uint4*>(local);
  p.slice_vecs = slice_bytes / 16;
  p.peer_data = peer_data;
  p.peer_flags = reinterpret_cast<uint32_t* const*>(peer_flags);
  p.my_flags = my_flags;
  p.state = state;
  p.rank = rank;
  p.nranks = nranks;
  p.timeout_cycles = 30000000000LL;  // ~15 s at 1.9 GHz: a missing peer aborts the kernel instead of hanging the GPU
  // enough CTAs to drive NVLink (a slice is 0.3 - 2.6 MB), few enough to leave the SMs to the compute stream
  long long ctas = (p.slice_vecs + 4095) / 4096;  // >= 64 KB per CTA
  if (ctas < 1) ctas = 1;
  if (ctas > 32) ctas = 32;
  allgather_push_wait_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  PF_CHECK_LAUNCH("allgather_push_wait_kernel");
  return PF_OK;
}
