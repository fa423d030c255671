// Ensemble-sharding exchange over NVLink 5 / NVSwitch peer memory: no NCCL launch on the update's critical
// path.  Payloads are a few KB (the target critics' Q rows, dL/da), so the exchange is pure latency: an NCCL all-gather
// captured in the update graph costs ~30-50 us per step (profiles/r1_08, r1_09); here every rank STORES its rows straight
// into all peers' symmetric buffers (the buffers' peer mappings come from torch.distributed._symmetric_memory: plumbing)
// and raises a per-source signal; the consumer spins on its own signals and gathers the rows into a fixed local tensor.
//
//   put  : one block copies `nbytes` from src to offset dst_off of every peer's buffer half (epoch parity) with 16-byte
//          peer stores, fences at system scope, then thread p raises sig[p][rank] = epoch (st.release.sys)
//   wait : thread p spins (ld.acquire.sys) until sig[rank][p] >= epoch, then the block copies the whole buffer half to
//          `out` (a tensor whose address is the same on every replay of a captured graph)
//
// Double buffering by epoch parity: a rank can only reach put(k+2) after it has seen every peer's put(k+1), which each
// peer issues (stream order) after the kernels that read exchange k have finished.  Epochs live in device memory, so
// the same two kernels replay inside CUDA graphs.  sm_100a.
#include "ssac_common.cuh"

namespace ssac {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// peer_bufs[p] / peer_sigs[p]: this rank's mappings of rank p's symmetric buffer (2 halves of half_bytes) / signal row
__global__ void __launch_bounds__(1024) peer_put_kernel(const uint8_t* __restrict__ src, int64_t nbytes, int64_t dst_off,
                                                        int64_t half_bytes, uint8_t* const* __restrict__ peer_bufs,
                                                        uint32_t* const* __restrict__ peer_sigs, int rank, int world,
                                                        uint32_t* __restrict__ epoch) {
  pdl_wait();
  pdl_trigger();
  const uint32_t e = *epoch + 1u;
  const int64_t base = (int64_t)(e & 1u) * half_bytes + dst_off;
  if (((nbytes | dst_off | half_bytes) & 15) == 0 && (((uintptr_t)src) & 15) == 0) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
    const int64_t n4 = nbytes >> 4;
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const int4 v = s4[i];
      for (int p = 0; p < world; ++p) reinterpret_cast<int4*>(peer_bufs[p] + base)[i] = v;
    }
  } else {
    for (int64_t i = threadIdx.x; i < nbytes; i += blockDim.x) {
      const uint8_t v = src[i];
      for (int p = 0; p < world; ++p) peer_bufs[p][base + i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) st_release_sys(peer_sigs[threadIdx.x] + rank, e);
  if (threadIdx.x == 0) *epoch = e;
}

__global__ void __launch_bounds__(1024) peer_wait_kernel(const uint8_t* __restrict__ my_buf, int64_t half_bytes,
                                                         int64_t nbytes, const uint32_t* __restrict__ my_sigs, int world,
                                                         const uint32_t* __restrict__ epoch, uint8_t* __restrict__ out,
                                                         const int32_t* __restrict__ row_index, int n_rows, int64_t row_bytes) {
  pdl_wait();
  pdl_trigger();
  const uint32_t e = *epoch;   // the put of this exchange (same stream, earlier) stored it
  if ((int)threadIdx.x < world) {
    // bounded (tens of seconds): a peer that never arrives -- a crashed rank, a schedule that deadlocks -- traps this kernel
    // instead of hanging the GPU
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(my_sigs + threadIdx.x) - e) < 0) {
      __nanosleep(100);
      if (++spins > 200000000u) __trap();
    }
  }
  __syncthreads();
  const uint8_t* half = my_buf + (int64_t)(e & 1u) * half_bytes;
  if (row_index != nullptr) {
    // gather a subset of the rows (the REDQ target subset: row_index[m] names a global critic)
    for (int m = 0; m < n_rows; ++m) {
      const uint8_t* srow = half + (int64_t)row_index[m] * row_bytes;
      uint8_t* orow = out + (int64_t)m * row_bytes;
      if ((row_bytes & 15) == 0 && (((uintptr_t)out) & 15) == 0 && (half_bytes & 15) == 0) {
        for (int64_t i = threadIdx.x; i < (row_bytes >> 4); i += blockDim.x)
          reinterpret_cast<int4*>(orow)[i] = __ldcv(reinterpret_cast<const int4*>(srow) + i);
      } else {
        for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) orow[i] = __ldcv(srow + i);
      }
    }
    return;
  }
  if (((nbytes | half_bytes) & 15) == 0 && (((uintptr_t)out) & 15) == 0) {
    const int4* s4 = reinterpret_cast<const int4*>(half);
    int4* o4 = reinterpret_cast<int4*>(out);
    for (int64_t i = threadIdx.x; i < (nbytes >> 4); i += blockDim.x) o4[i] = __ldcv(s4 + i);   // peers wrote it: no stale cache lines
  } else {
    for (int64_t i = threadIdx.x; i < nbytes; i += blockDim.x) out[i] = __ldcv(half + i);
  }
}

}  // namespace ssac

using namespace ssac;

extern "C" {

int ssac_peer_put(const void* src_dev, int64_t nbytes, int64_t dst_off, int64_t half_bytes, void* const* peer_bufs_dev,
                  uint32_t* const* peer_sigs_dev, int rank, int world, uint32_t* epoch_dev, void* stream) {
  SSAC_REQUIRE(src_dev && peer_bufs_dev && peer_sigs_dev && epoch_dev, "ssac_peer_put: null pointer");
  SSAC_REQUIRE(nbytes > 0 && dst_off >= 0 && dst_off + nbytes <= half_bytes && world > 0 && world <= 64 && rank >= 0 && rank < world,
               "ssac_peer_put: bad sizes");
  launch_pdl(peer_put_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, (const uint8_t*)src_dev, nbytes, dst_off, half_bytes,
             (uint8_t* const*)peer_bufs_dev, peer_sigs_dev, rank, world, epoch_dev);
  SSAC_CHECK_LAUNCH("ssac_peer_put");
  return 0;
}

int ssac_peer_wait(const void* my_buf_dev, int64_t half_bytes, int64_t nbytes, const uint32_t* my_sigs_dev, int world,
                   const uint32_t* epoch_dev, void* out_dev, const int32_t* row_index_dev, int n_rows, int64_t row_bytes,
                   void* stream) {
  SSAC_REQUIRE(my_buf_dev && my_sigs_dev && epoch_dev && out_dev, "ssac_peer_wait: null pointer");
  SSAC_REQUIRE(nbytes > 0 && nbytes <= half_bytes && world > 0 && world <= 64, "ssac_peer_wait: bad sizes");
  SSAC_REQUIRE(!row_index_dev || (n_rows > 0 && row_bytes > 0), "ssac_peer_wait: a row subset needs n_rows and row_bytes");
  launch_pdl(peer_wait_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, (const uint8_t*)my_buf_dev, half_bytes, nbytes,
             my_sigs_dev, world, epoch_dev, (uint8_t*)out_dev, row_index_dev, n_rows, row_bytes);
  SSAC_CHECK_LAUNCH("ssac_peer_wait");
  return 0;
}

}  // extern "C"
