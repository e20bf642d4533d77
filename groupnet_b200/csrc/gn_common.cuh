// Shared device helpers for libgroupnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/groupnet_b200.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "libgroupnet_b200 is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif
#endif

// SMs of the current device (148 on a B200): persistent grids are sized from it at launch time
namespace gn {
static inline int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) return n;
  return 148;
}
}  // namespace gn
#define GN_SM_COUNT (gn::sm_count())
#define GN_THREADS 256

namespace gn {

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// streaming 128-bit load that does not allocate in L1 (data read once)
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based generator.  One 128-bit block per 4 consecutive
// noise elements; element index = ((scene_global*E + e)*T + t), so a scene's
// noise does not depend on how the batch is sharded over GPUs.
// ---------------------------------------------------------------------------
struct Philox {
  static __device__ __forceinline__ uint4 block(uint64_t ctr, uint32_t stream, uint64_t seed) {
    uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32);
    uint32_t c2 = stream, c3 = 0x6e675f62u;  // "gn_b"
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  // uniform in [0,1) with 24 random bits, the same grid torch.rand(float32) uses
  static __device__ __forceinline__ float uniform(uint64_t elem, uint32_t stream, uint64_t seed) {
    uint4 b = block(elem >> 2, stream, seed);
    uint32_t w = (elem & 3) == 0 ? b.x : (elem & 3) == 1 ? b.y : (elem & 3) == 2 ? b.z : b.w;
    return static_cast<float>(w >> 8) * (1.0f / 16777216.0f);
  }
};

// g = -log(eps - log(U + eps))   (MS_HGNN_batch.py:446-455)
__device__ __forceinline__ float gumbel_from_uniform(float u) {
  return -logf(1e-10f - logf(u + 1e-10f));
}

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still draining its last wave; everything that reads the predecessor's output must come after pdl_wait()
// (blocks until the prerequisite grids have completed and their writes are visible).  pdl_trigger() lets the NEXT
// kernel's CTAs be scheduled as soon as SMs free up.  Static data (weights, constants) may be fetched before the wait:
// the kernels' prologues (TMEM allocation, barrier init, constants, first weight chunks) overlap the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t round_up_sz(size_t v, size_t m) { return (v + m - 1) / m * m; }

}  // namespace gn

namespace gn {
// Optional per-kernel CUDA-event timing (gn_profile_enable / gn_profile_collect).
// Disabled by default: the constructor is then a single relaxed load.
struct ProfScope {
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
  const char* name_; cudaStream_t st_; void* rec_;
};
}  // namespace gn

namespace gn {
// <<<grid, block, smem, st>>> with the programmatic-stream-serialization attribute (see pdl_wait above)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
}  // namespace gn

#define GN_LAUNCH_CHECK()                                  \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return static_cast<int>(e__);  \
  } while (0)
