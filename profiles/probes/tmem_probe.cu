// TMEM load / store throughput on one SM (B200): clocks to read (tcgen05.ld) / write (tcgen05.st) a [128 lanes x NCOL]
// fp32 block with 4, 8 or 16 warps (warp w may only touch lane quarter w % 4; warps sharing a quarter split the columns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I groupnet_b200/csrc -o profiles/probes/tmem_probe profiles/probes/tmem_probe.cu
#include <cstdio>
#include <vector>
#include "gn_tf32.cuh"
using namespace gn;

template <int X> __device__ __forceinline__ void ld(uint32_t t, uint32_t* r);
template <> __device__ __forceinline__ void ld<32>(uint32_t t, uint32_t* r) { tc::tmem_ld32_nowait(t, *reinterpret_cast<uint32_t(*)[32]>(r)); }
template <> __device__ __forceinline__ void ld<16>(uint32_t t, uint32_t* r) { tf::tmem_ld16_nowait(t, *reinterpret_cast<uint32_t(*)[16]>(r)); }
template <int X> __device__ __forceinline__ void st(uint32_t t, uint32_t* r);
template <> __device__ __forceinline__ void st<32>(uint32_t t, uint32_t* r) { tf::tmem_st32(t, *reinterpret_cast<uint32_t(*)[32]>(r)); }
template <> __device__ __forceinline__ void st<16>(uint32_t t, uint32_t* r) { tf::tmem_st16(t, *reinterpret_cast<uint32_t(*)[16]>(r)); }

// mode 0: loads only; 1: stores only; 2: load then store of the same columns (a drain)
template <int X, int MODE>
__device__ void run(uint32_t tmem, int nwarps, int ncol, int reps, long long* out, int slot) {
  const int warp = threadIdx.x >> 5, q = warp & 3, sl = warp >> 2, nsl = nwarps / 4;
  const uint32_t base = tmem + (static_cast<uint32_t>(q * 32) << 16);
  uint32_t r[X];
#pragma unroll
  for (int j = 0; j < X; ++j) r[j] = j;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    unsigned acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
      for (int c = sl * X; c < ncol; c += nsl * X) {
        if (MODE != 1) { ld<X>(base + c, r); tc::tmem_ld_wait(); }
#pragma unroll
        for (int j = 0; j < X; ++j) acc += r[j];
        if (MODE != 0) st<X>(base + 256 + c, r);
      }
      if (MODE != 0) tf::tmem_st_wait();
    }
    if (acc == 0x12345678u) out[1023] = acc;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[slot] = (t1 - t0) / reps;
}

__global__ void __launch_bounds__(512, 1) probe(long long* out) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = slot;
  int s = 0;
  for (int nw = 4; nw <= 16; nw *= 2) {
    run<32, 0>(tmem, nw, 128, 8, out, s++); run<16, 0>(tmem, nw, 128, 8, out, s++);
    run<32, 1>(tmem, nw, 128, 8, out, s++); run<16, 1>(tmem, nw, 128, 8, out, s++);
    run<32, 2>(tmem, nw, 128, 8, out, s++); run<16, 2>(tmem, nw, 128, 8, out, s++);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc::fence_after_thread_sync(); tc::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 1024 * sizeof(long long));
  probe<<<1, 512>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> h(1024);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  const char* modes[] = {"ld", "st", "ld+st"};
  int s = 0;
  printf("[128 lanes x 128 columns] fp32 = 64 KB per pass\nwarps shape mode  clk/pass  bytes/clk\n");
  for (int nw = 4; nw <= 16; nw *= 2)
    for (int m = 0; m < 3; ++m)
      for (int x = 32; x >= 16; x /= 2) {
        const long long c = h[s++];
        printf("%5d  x%-3d %-6s %8lld %9.1f\n", nw, x, modes[m], c, (m == 2 ? 131072.0 : 65536.0) / c);
      }
  return 0;
}
