// Second MMA probe: fully unrolled, warp-uniform issue (elect.sync once per batch), compile-time shapes — the
// per-MMA cost of the tensor pipe itself, without loop or descriptor overhead in the issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I groupnet_b200/csrc -o profiles/probes/mma_probe2 profiles/probes/mma_probe2.cu
#include <cstdio>
#include <vector>
#include "gn_tf32.cuh"
using namespace gn;

template <int KIND, int N, int TS, int NACC>
__device__ __forceinline__ void batch(uint32_t tmem, uint32_t sa, uint32_t sb, uint64_t* bar, uint32_t& ph, long long* out) {
  using namespace tc;
  constexpr int NM = 64;
  const uint32_t idesc = KIND == 0 ? tf::make_idesc_tf32(128, N) : make_idesc_bf16(128, N);
  const uint64_t da = make_smem_desc(sa, 128 * 16, 128), db = make_smem_desc(sb, N * 16, 128);
  long long best_i = 1LL << 60, best_t = 1LL << 60;
  for (int rep = 0; rep < 4; ++rep) {
    __syncwarp();
    const long long t0 = clock64();
    if (elect_one()) {
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        const uint32_t d = tmem + (i % NACC) * N;
        if (KIND == 0) {
          if (TS) tf::mma_tf32_ts(d, tmem + 448 + (i & 7) * 8, db, idesc, i >= NACC ? 1u : 0u);
          else tf::mma_tf32_ss(d, da + ((i & 7) * 256), db + ((i & 7) * ((2 * N * 16) >> 4)), idesc, i >= NACC ? 1u : 0u);
        } else {
          mma_bf16_ss(d, da + ((i & 7) * 256), db + ((i & 7) * ((2 * N * 16) >> 4)), idesc, i >= NACC ? 1u : 0u);
        }
      }
    }
    __syncwarp();
    const long long t1 = clock64();
    if (elect_one()) mma_commit(bar);
    __syncwarp();
    mbar_wait(bar, ph); ph ^= 1;
    const long long t2 = clock64();
    if (t1 - t0 < best_i) best_i = t1 - t0;
    if (t2 - t0 < best_t) best_t = t2 - t0;
  }
  if ((threadIdx.x & 31) == 0) { out[0] = best_i; out[1] = best_t; out[2] = NM; out[3] = KIND * 1000000 + TS * 100000 + NACC * 1000 + N; }
}

__global__ void __launch_bounds__(192, 1) probe(long long* out) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < (128 * 64 * 4 + 256 * 64 * 4) / 16; i += 192) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(&bar, 1);
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = slot, sa = smem_u32(smem), sb = sa + 128 * 64 * 4;
  if (tid >> 5 == 5) {
    uint32_t ph = 0;
    int c = 0;
#define RUN(K, N, TS, NA) batch<K, N, TS, NA>(tmem, sa, sb, &bar, ph, out + 4 * (c++));
    RUN(0, 16, 0, 1) RUN(0, 32, 0, 1) RUN(0, 64, 0, 1) RUN(0, 128, 0, 1) RUN(0, 256, 0, 1)
    RUN(0, 64, 0, 2) RUN(0, 128, 0, 2) RUN(0, 64, 0, 4)
    RUN(0, 16, 1, 1) RUN(0, 64, 1, 1) RUN(0, 128, 1, 1) RUN(0, 256, 1, 1) RUN(0, 128, 1, 2)
    RUN(1, 16, 0, 1) RUN(1, 32, 0, 1) RUN(1, 64, 0, 1) RUN(1, 128, 0, 1) RUN(1, 256, 0, 1)
    RUN(1, 64, 0, 2) RUN(1, 128, 0, 2) RUN(1, 64, 0, 4)
    if ((tid & 31) == 0) out[4 * c] = -1;
  }
  fence_before_thread_sync();
  __syncthreads();
  if (tid < 32) { fence_after_thread_sync(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 4096 * sizeof(long long));
  cudaMemset(dout, 0, 4096 * sizeof(long long));
  const int smem = 128 * 64 * 4 + 256 * 64 * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 192, smem>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> h(4096);
  cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("kind mode nacc    N | issue clk/MMA | total clk/MMA | MAC/clk (64 unrolled MMAs, M = 128, best of 4)\n");
  for (int c = 0; h[4 * c] != -1 && c < 1000; ++c) {
    const long long id = h[4 * c + 3];
    const int kind = id / 1000000, ts = (id / 100000) % 10, nacc = (id / 1000) % 100, N = id % 1000;
    const double per = double(h[4 * c + 1]) / h[4 * c + 2];
    printf("%-4s %-4s %4d %4d | %8.1f      | %8.1f      | %8.0f\n", kind ? "bf16" : "tf32", ts ? "TS" : "SS", nacc, N,
           double(h[4 * c]) / h[4 * c + 2], per, 128.0 * N * (kind ? 16 : 8) / per);
  }
  return 0;
}
