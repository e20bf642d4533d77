// Microbenchmark of tcgen05.mma issue / completion on one SM (B200): clocks per MMA for chains that accumulate into
// the same TMEM accumulator vs. several accumulators, SS vs TS operands, kind::tf32 vs kind::f16, N = 16..256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I groupnet_b200/csrc -o profiles/probes/mma_probe profiles/probes/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "gn_tf32.cuh"

using namespace gn;

struct Cfg { int kind; int ts; int N; int nacc; int nmma; int acc_stride; int uniform; };   // kind 0 = tf32, 1 = bf16; uniform: whole warp + elect.sync

__global__ void __launch_bounds__(192, 1) probe(const Cfg* cfgs, int ncfg, long long* out) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  // A [128 x 64] and B [256 x 64] 32-bit canonical operands (zeros are fine for timing)
  for (int i = tid; i < (128 * 64 * 4 + 256 * 64 * 4) / 16; i += 192) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(&bar, 1);
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = slot, sa = smem_u32(smem), sb = sa + 128 * 64 * 4;
  if (tid >> 5 == 5) {
    // warp-uniform issue: every lane runs the loop, one elected lane issues (descriptors stay in uniform registers)
    uint32_t ph = 0;
    for (int c = 0; c < ncfg; ++c) {
      const Cfg cf = cfgs[c];
      if (!cf.uniform) continue;
      for (int rep = 0; rep < 3; ++rep) {
        const uint32_t idesc = cf.kind == 0 ? tf::make_idesc_tf32(128, cf.N) : make_idesc_bf16(128, cf.N);
        const uint64_t da = make_smem_desc(sa, 128 * 16, 128), db = make_smem_desc(sb, cf.N * 16, 128);
        __syncwarp();
        long long t0 = clock64();
        for (int i = 0; i < cf.nmma; ++i) {
          const uint32_t d = tmem + (cf.nacc == 1 ? 0 : (i & (cf.nacc - 1)) * cf.acc_stride);
          const uint32_t accum = i >= cf.nacc ? 1u : 0u;
          if (elect_one()) {
            if (cf.kind == 0) {
              if (cf.ts) tf::mma_tf32_ts(d, tmem + 448 + (i & 7) * 8, db, idesc, accum);
              else tf::mma_tf32_ss(d, da, db, idesc, accum);
            } else {
              mma_bf16_ss(d, da, db, idesc, accum);
            }
          }
        }
        long long t1 = clock64();
        if (elect_one()) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, ph); ph ^= 1;
        long long t2 = clock64();
        if ((tid & 31) == 0) {
          out[(c * 3 + rep) * 2] = t1 - t0;
          out[(c * 3 + rep) * 2 + 1] = t2 - t0;
        }
      }
    }
  }
  __syncthreads();
  if (tid == 128) {
    uint32_t ph = 0;
    for (int c = 0; c < ncfg; ++c) if (cfgs[c].uniform) ph ^= 1;      // three reps each flip the phase: net one flip
    for (int c = 0; c < ncfg; ++c) {
      const Cfg cf = cfgs[c];
      if (cf.uniform) continue;
      for (int rep = 0; rep < 3; ++rep) {
        const uint32_t idesc = cf.kind == 0 ? tf::make_idesc_tf32(128, cf.N) : make_idesc_bf16(128, cf.N);
        const uint64_t da = make_smem_desc(sa, 128 * 16, 128), db = make_smem_desc(sb, cf.N * 16, 128);
        long long t0 = clock64();
        for (int i = 0; i < cf.nmma; ++i) {
          const uint32_t d = tmem + (i % cf.nacc) * cf.acc_stride;
          const uint32_t accum = i >= cf.nacc ? 1u : 0u;
          if (cf.kind == 0) {
            if (cf.ts) tf::mma_tf32_ts(d, tmem + 448 + (i & 7) * 8, db, idesc, accum);
            else tf::mma_tf32_ss(d, da, db, idesc, accum);
          } else {
            mma_bf16_ss(d, da, db, idesc, accum);
          }
        }
        long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, ph); ph ^= 1;
        long long t2 = clock64();
        out[(c * 3 + rep) * 2] = t1 - t0;
        out[(c * 3 + rep) * 2 + 1] = t2 - t0;
      }
    }
  }
  fence_before_thread_sync();
  __syncthreads();
  if (tid < 32) { fence_after_thread_sync(); tmem_dealloc(tmem, 512); }
}

int main() {
  std::vector<Cfg> cfgs;
  const int Ns[] = {16, 64, 128, 256};
  for (int kind = 0; kind < 2; ++kind)
    for (int N : Ns)
      for (int nacc : {1, 2, 4}) {
        if (N * nacc > 448) continue;
        if (nacc == 1) cfgs.push_back({kind, 0, N, nacc, 48, N, 0});
        cfgs.push_back({kind, 0, N, nacc, 48, N, 1});
      }
  for (int N : Ns) cfgs.push_back({0, 1, N, 1, 48, N, 1});          // TS, same accumulator
  cfgs.push_back({0, 1, 128, 2, 48, 128, 1});
  cfgs.push_back({0, 0, 128, 1, 192, 128, 1});                      // longer chain
  cfgs.push_back({0, 0, 256, 1, 192, 256, 1});
  cfgs.push_back({1, 0, 256, 1, 192, 256, 1});
  cfgs.push_back({0, 0, 128, 1, 8, 128, 1});                        // short chain (fixed costs)
  cfgs.push_back({0, 0, 128, 1, 1, 128, 1});
  Cfg* dc; long long* dout;
  cudaMalloc(&dc, cfgs.size() * sizeof(Cfg));
  cudaMalloc(&dout, cfgs.size() * 6 * sizeof(long long));
  cudaMemcpy(dc, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  const int smem = 128 * 64 * 4 + 256 * 64 * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 192, smem>>>(dc, static_cast<int>(cfgs.size()), dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> h(cfgs.size() * 6);
  cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("kind   mode issue      N   nacc nmma | issue clk/MMA | total clk/MMA  (best of 3)   ideal N/2 (bf16) or N (tf32?)\n");
  for (size_t c = 0; c < cfgs.size(); ++c) {
    long long bi = 1LL << 60, bt = 1LL << 60;
    for (int r = 0; r < 3; ++r) { bi = std::min(bi, h[(c * 3 + r) * 2]); bt = std::min(bt, h[(c * 3 + r) * 2 + 1]); }
    printf("%-6s %-4s %-8s %4d %3d %5d | %8.1f      | %8.1f\n", cfgs[c].kind ? "bf16" : "tf32", cfgs[c].ts ? "TS" : "SS",
           cfgs[c].uniform ? "elect" : "tid==k", cfgs[c].N, cfgs[c].nacc, cfgs[c].nmma, double(bi) / cfgs[c].nmma, double(bt) / cfgs[c].nmma);
  }
  return 0;
}
