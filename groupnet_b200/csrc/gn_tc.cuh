// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a).
//
// Operand layout used by every tensor-core kernel in this library: the
// canonical K-major, no-swizzle UMMA layout (cute::UMMA::LayoutType::SWIZZLE_NONE,
// "INTERLEAVE").  For an operand with R rows (M or N) and K columns of bf16:
//
//     byte(r, k) = (k / 8) * (R * 16) + r * 16 + (k % 8) * 2
//
// i.e. 8x8 "core matrices" of 128 contiguous bytes, consecutive 8-row groups
// 128 B apart (stride byte offset, SBO) and consecutive 8-column groups R*16 B
// apart (leading byte offset, LBO).  A thread that owns one row writes 16 B per
// k-group and a warp writes 512 contiguous bytes: conflict-free st.shared.v4.
// Weights are pre-arranged in this layout by the host (packing.py), so staging
// them is a flat copy.
#pragma once
#include <cuda_bf16.h>
#include "gn_common.cuh"

namespace gn { namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a CONVERGED warp (elect.sync).  tcgen05.mma / tcgen05.commit take their operands from the uniform
// datapath: issue them from warp-uniform control flow, predicated by this, so descriptors stay in uniform
// registers.  Issued from a divergent `if (tid == k)` branch instead, every MMA costs a register->uniform
// transfer sequence plus an election loop: measured 160 clk (kind::f16) / 287 clk (kind::tf32) per MMA whatever
// its N (profiles/probes/mma_probe.cu), i.e. several times the tensor-pipe time of the small MMAs used here.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout) ----
//  [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1
//  [49,52) base offset = 0 | [52] lbo mode = 0 | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// ---- instruction descriptor, kind::f16: bf16 x bf16 -> f32, both operands K-major ----
//  [4,6) c_format = 1 (f32) | [7,10) a_format = 1 (bf16) | [10,13) b_format = 1 (bf16)
//  [15] a_major = 0 (K) | [16] b_major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM allocation (one full warp) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// issue-only variant: several loads may be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: a descriptor / protocol bug must surface as a CUDA error, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = smem_u32(mbar);
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// one arrival + `bytes` of pending transaction count (completed by the bulk copies that name this barrier)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, no tensor map): 16-byte aligned addresses and size, completion counted in
// bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// pack two floats into one bf16x2 word (lo = a, hi = b), round-to-nearest-even
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// {lo = a, hi = b} -> bf16x2 with ReLU fused into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16_fast(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// byte offset of the 16-byte k-group `k8` of row `r` in an R-row canonical operand
__device__ __forceinline__ uint32_t canon_off(int r, int k8, int R) {
  return static_cast<uint32_t>(k8) * (R * 16) + static_cast<uint32_t>(r) * 16;
}

// Bias through the tensor core: D[128 x N] = ones[128 x 16] * biasB[N x 16]^T where ones has 1.0 in
// k = 0,1 and biasB holds bf16(b) in k = 0 and bf16(b - bf16(b)) in k = 1 (bias exact to ~16 bits).
// The GEMM proper then accumulates on top, and the epilogue needs no bias load / add.
__device__ __forceinline__ void issue_bias(uint32_t tmem_d, uint32_t ones_saddr, uint32_t biasb_saddr, int N) {
  const uint32_t idesc = make_idesc_bf16(128, N);
  uint64_t da = make_smem_desc(ones_saddr, 128 * 16, 128);
  uint64_t db = make_smem_desc(biasb_saddr, static_cast<uint32_t>(N) * 16, 128);
  mma_bf16_ss(tmem_d, da, db, idesc, 0u);
}

// build the [N x 16] bias operand (canonical layout) from fp32 biases; called by all threads of the CTA
__device__ __forceinline__ void build_bias_operand(unsigned char* dst, const float* __restrict__ b, int N,
                                                   int tid, int nthreads) {
  for (int n = tid; n < N; n += nthreads) {
    float v = __ldg(b + n);
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    uint32_t w0 = static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&hi)) |
                  (static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&lo)) << 16);
    *reinterpret_cast<uint4*>(dst + n * 16) = make_uint4(w0, 0u, 0u, 0u);            // k-group 0
    *reinterpret_cast<uint4*>(dst + N * 16 + n * 16) = make_uint4(0u, 0u, 0u, 0u);   // k-group 1
  }
}
__device__ __forceinline__ void build_ones_operand(unsigned char* dst, int tid, int nthreads) {
  for (int r = tid; r < 128; r += nthreads) {
    *reinterpret_cast<uint4*>(dst + r * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);   // bf16 1.0, 1.0
    *reinterpret_cast<uint4*>(dst + 128 * 16 + r * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Issue the K/16 MMAs of one GEMM:  D[128 x N] (+)= A[128 x K] * B[N x K]^T
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr,
                                           int N, int K, bool accumulate_first) {
  const uint32_t idesc = make_idesc_bf16(128, N);
  const uint32_t a_lbo = 128 * 16, b_lbo = static_cast<uint32_t>(N) * 16;
  // one descriptor per operand, then the start-address field (bytes >> 4) advances by two k-groups per MMA:
  // the issuing thread spends ~4 instructions per MMA instead of rebuilding both descriptors
  uint64_t da = make_smem_desc(a_saddr, a_lbo, 128), db = make_smem_desc(b_saddr, b_lbo, 128);
  const uint64_t ia = (2u * a_lbo) >> 4, ib = (2u * b_lbo) >> 4;
  for (int k = 0; k < K; k += 16) {
    mma_bf16_ss(tmem_d, da, db, idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    da += ia; db += ib;
  }
}

// named barrier over the 256 staging / drain threads of the warp-specialised kernels (warps 0-7)
__device__ __forceinline__ void drain_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Coalesced copy of a tile's incidence blocks (ns scenes x N x N fp32, scene stride hstride) into shared memory,
// row stride ldr floats ((N rounded up to 4) + 4: a thread reading "its" row with 128-bit loads does not
// bank-conflict).  Called by the 256 staging threads.
__device__ __forceinline__ void stage_raw_H(float* raw, const float* __restrict__ H, long long hstride,
                                            int b0s, int ns, int N, int ldr, int tid) {
  const int per = N * N;
  // (scene, row, column) of the flat index advance incrementally: no integer division in the loops
  if ((N & 3) == 0 && (hstride & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0) {
    const int n4 = N >> 2, per4 = per >> 2, total = ns * per4;
    int sc = tid / per4, r4 = tid - sc * per4;
    int e = r4 / n4, c4 = r4 - e * n4;
    const int de = 256 / n4, dc = 256 - de * n4;         // 256 float4s further: de rows and dc columns
#pragma unroll 4
    for (int i = tid; i < total; i += 256) {
      *reinterpret_cast<float4*>(raw + (sc * N + e) * ldr + 4 * c4) =
          ldg_f4(H + static_cast<size_t>(b0s + sc) * hstride + 4 * (e * n4 + c4));
      c4 += dc; e += de;
      if (c4 >= n4) { c4 -= n4; ++e; }
      while (e >= N) { e -= N; ++sc; }
    }
  } else {
    const int total = ns * per;
    int sc = tid / per, r = tid - sc * per;
    int e = r / N, n = r - e * N;
    const int de = 256 / N, dn = 256 - de * N;
#pragma unroll 4
    for (int i = tid; i < total; i += 256) {
      raw[(sc * N + e) * ldr + n] = __ldg(H + static_cast<size_t>(b0s + sc) * hstride + e * N + n);
      n += dn; e += de;
      if (n >= N) { n -= N; ++e; }
      while (e >= N) { e -= N; ++sc; }
    }
  }
}

}}  // namespace gn::tc

// Warp-uniform forms for dedicated producer / issuer warps: EVERY lane of the warp calls them from converged control
// flow, one elected lane executes the asynchronous instruction (see tc::elect_one).
namespace gn { namespace tcu {
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, int N, int K,
                                           bool accumulate_first) {
  if (tc::elect_one()) tc::issue_gemm(tmem_d, a_saddr, b_saddr, N, K, accumulate_first);
  __syncwarp();
}
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  if (tc::elect_one()) tc::mma_commit(mbar);
  __syncwarp();
}
__device__ __forceinline__ void expect_tx(uint64_t* mbar, uint32_t bytes) {
  if (tc::elect_one()) tc::mbar_expect_tx(mbar, bytes);
  __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* mbar) {
  if (tc::elect_one()) tc::bulk_g2s(dst_saddr, src, bytes, mbar);
  __syncwarp();
}
}}  // namespace gn::tcu
