// 3xTF32 ("fp32-grade") tensor-core primitives on tcgen05 / TMEM (inline PTX, sm_100a).
//
// The reference computes every Linear in fp32 (model/MS_HGNN_batch.py:201-229, torch addmm).  A plain
// kind::tf32 MMA keeps 11 significand bits (~1e-3), so every fp32 operand x is split into two tf32 words
//
//     hi = round_to_tf32(x)            (11 significand bits, low 13 bits of the fp32 word zero)
//     lo = round_to_tf32(x - hi)       (x - hi is exact in fp32; rounding it keeps the split unbiased)
//
// and a product A * B is issued as THREE kind::tf32 MMAs into one fp32 TMEM accumulator:
//
//     D += A_lo * B_hi;   D += A_hi * B_lo;   D += A_hi * B_hi          (A_lo * B_lo ~ 2^-22 is dropped)
//
// which carries ~21 significand bits per product (a CPU emulation of this plan: tests/test_tf32_cpu.py; the GPU
// parity tests hold the layer outputs to the 1e-5 * max|ref| bound of the fp32 path).
//
// Operand layouts
//   smem (SS mode, and every B operand): canonical K-major no-swizzle layout with 32-bit elements,
//       byte(r, k) = (k / 4) * (R * 16) + r * 16 + (k % 4) * 4          (k-groups of 4, LBO = R*16, SBO = 128)
//   TMEM (TS mode, A operand): row r of the 128-row tile = TMEM lane r, element k = column a_col + k
//       (one 32-bit column per tf32 element); hi and lo copies live in separate column ranges.
// One kind::tf32 MMA consumes K = 8 (32 bytes of K, as every tcgen05 kind does).
#pragma once
#include "gn_tc.cuh"

namespace gn { namespace tf {

// kind::tf32 instruction descriptor: tf32 x tf32 -> f32, both operands K-major
//  [4,6) c_format = 1 (f32) | [7,10) a_format = 2 (tf32) | [10,13) b_format = 2 | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T   (A: lane = row, one column per k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      :: "r"(taddr),
         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
         "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
         "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :: "r"(taddr),
         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2, IEEE per lane): one issue slot per two columns.  The row threads of
// the chains are bound by issue slots and dependent-instruction latency, not by the FMA pipe alone.
__device__ __forceinline__ unsigned long long pk2u(uint32_t lo, uint32_t hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long pk2f(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2u(unsigned long long v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ void upk2f(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// x ~ hi + lo for an ACTIVATION operand (measured on the goldens, profiles/tf32_error_margin.py: worst max|d| / max|ref|
// 2.5e-6 with this split against 2.2e-6 with nearest rounding of both halves (-DGN_TF32_SPLIT_RN), bar 1e-5; the node
// chains run 8-10 % faster for the three instructions saved per staged value)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
#ifdef GN_TF32_SPLIT_RN
  const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  hi = h;
  lo = (__float_as_uint(x - __uint_as_float(h)) + 0x1000u) & 0xFFFFE000u;
#else
  // ACTIVATION split, two instructions instead of five: hi = x truncated to tf32 (exactly representable, so the tensor
  // core's treatment of the low 13 bits does not matter), lo = x - hi exactly (<= 13 significant bits, of which the
  // tensor core keeps the top 11: |x - hi - lo_used| < 2^-21 |x|, twice the bound of the nearest-rounding split and
  // still under the dropped lo * lo term's order).  Weights are split with round-to-nearest on the host.
  const uint32_t h = __float_as_uint(x) & 0xFFFFE000u;
  hi = h;
  lo = __float_as_uint(x - __uint_as_float(h));
#endif
}

// byte offset of the 16-byte k-group `k4` (4 tf32 elements) of row `r` in an R-row canonical operand
__device__ __forceinline__ uint32_t canon_off32(int r, int k4, int R) {
  return static_cast<uint32_t>(k4) * (R * 16) + static_cast<uint32_t>(r) * 16;
}

// Three-term product over K (multiple of 8) with A in shared memory:
//   D[128 x N] (+)= (A_hi + A_lo)[128 x K] * (B_hi + B_lo)[N x K]^T   minus the lo*lo term
// a_*: canonical 128-row operands; b_*: canonical N-row operands (one weight chunk).
// STEPS > 0: K = 8 * STEPS known at compile time (fully unrolled: the issuing lane then spends ~10 clk per MMA instead
// of ~50 in the rolled loop, which matters for N <= 64 where an MMA occupies the tensor pipe for only N / 2 clk).
template <int STEPS>
__device__ __forceinline__ void issue_x3_ss_n(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                              int N, int K, bool accumulate_first) {
  const uint32_t idesc = make_idesc_tf32(128, N);
  const uint32_t a_lbo = 128 * 16, b_lbo = static_cast<uint32_t>(N) * 16;
  uint64_t dah = tc::make_smem_desc(a_hi, a_lbo, 128), dal = tc::make_smem_desc(a_lo, a_lbo, 128);
  uint64_t dbh = tc::make_smem_desc(b_hi, b_lbo, 128), dbl = tc::make_smem_desc(b_lo, b_lbo, 128);
  const uint64_t ia = (2u * a_lbo) >> 4, ib = (2u * b_lbo) >> 4;      // two k-groups (K = 8) per MMA
  const int steps = STEPS > 0 ? STEPS : (K >> 3);
#pragma unroll
  for (int i = 0; i < steps; ++i) {
    mma_tf32_ss(tmem_d, dal, dbh, idesc, (i > 0 || accumulate_first) ? 1u : 0u);
    mma_tf32_ss(tmem_d, dah, dbl, idesc, 1u);
    mma_tf32_ss(tmem_d, dah, dbh, idesc, 1u);
    dah += ia; dal += ia; dbh += ib; dbl += ib;
  }
}
__device__ __forceinline__ void issue_x3_ss(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            int N, int K, bool accumulate_first) {
  if (K == 64) issue_x3_ss_n<8>(tmem_d, a_hi, a_lo, b_hi, b_lo, N, K, accumulate_first);
  else if (K == 128) issue_x3_ss_n<16>(tmem_d, a_hi, a_lo, b_hi, b_lo, N, K, accumulate_first);
  else issue_x3_ss_n<0>(tmem_d, a_hi, a_lo, b_hi, b_lo, N, K, accumulate_first);
}

// Same with A in tensor memory: ta_hi / ta_lo = TMEM addresses (lane 0 | column) of element k = 0
template <int STEPS>
__device__ __forceinline__ void issue_x3_ts_n(uint32_t tmem_d, uint32_t ta_hi, uint32_t ta_lo, uint32_t b_hi, uint32_t b_lo,
                                              int N, int K, bool accumulate_first) {
  const uint32_t idesc = make_idesc_tf32(128, N);
  const uint32_t b_lbo = static_cast<uint32_t>(N) * 16;
  uint64_t dbh = tc::make_smem_desc(b_hi, b_lbo, 128), dbl = tc::make_smem_desc(b_lo, b_lbo, 128);
  const uint64_t ib = (2u * b_lbo) >> 4;
  const int steps = STEPS > 0 ? STEPS : (K >> 3);
#pragma unroll
  for (int i = 0; i < steps; ++i) {
    mma_tf32_ts(tmem_d, ta_lo + 8 * i, dbh, idesc, (i > 0 || accumulate_first) ? 1u : 0u);
    mma_tf32_ts(tmem_d, ta_hi + 8 * i, dbl, idesc, 1u);
    mma_tf32_ts(tmem_d, ta_hi + 8 * i, dbh, idesc, 1u);
    dbh += ib; dbl += ib;
  }
}
__device__ __forceinline__ void issue_x3_ts(uint32_t tmem_d, uint32_t ta_hi, uint32_t ta_lo, uint32_t b_hi, uint32_t b_lo,
                                            int N, int K, bool accumulate_first) {
  if (K == 64) issue_x3_ts_n<8>(tmem_d, ta_hi, ta_lo, b_hi, b_lo, N, K, accumulate_first);
  else if (K == 128) issue_x3_ts_n<16>(tmem_d, ta_hi, ta_lo, b_hi, b_lo, N, K, accumulate_first);
  else issue_x3_ts_n<0>(tmem_d, ta_hi, ta_lo, b_hi, b_lo, N, K, accumulate_first);
}

}}  // namespace gn::tf
