// Shared device helpers of the Circle FFT kernels (fft.cu: per-pass tile kernels; fft_fused.cu: the commit pipeline
// with the fused iFFT-tail / LDE-head kernel).  See fft.cu for the algorithm and the reference call sites.
#pragma once
#include "common.cuh"

namespace nb {

struct FftPass {
  const u32* src;     // source columns (column c at src + c*src_stride); zero-extended beyond src_len
  u32* dst;           // destination columns
  size_t src_stride, dst_stride;
  size_t src_len;     // valid words per source column
  const u32* tw;      // twiddle (or inverse twiddle) bank
  const u32* tw2;     // the same bank doubled (2t), for m31_mul_dbl
  const u32* ctw2;    // DOUBLED circle (layer 0) twiddles of this transform size, 2^(n-1) words
  u32 tw_len;         // bank length (2^k)
  u32 n_cols;
  u32 n;              // log size of the transform
  u32 lo;             // first layer of this pass
  u32 T, W;           // tile log, width log (L = T - W layers)
  u32 cb;             // columns per CTA
  u32 scale;          // multiply outputs by this (interpolate last pass) if apply_scale
  u32 apply_scale;
  u32 tn;             // log size of the canonic domain whose twiddle arrays are used (= n, or n + 1 for the half-domain transforms)
  u32 ztop;           // forward transforms of zero-extended input: layers >= ztop are copies (= log2 of the source length)
};

__device__ __forceinline__ void butterfly(u32& v0, u32& v1, u32 t) {
  u32 tmp = m31_mul(v1, t);
  v1 = m31_sub(v0, tmp);
  v0 = m31_add(v0, tmp);
}
__device__ __forceinline__ void ibutterfly(u32& v0, u32& v1, u32 it) {
  u32 tmp = v0;
  v0 = m31_add(tmp, v1);
  v1 = m31_mul(m31_sub(tmp, v1), it);
}

// twiddle of layer i (>= 1) at index h for a transform of log size n
__device__ __forceinline__ u32 line_tw(const u32* __restrict__ tw, u32 tw_len, u32 n, u32 i, u32 h) {
  return __ldg(tw + (tw_len - (1u << (n - i)) + h));
}
// circle twiddle (layer 0) at index h: from the first line layer, [x, y] -> [y, -y, -x, x]
__device__ __forceinline__ u32 circle_tw(const u32* __restrict__ tw, u32 tw_len, u32 n, u32 h) {
  const u32* l1 = tw + (tw_len - (1u << (n - 1)));
  u32 q = h >> 2, r = h & 3u;
  u32 x = __ldg(l1 + 2 * q), y = __ldg(l1 + 2 * q + 1);
  u32 v = (r < 2) ? y : x;
  return (r == 1 || r == 2) ? (P31 - v) : v;
}
static __global__ void circle_table_kernel(const u32* __restrict__ tw, u32 tw_len, u32 n, u32* __restrict__ out) {
  u32 h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < (1u << (n - 1))) out[h] = circle_tw(tw, tw_len, n, h) << 1;  // doubled, see m31_mul_dbl
}

// =====================================================================================================
// fast path: compile-time tile shape
// =====================================================================================================
// swizzle keeps aligned groups of 4 words intact (128-bit accesses) and is conflict-free for: 128-bit staging,
// the 128-bit round at bit 0, and the 32-bit rounds at every bit position used by the schedules below.
__device__ __forceinline__ u32 swz2(u32 s) { return s ^ (((s >> 5) & 3u) << 2) ^ (((s >> 8) & 1u) << 4); }

// a * t mod P with the twiddle pre-doubled (t2 = 2t < 2^32): the 64-bit product a * t2 has (a*t) >> 31 in its high
// word and 2 * ((a*t) mod 2^31) in its low word, so the Mersenne fold is one shifted add (LEA.HI) and one min — the
// mask and the funnel shift of the plain form disappear, which matters because the ALU pipe is the binding one.
__device__ __forceinline__ u32 m31_mul_dbl(u32 a, u32 t2) {
  u64 p = (u64)a * t2;
  u32 s = ((u32)p >> 1) + (u32)(p >> 32);
  return umin32(s, s - P31);
}
__device__ __forceinline__ void butterfly_dbl(u32& v0, u32& v1, u32 t2) {
  u32 tmp = m31_mul_dbl(v1, t2);
  v1 = m31_sub(v0, tmp);
  v0 = m31_add(v0, tmp);
}
__device__ __forceinline__ void ibutterfly_dbl(u32& v0, u32& v1, u32 it2) {
  u32 tmp = v0;
  v0 = m31_add(tmp, v1);
  v1 = m31_mul_dbl(m31_sub(tmp, v1), it2);
}

template <bool INV>
__device__ __forceinline__ void radix16(u32 (&v)[16], const u32 (&tw)[15], const int jlo, const u32 triv = 0u) {
  // tw holds DOUBLED twiddles.  triv bit j (forward only): layer j of this round sits at or above the zero-extension
  // boundary, its odd inputs are known zeros, so the butterfly degenerates to a copy (no arithmetic).
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = INV ? jj : 3 - jj;
    if (j >= jlo) {
      if (!INV && ((triv >> j) & 1u)) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int k0 = ((m >> j) << (j + 1)) | (m & ((1 << j) - 1));
          v[k0 | (1 << j)] = v[k0];
        }
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int k0 = ((m >> j) << (j + 1)) | (m & ((1 << j) - 1));
          const int k1 = k0 | (1 << j);
          const int off = (j == 0 ? 0 : j == 1 ? 8 : j == 2 ? 12 : 14) + (m >> j);
          if (INV) ibutterfly_dbl(v[k0], v[k1], tw[off]);
          else butterfly_dbl(v[k0], v[k1], tw[off]);
        }
      }
    }
  }
}


// ---- asynchronous global -> shared copies (LDGSTS): no registers, no issue slots between request and use ----
__device__ __forceinline__ void cp_async16(u32 smem_addr, const void* g, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// circle (layer 0) twiddle tables of a transform size, cached per ctx (fft.cu)
nb200_status fft_circle_tables(nb200_ctx* ctx, u32 n, const u32** fwd, const u32** inv);

}  // namespace nb
