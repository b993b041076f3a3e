// Twiddle bank: PolyOps::precompute_twiddles for the root half coset of the largest canonic circle domain.
// Replaces SimdBackend::precompute_twiddles at /root/reference prover/src/machine.rs:186-194.
// Layout (identical to stwo's TwiddleTree buffer so every smaller domain's twiddles are suffix slices):
//   for l in 0..k: 2^(k-1-l) x-coordinates of the first half of (root doubled l times), bit-reversed; then one pad word 1.
// Generated on the device: each thread turns its point index into a point with <= 31 group additions from a
// constant-memory table of generator doublings, and inverts it by Fermat.  Doubled copies (2t, 2/t) feed the FFT's
// multiply (stwo's SIMD backend keeps `twiddle_dbl` for the same reason).
#include "common.cuh"
#include "circle_host.h"

namespace nb {

__constant__ cpoint c_gen_pow2[31];

__device__ __forceinline__ cpoint dev_index_to_point(u32 idx) {
  cpoint r{1, 0};
#pragma unroll 1
  for (int b = 0; b < 31; ++b)
    if ((idx >> b) & 1u) r = cp_add(r, c_gen_pow2[b]);
  return r;
}

__global__ void twiddle_bank_kernel(u32* __restrict__ tw, u32* __restrict__ itw, u32* __restrict__ tw2, u32* __restrict__ itw2, u32 k) {
  u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  u32 len = 1u << k;
  if (e >= len) return;
  u32 m = len - e;                       // in [1, 2^k]
  u32 cl = 32 - __clz(m - 1);            // ceil_log2(m); m == 1 -> 0
  if (m == 1) cl = 0;
  u32 l = k - cl;
  u32 x;
  if (l >= k) {
    x = 1;  // pad
  } else {
    u32 off = len - (1u << (k - l));
    u32 j = e - off;
    u32 lg = k - l - 1;                  // log of the half-layer length
    u32 jr = lg ? (__brev(j) >> (32 - lg)) : 0;
    // coset_l = half_odds(k - l): initial = subgroup_gen(k-l+2), step = subgroup_gen(k-l)
    u32 init = 1u << (31 - (k - l + 2));
    u32 step = 1u << (31 - (k - l));
    u32 idx = (init + (u32)(((u64)step * jr) & 0x7fffffffu)) & 0x7fffffffu;
    x = dev_index_to_point(idx).x;
  }
  u32 xi = m31_inv(x);
  tw[e] = x; itw[e] = xi; tw2[e] = x << 1; itw2[e] = xi << 1;
}

// product bank: element e of layer l (index j inside the layer) times element j >> 1 of layer l + 1; odd j negated; doubled
__global__ void twiddle_product_kernel(const u32* __restrict__ tw, u32* __restrict__ out, u32 k) {
  u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  u32 len = 1u << k;
  if (e >= len) return;
  u32 m = len - e;                       // in [1, 2^k]
  u32 cl = (m == 1) ? 0 : 32 - __clz(m - 1);
  if (cl == 0) { out[e] = tw[e] << 1; return; }   // pad word
  u32 off = len - (1u << cl);            // start of this layer: 2^(cl-1) entries
  u32 j = e - off;
  u32 noff = len - (1u << (cl - 1));     // start of the next layer (or the pad)
  u32 nxt = (cl - 1 == 0) ? tw[len - 1] : tw[noff + (j >> 1)];
  u32 p = m31_mul(tw[e], nxt);
  if (j & 1u) p = m31_neg(p);
  out[e] = p << 1;
}

nb200_status twiddles_prepare(nb200_ctx* ctx, u32 max_domain_log) {
  NB_ARG(ctx, max_domain_log >= 1 && max_domain_log <= 30, "twiddles: domain log out of range");
  u32 k = max_domain_log - 1;
  if (ctx->tw.d_tw && ctx->tw.half_log >= k) return NB200_OK;
  if (ctx->tw.d_tw) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->tw.d_tw); cudaFree(ctx->tw.d_itw); cudaFree(ctx->tw.d_tw2); cudaFree(ctx->tw.d_itw2); cudaFree(ctx->tw.d_ptw2); cudaFree(ctx->tw.d_iptw2);
    ctx->tw = TwiddleBank();
  }
  static bool table_uploaded[64] = {false};
  if (!table_uploaded[ctx->device & 63]) {
    NB_CUDA(ctx, cudaMemcpyToSymbol(c_gen_pow2, gen_table().pow2, sizeof(cpoint) * 31));
    table_uploaded[ctx->device & 63] = true;
  }
  size_t len = (size_t)1 << k;
  NB_CUDA(ctx, cudaMalloc(&ctx->tw.d_tw, len * 4));
  NB_CUDA(ctx, cudaMalloc(&ctx->tw.d_itw, len * 4));
  NB_CUDA(ctx, cudaMalloc(&ctx->tw.d_tw2, len * 4));
  NB_CUDA(ctx, cudaMalloc(&ctx->tw.d_itw2, len * 4));
  NB_CUDA(ctx, cudaMalloc(&ctx->tw.d_ptw2, len * 4));
  NB_CUDA(ctx, cudaMalloc(&ctx->tw.d_iptw2, len * 4));
  ctx->tw.half_log = k;
  u32 threads = 256, blocks = (u32)((len + threads - 1) / threads);
  twiddle_bank_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->tw.d_tw, ctx->tw.d_itw, ctx->tw.d_tw2, ctx->tw.d_itw2, k);
  NB_LAUNCH_CHECK(ctx);
  twiddle_product_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->tw.d_tw, ctx->tw.d_ptw2, k);
  NB_LAUNCH_CHECK(ctx);
  twiddle_product_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->tw.d_itw, ctx->tw.d_iptw2, k);
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}

}  // namespace nb
