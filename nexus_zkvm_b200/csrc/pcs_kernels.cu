// Kernels of the polynomial commitment scheme's opening phase:
//   - DEEP / FRI quotients   (QuotientOps::accumulate_quotients,  stwo prover/backend/*/quotients.rs, core/pcs/quotients.rs)
//   - FRI folding            (FriOps::fold_circle_into_line / fold_line, stwo prover/backend/*/fri.rs)
//   - proof of work          (GrindOps::grind)
//   - secure-column helpers  (AccumulationOps::accumulate, domain point tables, gathers)
// all reached from stwo::prover::prove at /root/reference prover/src/machine.rs:286-290.
// Every kernel is one thread per row/pair with coalesced 32-bit column accesses (columns are SoA: a QM31 column is 4
// coordinate columns, exactly the layout the Merkle kernels hash).  Field arithmetic is exact, so the re-association
// of sums used here (per-batch sum of line terms hoisted to the host) is bit-exact with the reference formula.
#include "pcs.h"
#include "blake2s.cuh"
#include "circle_host.h"

namespace nb {

__constant__ cpoint c_gen_pow2_pcs[31];

static nb200_status ensure_gen_table(nb200_ctx* ctx) {
  static bool up[64] = {false};
  if (!up[ctx->device & 63]) {
    NB_CUDA(ctx, cudaMemcpyToSymbol(c_gen_pow2_pcs, gen_table().pow2, sizeof(cpoint) * 31));
    up[ctx->device & 63] = true;
  }
  return NB200_OK;
}

// x/y of CanonicCoset(log).circle_domain() in bit-reversed order
__global__ void domain_points_kernel(u32 log_size, u32* __restrict__ xs, u32* __restrict__ ys) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 n = 1u << log_size;
  if (i >= n) return;
  u32 j = __brev(i) >> (32 - log_size);
  u32 half = n >> 1;
  // half coset = half_odds(log-1): initial = subgroup_gen(log+1), step = subgroup_gen(log-1)
  u32 init = 1u << (31 - (log_size + 1));
  u32 step = (log_size - 1 == 0) ? 0u : (1u << (31 - (log_size - 1)));
  u32 jj = j < half ? j : j - half;
  u32 idx = (init + (u32)(((u64)step * jj) & 0x7fffffffu)) & 0x7fffffffu;
  cpoint r{1, 0};
#pragma unroll 1
  for (int b = 0; b < 31; ++b) if ((idx >> b) & 1u) r = cp_add(r, c_gen_pow2_pcs[b]);
  xs[i] = r.x;
  ys[i] = j < half ? r.y : m31_neg(r.y);
}
nb200_status domain_points(nb200_ctx* ctx, u32 log_size, u32* d_x, u32* d_y) {
  NB_TRY(ensure_gen_table(ctx));
  NB_ARG(ctx, log_size >= 1 && log_size <= 30, "domain_points: log size");
  u32 n = 1u << log_size, thr = 256;
  domain_points_kernel<<<(n + thr - 1) / thr, thr, 0, ctx->stream>>>(log_size, d_x, d_y);
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}

// ---- DEEP quotients ----
__global__ void __launch_bounds__(256) quotients_kernel(const QBatchDev* __restrict__ batches, u32 n_batches, const QEntryDev* __restrict__ entries,
                                                        const u32* __restrict__ dom_x, const u32* __restrict__ dom_y, u32 log_size,
                                                        u32* __restrict__ o0, u32* __restrict__ o1, u32* __restrict__ o2, u32* __restrict__ o3, u32 row0, u32 row_end) {
  const u32 row = row0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= row_end) return;
  const u32 x = __ldg(dom_x + row), y = __ldg(dom_y + row);
  qm31 acc = qm31_zero();
  for (u32 b = 0; b < n_batches; ++b) {
    const QBatchDev* qb = batches + b;
    // numerator = sum_k c_k f_k(row) - (A y + B)
    // sums of 64-bit products: four products (< 2^62 each) plus a reduced carry-in fit in 64 bits, so reduce after every fourth
    // entry.  The four column loads of a group are issued before any arithmetic (the loop is otherwise latency bound: one dependent
    // pointer + data load per entry).
    u64 s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const u32 first = qb->first, count = qb->count;
    const QEntryDev* en = entries + first;
    u32 e = 0;
    for (; e + 4 <= count; e += 4) {
      const u32* c0p = en[e].col; const u32* c1p = en[e + 1].col; const u32* c2p = en[e + 2].col; const u32* c3p = en[e + 3].col;
      const u32 f0 = __ldg(c0p + row), f1 = __ldg(c1p + row), f2 = __ldg(c2p + row), f3 = __ldg(c3p + row);
      const uint4 k0 = __ldg(reinterpret_cast<const uint4*>(en[e].c)), k1 = __ldg(reinterpret_cast<const uint4*>(en[e + 1].c));
      const uint4 k2 = __ldg(reinterpret_cast<const uint4*>(en[e + 2].c)), k3 = __ldg(reinterpret_cast<const uint4*>(en[e + 3].c));
      s0 += (u64)f0 * k0.x + (u64)f1 * k1.x + (u64)f2 * k2.x + (u64)f3 * k3.x;
      s1 += (u64)f0 * k0.y + (u64)f1 * k1.y + (u64)f2 * k2.y + (u64)f3 * k3.y;
      s2 += (u64)f0 * k0.z + (u64)f1 * k1.z + (u64)f2 * k2.z + (u64)f3 * k3.z;
      s3 += (u64)f0 * k0.w + (u64)f1 * k1.w + (u64)f2 * k2.w + (u64)f3 * k3.w;
      s0 = m31_red64(s0); s1 = m31_red64(s1); s2 = m31_red64(s2); s3 = m31_red64(s3);
    }
    for (; e < count; ++e) {   // at most three left: still within 64 bits
      const uint4 c = __ldg(reinterpret_cast<const uint4*>(en[e].c));
      const u32 f = __ldg(en[e].col + row);
      s0 += (u64)f * c.x; s1 += (u64)f * c.y; s2 += (u64)f * c.z; s3 += (u64)f * c.w;
    }
    qm31 numer = qm31_make(m31_red64(s0), m31_red64(s1), m31_red64(s2), m31_red64(s3));
    qm31 A = qm31_make(qb->A[0], qb->A[1], qb->A[2], qb->A[3]);
    qm31 B = qm31_make(qb->B[0], qb->B[1], qb->B[2], qb->B[3]);
    numer = qm31_sub(numer, qm31_add(qm31_mul_m31(A, y), B));
    // denominator = (prx - x) * piy - (pry - y) * pix   (CM31)
    cm31 prx{qb->prx[0], qb->prx[1]}, pry{qb->pry[0], qb->pry[1]}, pix{qb->pix[0], qb->pix[1]}, piy{qb->piy[0], qb->piy[1]};
    cm31 dx{m31_sub(prx.a, x), prx.b}, dy{m31_sub(pry.a, y), pry.b};
    cm31 den = cm31_sub(cm31_mul(dx, piy), cm31_mul(dy, pix));
    qm31 cf = qm31_make(qb->coeff[0], qb->coeff[1], qb->coeff[2], qb->coeff[3]);
    acc = qm31_add(qm31_mul(acc, cf), qm31_mul_cm31(numer, cm31_inv(den)));
  }
  o0[row] = acc.c[0]; o1[row] = acc.c[1]; o2[row] = acc.c[2]; o3[row] = acc.c[3];
}

// Four consecutive rows per thread: every column is read with 128-bit loads (512 contiguous bytes per warp and column instead of
// 128: the sweep jumps 2^log_size words from column to column, so longer bursts matter for DRAM efficiency) and the per-entry
// constants are fetched once for four rows.  Same arithmetic as quotients_kernel.
__global__ void __launch_bounds__(128) quotients_kernel_x4(const QBatchDev* __restrict__ batches, u32 n_batches, const QEntryDev* __restrict__ entries,
                                                           const u32* __restrict__ dom_x, const u32* __restrict__ dom_y, u32 log_size,
                                                           u32* __restrict__ o0, u32* __restrict__ o1, u32* __restrict__ o2, u32* __restrict__ o3, u32 row0, u32 row_end) {
  const u32 row = row0 + (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  if (row >= row_end) return;
  const uint4 xv = __ldg(reinterpret_cast<const uint4*>(dom_x + row)), yv = __ldg(reinterpret_cast<const uint4*>(dom_y + row));
  const u32 xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
  qm31 acc[4] = {qm31_zero(), qm31_zero(), qm31_zero(), qm31_zero()};
  for (u32 b = 0; b < n_batches; ++b) {
    const QBatchDev* qb = batches + b;
    u64 s[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[r][j] = 0;
    const u32 first = qb->first, count = qb->count;
    const QEntryDev* en = entries + first;
    u32 e = 0;
    for (; e + 8 <= count; e += 8) {   // eight columns (128 bytes) in flight per thread: two groups of four, one reduction per group
      uint4 f[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = __ldg(reinterpret_cast<const uint4*>(en[e + u].col + row));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) k[u] = __ldg(reinterpret_cast<const uint4*>(en[e + 4 * h + u].c));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const u32 fr[4] = {f[4 * h + u].x, f[4 * h + u].y, f[4 * h + u].z, f[4 * h + u].w};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            s[r][0] += (u64)fr[r] * k[u].x; s[r][1] += (u64)fr[r] * k[u].y; s[r][2] += (u64)fr[r] * k[u].z; s[r][3] += (u64)fr[r] * k[u].w;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[r][j] = m31_red64(s[r][j]);
      }
    }
    for (; e + 4 <= count; e += 4) {   // four columns in flight per thread (the loop is latency bound otherwise), then one reduction
      uint4 f[4], k[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) f[u] = __ldg(reinterpret_cast<const uint4*>(en[e + u].col + row));
#pragma unroll
      for (int u = 0; u < 4; ++u) k[u] = __ldg(reinterpret_cast<const uint4*>(en[e + u].c));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u32 fr[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          s[r][0] += (u64)fr[r] * k[u].x; s[r][1] += (u64)fr[r] * k[u].y; s[r][2] += (u64)fr[r] * k[u].z; s[r][3] += (u64)fr[r] * k[u].w;
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[r][j] = m31_red64(s[r][j]);
    }
    for (; e < count; ++e) {           // at most three left: still within 64 bits
      const uint4 f = __ldg(reinterpret_cast<const uint4*>(en[e].col + row));
      const uint4 k = __ldg(reinterpret_cast<const uint4*>(en[e].c));
      const u32 fr[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        s[r][0] += (u64)fr[r] * k.x; s[r][1] += (u64)fr[r] * k.y; s[r][2] += (u64)fr[r] * k.z; s[r][3] += (u64)fr[r] * k.w;
      }
    }
    const qm31 A = qm31_make(qb->A[0], qb->A[1], qb->A[2], qb->A[3]);
    const qm31 B = qm31_make(qb->B[0], qb->B[1], qb->B[2], qb->B[3]);
    const cm31 prx{qb->prx[0], qb->prx[1]}, pry{qb->pry[0], qb->pry[1]}, pix{qb->pix[0], qb->pix[1]}, piy{qb->piy[0], qb->piy[1]};
    const qm31 cf = qm31_make(qb->coeff[0], qb->coeff[1], qb->coeff[2], qb->coeff[3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      qm31 numer = qm31_make(m31_red64(s[r][0]), m31_red64(s[r][1]), m31_red64(s[r][2]), m31_red64(s[r][3]));
      numer = qm31_sub(numer, qm31_add(qm31_mul_m31(A, ys[r]), B));
      cm31 dx{m31_sub(prx.a, xs[r]), prx.b}, dy{m31_sub(pry.a, ys[r]), pry.b};
      cm31 den = cm31_sub(cm31_mul(dx, piy), cm31_mul(dy, pix));
      acc[r] = qm31_add(qm31_mul(acc[r], cf), qm31_mul_cm31(numer, cm31_inv(den)));
    }
  }
  *reinterpret_cast<uint4*>(o0 + row) = make_uint4(acc[0].c[0], acc[1].c[0], acc[2].c[0], acc[3].c[0]);
  *reinterpret_cast<uint4*>(o1 + row) = make_uint4(acc[0].c[1], acc[1].c[1], acc[2].c[1], acc[3].c[1]);
  *reinterpret_cast<uint4*>(o2 + row) = make_uint4(acc[0].c[2], acc[1].c[2], acc[2].c[2], acc[3].c[2]);
  *reinterpret_cast<uint4*>(o3 + row) = make_uint4(acc[0].c[3], acc[1].c[3], acc[2].c[3], acc[3].c[3]);
}

nb200_status quotients_launch(nb200_ctx* ctx, const QBatchDev* h_batches, size_t n_batches, const QEntryDev* h_entries, size_t n_entries,
                              const u32* dom_x, const u32* dom_y, u32 log_size, u32* out /* 4 columns */, u32 row0, size_t n_rows) {
  QBatchDev* d_b = nullptr; QEntryDev* d_e = nullptr;
  NB_CUDA(ctx, dmalloc(ctx, (void**)&d_b, n_batches * sizeof(QBatchDev)));
  NB_CUDA(ctx, dmalloc(ctx, (void**)&d_e, n_entries * sizeof(QEntryDev)));
  NB_CUDA(ctx, cudaMemcpyAsync(d_b, h_batches, n_batches * sizeof(QBatchDev), cudaMemcpyHostToDevice, ctx->stream));
  NB_CUDA(ctx, cudaMemcpyAsync(d_e, h_entries, n_entries * sizeof(QEntryDev), cudaMemcpyHostToDevice, ctx->stream));
  size_t n = (size_t)1 << log_size;
  const size_t nr = n_rows ? n_rows : n;      // rows [row0, row0 + nr) of the domain (a rank's slice in a multi-GPU proof); pointers are indexed by the global row
  const u32 row_end = (u32)(row0 + nr);
  bool aligned = log_size >= 2 && (row0 % 4 == 0) && (nr % 4 == 0) && ((((uintptr_t)out) | ((uintptr_t)dom_x) | ((uintptr_t)dom_y)) & 15u) == 0;
  for (size_t e = 0; e < n_entries && aligned; ++e) aligned = (((uintptr_t)h_entries[e].col) & 15u) == 0;
  if (aligned) {
    u32 thr = 128; size_t nt = nr / 4;
    quotients_kernel_x4<<<(u32)((nt + thr - 1) / thr), thr, 0, ctx->stream>>>(d_b, (u32)n_batches, d_e, dom_x, dom_y, log_size,
                                                                             out, out + n, out + 2 * n, out + 3 * n, row0, row_end);
  } else {
    u32 thr = 256;
    quotients_kernel<<<(u32)((nr + thr - 1) / thr), thr, 0, ctx->stream>>>(d_b, (u32)n_batches, d_e, dom_x, dom_y, log_size,
                                                                          out, out + n, out + 2 * n, out + 3 * n, row0, row_end);
  }
  NB_LAUNCH_CHECK(ctx);
  NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_b); dfree(ctx, d_e);
  return NB200_OK;
}

// ---- FRI folds ----
__device__ __forceinline__ qm31 ld_q(const u32* c, size_t n, size_t i) { return qm31_make(c[i], c[n + i], c[2 * n + i], c[3 * n + i]); }
__device__ __forceinline__ void st_q(u32* c, size_t n, size_t i, qm31 v) { c[i] = v.c[0]; c[n + i] = v.c[1]; c[2 * n + i] = v.c[2]; c[3 * n + i] = v.c[3]; }

// dst[i] = dst[i] * alpha^2 + (f0' + alpha f1'),  (f0', f1') = ibutterfly(src[2i], src[2i+1], 1/y_i)
__global__ void fold_circle_kernel(u32* __restrict__ dst, const u32* __restrict__ src, u32 src_log, const u32* __restrict__ itw, u32 tw_len,
                                   qm31 alpha, qm31 alpha_sq) {
  const size_t n = (size_t)1 << src_log, m = n >> 1;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  // circle inverse twiddle of layer 0 at h = i (same derivation as the FFT: [x, y] -> [y, -y, -x, x])
  u32 t;
  if (src_log <= 2) {
    t = 0;  // handled by the caller through explicit twiddles (tiny domains never reach FRI in practice)
  } else {
    const u32* l1 = itw + (tw_len - (1u << (src_log - 1)));
    u32 q = (u32)i >> 2, r = (u32)i & 3u;
    u32 x = __ldg(l1 + 2 * q), y = __ldg(l1 + 2 * q + 1);
    u32 v = (r < 2) ? y : x;
    t = (r == 1 || r == 2) ? (P31 - v) : v;
  }
  qm31 f0 = ld_q(src, n, 2 * i), f1 = ld_q(src, n, 2 * i + 1);
  qm31 s = qm31_add(f0, f1), d = qm31_mul_m31(qm31_sub(f0, f1), t);
  qm31 fp = qm31_add(qm31_mul(alpha, d), s);
  qm31 cur = ld_q(dst, m, i);
  st_q(dst, m, i, qm31_add(qm31_mul(cur, alpha_sq), fp));
}
// dst[i] = f0' + alpha f1',  (f0', f1') = ibutterfly(src[2i], src[2i+1], 1/x_i)
__global__ void fold_line_kernel(u32* __restrict__ dst, const u32* __restrict__ src, u32 src_log, const u32* __restrict__ itw, u32 tw_len, qm31 alpha) {
  const size_t n = (size_t)1 << src_log, m = n >> 1;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  u32 t = __ldg(itw + (tw_len - (1u << src_log)) + i);
  qm31 f0 = ld_q(src, n, 2 * i), f1 = ld_q(src, n, 2 * i + 1);
  qm31 s = qm31_add(f0, f1), d = qm31_mul_m31(qm31_sub(f0, f1), t);
  st_q(dst, m, i, qm31_add(s, qm31_mul(alpha, d)));
}
nb200_status fold_circle_into_line(nb200_ctx* ctx, u32* dst, const u32* src, u32 src_log, qm31 alpha) {
  NB_ARG(ctx, src_log >= 3, "fold_circle_into_line: domain too small");
  NB_ARG(ctx, ctx->tw.d_itw && ctx->tw.half_log + 1 >= src_log, "fold_circle_into_line: twiddles");
  size_t m = (size_t)1 << (src_log - 1);
  u32 thr = 256;
  fold_circle_kernel<<<(u32)((m + thr - 1) / thr), thr, 0, ctx->stream>>>(dst, src, src_log, ctx->tw.d_itw, 1u << ctx->tw.half_log, alpha, qm31_mul(alpha, alpha));
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}
nb200_status fold_line(nb200_ctx* ctx, u32* dst, const u32* src, u32 src_log, qm31 alpha) {
  NB_ARG(ctx, src_log >= 1, "fold_line: domain too small");
  NB_ARG(ctx, ctx->tw.d_itw && ctx->tw.half_log >= src_log, "fold_line: twiddles");
  size_t m = (size_t)1 << (src_log - 1);
  u32 thr = 256;
  fold_line_kernel<<<(u32)((m + thr - 1) / thr), thr, 0, ctx->stream>>>(dst, src, src_log, ctx->tw.d_itw, 1u << ctx->tw.half_log, alpha);
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}

// ---- accumulate (AccumulationOps::accumulate): a[i] += b[i] over 4-coordinate columns ----
__global__ void add_inplace_kernel(u32* __restrict__ a, const u32* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = m31_add(a[i], b[i]);
}
nb200_status add_inplace(nb200_ctx* ctx, u32* a, const u32* b, size_t n) {
  u32 thr = 256;
  add_inplace_kernel<<<(u32)((n + thr - 1) / thr), thr, 0, ctx->stream>>>(a, b, n);
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}

// a[i] = (a[i] - b[i]) / t, t = the top-layer (line layer tw_log - 1) twiddle of CanonicCoset(tw_log)'s FFT on the first half of the domain.
// Used to split a polynomial of 2^tw_log coefficients into its low and high halves from evaluations on two half-size domains (prove.cu).
__global__ void sub_scale_kernel(u32* __restrict__ a, const u32* __restrict__ b, size_t n, const u32* __restrict__ itw_top) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 it = __ldg(itw_top);
  if (i < n) a[i] = m31_mul(m31_sub(a[i], b[i]), it);
}
nb200_status sub_scale_top_twiddle(nb200_ctx* ctx, u32* a, const u32* b, size_t n, u32 tw_log) {
  NB_ARG(ctx, ctx->tw.d_itw && tw_log >= 2 && ctx->tw.half_log + 1 >= tw_log, "sub_scale_top_twiddle: twiddles");
  // TwiddleTree layout: the array of layer i of canonic(n) starts at tw_len - 2^(n - i); layer n - 1 has one entry
  const u32* it = ctx->tw.d_itw + (((size_t)1 << ctx->tw.half_log) - 2);
  u32 thr = 256;
  sub_scale_kernel<<<(u32)((n + thr - 1) / thr), thr, 0, ctx->stream>>>(a, b, n, it);
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}
// dst[c * dst_stride + i] += src[c * src_stride + i], i < len
__global__ void add_cols_kernel(u32* __restrict__ dst, size_t dst_stride, const u32* __restrict__ src, size_t src_stride, size_t len) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) { u32* d = dst + blockIdx.y * dst_stride + i; *d = m31_add(*d, src[blockIdx.y * src_stride + i]); }
}
nb200_status add_cols_strided(nb200_ctx* ctx, u32* dst, size_t dst_stride, const u32* src, size_t src_stride, size_t len, size_t n_cols) {
  u32 thr = 256;
  add_cols_kernel<<<dim3((u32)((len + thr - 1) / thr), (u32)n_cols), thr, 0, ctx->stream>>>(dst, dst_stride, src, src_stride, len);
  NB_LAUNCH_CHECK(ctx);
  return NB200_OK;
}

// ---- proof of work: smallest nonce with trailing_zeros(Blake2s(digest || nonce_le)) >= pow_bits ----
__global__ void grind_kernel(const u32 d0, const u32 d1, const u32 d2, const u32 d3, const u32 d4, const u32 d5, const u32 d6, const u32 d7,
                             u64 base, u32 pow_bits, unsigned long long* __restrict__ result) {
  u64 nonce = base + (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u32 h[8]; b2s_init(h);
  u32 m[16] = {d0, d1, d2, d3, d4, d5, d6, d7, (u32)nonce, (u32)(nonce >> 32), 0, 0, 0, 0, 0, 0};
  b2s_compress(h, m, 40, 0, 0xFFFFFFFFu, 0);
  // trailing zeros of the first 16 bytes as a little-endian u128
  u32 tz;
  if (h[0]) tz = __ffs(h[0]) - 1;
  else if (h[1]) tz = 32 + __ffs(h[1]) - 1;
  else if (h[2]) tz = 64 + __ffs(h[2]) - 1;
  else if (h[3]) tz = 96 + __ffs(h[3]) - 1;
  else tz = 128;
  if (tz >= pow_bits) atomicMin(result, (unsigned long long)nonce);
}
nb200_status grind(nb200_ctx* ctx, const uint8_t digest[32], u32 pow_bits, uint64_t* nonce_out) {
  NB_ARG(ctx, ctx->pow_variant == 0, "grind: only pow_variant 0 is implemented");
  NB_ARG(ctx, pow_bits <= 64, "grind: pow_bits too large");
  u32 d[8];
  for (int i = 0; i < 8; ++i) d[i] = (u32)digest[4 * i] | ((u32)digest[4 * i + 1] << 8) | ((u32)digest[4 * i + 2] << 16) | ((u32)digest[4 * i + 3] << 24);
  unsigned long long* d_res = nullptr;
  NB_CUDA(ctx, dmalloc(ctx, (void**)&d_res, 8));
  const unsigned long long none = ~0ull;
  u64 base = 0;
  // chunk size grows with the expected work so tiny pow_bits stay cheap
  u32 blocks = pow_bits <= 12 ? 64 : 148 * 32;
  const u32 thr = 256;
  while (true) {
    NB_CUDA(ctx, cudaMemcpyAsync(d_res, &none, 8, cudaMemcpyHostToDevice, ctx->stream));
    grind_kernel<<<blocks, thr, 0, ctx->stream>>>(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], base, pow_bits, d_res);
    NB_LAUNCH_CHECK(ctx);
    unsigned long long r;
    NB_CUDA(ctx, cudaMemcpyAsync(&r, d_res, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (r != none) { *nonce_out = r; break; }
    base += (u64)blocks * thr;
  }
  dfree(ctx, d_res);
  return NB200_OK;
}

}  // namespace nb
