// M31 / CM31 / QM31 arithmetic for host and device (sm_100a).
// Field definitions: reference spec zkvm-spec-3.0.pdf §3.1 p.14; stwo core/fields/{m31,cm31,qm31}.rs.
// All values are canonical u32 in [0, P).  Reductions use the Mersenne structure of P = 2^31 - 1:
//   x mod P = (x & P) + (x >> 31) folded once, then one conditional subtract done as an unsigned min.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#define NB_D __device__ __forceinline__
#else
#define NB_HD inline
#define NB_D inline
#endif

namespace nb {

typedef uint32_t u32;
typedef uint64_t u64;

static constexpr u32 P31 = 0x7fffffffu;

NB_HD u32 umin32(u32 a, u32 b) { return a < b ? a : b; }

NB_HD u32 m31_add(u32 a, u32 b) { u32 s = a + b; return umin32(s, s - P31); }
NB_HD u32 m31_sub(u32 a, u32 b) { u32 d = a - b; return umin32(d, d + P31); }
NB_HD u32 m31_neg(u32 a) { return a == 0 ? 0 : P31 - a; }
NB_HD u32 m31_dbl(u32 a) { return m31_add(a, a); }
NB_HD u32 m31_mul(u32 a, u32 b) {
  u64 p = (u64)a * b;
  u32 s = ((u32)p & P31) + (u32)(p >> 31);
  return umin32(s, s - P31);
}
// reduce an arbitrary u64 < 2^62+... (used by host code / accumulations)
NB_HD u32 m31_reduce64(u64 x) {
  u64 s = (x & P31) + (x >> 31);          // < 2^33+2^31
  u32 t = (u32)(s & P31) + (u32)(s >> 31);  // < 2^31 + 4
  return umin32(t, t - P31);
}
#ifdef __CUDACC__
// device: canonical residue of any u64 in 5 instructions (2^32 == 2, 2^31 == 1 mod P): y = 2 hi + lo, then fold bit 31 and above
__device__ __forceinline__ u32 m31_red64(u64 x) {
  u64 y;
  asm("mad.wide.u32 %0, %1, 2, %2;" : "=l"(y) : "r"((u32)(x >> 32)), "l"((u64)(u32)x));
  u32 yl = (u32)y, yh = (u32)(y >> 32);
  u32 s = (yl & P31) + __funnelshift_r(yl, yh, 31);
  return umin32(s, s - P31);
}
#endif
NB_HD u32 m31_pow(u32 a, u32 e) {
  u32 r = 1;
  while (e) { if (e & 1) r = m31_mul(r, a); a = m31_mul(a, a); e >>= 1; }
  return r;
}
NB_HD u32 m31_inv(u32 a) {
  return m31_pow(a, P31 - 2);  // Fermat
}

struct cm31 { u32 a, b; };  // a + b i
NB_HD cm31 cm31_add(cm31 x, cm31 y) { return cm31{m31_add(x.a, y.a), m31_add(x.b, y.b)}; }
NB_HD cm31 cm31_sub(cm31 x, cm31 y) { return cm31{m31_sub(x.a, y.a), m31_sub(x.b, y.b)}; }
NB_HD cm31 cm31_neg(cm31 x) { return cm31{m31_neg(x.a), m31_neg(x.b)}; }
NB_HD cm31 cm31_mul(cm31 x, cm31 y) {
  return cm31{m31_sub(m31_mul(x.a, y.a), m31_mul(x.b, y.b)), m31_add(m31_mul(x.a, y.b), m31_mul(x.b, y.a))};
}
NB_HD cm31 cm31_mul_m31(cm31 x, u32 y) { return cm31{m31_mul(x.a, y), m31_mul(x.b, y)}; }
NB_HD cm31 cm31_inv(cm31 x) {
  u32 n = m31_inv(m31_add(m31_mul(x.a, x.a), m31_mul(x.b, x.b)));
  return cm31{m31_mul(x.a, n), m31_neg(m31_mul(x.b, n))};
}
// multiply by R = 2 + i :  (a+bi)(2+i) = (2a - b) + (a + 2b) i
NB_HD cm31 cm31_mul_R(cm31 x) {
  return cm31{m31_sub(m31_dbl(x.a), x.b), m31_add(x.a, m31_dbl(x.b))};
}

struct qm31 { u32 c[4]; };  // (c0 + c1 i) + (c2 + c3 i) u
NB_HD qm31 qm31_make(u32 a, u32 b, u32 c, u32 d) { qm31 r; r.c[0] = a; r.c[1] = b; r.c[2] = c; r.c[3] = d; return r; }
NB_HD qm31 qm31_zero() { return qm31_make(0, 0, 0, 0); }
NB_HD qm31 qm31_one() { return qm31_make(1, 0, 0, 0); }
NB_HD qm31 qm31_from_m31(u32 x) { return qm31_make(x, 0, 0, 0); }
NB_HD cm31 qm31_lo(qm31 x) { return cm31{x.c[0], x.c[1]}; }
NB_HD cm31 qm31_hi(qm31 x) { return cm31{x.c[2], x.c[3]}; }
NB_HD qm31 qm31_from_cm31(cm31 lo, cm31 hi) { return qm31_make(lo.a, lo.b, hi.a, hi.b); }
NB_HD bool qm31_eq(qm31 x, qm31 y) { return x.c[0] == y.c[0] && x.c[1] == y.c[1] && x.c[2] == y.c[2] && x.c[3] == y.c[3]; }
NB_HD bool qm31_is_zero(qm31 x) { return (x.c[0] | x.c[1] | x.c[2] | x.c[3]) == 0; }
NB_HD qm31 qm31_add(qm31 x, qm31 y) { return qm31_make(m31_add(x.c[0], y.c[0]), m31_add(x.c[1], y.c[1]), m31_add(x.c[2], y.c[2]), m31_add(x.c[3], y.c[3])); }
NB_HD qm31 qm31_sub(qm31 x, qm31 y) { return qm31_make(m31_sub(x.c[0], y.c[0]), m31_sub(x.c[1], y.c[1]), m31_sub(x.c[2], y.c[2]), m31_sub(x.c[3], y.c[3])); }
NB_HD qm31 qm31_neg(qm31 x) { return qm31_make(m31_neg(x.c[0]), m31_neg(x.c[1]), m31_neg(x.c[2]), m31_neg(x.c[3])); }
NB_HD qm31 qm31_mul(qm31 x, qm31 y) {
  cm31 a = qm31_lo(x), b = qm31_hi(x), c = qm31_lo(y), d = qm31_hi(y);
  cm31 lo = cm31_add(cm31_mul(a, c), cm31_mul_R(cm31_mul(b, d)));
  cm31 hi = cm31_add(cm31_mul(a, d), cm31_mul(b, c));
  return qm31_from_cm31(lo, hi);
}
NB_HD qm31 qm31_mul_m31(qm31 x, u32 y) { return qm31_make(m31_mul(x.c[0], y), m31_mul(x.c[1], y), m31_mul(x.c[2], y), m31_mul(x.c[3], y)); }
NB_HD qm31 qm31_mul_cm31(qm31 x, cm31 y) { return qm31_from_cm31(cm31_mul(qm31_lo(x), y), cm31_mul(qm31_hi(x), y)); }
NB_HD qm31 qm31_add_m31(qm31 x, u32 y) { x.c[0] = m31_add(x.c[0], y); return x; }
NB_HD qm31 qm31_sub_m31(qm31 x, u32 y) { x.c[0] = m31_sub(x.c[0], y); return x; }
NB_HD qm31 qm31_sqr(qm31 x) { return qm31_mul(x, x); }
NB_HD qm31 qm31_conj(qm31 x) { return qm31_make(x.c[0], x.c[1], m31_neg(x.c[2]), m31_neg(x.c[3])); }  // a + bu -> a - bu
NB_HD qm31 qm31_inv(qm31 x) {
  cm31 a = qm31_lo(x), b = qm31_hi(x);
  cm31 denom = cm31_sub(cm31_mul(a, a), cm31_mul_R(cm31_mul(b, b)));
  cm31 di = cm31_inv(denom);
  return qm31_from_cm31(cm31_mul(a, di), cm31_neg(cm31_mul(b, di)));
}
NB_HD qm31 qm31_pow(qm31 a, u64 e) {
  qm31 r = qm31_one();
  while (e) { if (e & 1) r = qm31_mul(r, a); a = qm31_mul(a, a); e >>= 1; }
  return r;
}

// ---- circle group over M31 ----
struct cpoint { u32 x, y; };
NB_HD cpoint cp_add(cpoint p, cpoint q) {
  return cpoint{m31_sub(m31_mul(p.x, q.x), m31_mul(p.y, q.y)), m31_add(m31_mul(p.x, q.y), m31_mul(p.y, q.x))};
}
NB_HD cpoint cp_neg(cpoint p) { return cpoint{p.x, m31_neg(p.y)}; }
NB_HD u32 m31_double_x(u32 x) { u32 s = m31_mul(x, x); return m31_sub(m31_dbl(s), 1); }

// circle over QM31 (OODS points)
struct qpoint { qm31 x, y; };
NB_HD qpoint qp_add(qpoint p, qpoint q) {
  return qpoint{qm31_sub(qm31_mul(p.x, q.x), qm31_mul(p.y, q.y)), qm31_add(qm31_mul(p.x, q.y), qm31_mul(p.y, q.x))};
}
NB_HD qpoint qp_from_m31(cpoint p) { return qpoint{qm31_from_m31(p.x), qm31_from_m31(p.y)}; }
NB_HD qm31 qm31_double_x(qm31 x) { qm31 s = qm31_sqr(x); return qm31_sub_m31(qm31_add(s, s), 1); }

NB_HD u32 bit_reverse_u32(u32 i, u32 log_size) {
  if (log_size == 0) return i;
#if defined(__CUDA_ARCH__)
  return __brev(i) >> (32 - log_size);
#else
  u32 r = 0;
  for (u32 b = 0; b < log_size; ++b) r |= ((i >> b) & 1u) << (log_size - 1 - b);
  return r;
#endif
}

}  // namespace nb
