// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the product path.
// PARITY UNPINNED for Stwo-internal details (see oracle/README.md): the arithmetic below is a CPU
// restatement of starkware-libs/stwo @ rev 0790eba (Cargo.toml:39-48 of the reference), which is NOT
// present under /root/reference.  Field definitions ARE pinned by the reference's spec
// (specification/zkvm-spec-3.0.pdf §3.1 p.14: M31, CM31 = M31[i]/(i^2+1), QM31 = CM31[u]/(u^2-(2+i))).
//
// Deliberately naive: every product is a u64 `%`, no Mersenne tricks, so the oracle shares no
// reduction code with the CUDA path it checks.
#pragma once
#include <cstdint>
#include <vector>
#include <cassert>

namespace orc {

static constexpr uint32_t P = 2147483647u;  // 2^31 - 1   (stwo core/fields/m31.rs: `pub const P`)

struct M31 {
  uint32_t v;
  M31() : v(0) {}
  explicit M31(uint32_t x) : v(x % P) {}
  static M31 raw(uint32_t x) { M31 r; r.v = x; return r; }
  // stwo M31::reduce(u64) — value mod P
  static M31 reduce(uint64_t x) { M31 r; r.v = (uint32_t)(x % P); return r; }
  static M31 from_i64(int64_t x) { int64_t m = x % (int64_t)P; if (m < 0) m += P; return raw((uint32_t)m); }
  bool operator==(const M31& o) const { return v == o.v; }
  bool operator!=(const M31& o) const { return v != o.v; }
  bool is_zero() const { return v == 0; }
};
inline M31 operator+(M31 a, M31 b) { return M31::raw((uint32_t)(((uint64_t)a.v + b.v) % P)); }
inline M31 operator-(M31 a, M31 b) { return M31::raw((uint32_t)(((uint64_t)a.v + P - b.v) % P)); }
inline M31 operator-(M31 a) { return M31::raw(a.v == 0 ? 0 : P - a.v); }
inline M31 operator*(M31 a, M31 b) { return M31::raw((uint32_t)(((uint64_t)a.v * b.v) % P)); }
inline M31 pow(M31 a, uint64_t e) {
  M31 r = M31::raw(1);
  while (e) { if (e & 1) r = r * a; a = a * a; e >>= 1; }
  return r;
}
inline M31 inv(M31 a) { assert(a.v != 0); return pow(a, P - 2); }

// CM31 = M31[i]/(i^2+1)       (stwo core/fields/cm31.rs)
struct CM31 {
  M31 a, b;  // a + b i
  CM31() {}
  CM31(M31 a_, M31 b_) : a(a_), b(b_) {}
  bool operator==(const CM31& o) const { return a == o.a && b == o.b; }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
};
inline CM31 operator+(CM31 x, CM31 y) { return CM31(x.a + y.a, x.b + y.b); }
inline CM31 operator-(CM31 x, CM31 y) { return CM31(x.a - y.a, x.b - y.b); }
inline CM31 operator-(CM31 x) { return CM31(-x.a, -x.b); }
inline CM31 operator*(CM31 x, CM31 y) { return CM31(x.a * y.a - x.b * y.b, x.a * y.b + x.b * y.a); }
inline CM31 operator*(CM31 x, M31 y) { return CM31(x.a * y, x.b * y); }
inline CM31 inv(CM31 x) {
  // 1/(a+bi) = (a-bi)/(a^2+b^2)
  M31 n = inv(x.a * x.a + x.b * x.b);
  return CM31(x.a * n, -(x.b * n));
}

// QM31 = CM31[u]/(u^2 - (2+i))   (stwo core/fields/qm31.rs; `R = CM31(2,1)`)
struct QM31 {
  CM31 a, b;  // a + b u ; coordinates [a.a, a.b, b.a, b.b]
  QM31() {}
  QM31(CM31 a_, CM31 b_) : a(a_), b(b_) {}
  static QM31 from_u32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return QM31(CM31(M31::raw(c0), M31::raw(c1)), CM31(M31::raw(c2), M31::raw(c3)));
  }
  static QM31 from_m31(M31 x) { return QM31(CM31(x, M31()), CM31()); }
  static QM31 zero() { return QM31(); }
  static QM31 one() { return from_m31(M31::raw(1)); }
  bool operator==(const QM31& o) const { return a == o.a && b == o.b; }
  bool operator!=(const QM31& o) const { return !(*this == o); }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
  uint32_t coord(int k) const { return k == 0 ? a.a.v : k == 1 ? a.b.v : k == 2 ? b.a.v : b.b.v; }
  // "complex_conjugate" in stwo = conjugation over CM31: a + bu -> a - bu
  QM31 complex_conjugate() const { return QM31(a, -b); }
};
static inline CM31 QM31_R() { return CM31(M31::raw(2), M31::raw(1)); }
inline QM31 operator+(QM31 x, QM31 y) { return QM31(x.a + y.a, x.b + y.b); }
inline QM31 operator-(QM31 x, QM31 y) { return QM31(x.a - y.a, x.b - y.b); }
inline QM31 operator-(QM31 x) { return QM31(-x.a, -x.b); }
inline QM31 operator*(QM31 x, QM31 y) {
  // (a+bu)(c+du) = (ac + R bd) + (ad+bc)u
  return QM31(x.a * y.a + QM31_R() * (x.b * y.b), x.a * y.b + x.b * y.a);
}
inline QM31 operator*(QM31 x, M31 y) { return QM31(x.a * y, x.b * y); }
inline QM31 operator+(QM31 x, M31 y) { return QM31(CM31(x.a.a + y, x.a.b), x.b); }
inline QM31 operator-(QM31 x, M31 y) { return QM31(CM31(x.a.a - y, x.a.b), x.b); }
inline QM31 mul_cm31(QM31 x, CM31 y) { return QM31(x.a * y, x.b * y); }
inline QM31 inv(QM31 x) {
  // 1/(a+bu) = (a-bu)/(a^2 - R b^2)
  CM31 b2 = x.b * x.b;
  CM31 denom = x.a * x.a - QM31_R() * b2;
  CM31 di = inv(denom);
  return QM31(x.a * di, -(x.b * di));
}
inline QM31 pow(QM31 a, uint64_t e) {
  QM31 r = QM31::one();
  while (e) { if (e & 1) r = r * a; a = a * a; e >>= 1; }
  return r;
}

// Generic helpers used by templated circle code
inline M31 fone(M31) { return M31::raw(1); }
inline QM31 fone(QM31) { return QM31::one(); }
inline M31 fzero(M31) { return M31(); }
inline QM31 fzero(QM31) { return QM31::zero(); }
inline M31 fdouble(M31 x) { return x + x; }
inline QM31 fdouble(QM31 x) { return x + x; }

}  // namespace orc
