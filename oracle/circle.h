// ORACLE — TEST INFRASTRUCTURE ONLY (see fields.h header).
// Circle group over M31 / QM31, cosets, circle/line domains, index helpers.
// Restates stwo @0790eba core/circle.rs, core/poly/circle/{canonic,domain}.rs, core/poly/line.rs,
// core/utils.rs.  Pinned by the reference where noted:
//   - coset_order_to_circle_domain_order : /root/reference prover/src/trace/utils_external.rs:24-39
//   - test identity col[i] == vals[bit_reverse_index(coset_index_to_circle_domain_index(i,n),n)]
//                                          : /root/reference prover/src/trace/utils.rs:110-128
//   - CanonicCoset(log).circle_domain().half_coset as twiddle root : prover/src/machine.rs:186-194
#pragma once
#include "fields.h"

namespace orc {

static constexpr uint32_t M31_CIRCLE_LOG_ORDER = 31;

template <class F>
struct CirclePoint {
  F x, y;
  CirclePoint() {}
  CirclePoint(F x_, F y_) : x(x_), y(y_) {}
  static CirclePoint zero() { return CirclePoint(fone(F()), fzero(F())); }
  CirclePoint operator+(const CirclePoint& o) const {
    return CirclePoint(x * o.x - y * o.y, x * o.y + y * o.x);
  }
  CirclePoint conjugate() const { return CirclePoint(x, -y); }
  CirclePoint operator-() const { return conjugate(); }
  CirclePoint operator-(const CirclePoint& o) const { return *this + (-o); }
  CirclePoint dbl() const { return *this + *this; }
  CirclePoint mul(uint64_t k) const {
    CirclePoint res = zero(), cur = *this;
    while (k) { if (k & 1) res = res + cur; cur = cur.dbl(); k >>= 1; }
    return res;
  }
  bool operator==(const CirclePoint& o) const { return x == o.x && y == o.y; }
};
// x -> 2x^2 - 1
template <class F> inline F double_x(F x) { F sx = x * x; return sx + sx - fone(F()); }

// stwo: M31_CIRCLE_GEN = (2, 1268011823), order 2^31
inline CirclePoint<M31> circle_gen() { return CirclePoint<M31>(M31::raw(2), M31::raw(1268011823u)); }

// CirclePointIndex: integer mod 2^31, point = idx * G
struct PointIndex {
  uint32_t i;  // in [0, 2^31)
  PointIndex() : i(0) {}
  explicit PointIndex(uint64_t v) : i((uint32_t)(v & 0x7fffffffu)) {}
  static PointIndex subgroup_gen(uint32_t log_size) {
    assert(log_size <= M31_CIRCLE_LOG_ORDER);
    return PointIndex((uint64_t)1 << (M31_CIRCLE_LOG_ORDER - log_size));
  }
  PointIndex operator+(PointIndex o) const { return PointIndex((uint64_t)i + o.i); }
  PointIndex operator-(PointIndex o) const { return PointIndex((uint64_t)i + (1ull << 31) - o.i); }
  PointIndex operator-() const { return PointIndex((1ull << 31) - i); }
  PointIndex mul(uint64_t k) const { return PointIndex((uint64_t)i * (k & 0x7fffffffu)); }
  PointIndex half() const { assert((i & 1) == 0); return PointIndex(i >> 1); }
  CirclePoint<M31> to_point() const { return circle_gen().mul(i); }
};

struct Coset {
  PointIndex initial_index, step_size;
  uint32_t log_size;
  Coset() : log_size(0) {}
  Coset(PointIndex init, uint32_t log) : initial_index(init), step_size(PointIndex::subgroup_gen(log)), log_size(log) {}
  static Coset subgroup(uint32_t log) { return Coset(PointIndex(0), log); }
  static Coset odds(uint32_t log) { return Coset(PointIndex::subgroup_gen(log + 1), log); }
  static Coset half_odds(uint32_t log) { return Coset(PointIndex::subgroup_gen(log + 2), log); }
  size_t size() const { return (size_t)1 << log_size; }
  PointIndex index_at(size_t k) const { return initial_index + step_size.mul(k); }
  CirclePoint<M31> at(size_t k) const { return index_at(k).to_point(); }
  CirclePoint<M31> initial() const { return initial_index.to_point(); }
  CirclePoint<M31> step() const { return step_size.to_point(); }
  Coset dbl() const {
    assert(log_size > 0);
    Coset c; c.initial_index = initial_index.mul(2); c.step_size = step_size.mul(2); c.log_size = log_size - 1; return c;
  }
  Coset conjugate() const { Coset c; c.initial_index = -initial_index; c.step_size = -step_size; c.log_size = log_size; return c; }
  Coset shift(PointIndex s) const { Coset c = *this; c.initial_index = initial_index + s; return c; }
};

struct CircleDomain {
  Coset half_coset;
  CircleDomain() {}
  explicit CircleDomain(Coset h) : half_coset(h) {}
  uint32_t log_size() const { return half_coset.log_size + 1; }
  size_t size() const { return (size_t)1 << log_size(); }
  PointIndex index_at(size_t i) const {
    if (i < half_coset.size()) return half_coset.index_at(i);
    return -half_coset.index_at(i - half_coset.size());
  }
  CirclePoint<M31> at(size_t i) const { return index_at(i).to_point(); }
};

struct CanonicCoset {
  Coset coset;
  explicit CanonicCoset(uint32_t log) : coset(Coset::odds(log)) { assert(log > 0); }
  uint32_t log_size() const { return coset.log_size; }
  Coset half_coset() const { return Coset::half_odds(log_size() - 1); }
  CircleDomain circle_domain() const { return CircleDomain(half_coset()); }
  PointIndex step_size() const { return coset.step_size; }
  CirclePoint<M31> step() const { return coset.step(); }
};

// LineDomain (stwo core/poly/line.rs): x-coordinates of a coset
struct LineDomain {
  Coset coset;
  LineDomain() {}
  explicit LineDomain(Coset c) : coset(c) {}
  uint32_t log_size() const { return coset.log_size; }
  size_t size() const { return coset.size(); }
  M31 at(size_t i) const { return coset.at(i).x; }
  LineDomain dbl() const { return LineDomain(coset.dbl()); }
};

// stwo core/utils.rs
inline size_t bit_reverse_index(size_t i, uint32_t log_size) {
  if (log_size == 0) return i;
  size_t r = 0;
  for (uint32_t b = 0; b < log_size; ++b) r |= ((i >> b) & 1) << (log_size - 1 - b);
  return r;
}
inline size_t coset_index_to_circle_domain_index(size_t coset_index, uint32_t log_size) {
  if ((coset_index & 1) == 0) return coset_index / 2;
  return (((size_t)2 << log_size) - coset_index) / 2;
}
template <class T> inline void bit_reverse(std::vector<T>& v) {
  size_t n = v.size(); uint32_t lg = 0; while (((size_t)1 << lg) < n) ++lg;
  assert(((size_t)1 << lg) == n);
  for (size_t i = 0; i < n; ++i) { size_t j = bit_reverse_index(i, lg); if (i < j) std::swap(v[i], v[j]); }
}
// /root/reference prover/src/trace/utils_external.rs:24-39
template <class T> inline std::vector<T> coset_order_to_circle_domain_order(const std::vector<T>& values) {
  size_t n = values.size(), half = n / 2;
  std::vector<T> ret; ret.reserve(n);
  for (size_t i = 0; i < half; ++i) ret.push_back(values[i << 1]);
  for (size_t i = 0; i < half; ++i) ret.push_back(values[n - 1 - (i << 1)]);
  return ret;
}

// coset_vanishing (stwo core/constraints.rs): vanishing poly of a coset evaluated at p
template <class F>
inline F coset_vanishing(const Coset& coset, CirclePoint<F> p);

inline CirclePoint<QM31> to_ef(CirclePoint<M31> p) { return CirclePoint<QM31>(QM31::from_m31(p.x), QM31::from_m31(p.y)); }

template <> inline M31 coset_vanishing<M31>(const Coset& coset, CirclePoint<M31> p) {
  // Rotate the coset to the canonic one (initial = step/2), then double x log_size-1 times.
  p = p - coset.initial() + coset.step_size.half().to_point();
  M31 x = p.x;
  for (uint32_t i = 1; i < coset.log_size; ++i) x = double_x(x);
  return x;
}
template <> inline QM31 coset_vanishing<QM31>(const Coset& coset, CirclePoint<QM31> p) {
  p = p - to_ef(coset.initial()) + to_ef(coset.step_size.half().to_point());
  QM31 x = p.x;
  for (uint32_t i = 1; i < coset.log_size; ++i) x = double_x(x);
  return x;
}

}  // namespace orc
