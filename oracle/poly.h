// ORACLE — TEST INFRASTRUCTURE ONLY (see fields.h header).  PARITY UNPINNED (no golden vectors in the
// reference); mathematically pinned by the self-checks in tests/test_oracle_golden.py:
//   eval_at_point(interpolate(v), domain.at(bitrev(i))) == v[i], iFFT∘FFT = id, twiddles == domain points.
// Restates stwo @0790eba prover/backend/cpu/circle.rs (precompute_twiddles / interpolate / evaluate /
// eval_at_point), prover/poly/twiddles.rs, core/fft.rs (butterfly / ibutterfly), core/poly/utils.rs (fold).
// Reference call sites: prover/src/machine.rs:186-194 (twiddles), :208-263 (extend_evals -> interpolate,
// commit -> evaluate on the blown-up canonic domain).
#pragma once
#include "circle.h"
#include <algorithm>

namespace orc {

typedef std::vector<M31> Col;

// core/fft.rs
inline void butterfly(M31& v0, M31& v1, M31 twid) { M31 tmp = v1 * twid; v1 = v0 - tmp; v0 = v0 + tmp; }
inline void ibutterfly(M31& v0, M31& v1, M31 itwid) { M31 tmp = v0; v0 = tmp + v1; v1 = (tmp - v1) * itwid; }
inline void butterfly(QM31& v0, QM31& v1, M31 twid) { QM31 tmp = v1 * twid; v1 = v0 - tmp; v0 = v0 + tmp; }
inline void ibutterfly(QM31& v0, QM31& v1, M31 itwid) { QM31 tmp = v0; v0 = tmp + v1; v1 = (tmp - v1) * itwid; }

struct TwiddleTree {
  Coset root_coset;
  std::vector<M31> twiddles, itwiddles;
};

// backend/cpu/circle.rs: slow_precompute_twiddles
inline std::vector<M31> slow_precompute_twiddles(Coset coset) {
  std::vector<M31> tw; tw.reserve(coset.size());
  uint32_t logn = coset.log_size;
  for (uint32_t l = 0; l < logn; ++l) {
    size_t i0 = tw.size(), half = coset.size() / 2;
    std::vector<M31> xs(half);
    CirclePoint<M31> cur = coset.initial(), st = coset.step();
    for (size_t k = 0; k < half; ++k) { xs[k] = cur.x; cur = cur + st; }
    bit_reverse(xs);
    tw.insert(tw.end(), xs.begin(), xs.end());
    (void)i0;
    coset = coset.dbl();
  }
  tw.push_back(M31::raw(1));  // pad to a power of two
  return tw;
}
inline TwiddleTree precompute_twiddles(Coset coset) {
  TwiddleTree t; t.root_coset = coset;
  t.twiddles = slow_precompute_twiddles(coset);
  t.itwiddles.resize(t.twiddles.size());
  // batch inverse (Montgomery trick); element-wise inverse, result is order-independent
  std::vector<M31> pre(t.twiddles.size());
  M31 acc = M31::raw(1);
  for (size_t i = 0; i < pre.size(); ++i) { pre[i] = acc; acc = acc * t.twiddles[i]; }
  M31 ia = inv(acc);
  for (size_t i = pre.size(); i-- > 0;) { t.itwiddles[i] = ia * pre[i]; ia = ia * t.twiddles[i]; }
  return t;
}

// prover/poly/twiddles / backend/cpu/circle.rs: domain_line_twiddles_from_tree
// returns, for line layer k (k = 0 is the biggest), pointer+len into `buf`
struct Slice { const M31* p; size_t n; };
inline std::vector<Slice> domain_line_twiddles_from_tree(const CircleDomain& domain, const std::vector<M31>& buf) {
  uint32_t k = domain.half_coset.log_size;
  assert(((size_t)1 << k) <= buf.size());
  std::vector<Slice> out;
  for (uint32_t i = 0; i < k; ++i) {
    size_t len = (size_t)1 << i;
    out.push_back(Slice{buf.data() + buf.size() - len * 2, len});
  }
  std::reverse(out.begin(), out.end());
  return out;
}
// circle_twiddles_from_line_twiddles: [x, y] -> [y, -y, -x, x]
inline std::vector<M31> circle_twiddles_from_line_twiddles(Slice first) {
  std::vector<M31> out; out.reserve(first.n * 2);
  for (size_t i = 0; i + 1 < first.n; i += 2) {
    M31 x = first.p[i], y = first.p[i + 1];
    out.push_back(y); out.push_back(-y); out.push_back(-x); out.push_back(x);
  }
  return out;
}

template <class BF, class V>
inline void fft_layer_loop(std::vector<V>& values, uint32_t i, size_t h, M31 t, BF bf) {
  for (size_t l = 0; l < ((size_t)1 << i); ++l) {
    size_t idx0 = (h << (i + 1)) + l, idx1 = idx0 + ((size_t)1 << i);
    V a = values[idx0], b = values[idx1];
    bf(a, b, t);
    values[idx0] = a; values[idx1] = b;
  }
}

// interpolate: values on `domain` (bit-reversed) -> coefficients
inline Col interpolate(const CircleDomain& domain, Col values, const TwiddleTree& tw) {
  uint32_t log_size = domain.log_size();
  assert(values.size() == domain.size());
  if (log_size == 1) {
    M31 y = domain.half_coset.initial().y;
    M31 n = M31::raw(2);
    M31 yn_inv = inv(y * n), y_inv = yn_inv * n, n_inv = yn_inv * y;
    M31 v0 = values[0], v1 = values[1];
    ibutterfly(v0, v1, y_inv);
    return Col{v0 * n_inv, v1 * n_inv};
  }
  if (log_size == 2) {
    CirclePoint<M31> p = domain.half_coset.initial();
    M31 n = M31::raw(4);
    M31 xyn_inv = inv(p.x * p.y * n);
    M31 x_inv = xyn_inv * p.y * n, y_inv = xyn_inv * p.x * n, n_inv = xyn_inv * p.x * p.y;
    M31 v0 = values[0], v1 = values[1], v2 = values[2], v3 = values[3];
    ibutterfly(v0, v1, y_inv);
    ibutterfly(v2, v3, -y_inv);
    ibutterfly(v0, v2, x_inv);
    ibutterfly(v1, v3, x_inv);
    return Col{v0 * n_inv, v1 * n_inv, v2 * n_inv, v3 * n_inv};
  }
  std::vector<Slice> line = domain_line_twiddles_from_tree(domain, tw.itwiddles);
  std::vector<M31> circ = circle_twiddles_from_line_twiddles(line[0]);
  for (size_t h = 0; h < circ.size(); ++h)
    fft_layer_loop(values, 0, h, circ[h], [](M31& a, M31& b, M31 t) { ibutterfly(a, b, t); });
  for (size_t layer = 0; layer < line.size(); ++layer)
    for (size_t h = 0; h < line[layer].n; ++h)
      fft_layer_loop(values, (uint32_t)layer + 1, h, line[layer].p[h], [](M31& a, M31& b, M31 t) { ibutterfly(a, b, t); });
  M31 sc = inv(M31::raw((uint32_t)domain.size()));
  for (auto& v : values) v = v * sc;
  return values;
}

// evaluate: coefficients (log size <= domain log size; zero-extended) -> values on `domain` (bit-reversed)
inline Col evaluate(const CircleDomain& domain, const Col& coeffs, const TwiddleTree& tw) {
  uint32_t log_size = domain.log_size();
  assert(coeffs.size() <= domain.size());
  Col values = coeffs;
  values.resize(domain.size());  // CirclePoly::extend — zero pad
  if (log_size == 1) {
    M31 v0 = values[0], v1 = values[1];
    butterfly(v0, v1, domain.half_coset.initial().y);
    return Col{v0, v1};
  }
  if (log_size == 2) {
    CirclePoint<M31> p = domain.half_coset.initial();
    M31 v0 = values[0], v1 = values[1], v2 = values[2], v3 = values[3];
    butterfly(v0, v2, p.x);
    butterfly(v1, v3, p.x);
    butterfly(v0, v1, p.y);
    butterfly(v2, v3, -p.y);
    return Col{v0, v1, v2, v3};
  }
  std::vector<Slice> line = domain_line_twiddles_from_tree(domain, tw.twiddles);
  std::vector<M31> circ = circle_twiddles_from_line_twiddles(line[0]);
  for (size_t layer = line.size(); layer-- > 0;)
    for (size_t h = 0; h < line[layer].n; ++h)
      fft_layer_loop(values, (uint32_t)layer + 1, h, line[layer].p[h], [](M31& a, M31& b, M31 t) { butterfly(a, b, t); });
  for (size_t h = 0; h < circ.size(); ++h)
    fft_layer_loop(values, 0, h, circ[h], [](M31& a, M31& b, M31 t) { butterfly(a, b, t); });
  return values;
}

// core/poly/utils.rs fold + backend/cpu/circle.rs eval_at_point
inline QM31 fold_rec(const M31* values, size_t n, const QM31* factors) {
  if (n == 1) return QM31::from_m31(values[0]);
  QM31 l = fold_rec(values, n / 2, factors + 1);
  QM31 r = fold_rec(values + n / 2, n / 2, factors + 1);
  return l + r * factors[0];
}
inline QM31 eval_at_point(const Col& coeffs, CirclePoint<QM31> point) {
  uint32_t lg = 0; while (((size_t)1 << lg) < coeffs.size()) ++lg;
  if (lg == 0) return QM31::from_m31(coeffs[0]);
  std::vector<QM31> mappings; mappings.push_back(point.y);
  QM31 x = point.x;
  for (uint32_t i = 1; i < lg; ++i) { mappings.push_back(x); x = double_x(x); }
  std::reverse(mappings.begin(), mappings.end());
  return fold_rec(coeffs.data(), coeffs.size(), mappings.data());
}

}  // namespace orc
