// ORACLE — TEST INFRASTRUCTURE ONLY (see fields.h header).  C entry points for ctypes, used solely by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include "vcs.h"
#include "prove.h"
#include <memory>
#include <mutex>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

static std::map<uint32_t, std::shared_ptr<TwiddleTree>>& tw_cache() { static std::map<uint32_t, std::shared_ptr<TwiddleTree>> m; return m; }
static std::mutex tw_mu;
static const TwiddleTree& orc_get_twiddles(uint32_t domain_log);
namespace orc { const TwiddleTree& get_twiddles(uint32_t domain_log) { return orc_get_twiddles(domain_log); } }
// twiddle tree whose root is the half coset of the canonic circle domain of `domain_log`
static const TwiddleTree& orc_get_twiddles(uint32_t domain_log) {
  std::lock_guard<std::mutex> g(tw_mu);
  auto& m = tw_cache();
  // any cached bigger tree works (suffix property) but keep it simple/explicit: exact size
  auto it = m.find(domain_log);
  if (it == m.end()) {
    auto t = std::make_shared<TwiddleTree>(precompute_twiddles(CanonicCoset(domain_log).circle_domain().half_coset));
    it = m.emplace(domain_log, t).first;
  }
  return *it->second;
}

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
void orc_set_flavor(int merkle_hash, int draw_domain_sep, int pow_variant) {
  flavor().merkle_hash = merkle_hash; flavor().draw_domain_sep = draw_domain_sep; flavor().pow_variant = pow_variant;
}

// ---- fields (for the field-axiom tests) ----
uint32_t orc_m31_mul(uint32_t a, uint32_t b) { return (M31::raw(a) * M31::raw(b)).v; }
uint32_t orc_m31_inv(uint32_t a) { return inv(M31::raw(a)).v; }
void orc_qm31_mul(const uint32_t a[4], const uint32_t b[4], uint32_t out[4]) {
  QM31 r = QM31::from_u32(a[0], a[1], a[2], a[3]) * QM31::from_u32(b[0], b[1], b[2], b[3]);
  for (int k = 0; k < 4; ++k) out[k] = r.coord(k);
}
void orc_qm31_inv(const uint32_t a[4], uint32_t out[4]) {
  QM31 r = inv(QM31::from_u32(a[0], a[1], a[2], a[3]));
  for (int k = 0; k < 4; ++k) out[k] = r.coord(k);
}

// ---- circle ----
void orc_circle_domain_at(uint32_t log_size, uint64_t i, uint32_t out_xy[2]) {
  CirclePoint<M31> p = CanonicCoset(log_size).circle_domain().at(i);
  out_xy[0] = p.x.v; out_xy[1] = p.y.v;
}
uint64_t orc_bit_reverse_index(uint64_t i, uint32_t log_size) { return bit_reverse_index(i, log_size); }
uint64_t orc_coset_index_to_circle_domain_index(uint64_t i, uint32_t log_size) { return coset_index_to_circle_domain_index(i, log_size); }
// reference finalize_columns: coset order -> circle-domain order -> bit reverse
void orc_finalize_column(const uint32_t* in, uint32_t log_size, uint32_t* out) {
  size_t n = (size_t)1 << log_size;
  std::vector<uint32_t> v(in, in + n);
  std::vector<uint32_t> r = coset_order_to_circle_domain_order(v);
  bit_reverse(r);
  memcpy(out, r.data(), n * 4);
}

// ---- poly ----
void orc_twiddles(uint32_t domain_log, uint32_t* tw, uint32_t* itw) {
  const TwiddleTree& t = orc_get_twiddles(domain_log);
  for (size_t i = 0; i < t.twiddles.size(); ++i) { tw[i] = t.twiddles[i].v; itw[i] = t.itwiddles[i].v; }
}
void orc_interpolate(uint32_t log_size, const uint32_t* evals, uint32_t* coeffs) {
  const TwiddleTree& t = orc_get_twiddles(log_size);
  size_t n = (size_t)1 << log_size;
  Col v(n); for (size_t i = 0; i < n; ++i) v[i] = M31::raw(evals[i]);
  Col c = interpolate(CanonicCoset(log_size).circle_domain(), std::move(v), t);
  for (size_t i = 0; i < n; ++i) coeffs[i] = c[i].v;
}
void orc_evaluate(uint32_t coeff_log, uint32_t domain_log, const uint32_t* coeffs, uint32_t* evals) {
  const TwiddleTree& t = orc_get_twiddles(domain_log);
  size_t n = (size_t)1 << coeff_log;
  Col c(n); for (size_t i = 0; i < n; ++i) c[i] = M31::raw(coeffs[i]);
  Col v = evaluate(CanonicCoset(domain_log).circle_domain(), c, t);
  for (size_t i = 0; i < v.size(); ++i) evals[i] = v[i].v;
}
// batched, OpenMP over columns (cpu_baseline): evals (n_cols x 2^log) -> lde (n_cols x 2^(log+blowup)); column-major contiguous
void orc_interpolate_evaluate_batch(uint32_t log_size, uint32_t log_blowup, size_t n_cols, const uint32_t* evals, uint32_t* coeffs_out, uint32_t* lde_out) {
  const TwiddleTree& t1 = orc_get_twiddles(log_size);
  const TwiddleTree& t2 = orc_get_twiddles(log_size + log_blowup);
  size_t n = (size_t)1 << log_size, m = (size_t)1 << (log_size + log_blowup);
  CircleDomain d1 = CanonicCoset(log_size).circle_domain(), d2 = CanonicCoset(log_size + log_blowup).circle_domain();
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < n_cols; ++c) {
    Col v(n); for (size_t i = 0; i < n; ++i) v[i] = M31::raw(evals[c * n + i]);
    Col co = interpolate(d1, std::move(v), t1);
    if (coeffs_out) for (size_t i = 0; i < n; ++i) coeffs_out[c * n + i] = co[i].v;
    Col e = evaluate(d2, co, t2);
    for (size_t i = 0; i < m; ++i) lde_out[c * m + i] = e[i].v;
  }
}
void orc_eval_at_point(uint32_t log_size, const uint32_t* coeffs, const uint32_t px[4], const uint32_t py[4], uint32_t out[4]) {
  size_t n = (size_t)1 << log_size;
  Col c(n); for (size_t i = 0; i < n; ++i) c[i] = M31::raw(coeffs[i]);
  QM31 r = eval_at_point(c, CirclePoint<QM31>(QM31::from_u32(px[0], px[1], px[2], px[3]), QM31::from_u32(py[0], py[1], py[2], py[3])));
  for (int k = 0; k < 4; ++k) out[k] = r.coord(k);
}

// ---- hash / merkle ----
void orc_blake2s(const uint8_t* data, size_t len, uint8_t out[32]) { Hash32 h = blake2s_hash(data, len); memcpy(out, h.data(), 32); }
void orc_blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t f0, uint32_t f1) { b2s_compress(h, m, t0, t1, f0, f1); }

// columns: n_cols pointers, log sizes. layers_out (optional): concatenated layers from root (1 hash) to leaves.
void orc_merkle_commit(size_t n_cols, const uint32_t* const* cols, const uint32_t* log_sizes, uint8_t root[32], uint8_t* layers_out) {
  std::vector<Col> store(n_cols);
  std::vector<const Col*> ptrs(n_cols);
  for (size_t c = 0; c < n_cols; ++c) {
    size_t n = (size_t)1 << log_sizes[c];
    store[c].resize(n);
    for (size_t i = 0; i < n; ++i) store[c][i] = M31::raw(cols[c][i]);
    ptrs[c] = &store[c];
  }
  MerkleProver mp = MerkleProver::commit(ptrs);
  memcpy(root, mp.root().data(), 32);
  if (layers_out) {
    size_t off = 0;
    for (auto& L : mp.layers) { memcpy(layers_out + off, L.data(), L.size() * 32); off += L.size() * 32; }
  }
}

// Decommit: queries given per log size as (log_sizes_q[k], positions concatenated with counts).
// Outputs are written to caller buffers with lengths returned through *n_*.
void orc_merkle_decommit(size_t n_cols, const uint32_t* const* cols, const uint32_t* log_sizes,
                         size_t n_q_sizes, const uint32_t* q_log_sizes, const uint64_t* q_counts, const uint64_t* q_positions,
                         uint32_t* queried_values, size_t* n_queried, uint8_t* hash_witness, size_t* n_hash, uint32_t* column_witness, size_t* n_colw) {
  std::vector<Col> store(n_cols);
  std::vector<const Col*> ptrs(n_cols);
  for (size_t c = 0; c < n_cols; ++c) {
    size_t n = (size_t)1 << log_sizes[c];
    store[c].resize(n);
    for (size_t i = 0; i < n; ++i) store[c][i] = M31::raw(cols[c][i]);
    ptrs[c] = &store[c];
  }
  MerkleProver mp = MerkleProver::commit(ptrs);
  std::map<uint32_t, std::vector<size_t>> q;
  size_t off = 0;
  for (size_t k = 0; k < n_q_sizes; ++k) {
    std::vector<size_t> v(q_positions + off, q_positions + off + q_counts[k]);
    off += q_counts[k];
    q[q_log_sizes[k]] = v;
  }
  auto res = mp.decommit(q, ptrs);
  *n_queried = res.first.size(); *n_hash = res.second.hash_witness.size(); *n_colw = res.second.column_witness.size();
  for (size_t i = 0; i < res.first.size(); ++i) queried_values[i] = res.first[i].v;
  for (size_t i = 0; i < res.second.hash_witness.size(); ++i) memcpy(hash_witness + 32 * i, res.second.hash_witness[i].data(), 32);
  for (size_t i = 0; i < res.second.column_witness.size(); ++i) column_witness[i] = res.second.column_witness[i].v;
}

// ---- channel (opaque handle) ----
void* orc_channel_new() { return new Channel(); }
void orc_channel_free(void* c) { delete (Channel*)c; }
void orc_channel_digest(void* c, uint8_t out[32]) { memcpy(out, ((Channel*)c)->digest.data(), 32); }
void orc_channel_mix_u64(void* c, uint64_t v) { ((Channel*)c)->mix_u64(v); }
void orc_channel_mix_u32s(void* c, const uint32_t* w, size_t n) { ((Channel*)c)->mix_u32s(w, n); }
void orc_channel_mix_felts(void* c, const uint32_t* felts, size_t n) {
  std::vector<QM31> f(n);
  for (size_t i = 0; i < n; ++i) f[i] = QM31::from_u32(felts[4 * i], felts[4 * i + 1], felts[4 * i + 2], felts[4 * i + 3]);
  ((Channel*)c)->mix_felts(f);
}
void orc_channel_mix_root(void* c, const uint8_t root[32]) { Hash32 h; memcpy(h.data(), root, 32); ((Channel*)c)->mix_root(h); }
void orc_channel_draw_felt(void* c, uint32_t out[4]) { QM31 q = ((Channel*)c)->draw_felt(); for (int k = 0; k < 4; ++k) out[k] = q.coord(k); }
void orc_channel_draw_felts(void* c, size_t n, uint32_t* out) {
  auto v = ((Channel*)c)->draw_felts(n);
  for (size_t i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) out[4 * i + k] = v[i].coord(k);
}
void orc_channel_draw_random_bytes(void* c, uint8_t out[32]) { Hash32 h = ((Channel*)c)->draw_random_bytes(); memcpy(out, h.data(), 32); }

}  // extern "C"
