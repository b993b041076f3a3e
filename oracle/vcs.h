// ORACLE — TEST INFRASTRUCTURE ONLY (see fields.h header).  PARITY UNPINNED: the reference holds no
// expected roots / transcripts; hash-construction variants are behind `Flavor`.
// Restates stwo @0790eba:
//   core/vcs/blake2_merkle.rs  (Blake2sMerkleHasher::hash_node, Blake2sMerkleChannel::mix_root)
//   prover/vcs/prover.rs       (MerkleProver::commit / decommit)
//   core/channel/blake2s.rs    (Blake2sChannel)
// Reference call sites: prover/src/machine.rs:197-206 (channel mixes), :208-263 (tree commits).
#pragma once
#include "blake2s.h"
#include "poly.h"
#include <map>

namespace orc {

// Transcript-affecting variants with more than one plausible form in upstream history.
struct Flavor {
  // 0: hash_node = raw Blake2s compression chained from an all-zero state, zero counters/flags,
  //    first block = left||right, then 16 column words per block, zero padded (what SIMD compress16 computes)
  // 1: hash_node = standard Blake2s-256 over left||right||le_bytes(values)
  int merkle_hash = 0;
  // append a 0x00 domain-separation byte in draw_random_bytes (newer upstream revs)
  int draw_domain_sep = 0;
  // PoW predicate: 0 = trailing_zeros(H(digest || nonce_le)) >= bits ; 1 = prefixed variant
  int pow_variant = 0;
};
inline Flavor& flavor() { static Flavor f; return f; }

inline Hash32 hash_node(const Hash32* left, const Hash32* right, const M31* vals, size_t n_vals) {
  if (flavor().merkle_hash == 0) {
    uint32_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t m[16];
    if (left) {
      memcpy(m, left->data(), 32); memcpy(m + 8, right->data(), 32);
      b2s_compress(st, m, 0, 0, 0, 0);
    }
    for (size_t i = 0; i < n_vals; i += 16) {
      for (size_t j = 0; j < 16; ++j) m[j] = (i + j < n_vals) ? vals[i + j].v : 0;
      b2s_compress(st, m, 0, 0, 0, 0);
    }
    Hash32 out; memcpy(out.data(), st, 32);
    return out;
  } else {
    Blake2s b;
    if (left) { b.update(left->data(), 32); b.update(right->data(), 32); }
    for (size_t i = 0; i < n_vals; ++i) { uint32_t w = vals[i].v; b.update(&w, 4); }
    return b.finalize();
  }
}

struct MerkleDecommitment {
  std::vector<Hash32> hash_witness;
  std::vector<M31> column_witness;
};

struct MerkleProver {
  // layers[0] = root layer (1 hash) ... layers[max_log] = leaves
  std::vector<std::vector<Hash32>> layers;
  Hash32 root() const { return layers[0][0]; }

  static std::vector<const Col*> sort_cols(const std::vector<const Col*>& cols) {
    std::vector<const Col*> s = cols;
    std::stable_sort(s.begin(), s.end(), [](const Col* a, const Col* b) { return a->size() > b->size(); });
    return s;
  }
  static uint32_t ilog2(size_t n) { uint32_t l = 0; while (((size_t)1 << (l + 1)) <= n) ++l; return l; }

  static MerkleProver commit(const std::vector<const Col*>& columns) {
    MerkleProver mp;
    if (columns.empty()) {
      mp.layers.push_back({hash_node(nullptr, nullptr, nullptr, 0)});
      return mp;
    }
    std::vector<const Col*> s = sort_cols(columns);
    uint32_t max_log = ilog2(s[0]->size());
    size_t ci = 0;
    std::vector<std::vector<Hash32>> rev;  // from leaves down to the root
    for (int log_size = (int)max_log; log_size >= 0; --log_size) {
      std::vector<const Col*> layer_cols;
      while (ci < s.size() && ilog2(s[ci]->size()) == (uint32_t)log_size) layer_cols.push_back(s[ci++]);
      size_t n = (size_t)1 << log_size;
      std::vector<Hash32> layer(n);
      const std::vector<Hash32>* prev = rev.empty() ? nullptr : &rev.back();
      std::vector<M31> row(layer_cols.size());
#pragma omp parallel for firstprivate(row) schedule(static)
      for (size_t i = 0; i < n; ++i) {
        for (size_t c = 0; c < layer_cols.size(); ++c) row[c] = (*layer_cols[c])[i];
        if (prev) layer[i] = hash_node(&(*prev)[2 * i], &(*prev)[2 * i + 1], row.data(), row.size());
        else layer[i] = hash_node(nullptr, nullptr, row.data(), row.size());
      }
      rev.push_back(std::move(layer));
    }
    mp.layers.assign(rev.rbegin(), rev.rend());
    return mp;
  }

  // queries_per_log_size: log_size -> sorted positions.  Returns queried values (flat) + decommitment.
  std::pair<std::vector<M31>, MerkleDecommitment> decommit(const std::map<uint32_t, std::vector<size_t>>& queries_per_log_size,
                                                          const std::vector<const Col*>& columns) const {
    std::vector<M31> queried_values;
    MerkleDecommitment d;
    std::vector<const Col*> s = sort_cols(columns);
    size_t ci = 0;
    std::vector<size_t> last_layer_queries;
    for (int layer_log = (int)layers.size() - 1; layer_log >= 0; --layer_log) {
      std::vector<size_t> layer_total_queries;
      std::vector<const Col*> layer_cols;
      while (ci < s.size() && ilog2(s[ci]->size()) == (uint32_t)layer_log) layer_cols.push_back(s[ci++]);
      const std::vector<Hash32>* prev_hashes = ((size_t)layer_log + 1 < layers.size()) ? &layers[layer_log + 1] : nullptr;
      size_t pq = 0;  // cursor into last_layer_queries
      static const std::vector<size_t> empty;
      auto it = queries_per_log_size.find((uint32_t)layer_log);
      const std::vector<size_t>& lq = it == queries_per_log_size.end() ? empty : it->second;
      size_t cq = 0;
      while (true) {
        // next_decommitment_node: min(prev.peek()/2, layer.peek())
        bool has_p = pq < last_layer_queries.size(), has_c = cq < lq.size();
        if (!has_p && !has_c) break;
        size_t node;
        if (has_p && has_c) node = std::min(last_layer_queries[pq] / 2, lq[cq]);
        else if (has_p) node = last_layer_queries[pq] / 2;
        else node = lq[cq];
        if (prev_hashes) {
          if (pq < last_layer_queries.size() && last_layer_queries[pq] == 2 * node) ++pq;
          else d.hash_witness.push_back((*prev_hashes)[2 * node]);
          if (pq < last_layer_queries.size() && last_layer_queries[pq] == 2 * node + 1) ++pq;
          else d.hash_witness.push_back((*prev_hashes)[2 * node + 1]);
        }
        bool queried = cq < lq.size() && lq[cq] == node;
        if (queried) ++cq;
        for (const Col* c : layer_cols) {
          if (queried) queried_values.push_back((*c)[node]);
          else d.column_witness.push_back((*c)[node]);
        }
        layer_total_queries.push_back(node);
      }
      last_layer_queries = layer_total_queries;
    }
    return {queried_values, d};
  }
};

// core/channel/blake2s.rs
struct Channel {
  Hash32 digest;
  size_t n_challenges = 0, n_sent = 0;
  Channel() { digest.fill(0); }
  void update_digest(const Hash32& d) { digest = d; n_challenges += 1; n_sent = 0; }
  void mix_u32s(const uint32_t* w, size_t n) {
    Blake2s b; b.update(digest.data(), 32);
    for (size_t i = 0; i < n; ++i) b.update(&w[i], 4);
    update_digest(b.finalize());
  }
  void mix_u64(uint64_t v) { uint32_t w[2] = {(uint32_t)v, (uint32_t)(v >> 32)}; mix_u32s(w, 2); }
  void mix_felts(const std::vector<QM31>& f) {
    Blake2s b; b.update(digest.data(), 32);
    for (const QM31& q : f) for (int k = 0; k < 4; ++k) { uint32_t w = q.coord(k); b.update(&w, 4); }
    update_digest(b.finalize());
  }
  void mix_root(const Hash32& root) {  // Blake2sMerkleChannel::mix_root
    Blake2s b; b.update(digest.data(), 32); b.update(root.data(), 32);
    update_digest(b.finalize());
  }
  Hash32 draw_random_bytes() {
    uint8_t in[65]; memcpy(in, digest.data(), 32); memset(in + 32, 0, 33);
    uint64_t c = n_sent;
    for (int i = 0; i < 8; ++i) in[32 + i] = (uint8_t)(c >> (8 * i));
    n_sent += 1;
    return blake2s_hash(in, flavor().draw_domain_sep ? 65 : 64);
  }
  void draw_base_felts(M31 out[8]) {
    while (true) {
      Hash32 r = draw_random_bytes();
      uint32_t w[8]; memcpy(w, r.data(), 32);
      bool ok = true;
      for (int i = 0; i < 8; ++i) if (w[i] >= 2 * P) ok = false;
      if (ok) { for (int i = 0; i < 8; ++i) out[i] = M31::reduce(w[i]); return; }
    }
  }
  QM31 draw_felt() { M31 f[8]; draw_base_felts(f); return QM31(CM31(f[0], f[1]), CM31(f[2], f[3])); }
  std::vector<QM31> draw_felts(size_t n) {
    std::vector<QM31> out; M31 f[8]; int have = 0;
    std::vector<M31> pool;
    while (out.size() < n) {
      if (have == 0) { draw_base_felts(f); have = 8; }
      int o = 8 - have;
      out.push_back(QM31(CM31(f[o], f[o + 1]), CM31(f[o + 2], f[o + 3])));
      have -= 4;
    }
    return out;
  }
  uint32_t trailing_zeros() const {
    // u128 from the first 16 digest bytes, LE
    uint32_t tz = 0;
    for (int i = 0; i < 16; ++i) {
      uint8_t b = digest[i];
      if (b == 0) { tz += 8; continue; }
      while ((b & 1) == 0) { ++tz; b >>= 1; }
      return tz;
    }
    return 128;
  }
};

}  // namespace orc
