// Host-side circle-group geometry for the product library (cosets, canonic circle domains, line domains).
// Mirrors the *semantics* of stwo core/circle.rs + core/poly/circle/{canonic,domain}.rs + core/poly/line.rs
// as used by the reference at prover/src/machine.rs:186-194 and prover/src/trace/trace_builder.rs:156-164.
// Points are produced from integer point indices (multiples of the generator) through a 31-entry table of
// generator doublings, so any domain point costs <= 31 group additions.
#pragma once
#include "m31.cuh"
#include <cstddef>
#include <cassert>

namespace nb {

static constexpr u32 CIRCLE_LOG_ORDER = 31;
static constexpr u32 CIRCLE_GEN_X = 2, CIRCLE_GEN_Y = 1268011823u;

struct GenTable {
  cpoint pow2[31];  // pow2[b] = 2^b * G
  GenTable() {
    cpoint g{CIRCLE_GEN_X, CIRCLE_GEN_Y};
    for (int b = 0; b < 31; ++b) { pow2[b] = g; g = cp_add(g, g); }
  }
};
inline const GenTable& gen_table() { static GenTable t; return t; }

inline cpoint index_to_point(u32 idx) {
  idx &= 0x7fffffffu;
  cpoint r{1, 0};
  const GenTable& t = gen_table();
  for (int b = 0; b < 31; ++b) if ((idx >> b) & 1u) r = cp_add(r, t.pow2[b]);
  return r;
}
inline u32 idx_add(u32 a, u32 b) { return (a + b) & 0x7fffffffu; }
inline u32 idx_neg(u32 a) { return (0x80000000u - a) & 0x7fffffffu; }
inline u32 idx_mul(u32 a, u64 k) { return (u32)(((u64)a * (k & 0x7fffffffu)) & 0x7fffffffu); }
inline u32 subgroup_gen_index(u32 log_size) { assert(log_size <= 31); return (u32)(1ull << (31 - log_size)) & 0x7fffffffu; }

struct HCoset {
  u32 initial_index, step_index, log_size;
  static HCoset make(u32 init, u32 log) { return HCoset{init & 0x7fffffffu, subgroup_gen_index(log), log}; }
  static HCoset odds(u32 log) { return make(subgroup_gen_index(log + 1), log); }
  static HCoset half_odds(u32 log) { return make(subgroup_gen_index(log + 2), log); }
  size_t size() const { return (size_t)1 << log_size; }
  u32 index_at(size_t k) const { return idx_add(initial_index, idx_mul(step_index, k)); }
  cpoint at(size_t k) const { return index_to_point(index_at(k)); }
  HCoset dbl() const { return HCoset{idx_mul(initial_index, 2), idx_mul(step_index, 2), log_size - 1}; }
};

// CanonicCoset(log).circle_domain(): half_coset = half_odds(log-1); at(i) = half.at(i) for i < n/2 else -half.at(i-n/2)
struct HCircleDomain {
  HCoset half;
  static HCircleDomain canonic(u32 log_size) { assert(log_size >= 1); return HCircleDomain{HCoset::half_odds(log_size - 1)}; }
  u32 log_size() const { return half.log_size + 1; }
  size_t size() const { return (size_t)1 << log_size(); }
  u32 index_at(size_t i) const {
    size_t h = half.size();
    return i < h ? half.index_at(i) : idx_neg(half.index_at(i - h));
  }
  cpoint at(size_t i) const { return index_to_point(index_at(i)); }
};
// trace step of CanonicCoset(log): coset = odds(log), step index = subgroup_gen(log)
inline u32 canonic_step_index(u32 log_size) { return subgroup_gen_index(log_size); }

struct HLineDomain {
  HCoset coset;
  static HLineDomain make(HCoset c) { return HLineDomain{c}; }
  u32 log_size() const { return coset.log_size; }
  size_t size() const { return coset.size(); }
  u32 at(size_t i) const { return coset.at(i).x; }
  HLineDomain dbl() const { return HLineDomain{coset.dbl()}; }
};

}  // namespace nb
