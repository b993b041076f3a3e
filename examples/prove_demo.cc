// A native (C++) host for the C ABI of include/nb200.h: the `Machine::prove_with_extensions` sequence of the reference
// (/root/reference prover/src/machine.rs:130-297) driven without Python — what the Rust shim of INTEGRATION.md does, in the only
// compiled host language this image has.  Reads a job file (AIR bytecode, PcsConfig, host trace columns in coset order, the lookup
// relations to draw and the per-component logup metadata), proves on GPU 0 and writes the postcard proof bytes.
//
//   g++ -O2 -std=c++17 -Iinclude examples/prove_demo.cc -Lnexus_zkvm_b200 -lnexus_b200 -Wl,-rpath,'$ORIGIN/../nexus_zkvm_b200' -o examples/prove_demo
//   examples/prove_demo job.bin proof.bin
//
// Job file (little-endian u32 words), written by tests/test_gpu_native_host.py:
//   'NBJB', 1, n_words, words[n_words], pow_bits, log_blowup, log_last, n_queries,
//   n_assoc, assoc bytes (one per word), n_log_sizes, log_sizes[],
//   n_relations, per relation: z_param, size, alpha_param[size],
//   n_components, per component: log_size, n_logup_cols, cumsum_shift_param,
//   2 trees (preprocessed, main), per tree: n_batches, per batch: n_cols, log_size, data[n_cols << log_size].
#include "nb200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
const uint32_t P = 0x7fffffffu;
// host M31 / QM31 arithmetic for the two protocol-level computations the host owns (alpha powers, claimed_sum / 2^n)
uint32_t madd(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; return (uint32_t)(s >= P ? s - P : s); }
uint32_t msub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + P - b; }
uint32_t mmul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
uint32_t mpow(uint32_t a, uint32_t e) { uint32_t r = 1; while (e) { if (e & 1) r = mmul(r, a); a = mmul(a, a); e >>= 1; } return r; }
struct Q { uint32_t c[4]; };
Q qmul(const Q& x, const Q& y) {  // (a + bu)(c + du), u^2 = 2 + i, i^2 = -1
  auto cmul = [](uint32_t ar, uint32_t ai, uint32_t br, uint32_t bi, uint32_t& r, uint32_t& i) {
    r = msub(mmul(ar, br), mmul(ai, bi)); i = madd(mmul(ar, bi), mmul(ai, br));
  };
  uint32_t acr, aci, bdr, bdi, adr, adi, bcr, bci;
  cmul(x.c[0], x.c[1], y.c[0], y.c[1], acr, aci);
  cmul(x.c[2], x.c[3], y.c[2], y.c[3], bdr, bdi);
  cmul(x.c[0], x.c[1], y.c[2], y.c[3], adr, adi);
  cmul(x.c[2], x.c[3], y.c[0], y.c[1], bcr, bci);
  uint32_t rr = msub(madd(bdr, bdr), bdi), ri = madd(bdr, madd(bdi, bdi));  // (2 + i) * bd
  return Q{{madd(acr, rr), madd(aci, ri), madd(adr, bcr), madd(adi, bci)}};
}

struct Reader {
  std::vector<uint32_t> w; size_t i = 0;
  uint32_t next() { if (i >= w.size()) { fprintf(stderr, "job file truncated\n"); exit(2); } return w[i++]; }
  const uint32_t* take(size_t n) { if (i + n > w.size()) { fprintf(stderr, "job file truncated\n"); exit(2); } const uint32_t* p = &w[i]; i += n; return p; }
};

#define CHECK(ctx, expr) do { nb200_status _s = (expr); if (_s != NB200_OK) { fprintf(stderr, "%s -> status %d: %s\n", #expr, _s, nb200_last_error(ctx)); exit(10 + _s); } } while (0)
}  // namespace

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s job.bin proof.bin\n", argv[0]); return 2; }
  Reader rd;
  {
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    rd.w.resize((size_t)n / 4);
    if (fread(rd.w.data(), 4, rd.w.size(), f) != rd.w.size()) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);
  }
  if (rd.next() != 0x424a424eu /* 'NBJB' */ || rd.next() != 1) { fprintf(stderr, "not a job file\n"); return 2; }
  const uint32_t n_words = rd.next();
  const uint32_t* words = rd.take(n_words);
  const uint32_t pow_bits = rd.next(), log_blowup = rd.next(), log_last = rd.next(), n_queries = rd.next();

  nb200_ctx* ctx = nullptr;
  if (nb200_ctx_create(0, &ctx) != NB200_OK) { fprintf(stderr, "no context: %s\n", nb200_last_error(nullptr)); return 3; }
  nb200_channel* ch = nullptr;
  CHECK(ctx, nb200_channel_new(ctx, &ch));
  for (uint32_t n = rd.next(), k = 0; k < n; ++k) nb200_channel_mix_u64(ch, rd.next());      // associated data, machine.rs:197-200
  for (uint32_t n = rd.next(), k = 0; k < n; ++k) nb200_channel_mix_u64(ch, rd.next());      // log sizes, machine.rs:204-206

  nb200_air* air = nullptr;
  CHECK(ctx, nb200_air_load(ctx, words, n_words, &air));
  nb200_scheme* scheme = nullptr;
  CHECK(ctx, nb200_scheme_new(ctx, pow_bits, log_blowup, log_last, n_queries, &scheme));
  CHECK(ctx, nb200_scheme_set_constraint_log_degree(scheme, nb200_air_max_log_expand(air)));

  struct Rel { uint32_t z, size; std::vector<uint32_t> alpha; };
  std::vector<Rel> rels(rd.next());
  for (auto& r : rels) { r.z = rd.next(); r.size = rd.next(); const uint32_t* a = rd.take(r.size); r.alpha.assign(a, a + r.size); }
  struct Comp { uint32_t log_size, n_logup_cols, shift_param; };
  std::vector<Comp> comps(rd.next());
  for (auto& c : comps) { c.log_size = rd.next(); c.n_logup_cols = rd.next(); c.shift_param = rd.next(); }

  // trees 0 and 1 from host columns (machine.rs:208-237): pinned staging so that the copies run at link speed
  std::vector<std::vector<nb200_cols*>> evals(2);
  for (int t = 0; t < 2; ++t) {
    const uint32_t nb = rd.next();
    std::vector<const uint32_t*> ptrs(nb); std::vector<size_t> ncols(nb); std::vector<uint32_t> logs(nb); std::vector<void*> pinned(nb);
    for (uint32_t b = 0; b < nb; ++b) {
      ncols[b] = rd.next(); logs[b] = rd.next();
      const size_t n = ncols[b] << logs[b];
      const uint32_t* src = rd.take(n);
      if (nb200_host_alloc(n * 4, &pinned[b]) != NB200_OK) { fprintf(stderr, "pinned alloc failed\n"); return 4; }
      memcpy(pinned[b], src, n * 4);
      ptrs[b] = (const uint32_t*)pinned[b];
    }
    evals[t].resize(nb);
    uint8_t root[32];
    CHECK(ctx, nb200_scheme_commit_host(scheme, ptrs.data(), ncols.data(), logs.data(), nb, /*coset_order=*/1, ch, root, evals[t].data()));
    CHECK(ctx, nb200_sync(ctx));
    for (void* p : pinned) nb200_host_free(p);
  }

  // lookup elements (machine.rs:239-240): [z, alpha] = draw_felts(2) per relation, alpha powers on the host
  std::vector<Q> params(nb200_air_n_params(air), Q{{0, 0, 0, 0}});
  for (auto& r : rels) {
    uint32_t za[8];
    nb200_channel_draw_felts(ch, 2, za);
    memcpy(params[r.z].c, za, 16);
    Q alpha, cur{{1, 0, 0, 0}};
    memcpy(alpha.c, za + 4, 16);
    for (uint32_t i = 0; i < r.size; ++i) { params[r.alpha[i]] = cur; cur = qmul(cur, alpha); }
  }

  // interaction trace per component (machine.rs:242-260); claimed sums are mixed before the commit (machine.rs:262-263)
  std::vector<nb200_cols*> inter;
  std::vector<uint32_t> claimed;
  for (size_t k = 0; k < comps.size(); ++k) {
    nb200_cols* out = nullptr; uint32_t cs[4];
    CHECK(ctx, nb200_gen_interaction_trace(ctx, air, (uint32_t)k, evals[0].data(), evals[0].size(), evals[1].data(), evals[1].size(),
                                           &params[0].c[0], params.size(), &out, cs));
    inter.push_back(out);
    claimed.insert(claimed.end(), cs, cs + 4);
    const uint32_t inv_n = mpow((uint32_t)((1ull << comps[k].log_size) % P), P - 2);
    if (comps[k].shift_param != 0xffffffffu)
      for (int j = 0; j < 4; ++j) params[comps[k].shift_param].c[j] = mmul(cs[j], inv_n);
  }
  nb200_channel_mix_felts(ch, claimed.data(), comps.size());
  {
    uint8_t root[32];
    CHECK(ctx, nb200_scheme_commit(scheme, inter.data(), inter.size(), ch, root));
  }

  uint8_t* proof = nullptr; size_t len = 0;
  CHECK(ctx, nb200_prove(scheme, air, &params[0].c[0], params.size(), ch, &proof, &len));
  FILE* o = fopen(argv[2], "wb");
  if (!o || fwrite(proof, 1, len, o) != len) { perror(argv[2]); return 5; }
  fclose(o);
  fprintf(stderr, "proof: %zu bytes\n", len);
  nb200_free(proof);
  for (auto* c : inter) nb200_cols_free(ctx, c);
  for (auto& t : evals) for (auto* c : t) nb200_cols_free(ctx, c);
  nb200_scheme_free(scheme);
  nb200_air_free(air);
  nb200_channel_free(ch);
  nb200_ctx_destroy(ctx);
  return 0;
}
