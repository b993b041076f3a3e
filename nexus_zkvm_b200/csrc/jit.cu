// Runtime specialisation of an AIR component's constraint program: the SSA bytecode (air.h) is translated to CUDA C,
// compiled once per loaded AIR with NVRTC for sm_100a and launched instead of the bytecode interpreter
// (interp.cu).  Same arithmetic, same accumulation order => bit-identical results; the point is that the virtual
// registers become machine registers and the decode loop disappears (SURVEY.md §7.3-4 "NVRTC-specialised kernel").
// Replaces, like the interpreter, FrameworkComponent::evaluate_constraint_quotients_on_domain reached from
// stwo::prover::prove at /root/reference prover/src/machine.rs:286-290.
//
// libnvrtc is opened with dlopen at first use; if it is missing (or NB200_JIT=0) the interpreter kernels — also CUDA —
// are used.  The generated function is cut into __noinline__ chunks of a few hundred statements: NVVM's optimiser is
// super-linear in function size (27 s for one 4400-statement function vs 7 s chunked, measured).
#include "pcs.h"
#include "jit.h"
#include <dlfcn.h>
#include <unistd.h>
#include <cstring>
#include <cstdio>
#include <nvrtc.h>
#include <sstream>
#include <cstdlib>

namespace nb {

// threads per CTA the constraint kernels are compiled for (their __launch_bounds__): 1024 -> <= 64 registers per thread; NB200_JIT_BOUND=512 lets
// the compiler use up to 128 (fewer spills of the chunk state, half the warps per SM) — a tuning knob, part of the generated source and so of its cache key
static u32 jit_bound() {
  static u32 v = 0;
  if (!v) { v = JIT_BLOCK; if (const char* e = getenv("NB200_JIT_BOUND")) { int x = atoi(e); if (x == 256 || x == 512 || x == 1024) v = (u32)x; } }
  return v;
}


namespace {
struct Nvrtc {
  void* h = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  bool ok = false;
};
Nvrtc& nvrtc() {
  static Nvrtc n;
  static bool tried = false;
  if (tried) return n;
  tried = true;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
  for (const char* nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (n.h) break; }
  if (!n.h) return n;
#define NB_SYM(f) n.f = (decltype(n.f))dlsym(n.h, "nvrtc" #f); if (!n.f) return n;
  NB_SYM(CreateProgram) NB_SYM(CompileProgram) NB_SYM(GetCUBINSize) NB_SYM(GetCUBIN) NB_SYM(GetProgramLogSize) NB_SYM(GetProgramLog) NB_SYM(DestroyProgram)
#undef NB_SYM
  n.ok = true;
  return n;
}

const char* kPrelude = R"SRC(
typedef unsigned int u32; typedef unsigned long long u64;
#define P31 0x7fffffffu
struct Q { u32 c0, c1, c2, c3; };
__device__ __forceinline__ u32 mn(u32 a, u32 b) { return a < b ? a : b; }
__device__ __forceinline__ u32 add(u32 a, u32 b) { u32 s = a + b; return mn(s, s - P31); }
__device__ __forceinline__ u32 sub(u32 a, u32 b) { u32 d = a - b; return mn(d, d + P31); }
__device__ __forceinline__ u32 neg(u32 a) { return a ? P31 - a : 0u; }
// one product: a * 2b = hi * 2^32 + lo  =>  a * b = hi * 2^31 + lo / 2 == hi + lo / 2 (mod P), below 2P
__device__ __forceinline__ u32 mul2(u32 a, u32 b2) { u64 p = (u64)a * (u64)b2; u32 s = (u32)(p >> 32) + ((u32)p >> 1); return mn(s, s - P31); }
__device__ __forceinline__ u32 mul(u32 a, u32 b) { return mul2(a, b + b); }
// canonical residue of any 64-bit value: 2^32 == 2 and 2^31 == 1 (mod P)
__device__ __forceinline__ u32 red64(u64 x) {
  u64 y;                                                // y = 2 * hi + lo < 3 * 2^32 (mad.wide keeps NVVM from expanding it into a carry chain)
  asm("mad.wide.u32 %0, %1, 2, %2;" : "=l"(y) : "r"((u32)(x >> 32)), "l"((u64)(u32)x));
  u32 yl = (u32)y, yh = (u32)(y >> 32);
  u32 s = (yl & P31) + __funnelshift_r(yl, yh, 31);     // <= P + 5
  return mn(s, s - P31);
}
__device__ __forceinline__ Q qadd(Q x, Q y) { return Q{add(x.c0, y.c0), add(x.c1, y.c1), add(x.c2, y.c2), add(x.c3, y.c3)}; }
__device__ __forceinline__ Q qsub(Q x, Q y) { return Q{sub(x.c0, y.c0), sub(x.c1, y.c1), sub(x.c2, y.c2), sub(x.c3, y.c3)}; }
__device__ __forceinline__ Q qneg(Q x) { return Q{neg(x.c0), neg(x.c1), neg(x.c2), neg(x.c3)}; }
__device__ __forceinline__ Q qmulb(Q x, u32 b) { u32 b2 = b + b; return Q{mul2(x.c0, b2), mul2(x.c1, b2), mul2(x.c2, b2), mul2(x.c3, b2)}; }
__device__ __forceinline__ Q qaddb(Q x, u32 b) { x.c0 = add(x.c0, b); return x; }
__device__ __forceinline__ Q qsubb(Q x, u32 b) { x.c0 = sub(x.c0, b); return x; }
// acc + x * y in QM31 with every coordinate accumulated in 64 bits and reduced once.  (a + bu)(c + du) = (ac + R bd) + (ad + bc)u,
// R = 2 + i (the formula of m31.cuh qm31_mul), regrouped per coordinate of x so that each output is four products:
//   r0 = x0 y0 - x1 y1 + x2 (2 y2 - y3) - x3 (y2 + 2 y3)      r1 = x0 y1 + x1 y0 + x2 (y2 + 2 y3) + x3 (2 y2 - y3)
//   r2 = x0 y2 - x1 y3 + x2 y0 - x3 y1                        r3 = x0 y3 + x1 y2 + x2 y1 + x3 y0
// ny1 = P - y1, ny3 = P - y3, g = 2 y2 - y3, gp = y2 + 2 y3, h = P - gp: all <= P, so 4 products + acc < 2^64.
__device__ __forceinline__ Q qmac(Q acc, Q x, u32 y0, u32 y1, u32 y2, u32 y3, u32 ny1, u32 ny3, u32 g, u32 gp, u32 h) {
  u64 r0 = (u64)acc.c0 + (u64)x.c0 * y0 + (u64)x.c1 * ny1 + (u64)x.c2 * g + (u64)x.c3 * h;
  u64 r1 = (u64)acc.c1 + (u64)x.c0 * y1 + (u64)x.c1 * y0 + (u64)x.c2 * gp + (u64)x.c3 * g;
  u64 r2 = (u64)acc.c2 + (u64)x.c0 * y2 + (u64)x.c1 * ny3 + (u64)x.c2 * y0 + (u64)x.c3 * ny1;
  u64 r3 = (u64)acc.c3 + (u64)x.c0 * y3 + (u64)x.c1 * y2 + (u64)x.c2 * y1 + (u64)x.c3 * y0;
  return Q{red64(r0), red64(r1), red64(r2), red64(r3)};
}
__device__ __forceinline__ Q qmul(Q x, Q y) {
  u32 gp = add(add(y.c3, y.c3), y.c2), g = sub(add(y.c2, y.c2), y.c3);
  return qmac(Q{0u, 0u, 0u, 0u}, x, y.c0, y.c1, y.c2, y.c3, P31 - y.c1, P31 - y.c3, g, gp, P31 - gp);
}
// x^(P-2): 30 squarings + 8 products
__device__ __forceinline__ u32 sqn(u32 x, int n) { for (int i = 0; i < n; ++i) x = mul(x, x); return x; }
__device__ __forceinline__ u32 minv(u32 x) {
  u32 a2 = mul(sqn(x, 1), x), a4 = mul(sqn(a2, 2), a2), a8 = mul(sqn(a4, 4), a4), a16 = mul(sqn(a8, 8), a8);
  u32 a24 = mul(sqn(a16, 8), a8), a28 = mul(sqn(a24, 4), a4), a29 = mul(sqn(a28, 1), x);
  return mul(sqn(a29, 2), x);
}
// (a + bu)^-1 = (a - bu) / (a^2 - (2 + i) b^2), a, b in CM31 (the formula of m31.cuh qm31_inv); the CM31 inverse is conj / norm
__device__ __noinline__ Q qinv(Q q) {
  u32 b2r = sub(mul(q.c2, q.c2), mul(q.c3, q.c3)), b2i = mul(add(q.c2, q.c2), q.c3);           // b^2
  u32 a2r = sub(mul(q.c0, q.c0), mul(q.c1, q.c1)), a2i = mul(add(q.c0, q.c0), q.c1);           // a^2
  u32 dr = sub(a2r, sub(add(b2r, b2r), b2i)), di = sub(a2i, add(add(b2i, b2i), b2r));          // a^2 - (2 b^2 + i b^2)
  u32 ni = minv(add(mul(dr, dr), mul(di, di)));
  u32 ir = mul(dr, ni), ii = neg(mul(di, ni));                                                 // 1 / denom
  return Q{sub(mul(q.c0, ir), mul(q.c1, ii)), add(mul(q.c0, ii), mul(q.c1, ir)),
           neg(sub(mul(q.c2, ir), mul(q.c3, ii))), neg(add(mul(q.c2, ii), mul(q.c3, ir)))};
}
__device__ __forceinline__ Q ldq(const u32* p) { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); return Q{v.x, v.y, v.z, v.w}; }
// rr + coeff * x with the coefficient's derived multipliers precomputed by the host (12 words per constraint, see jit_coeff_table)
__device__ __forceinline__ Q qmac_tab(Q acc, Q x, const u32* t) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(t)), b = __ldg(reinterpret_cast<const uint4*>(t) + 1), c = __ldg(reinterpret_cast<const uint4*>(t) + 2);
  return qmac(acc, x, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x);
}
)SRC";

// the arithmetic / load opcodes shared by the constraint and the logup programs
template <class Ld>
void emit_op(std::ostringstream& o, const AirInstr& in, Ld ld) {
  switch (in.op) {
    case OP_LOADM: o << "b[" << in.dst << "] = " << ld(in.a) << ";"; break;
    case OP_CONSTB: o << "b[" << in.dst << "] = " << in.a << "u;"; break;
    case OP_ADDB: o << "b[" << in.dst << "] = add(b[" << in.a << "], b[" << in.b << "]);"; break;
    case OP_SUBB: o << "b[" << in.dst << "] = sub(b[" << in.a << "], b[" << in.b << "]);"; break;
    case OP_MULB: o << "b[" << in.dst << "] = mul(b[" << in.a << "], b[" << in.b << "]);"; break;
    case OP_NEGB: o << "b[" << in.dst << "] = neg(b[" << in.a << "]);"; break;
    case OP_PARAME: o << "e[" << in.dst << "] = ldq(params + " << 4 * in.a << ");"; break;
    case OP_ADDE: o << "e[" << in.dst << "] = qadd(e[" << in.a << "], e[" << in.b << "]);"; break;
    case OP_SUBE: o << "e[" << in.dst << "] = qsub(e[" << in.a << "], e[" << in.b << "]);"; break;
    case OP_MULE: o << "e[" << in.dst << "] = qmul(e[" << in.a << "], e[" << in.b << "]);"; break;
    case OP_NEGE: o << "e[" << in.dst << "] = qneg(e[" << in.a << "]);"; break;
    case OP_ADDEB: o << "e[" << in.dst << "] = qaddb(e[" << in.a << "], b[" << in.b << "]);"; break;
    case OP_SUBEB: o << "e[" << in.dst << "] = qsubb(e[" << in.a << "], b[" << in.b << "]);"; break;
    case OP_MULEB: o << "e[" << in.dst << "] = qmulb(e[" << in.a << "], b[" << in.b << "]);"; break;
    case OP_BTOE: o << "e[" << in.dst << "] = Q{b[" << in.a << "], 0u, 0u, 0u};"; break;
    case OP_LOADME: o << "e[" << in.dst << "] = Q{" << ld(in.a) << ", " << ld(in.a + 1) << ", " << ld(in.a + 2) << ", " << ld(in.a + 3) << "};"; break;
    default: break;
  }
}

// Which virtual registers cross a chunk boundary?  live_in[ci] = registers read in chunk ci or later before they are rewritten.  A chunk loads only
// those from the state struct and stores only the ones it wrote that a later chunk still reads: the real v1 AIR keeps 26 base + 15 secure
// registers alive somewhere (shared selectors such as IsTypeR), and copying all 90 words in and out of local memory at each of its 19 chunk
// boundaries cost more than the arithmetic (the synthetic ADD machine has 4 + 4 and never showed it).
struct ChunkLive { std::vector<std::vector<u32>> in_b, in_e, out_b, out_e; };
static void instr_regs(const AirInstr& in, std::vector<u32>& rb, std::vector<u32>& re, int& wb, int& we) {
  wb = we = -1;
  switch (in.op) {
    case OP_LOADM: case OP_CONSTB: wb = (int)in.dst; break;
    case OP_ADDB: case OP_SUBB: case OP_MULB: rb = {in.a, in.b}; wb = (int)in.dst; break;
    case OP_NEGB: rb = {in.a}; wb = (int)in.dst; break;
    case OP_PARAME: case OP_LOADME: we = (int)in.dst; break;
    case OP_ADDE: case OP_SUBE: case OP_MULE: re = {in.a, in.b}; we = (int)in.dst; break;
    case OP_NEGE: re = {in.a}; we = (int)in.dst; break;
    case OP_ADDEB: case OP_SUBEB: case OP_MULEB: re = {in.a}; rb = {in.b}; we = (int)in.dst; break;
    case OP_BTOE: rb = {in.a}; we = (int)in.dst; break;
    case OP_CONSTRB: rb = {in.a}; break;
    case OP_CONSTRE: re = {in.a}; break;
    case OP_FRAC: re = {in.a, in.b}; break;
    default: break;
  }
}
static ChunkLive chunk_liveness(const std::vector<AirInstr>& prog, size_t CH, u32 nb, u32 ne) {
  const size_t n_chunks = (prog.size() + CH - 1) / CH;
  ChunkLive L; L.in_b.resize(n_chunks); L.in_e.resize(n_chunks); L.out_b.resize(n_chunks); L.out_e.resize(n_chunks);
  std::vector<char> lb(nb + 1, 0), le(ne + 1, 0);
  std::vector<std::vector<char>> after_b(n_chunks), after_e(n_chunks);   // live sets at the END of each chunk
  for (size_t ci = n_chunks; ci-- > 0;) {
    after_b[ci] = lb; after_e[ci] = le;
    for (size_t pc = std::min(prog.size(), (ci + 1) * CH); pc-- > ci * CH;) {
      std::vector<u32> rb, re; int wb, we;
      instr_regs(prog[pc], rb, re, wb, we);
      if (wb >= 0) lb[wb] = 0;
      if (we >= 0) le[we] = 0;
      for (u32 r : rb) lb[r] = 1;
      for (u32 r : re) le[r] = 1;
    }
    for (u32 r = 0; r < nb; ++r) if (lb[r]) L.in_b[ci].push_back(r);
    for (u32 r = 0; r < ne; ++r) if (le[r]) L.in_e[ci].push_back(r);
  }
  for (size_t ci = 0; ci < n_chunks; ++ci) {
    std::vector<char> wrb(nb + 1, 0), wre(ne + 1, 0);
    for (size_t pc = ci * CH; pc < std::min(prog.size(), (ci + 1) * CH); ++pc) {
      std::vector<u32> rb, re; int wb, we;
      instr_regs(prog[pc], rb, re, wb, we);
      if (wb >= 0) wrb[wb] = 1;
      if (we >= 0) wre[we] = 1;
    }
    for (u32 r = 0; r < nb; ++r) if (wrb[r] && after_b[ci][r]) L.out_b[ci].push_back(r);
    for (u32 r = 0; r < ne; ++r) if (wre[r] && after_e[ci][r]) L.out_e[ci].push_back(r);
  }
  return L;
}

std::string gen_source(const AirComponent& c) {
  std::ostringstream o;
  o << kPrelude;
  const u32 DL = c.log_size;  // EL (the canonic domain the rows belong to) is a kernel argument: whole domains and half domains share the kernel
  // offset_bit_reversed_circle_domain_index with the domain sizes baked in
  o << "__device__ __forceinline__ u32 offrow(u32 i, int off, u32 EL) { const u32 DL = " << DL << ";\n"
    << "  u32 prev = __brev(i) >> (32 - EL); u32 half = 1u << (EL - 1); long long step = (long long)off * (1ll << (EL - DL - 1)); long long v;\n"
    << "  if (prev < half) { v = ((long long)prev + step) % (long long)half; if (v < 0) v += half; }\n"
    << "  else { v = ((long long)prev - step) % (long long)half; if (v < 0) v += half; v += half; }\n"
    << "  return __brev((u32)v) >> (32 - EL); }\n";
  const u32 nb = c.n_base_regs ? c.n_base_regs : 1, ne = c.n_ext_regs ? c.n_ext_regs : 1;
  o << "struct St { u32 b[" << nb << "]; Q e[" << ne << "]; Q rr; };\n";
  // column base pointers live in constant memory (filled before every launch): an access costs no pointer load from global memory
  // (the real AIR reads 2765 mask values per row: one dependent global load less per value)
  o << "#define NB_NMASKS " << c.masks.size() << "\n__constant__ const u32* ccols[NB_NMASKS > 0 ? NB_NMASKS : 1];\n";
  auto ld = [&](u32 m) {
    std::ostringstream s;
    if (c.masks[m].off == 0) s << "__ldg(ccols[" << m << "] + row)";
    else s << "__ldg(ccols[" << m << "] + offrow(row, " << c.masks[m].off << ", EL))";
    return s.str();
  };
  size_t CH = 250;
  if (const char* e = getenv("NB200_JIT_CHUNK")) { long v = atol(e); if (v >= 16 && v <= 100000) CH = (size_t)v; }
  size_t n_chunks = (c.prog.size() + CH - 1) / CH;
  const ChunkLive live = chunk_liveness(c.prog, CH, nb, ne);
  u32 k = 0;
  for (size_t ci = 0; ci < n_chunks; ++ci) {
    o << "__device__ __noinline__ void chunk" << ci << "(St& s, const u32* const* __restrict__ cols, const u32* __restrict__ params, const u32* __restrict__ coeff, u32 row, u32 EL) {\n";
    o << "  u32 b[" << nb << "]; Q e[" << ne << "]; Q rr = s.rr;\n";
    for (u32 r : live.in_b[ci]) o << "  b[" << r << "] = s.b[" << r << "];";
    for (u32 r : live.in_e[ci]) o << "  e[" << r << "] = s.e[" << r << "];";
    o << "\n";
    for (size_t pc = ci * CH; pc < std::min(c.prog.size(), (ci + 1) * CH); ++pc) {
      const AirInstr& in = c.prog[pc];
      o << "  ";
      switch (in.op) {
        case OP_CONSTRB: o << "rr = qadd(rr, qmulb(ldq(coeff + " << JIT_COEFF_WORDS * k << "), b[" << in.a << "]));"; ++k; break;
        case OP_CONSTRE: o << "rr = qmac_tab(rr, e[" << in.a << "], coeff + " << JIT_COEFF_WORDS * k << ");"; ++k; break;
        default: emit_op(o, in, ld); break;
      }
      o << "\n";
    }
    for (u32 r : live.out_b[ci]) o << "  s.b[" << r << "] = b[" << r << "];";
    for (u32 r : live.out_e[ci]) o << "  s.e[" << r << "] = e[" << r << "];";
    o << "\n  s.rr = rr;\n}\n";
  }
  // The CTAs re-converge (__syncthreads) after every chunk: the warps of a CTA then execute the same few tens of KB of straight-line
  // code at a time and the instruction cache serves them from one fetch.  Without the barriers the warps drift apart over the ~0.7 MB
  // program and the kernel is instruction-fetch bound (ncu: stall_no_instruction 11 per issue, icc hit rate 53 %).
  o << "extern \"C\" __global__ void __launch_bounds__(" << jit_bound() << ", 1) nbjit(const u32* const* __restrict__ cols, const u32* __restrict__ params, const u32* __restrict__ coeff,\n"
    << "    const u32* __restrict__ dinv, u32* __restrict__ a0, u32* __restrict__ a1, u32* __restrict__ a2, u32* __restrict__ a3, u32 EL, u32 row0) {\n"
    << "  const u32 row = row0 + blockIdx.x * blockDim.x + threadIdx.x;   // row0: a rank of a multi-GPU proof evaluates its slice of the domain's rows\n  St s;\n"
    << "  for (int i = 0; i < " << nb << "; ++i) s.b[i] = 0u;\n  for (int i = 0; i < " << ne << "; ++i) s.e[i] = Q{0u, 0u, 0u, 0u};\n  s.rr = Q{0u, 0u, 0u, 0u};\n";
  for (size_t ci = 0; ci < n_chunks; ++ci) o << "  chunk" << ci << "(s, cols, params, coeff, row, EL);\n  __syncthreads();\n";
  o << "  const u32 di = __ldg(dinv + (row >> " << DL << "));\n"
    << "  a0[row] = add(a0[row], mul(s.rr.c0, di)); a1[row] = add(a1[row], mul(s.rr.c1, di));\n"
    << "  a2[row] = add(a2[row], mul(s.rr.c2, di)); a3[row] = add(a3[row], mul(s.rr.c3, di));\n}\n";
  return o.str();
}

// LogupTraceGenerator for one component (interp.cu interp_kernel<.., LOGUP> is the bytecode version): run the logup program, combine the
// fractions of each batch (write_frac), add to the running row sum and store the 4 coordinate columns of the batch (finalize_col).
// The prefix sum over rows of the last column stays in logup_generate.
std::string gen_logup_source(const AirComponent& c) {
  std::ostringstream o;
  o << kPrelude;
  const u32 nb = c.lg_base_regs ? c.lg_base_regs : 1, ne = c.lg_ext_regs ? c.lg_ext_regs : 1;
  o << "struct St { u32 b[" << nb << "]; Q e[" << ne << "]; Q fn, fd, run; };\n";
  o << "#define NB_NMASKS " << c.masks.size() << "\n__constant__ const u32* ccols[NB_NMASKS > 0 ? NB_NMASKS : 1];\n";
  auto ld = [&](u32 m) {
    std::ostringstream s;
    // next-row masks and interaction-trace masks are not inputs of the logup program (they read as zero, as in the interpreter)
    if (m < c.masks.size() && c.masks[m].off == 0 && c.masks[m].tree != 2) s << "__ldg(ccols[" << m << "] + row)";
    else s << "0u";
    return s.str();
  };
  size_t CH = 150;
  size_t n_chunks = (c.logup_prog.size() + CH - 1) / CH;
  const ChunkLive live = chunk_liveness(c.logup_prog, CH, nb, ne);
  // The QM31 inverse (one per batch of fractions; 38 M31 products for x^(P-2) alone) is batched over G consecutive batches of the same
  // row: prefix products, ONE inverse, back-substitution (3 products per extra batch).  The inverse is unique, so the values are the
  // interpreter's.  `pending` batches wait in pfn/pfd until the group is full; the running row sum is then advanced batch by batch.
  const int G = 4;
  u32 k = 0; bool have = false; u32 cur_batch = 0;
  std::vector<u32> pending;   // batch ids waiting in slots 0..pending.size()-1 (tracked at code-generation time)
  auto flush = [&](std::ostringstream& o2) {
    const int n = (int)pending.size();
    if (n == 0) return;
    o2 << "  {\n";
    for (int g = 1; g < n; ++g) o2 << "    Q pp" << g << " = qmul(" << (g == 1 ? std::string("pfd0") : "pp" + std::to_string(g - 1)) << ", pfd" << g << ");\n";
    o2 << "    Q iv = qinv(" << (n == 1 ? std::string("pfd0") : "pp" + std::to_string(n - 1)) << ");\n";
    for (int g = n - 1; g >= 1; --g)
      o2 << "    { Q t = qmul(iv, " << (g == 1 ? std::string("pfd0") : "pp" + std::to_string(g - 1)) << "); iv = qmul(iv, pfd" << g << "); pfd" << g << " = t; }\n";
    o2 << "    pfd0 = iv;\n";
    for (int g = 0; g < n; ++g) {
      o2 << "    run = qadd(run, qmul(pfn" << g << ", pfd" << g << "));\n";
      for (int cc = 0; cc < 4; ++cc) o2 << "    out[(" << (4 * (size_t)pending[g] + cc) << "ull << LS) + row] = run.c" << cc << ";\n";
    }
    o2 << "  }\n";
    pending.clear();
  };
  auto finalize = [&](std::ostringstream& o2) {   // the current batch's combined fraction fn / fd is complete
    o2 << "  pfn" << pending.size() << " = fn; pfd" << pending.size() << " = fd;\n";
    pending.push_back(cur_batch);
    if ((int)pending.size() == G) flush(o2);
  };
  o << "struct Pend { Q n[" << G << "], d[" << G << "]; };\n";
  for (size_t ci = 0; ci < n_chunks; ++ci) {
    o << "__device__ __noinline__ void chunk" << ci << "(St& s, Pend& pe, const u32* const* __restrict__ cols, const u32* __restrict__ params, u32* __restrict__ out, u32 row, u32 LS) {\n";
    o << "  u32 b[" << nb << "]; Q e[" << ne << "]; Q fn = s.fn, fd = s.fd, run = s.run;\n";
    for (int g = 0; g < G; ++g) o << "  Q pfn" << g << " = pe.n[" << g << "], pfd" << g << " = pe.d[" << g << "];\n";
    for (u32 r : live.in_b[ci]) o << "  b[" << r << "] = s.b[" << r << "];";
    for (u32 r : live.in_e[ci]) o << "  e[" << r << "] = s.e[" << r << "];";
    o << "\n";
    for (size_t pc = ci * CH; pc < std::min(c.logup_prog.size(), (ci + 1) * CH); ++pc) {
      const AirInstr& in = c.logup_prog[pc];
      if (in.op == OP_FRAC) {
        u32 bt = c.batching[k];
        if (have && bt != cur_batch) { finalize(o); have = false; }
        if (!have) o << "  fn = e[" << in.a << "]; fd = e[" << in.b << "];\n";                      // first fraction of the batch: 0/1 + n/d
        else o << "  fn = qadd(qmul(fn, e[" << in.b << "]), qmul(e[" << in.a << "], fd)); fd = qmul(fd, e[" << in.b << "]);\n";
        cur_batch = bt; have = true; ++k;
      } else {
        o << "  "; emit_op(o, in, ld); o << "\n";
      }
    }
    if (ci + 1 == n_chunks) { if (have) finalize(o); flush(o); }
    for (u32 r : live.out_b[ci]) o << "  s.b[" << r << "] = b[" << r << "];";
    for (u32 r : live.out_e[ci]) o << "  s.e[" << r << "] = e[" << r << "];";
    o << "\n  s.fn = fn; s.fd = fd; s.run = run;\n";
    for (int g = 0; g < G; ++g) o << "  pe.n[" << g << "] = pfn" << g << "; pe.d[" << g << "] = pfd" << g << ";\n";
    o << "}\n";
  }
  o << "extern \"C\" __global__ void __launch_bounds__(" << JIT_BLOCK << ", 1) nbjit(const u32* const* __restrict__ cols, const u32* __restrict__ params, u32* __restrict__ out, u32 LS) {\n"
    << "  const u32 row = blockIdx.x * blockDim.x + threadIdx.x;\n  St s;\n"
    << "  for (int i = 0; i < " << nb << "; ++i) s.b[i] = 0u;\n  for (int i = 0; i < " << ne << "; ++i) s.e[i] = Q{0u, 0u, 0u, 0u};\n"
    << "  s.fn = Q{0u, 0u, 0u, 0u}; s.fd = Q{1u, 0u, 0u, 0u}; s.run = Q{0u, 0u, 0u, 0u};\n"
    << "  Pend pe; for (int i = 0; i < " << 4 << "; ++i) { pe.n[i] = Q{0u, 0u, 0u, 0u}; pe.d[i] = Q{1u, 0u, 0u, 0u}; }\n";
  for (size_t ci = 0; ci < n_chunks; ++ci) o << "  chunk" << ci << "(s, pe, cols, params, out, row, LS);\n  __syncthreads();\n";
  o << "}\n";
  return o.str();
}
}  // namespace

std::string jit_source(const AirComponent& c) { return gen_source(c); }
std::string jit_logup_source(const AirComponent& c) { return gen_logup_source(c); }

// threads per CTA at launch: the code is compiled for up to JIT_BLOCK threads at <= 64 registers; 512 runs as two CTAs per SM, which
// measured slightly faster than one CTA of 1024 (10.7 vs 11.3 ms: the two CTAs sit in different phases of the program and share the pipes better)
static u32 jit_block() {
  static u32 b = 0;
  if (!b) { b = 512; if (const char* e = getenv("NB200_JIT_BLOCK")) { int v = atoi(e); if (v == 256 || v == 512 || v == 1024) b = (u32)v; } }
  return b < jit_bound() ? b : jit_bound();
}

bool jit_enabled() {
  const char* e = getenv("NB200_JIT");
  return !(e && e[0] == '0');   // libnvrtc itself is only opened when a kernel is missing from the cubin cache
}

void jit_release(JitKernel& jk) {
  if (jk.lib) cudaLibraryUnload((cudaLibrary_t)jk.lib);
  jk = JitKernel();
}

static nb200_status jit_compile_source(nb200_ctx* ctx, const AirComponent& c, const std::string& src, JitKernel* out);
nb200_status jit_compile_constraints(nb200_ctx* ctx, const AirComponent& c, JitKernel* out) { return jit_compile_source(ctx, c, gen_source(c), out); }
nb200_status jit_compile_logup(nb200_ctx* ctx, const AirComponent& c, JitKernel* out) { return jit_compile_source(ctx, c, gen_logup_source(c), out); }
// ---- cubin cache: <directory of this library>/jit_cache/<key>.cubin (or $NB200_JIT_CACHE).  `python -m nexus_zkvm_b200.build`
// fills it for the shipped machines with nvcc, so a fresh box neither loads libnvrtc nor compiles; kernels compiled at run time are
// added when the directory is writable.  The key covers the generated source and the target.
uint64_t jit_source_key(const std::string& src) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const char* p, size_t n) { for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; } };
  mix(src.data(), src.size());
  const char* tgt = "|sm_100a|nb200-jit-4";
  mix(tgt, strlen(tgt));
  return h;
}
static std::string jit_cache_dir() {
  if (const char* e = getenv("NB200_JIT_CACHE")) return e;
  Dl_info info;
  if (dladdr((const void*)&jit_source_key, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t k = p.rfind('/');
    return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/jit_cache";
  }
  return "";
}
static std::string jit_cache_path(const std::string& src) {
  std::string d = jit_cache_dir();
  if (d.empty()) return "";
  char name[32]; snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)jit_source_key(src));
  return d + name;
}
static bool read_file(const std::string& path, std::vector<char>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  bool ok = n > 0;
  if (ok) { out.resize((size_t)n); ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n; }
  fclose(f);
  return ok;
}
static void write_file_atomic(const std::string& path, const std::vector<char>& data) {
  std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}
static nb200_status jit_load_cubin(nb200_ctx* ctx, const AirComponent& c, const std::vector<char>& cubin, JitKernel* out) {
  cudaLibrary_t lib;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) return set_err(ctx, NB200_ERR_CUDA, std::string("jit: cudaLibraryLoadData: ") + cudaGetErrorString(e));
  cudaKernel_t k;
  e = cudaLibraryGetKernel(&k, lib, "nbjit");
  if (e != cudaSuccess) { cudaLibraryUnload(lib); return set_err(ctx, NB200_ERR_CUDA, std::string("jit: cudaLibraryGetKernel: ") + cudaGetErrorString(e)); }
  out->lib = (void*)lib; out->kernel = (void*)k; out->log_size = c.log_size; out->eval_log = c.eval_log();
  return NB200_OK;
}

static nb200_status jit_compile_source(nb200_ctx* ctx, const AirComponent& c, const std::string& src, JitKernel* out) {
  out->lib = nullptr; out->kernel = nullptr; out->tried = true;
  const std::string cache = jit_cache_path(src);
  if (!cache.empty()) {
    std::vector<char> cubin;
    if (read_file(cache, cubin) && jit_load_cubin(ctx, c, cubin, out) == NB200_OK) return NB200_OK;
    cudaGetLastError();  // a stale or foreign file: fall through to the compiler
  }
  Nvrtc& n = nvrtc();
  if (!n.ok) return set_err(ctx, NB200_ERR_STATE, "jit: libnvrtc not available");
  nvrtcProgram prog;
  if (n.CreateProgram(&prog, src.c_str(), "nb200_air.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return set_err(ctx, NB200_ERR_STATE, "jit: nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  nvrtcResult r = n.CompileProgram(prog, 3, opts);
  if (r != NVRTC_SUCCESS) {
    size_t ls = 0; n.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0'); if (ls) n.GetProgramLog(prog, &log[0]);
    n.DestroyProgram(&prog);
    return set_err(ctx, NB200_ERR_STATE, "jit: compile failed: " + log.substr(0, 2000));
  }
  size_t cs = 0; n.GetCUBINSize(prog, &cs);
  std::vector<char> cubin(cs);
  n.GetCUBIN(prog, cubin.data());
  n.DestroyProgram(&prog);
  if (!cache.empty()) write_file_atomic(cache, cubin);
  return jit_load_cubin(ctx, c, cubin, out);
}

// per constraint: y0 y1 y2 y3 | P-y1 P-y3 2y2-y3 y2+2y3 | P-(y2+2y3) 0 0 0   (the multipliers of qmac in the generated code)
void jit_coeff_table(const std::vector<qm31>& coeffs, std::vector<u32>& out) {
  out.assign(coeffs.size() * JIT_COEFF_WORDS, 0u);
  for (size_t k = 0; k < coeffs.size(); ++k) {
    const u32* y = coeffs[k].c;
    u32* t = &out[k * JIT_COEFF_WORDS];
    u32 gp = m31_add(m31_add(y[3], y[3]), y[2]), g = m31_sub(m31_add(y[2], y[2]), y[3]);
    t[0] = y[0]; t[1] = y[1]; t[2] = y[2]; t[3] = y[3];
    t[4] = P31 - y[1]; t[5] = P31 - y[3]; t[6] = g; t[7] = gp; t[8] = P31 - gp;
  }
}

// the kernel's column pointers: device array -> the module's __constant__ table, in stream order
static nb200_status jit_set_cols(nb200_ctx* ctx, const JitKernel& jk, const u32* const* d_cols) {
  void* dptr = nullptr; size_t bytes = 0;
  cudaError_t e = cudaLibraryGetGlobal(&dptr, &bytes, (cudaLibrary_t)jk.lib, "ccols");
  if (e != cudaSuccess) return set_err(ctx, NB200_ERR_CUDA, std::string("jit: cudaLibraryGetGlobal(ccols): ") + cudaGetErrorString(e));
  e = cudaMemcpyAsync(dptr, d_cols, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) return set_err(ctx, NB200_ERR_CUDA, std::string("jit: ccols copy: ") + cudaGetErrorString(e));
  return NB200_OK;
}

nb200_status jit_launch_logup(nb200_ctx* ctx, const JitKernel& jk, const u32* const* d_cols, const u32* d_params, u32* d_out, u32 log_size) {
  // log_size = log2 of the rows of THIS launch (the whole trace domain, or a rank's slice of it: the kernel is row-local)
  size_t rows = (size_t)1 << log_size;
  if (rows < JIT_BLOCK) return set_err(ctx, NB200_ERR_STATE, "jit: domain too small");
  NB_TRY(jit_set_cols(ctx, jk, d_cols));
  void* args[] = {(void*)&d_cols, (void*)&d_params, (void*)&d_out, (void*)&log_size};
  cudaError_t e = cudaLaunchKernel((const void*)jk.kernel, dim3((u32)(rows / jit_block())), dim3(jit_block()), args, 0, ctx->stream);
  ctx->launches += 1;
  if (e != cudaSuccess) return set_err(ctx, NB200_ERR_CUDA, std::string("jit launch: ") + cudaGetErrorString(e));
  return NB200_OK;
}

nb200_status jit_launch_constraints(nb200_ctx* ctx, const JitKernel& jk, const u32* const* d_cols, const u32* d_params, const u32* d_coeff, const u32* d_dinv, u32* const acc[4],
                                    u32 rows_log, u32 dom_log, u32 row0, size_t n_rows) {
  size_t rows = n_rows ? n_rows : (size_t)1 << rows_log;
  if (rows < JIT_BLOCK || rows % JIT_BLOCK != 0) return set_err(ctx, NB200_ERR_STATE, "jit: row range too small");
  NB_TRY(jit_set_cols(ctx, jk, d_cols));
  u32 el = dom_log;
  u32* a0 = acc[0]; u32* a1 = acc[1]; u32* a2 = acc[2]; u32* a3 = acc[3];
  void* args[] = {(void*)&d_cols, (void*)&d_params, (void*)&d_coeff, (void*)&d_dinv, (void*)&a0, (void*)&a1, (void*)&a2, (void*)&a3, (void*)&el, (void*)&row0};
  cudaError_t e = cudaLaunchKernel((const void*)jk.kernel, dim3((u32)(rows / jit_block())), dim3(jit_block()), args, 0, ctx->stream);
  ctx->launches += 1;
  if (e != cudaSuccess) return set_err(ctx, NB200_ERR_CUDA, std::string("jit launch: ") + cudaGetErrorString(e));
  return NB200_OK;
}

}  // namespace nb
