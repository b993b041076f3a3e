// NVRTC specialisation of AIR constraint programs (see jit.cu).
#pragma once
#include "common.cuh"
#include "air.h"
#include <string>
#include <vector>

namespace nb {

static const u32 JIT_BLOCK = 1024;       // the generated kernels are compiled for up to this many threads per CTA (<= 64 registers)
static const u32 JIT_MIN_INSTR = 64;      // shorter programs stay on the bytecode interpreter
static const u32 JIT_COEFF_WORDS = 12;   // words per constraint in the coefficient table the generated kernel reads

struct JitKernel {
  void* lib = nullptr;     // cudaLibrary_t
  void* kernel = nullptr;  // cudaKernel_t
  u32 log_size = 0, eval_log = 0;
  bool tried = false;      // compilation attempted (failed attempts fall back to the interpreter)
};

bool jit_enabled();
uint64_t jit_source_key(const std::string& src);   // name of the kernel's file in the cubin cache
std::string jit_source(const AirComponent& c);  // the CUDA C the component is specialised to (inspection / offline ptxas checks)
nb200_status jit_compile_constraints(nb200_ctx* ctx, const AirComponent& c, JitKernel* out);
nb200_status jit_compile_logup(nb200_ctx* ctx, const AirComponent& c, JitKernel* out);
nb200_status jit_launch_logup(nb200_ctx* ctx, const JitKernel& jk, const u32* const* d_cols, const u32* d_params, u32* d_out, u32 log_size);
std::string jit_logup_source(const AirComponent& c);
nb200_status jit_launch_constraints(nb200_ctx* ctx, const JitKernel& jk, const u32* const* d_cols, const u32* d_params, const u32* d_coeff,
                                    const u32* d_dinv, u32* const acc[4], u32 rows_log, u32 dom_log, u32 row0 = 0, size_t n_rows = 0);
void jit_release(JitKernel& jk);
void jit_coeff_table(const std::vector<qm31>& coeffs, std::vector<u32>& out);

}  // namespace nb
