// Instruction-throughput microbenchmark for the M31 butterfly's building blocks on sm_100a (B200): which pipe does each op use and at what rate?
// Each kernel runs ITERS x 8 independent chains per thread of one op pattern; result = warp-instructions per cycle per SM sub-partition (SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu && tools/ubench/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define P31 0x7fffffffu
constexpr int ITERS = 4096, NCH = 8;

template <int OP> __device__ __forceinline__ void step(u32& a, u32& b, const u32 c) {
  if (OP == 0) { u64 p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(c)); a = (u32)p ^ (u32)(p >> 32); }   // IMAD.WIDE (+ LOP3)
  if (OP == 1) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(c)); }                                       // IMAD.HI
  if (OP == 2) { asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(a), "r"(c), "r"(b)); }                             // IMAD
  if (OP == 3) { u32 s = a + c; a = min(s, s - P31); }                                                                        // IADD + VIADDMNMX
  if (OP == 4) { asm volatile("min.u32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(c)); a += 1; }                                    // IMNMX + IADD
  if (OP == 5) { a = (a >> 1) + b; }                                                                                           // LEA.HI-like
  if (OP == 6) { asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a) : "r"(a), "r"(b), "r"(c)); }                        // LOP3
  if (OP == 7) { u64 p = (u64)a * c; u32 s = ((u32)p >> 1) + (u32)(p >> 32); a = min(s, s - P31); }                           // m31_mul_dbl: WIDE + LEA.HI + VIADDMNMX
  if (OP == 8) { u64 p = (u64)b * c; u32 s = ((u32)p >> 1) + (u32)(p >> 32); u32 t = min(s, s - P31);                         // full butterfly (7 instr)
                 u32 x = a + t; u32 y = a - t; a = min(x, x - P31); b = min(y, y + P31); }
  if (OP == 9) { asm volatile("add.u32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(c)); }                                            // IADD (which pipe?)
  if (OP == 10) { u64 p; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(a), "r"(c), "l"((u64)b << 32 | a)); a = (u32)(p >> 32); }  // IMAD.WIDE with 64-bit addend
  // butterfly variants that move the two plain adds to the FMA pipe as integer multiply-adds (x * (+-1) + y); the constants sit in registers
  if (OP == 12) { u64 p = (u64)b * c; u32 s = ((u32)p >> 1) + (u32)(p >> 32); u32 t = min(s, s - P31); u32 x, y;
                  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x) : "r"(t), "r"(1u), "r"(a));
                  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(t), "r"(0xffffffffu), "r"(a));
                  a = min(x, x - P31); b = min(y, y + P31); }
  if (OP == 13) { u64 p = (u64)b * c; u32 s = ((u32)p >> 1) + (u32)(p >> 32); u32 t = min(s, s - P31); u32 y;                    // only the subtraction on the FMA pipe
                  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(t), "r"(0xffffffffu), "r"(a));
                  u32 x = a + t; a = min(x, x - P31); b = min(y, y + P31); }
  if (OP == 14) { u64 p = (u64)b * c; u32 lo = (u32)p, hi = (u32)(p >> 32), s;                                                     // ... and the fold as shift (ALU) + IMAD add instead of LEA.HI
                  u32 h = lo >> 1; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(h), "r"(1u), "r"(hi));
                  u32 t = min(s, s - P31); u32 x, y;
                  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x) : "r"(t), "r"(1u), "r"(a));
                  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(t), "r"(0xffffffffu), "r"(a));
                  a = min(x, x - P31); b = min(y, y + P31); }
  if (OP == 11) { u32 h; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(a), "r"(c)); u32 l = a * (c >> 1); asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(h), "r"(0x80000001u), "r"(l)); }  // Shoup-style product: HI + LO + MAD (all FMA pipe)
}
template <int OP> __global__ void k(u32* out, u32 c0, long long* cyc) {
  u32 a[NCH], b[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { a[i] = threadIdx.x * 2654435761u + i; b[i] = a[i] ^ 0x5bd1e995u; }
  const u32 c = c0 | 1u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) step<OP>(a[i], b[i], c);
  }
  long long t1 = clock64();
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) r ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int ops_per_step, u32* d_out, long long* d_cyc, int warps_per_smsp) {
  const int threads = 128 * warps_per_smsp;   // 4 SMSPs x warps_per_smsp warps
  k<OP><<<148, threads>>>(d_out, 12345u, d_cyc);
  cudaDeviceSynchronize();
  long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  double steps = (double)ITERS * NCH * warps_per_smsp;   // warp-steps per SMSP
  printf("%-44s warps/SMSP %d  cycles/step %.3f  (~%d instr/step -> %.2f instr/clk/SMSP)\n", name, warps_per_smsp, cyc / steps, ops_per_step, ops_per_step * steps / cyc);
}
int main() {
  u32* d_out; long long* d_cyc; cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 8);
  for (int w : {1, 4, 8}) {
    run<0>("mul.wide.u32 + lop3", 2, d_out, d_cyc, w);
    run<1>("mul.hi.u32", 1, d_out, d_cyc, w);
    run<2>("mad.lo.u32 (IMAD)", 1, d_out, d_cyc, w);
    run<9>("add.u32", 1, d_out, d_cyc, w);
    run<3>("add + min(s, s-P)  (m31_add)", 2, d_out, d_cyc, w);
    run<4>("min.u32 + add", 2, d_out, d_cyc, w);
    run<5>("(a>>1)+b  (LEA.HI)", 1, d_out, d_cyc, w);
    run<6>("lop3", 1, d_out, d_cyc, w);
    run<10>("mad.wide.u32 with 64-bit addend", 1, d_out, d_cyc, w);
    run<7>("m31_mul_dbl (WIDE+LEA.HI+VIADDMNMX)", 3, d_out, d_cyc, w);
    run<11>("Shoup product (HI+LO+MAD)", 3, d_out, d_cyc, w);
    run<8>("butterfly (7 instr)", 7, d_out, d_cyc, w);
    run<12>("butterfly, add and sub as IMAD (FMA pipe)", 7, d_out, d_cyc, w);
    run<13>("butterfly, sub as IMAD", 7, d_out, d_cyc, w);
    run<14>("butterfly, add/sub/fold-add as IMAD", 8, d_out, d_cyc, w);
    printf("\n");
  }
  return 0;
}
