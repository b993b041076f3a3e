// Blake2s (RFC 7693) for host and device: the raw compression function F used by the Merkle hasher and a
// small streaming hasher for the Fiat-Shamir channel / proof of work.
// Replaces stwo core/vcs/{blake2_hash,blake2s_refs}.rs + the SIMD `compress16` used by
// MerkleOps<Blake2sMerkleHasher> for SimdBackend (reached from /root/reference prover/src/machine.rs:228,237,263).
#pragma once
#include "m31.cuh"
#include <cstring>

namespace nb {

#if defined(__CUDA_ARCH__)
#define NB_ROTR(x, n) __funnelshift_r((x), (x), (n))
#else
#define NB_ROTR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
#endif

#define NB_B2S_IV0 0x6A09E667u
#define NB_B2S_IV1 0xBB67AE85u
#define NB_B2S_IV2 0x3C6EF372u
#define NB_B2S_IV3 0xA54FF53Au
#define NB_B2S_IV4 0x510E527Fu
#define NB_B2S_IV5 0x9B05688Cu
#define NB_B2S_IV6 0x1F83D9ABu
#define NB_B2S_IV7 0x5BE0CD19u

#define NB_G(a, b, c, d, x, y)             \
  a = a + b + (x); d = NB_ROTR(d ^ a, 16); \
  c = c + d;       b = NB_ROTR(b ^ c, 12); \
  a = a + b + (y); d = NB_ROTR(d ^ a, 8);  \
  c = c + d;       b = NB_ROTR(b ^ c, 7);

#define NB_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  NB_G(v0, v4, v8, v12, m[s0], m[s1])                                                  \
  NB_G(v1, v5, v9, v13, m[s2], m[s3])                                                  \
  NB_G(v2, v6, v10, v14, m[s4], m[s5])                                                 \
  NB_G(v3, v7, v11, v15, m[s6], m[s7])                                                 \
  NB_G(v0, v5, v10, v15, m[s8], m[s9])                                                 \
  NB_G(v1, v6, v11, v12, m[s10], m[s11])                                               \
  NB_G(v2, v7, v8, v13, m[s12], m[s13])                                                \
  NB_G(v3, v4, v9, v14, m[s14], m[s15])

// h <- F(h, m, t0, t1, f0, f1); fully unrolled, message schedule resolved at compile time.
NB_HD void b2s_compress(u32 (&h)[8], const u32 (&m)[16], u32 t0, u32 t1, u32 f0, u32 f1) {
  u32 v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  u32 v8 = NB_B2S_IV0, v9 = NB_B2S_IV1, v10 = NB_B2S_IV2, v11 = NB_B2S_IV3;
  u32 v12 = NB_B2S_IV4 ^ t0, v13 = NB_B2S_IV5 ^ t1, v14 = NB_B2S_IV6 ^ f0, v15 = NB_B2S_IV7 ^ f1;
  NB_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
  NB_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
  NB_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
  NB_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
  NB_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
  NB_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
  NB_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
  NB_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
  NB_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
  NB_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
  h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
  h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

#if defined(__CUDACC__)
// Device variant for the ALU-pipe-bound Merkle kernels (ncu: INT ALU pipe 95 % busy, FMA pipe 10 %): the 3-input adds
// a + b + m are issued as two integer multiply-adds (x * one + y) on the FMA pipe.  `one` must be the value 1 held in a
// register the compiler cannot fold (a kernel parameter); results are identical to b2s_compress.
#define NB_MAD1(x, y) ((x) * one + (y))
#define NB_GF(a, b, c, d, x, y)                          \
  a = NB_MAD1(NB_MAD1(a, b), (x)); d = NB_ROTR(d ^ a, 16); \
  c = c + d;                       b = NB_ROTR(b ^ c, 12); \
  a = NB_MAD1(NB_MAD1(a, b), (y)); d = NB_ROTR(d ^ a, 8);  \
  c = c + d;                       b = NB_ROTR(b ^ c, 7);
#define NB_ROUNDF(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  NB_GF(v0, v4, v8, v12, m[s0], m[s1])                                                  \
  NB_GF(v1, v5, v9, v13, m[s2], m[s3])                                                  \
  NB_GF(v2, v6, v10, v14, m[s4], m[s5])                                                 \
  NB_GF(v3, v7, v11, v15, m[s6], m[s7])                                                 \
  NB_GF(v0, v5, v10, v15, m[s8], m[s9])                                                 \
  NB_GF(v1, v6, v11, v12, m[s10], m[s11])                                               \
  NB_GF(v2, v7, v8, v13, m[s12], m[s13])                                                \
  NB_GF(v3, v4, v9, v14, m[s14], m[s15])
__device__ __forceinline__ void b2s_compress_fma(u32 (&h)[8], const u32 (&m)[16], u32 t0, u32 t1, u32 f0, u32 f1, const u32 one) {
  u32 v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  u32 v8 = NB_B2S_IV0, v9 = NB_B2S_IV1, v10 = NB_B2S_IV2, v11 = NB_B2S_IV3;
  u32 v12 = NB_B2S_IV4 ^ t0, v13 = NB_B2S_IV5 ^ t1, v14 = NB_B2S_IV6 ^ f0, v15 = NB_B2S_IV7 ^ f1;
  NB_ROUNDF(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
  NB_ROUNDF(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
  NB_ROUNDF(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
  NB_ROUNDF(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
  NB_ROUNDF(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
  NB_ROUNDF(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
  NB_ROUNDF(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
  NB_ROUNDF(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
  NB_ROUNDF(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
  NB_ROUNDF(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
  h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
  h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}
#endif

NB_HD void b2s_init(u32 (&h)[8]) {
  h[0] = NB_B2S_IV0 ^ 0x01010020u; h[1] = NB_B2S_IV1; h[2] = NB_B2S_IV2; h[3] = NB_B2S_IV3;
  h[4] = NB_B2S_IV4; h[5] = NB_B2S_IV5; h[6] = NB_B2S_IV6; h[7] = NB_B2S_IV7;
}

// Host-side streaming Blake2s-256 (channel, PoW check, small host hashing).
struct Blake2sHost {
  u32 h[8];
  uint8_t buf[64];
  size_t buflen = 0;
  u64 t = 0;
  Blake2sHost() { b2s_init(h); }
  void block(bool last) {
    u32 m[16];
    for (int i = 0; i < 16; ++i) m[i] = (u32)buf[4 * i] | ((u32)buf[4 * i + 1] << 8) | ((u32)buf[4 * i + 2] << 16) | ((u32)buf[4 * i + 3] << 24);
    b2s_compress(h, m, (u32)t, (u32)(t >> 32), last ? 0xFFFFFFFFu : 0u, 0u);
  }
  void update(const void* data, size_t len) {
    const uint8_t* p = (const uint8_t*)data;
    while (len) {
      if (buflen == 64) { t += 64; block(false); buflen = 0; }
      size_t take = 64 - buflen; if (take > len) take = len;
      memcpy(buf + buflen, p, take); buflen += take; p += take; len -= take;
    }
  }
  void finalize(uint8_t out[32]) {
    t += buflen;
    memset(buf + buflen, 0, 64 - buflen);
    block(true);
    for (int i = 0; i < 8; ++i) { out[4 * i] = h[i] & 0xff; out[4 * i + 1] = (h[i] >> 8) & 0xff; out[4 * i + 2] = (h[i] >> 16) & 0xff; out[4 * i + 3] = (h[i] >> 24) & 0xff; }
  }
};

}  // namespace nb
