// ORACLE (test infrastructure, never shipped): BLAKE2b (RFC 7693) with personalisation, streaming, clonable.
// halo2's transcript uses blake2b_simd (EXT dependency of halo2_proofs); checked against Python hashlib in tests.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {
struct Blake2b {
  uint64_t h[8]; uint64_t t; uint8_t buf[128]; size_t buflen; size_t outlen;
  static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  static const uint64_t* IV() {
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    return iv;
  }
  explicit Blake2b(size_t out = 64, const char* personal16 = nullptr) : t(0), buflen(0), outlen(out) {
    for (int i = 0; i < 8; ++i) h[i] = IV()[i];
    h[0] ^= 0x01010000ULL ^ (uint64_t)out;
    if (personal16) { uint64_t p[2]; memcpy(p, personal16, 16); h[6] ^= p[0]; h[7] ^= p[1]; }
  }
  void compress(const uint8_t* block, bool last) {
    static const uint8_t S[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    memcpy(m, block, 128);
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = IV()[i]; }
    v[12] ^= t; if (last) v[14] = ~v[14];
#define ORC_G(a, b, c, d, x, y) \
    v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; ++r) {
      const uint8_t* s = S[r];
      ORC_G(0, 4, 8, 12, m[s[0]], m[s[1]]) ORC_G(1, 5, 9, 13, m[s[2]], m[s[3]]) ORC_G(2, 6, 10, 14, m[s[4]], m[s[5]]) ORC_G(3, 7, 11, 15, m[s[6]], m[s[7]])
      ORC_G(0, 5, 10, 15, m[s[8]], m[s[9]]) ORC_G(1, 6, 11, 12, m[s[10]], m[s[11]]) ORC_G(2, 7, 8, 13, m[s[12]], m[s[13]]) ORC_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef ORC_G
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
  }
  void update(const void* data, size_t len) {
    const uint8_t* p = (const uint8_t*)data;
    while (len) {
      if (buflen == 128) { t += 128; compress(buf, false); buflen = 0; }
      size_t take = 128 - buflen; if (take > len) take = len;
      memcpy(buf + buflen, p, take); buflen += take; p += take; len -= take;
    }
  }
  void finalize(uint8_t* out) const {  // const: works on a copy, like blake2b_simd's State::clone().finalize()
    Blake2b c = *this;
    c.t += c.buflen;
    memset(c.buf + c.buflen, 0, 128 - c.buflen);
    c.compress(c.buf, true);
    memcpy(out, c.h, c.outlen);
  }
};
}  // namespace orc
