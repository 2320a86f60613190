// Merlin transcript (STROBE-128 over Keccak-f[1600]) operating in place on the caller's 203-byte
// `sxt_transcript` ({state[200], pos, pos_begin, cur_flags}). Replaces sxt/proof/transcript/
// {keccakf,strobe128,transcript,transcript_utility}.cc for the inner-product argument: the byte
// layout and every absorbed / squeezed byte must equal the reference's, because the transcript is
// created by the caller and shared with the verifier.
#pragma once
#include <cstdint>
#include <cstring>

#include "scalar25.h"

namespace b200 {

namespace merlin {
inline uint64_t rotl(uint64_t x, int s) { return (x << s) | (x >> (64 - s)); }
// Keccak-f[1600], 24 rounds (FIPS 202)
inline void keccakf(uint8_t state_bytes[200]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int RHO[24] = {1,  3,  6,  10, 15, 21, 28, 36, 45, 55, 2,  14,
                              27, 41, 56, 8,  25, 43, 62, 18, 39, 61, 20, 44};
  static const int PI[24] = {10, 7,  11, 17, 18, 3, 5,  16, 8,  21, 24, 4,
                             15, 23, 19, 13, 12, 2, 20, 14, 22, 9,  6,  1};
  uint64_t a[25];
  std::memcpy(a, state_bytes, 200);
  for (int round = 0; round < 24; ++round) {
    uint64_t c[5];
    for (int x = 0; x < 5; ++x)
      c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; ++x) {
      uint64_t d = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
      for (int y = 0; y < 25; y += 5)
        a[y + x] ^= d;
    }
    uint64_t t = a[1];
    for (int i = 0; i < 24; ++i) {
      uint64_t u = a[PI[i]];
      a[PI[i]] = rotl(t, RHO[i]);
      t = u;
    }
    for (int y = 0; y < 25; y += 5) {
      uint64_t r[5];
      for (int x = 0; x < 5; ++x)
        r[x] = a[y + x];
      for (int x = 0; x < 5; ++x)
        a[y + x] = r[x] ^ (~r[(x + 1) % 5] & r[(x + 2) % 5]);
    }
    a[0] ^= RC[round];
  }
  std::memcpy(state_bytes, a, 200);
}
}  // namespace merlin

// view over the caller's 203 bytes
class Transcript {
public:
  explicit Transcript(uint8_t* bytes203) : s_(bytes203) {}

  void append_message(const char* label, const uint8_t* msg, size_t len) {
    uint32_t data_len = (uint32_t)len;
    meta_ad((const uint8_t*)label, std::strlen(label), false);
    meta_ad((const uint8_t*)&data_len, 4, true);
    begin_op(kA, false);
    absorb(msg, len);
  }
  void challenge_bytes(uint8_t* dest, size_t len, const char* label) {
    uint32_t data_len = (uint32_t)len;
    meta_ad((const uint8_t*)label, std::strlen(label), false);
    meta_ad((const uint8_t*)&data_len, 4, true);
    begin_op(kI | kA | kC, false);
    squeeze(dest, len);
  }
  // 32 challenge bytes reduced modulo l (prft::challenge_value)
  Sc challenge_scalar(const char* label) {
    uint8_t buf[32];
    challenge_bytes(buf, 32, label);
    return sc_reduce(sc_load(buf));
  }

private:
  static constexpr uint8_t kR = 166, kI = 1, kA = 2, kC = 4, kT = 8, kM = 16, kK = 32;
  uint8_t* s_;
  uint8_t& pos() { return s_[200]; }
  uint8_t& pos_begin() { return s_[201]; }
  uint8_t& cur_flags() { return s_[202]; }

  void run_f() {
    s_[pos()] ^= pos_begin();
    s_[pos() + 1] ^= 0x04;
    s_[kR + 1] ^= 0x80;
    merlin::keccakf(s_);
    pos() = 0;
    pos_begin() = 0;
  }
  void absorb(const uint8_t* data, size_t len) {
    for (size_t i = 0; i < len; ++i) {
      s_[pos()] ^= data[i];
      pos() += 1;
      if (pos() == kR)
        run_f();
    }
  }
  void squeeze(uint8_t* data, size_t len) {
    for (size_t i = 0; i < len; ++i) {
      data[i] = s_[pos()];
      s_[pos()] = 0;
      pos() += 1;
      if (pos() == kR)
        run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more)
      return;
    uint8_t old_begin = pos_begin();
    pos_begin() = pos() + 1;
    cur_flags() = flags;
    uint8_t data[2] = {old_begin, flags};
    absorb(data, 2);
    if ((flags & (kC | kK)) && pos() != 0)
      run_f();
  }
  void meta_ad(const uint8_t* data, size_t len, bool more) {
    begin_op(kM | kA, more);
    absorb(data, len);
  }
};

}  // namespace b200
