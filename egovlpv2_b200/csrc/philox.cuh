// Philox4x32-10 (Salmon et al., SC'11) -- the counter-based generator torch's CUDA dropout uses.  A keep decision is a
// pure function of (key, element index): counter = index / 4, word = index % 4, keep iff the 24-bit uniform fraction of
// that word >= p.  The key of a dropout site = (device-resident step seed) + site * golden-ratio constant, so a captured
// CUDA graph draws fresh masks on every replay and the backward regenerates the forward's mask instead of storing it.
// tests/fake_kernels.py::FakeKernels.philox_keep is the host restatement.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace egv {

EGV_DEVINL uint4 philox4(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
EGV_DEVINL bool philox_keep_word(uint32_t w, float p_drop) { return (float)(w >> 8) * (1.0f / 16777216.0f) >= p_drop; }
EGV_DEVINL bool philox_keep(unsigned long long key, unsigned long long idx, float p_drop) {
  const unsigned long long ctr = idx >> 2;
  const uint4 r = philox4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)key, (uint32_t)(key >> 32));
  const uint32_t w = (idx & 3) == 0 ? r.x : ((idx & 3) == 1 ? r.y : ((idx & 3) == 2 ? r.z : r.w));
  return philox_keep_word(w, p_drop);
}
// key of one dropout site: seed_dev may be NULL (key = site constant only: tests with a fixed stream)
EGV_DEVINL unsigned long long philox_key(const unsigned long long* seed_dev, unsigned long long site) {
  return (seed_dev ? *seed_dev : 0ull) + site * 0x9E3779B97F4A7C15ull;
}

}  // namespace egv
