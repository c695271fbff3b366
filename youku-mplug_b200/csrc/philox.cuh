// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) - the
// counter-based generator behind the dropout of the GPT-3 decoder (reference: nn.Dropout / F.dropout at
// models/modeling_distributed_gpt3.py:631,732,1056-1078, drawn from torch's CUDA Philox stream).
//
// A dropout decision is a pure function of (seed, offset, site, row, column): no state, no mask tensor.  The
// forward kernel and the backward kernel regenerate the same bits, and the CPU oracle (oracle/philox.py)
// reproduces them, so GPU results are checked against the oracle WITH dropout active.
//   key     = (seed lo, seed hi)
//   counter = (column >> 2, row, site, offset)        word (column & 3) of the 4 outputs belongs to `column`
//   keep    = word >= floor(p * 2^32)                 kept values are scaled by 1 / (1 - p)
// `site` numbers the dropout call sites of one decoder pass (ymp.h: YMP_DROP_SITE_*), `offset` is the per-pass
// counter the host advances (like torch's Philox offset), `row` / `column` are the logical tensor coordinates
// (hidden states: row = b*S + s, column = feature; attention: row = (b*heads + h)*S_q + q, column = key).
#pragma once
#include <stdint.h>

namespace ymp {

struct PhiloxKey { uint32_t k0, k1; };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Dropout parameters as the kernels receive them: rng -> {seed, offset} in DEVICE memory (read at run time, so a
// captured CUDA graph draws fresh masks on every replay), the call site and the drop probability.
struct DropSpec {
  const uint64_t* rng;
  uint32_t site;
  float p;
};
struct DropState {
  uint32_t k0, k1, site, offset, thresh;
  float scale;
};
__device__ __forceinline__ DropState drop_state(const DropSpec& d) {
  DropState s;
  const uint64_t seed = d.rng[0], off = d.rng[1];
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
  s.site = d.site; s.offset = (uint32_t)off;
  const double t = (double)d.p * 4294967296.0;
  s.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  s.scale = 1.0f / (1.0f - d.p);
  return s;
}
// random words of columns 4*(col4) .. 4*(col4)+3 of `row`
__device__ __forceinline__ uint4 drop_words(const DropState& s, uint32_t row, uint32_t col4) {
  return philox4x32_10(col4, row, s.site, s.offset, s.k0, s.k1);
}
// in place on 4 consecutive columns (col % 4 == 0)
__device__ __forceinline__ void drop4(const DropState& s, uint32_t row, uint32_t col, float& a, float& b, float& c, float& d) {
  const uint4 w = drop_words(s, row, col >> 2);
  a = w.x >= s.thresh ? a * s.scale : 0.f;
  b = w.y >= s.thresh ? b * s.scale : 0.f;
  c = w.z >= s.thresh ? c * s.scale : 0.f;
  d = w.w >= s.thresh ? d * s.scale : 0.f;
}

}  // namespace ymp
