// oracle/philox.hpp — TEST INFRASTRUCTURE ONLY.
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11),
// restated from the published algorithm; pinned by the Random123 known-answer vectors
// in tests/test_oracle_kat.py.  The reference spawns with numpy's *global* RNG
// (reference mrp00:311-315,366-367; mrp02:307-308,324,356-359), which is not
// reproducible per env; the north star replaces it with counter-based Philox, so the
// oracle and the product must agree on this mapping (DESIGN.md "RNG").
#pragma once
#include <cstdint>

namespace orc {

struct Philox4 {
    uint32_t v[4];
};

inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

enum Stream : uint32_t { kStreamSpawn = 1, kStreamResetAction = 2, kStreamAction = 3 };

// d-th uniform double in [0,1) of the sequence keyed by (seed, stream, env, epoch)
inline double uniform53(uint64_t seed, uint32_t stream, uint64_t env, uint32_t epoch, uint32_t d) {
    Philox4 r = philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), epoch, d >> 1, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (stream * 0x9E3779B9u));
    uint32_t hi = r.v[(d & 1) * 2], lo = r.v[(d & 1) * 2 + 1];
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

}  // namespace orc
