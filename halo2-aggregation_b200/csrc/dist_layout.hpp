// dist_layout.hpp — who holds which rows of an extended-domain column when ONE proof is spread over several GPUs
// (plonk_prove.cu, native distribution).
//
// Rank r evaluates the quotient on rows [r * slice, (r + 1) * slice) of the m-row extended domain (slice = m / world) and a
// quotient row reads its columns at rows i + rot * step for the rotations of the circuit, so of every extended column a rank needs
// its slice and `halo` rows on either side (halo = step * largest |rotation|), modulo m.  The owner of a column sends each rank
// that window as ONE or TWO contiguous pieces placed at their own rows of the full-size array, so the quotient kernel indexes a
// column exactly as it does on a single GPU.  Windows are used when 2 * halo <= slice; otherwise whole columns are broadcast.
// tests/test_abi_cpu.py checks through h2a_dist_window, for many (m, world, halo), that a window covers every row its rank can
// reach and nothing is sent twice.
#pragma once
#include <cstdint>

struct DistPiece {
    uint32_t first;   // first row
    uint32_t count;   // rows
};

// the window of `rank`: returns the number of pieces (1 or 2) written to out[2]; m % world == 0, 2 * halo <= m / world, world >= 2
inline int dist_window(uint32_t m, int world, int rank, uint32_t halo, DistPiece out[2]) {
    const uint32_t slice = m / (uint32_t)world;
    const int64_t a = (int64_t)rank * slice - halo, b = (int64_t)(rank + 1) * slice + halo;
    if (a < 0) {
        out[0] = DistPiece{(uint32_t)(a + m), (uint32_t)(-a)};
        out[1] = DistPiece{0u, (uint32_t)b};
        return 2;
    }
    if (b > (int64_t)m) {
        out[0] = DistPiece{(uint32_t)a, (uint32_t)(m - a)};
        out[1] = DistPiece{0u, (uint32_t)(b - m)};
        return 2;
    }
    out[0] = DistPiece{(uint32_t)a, (uint32_t)(b - a)};
    return 1;
}
