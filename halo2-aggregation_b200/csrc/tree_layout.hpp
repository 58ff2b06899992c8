// tree_layout.hpp — where the two halves of the batched-affine addition tree (msm.cu) keep their points.
//
// The halves run on two streams that nothing orders against each other, so the layout has one rule: no region one half
// writes in any round may be touched by the other half in any round before the join.  Half h ping-pongs between its own
// part of array A (half_out0 points) and of array B (half_out0 / 2 points); only the LAST round writes array C, half 0
// in front of half 1, and C is read after the join only.  tests/test_abi_cpu.py checks the rule through
// h2a_tree_layout for many sizes; test_msm_deep_tree_two_streams is the device case that exposed a shared layout.
#pragma once
#include <cstdint>

struct TreeSpan {
    int array;        // 0 = A, 1 = B, 2 = C, -1 = the sorted entries (input of round 0)
    uint64_t first;   // first point of the span in that array
    uint64_t count;   // points
};

// total_padded: slots of the whole tree (a multiple of 2^(R+1)); R >= 1 rounds; half 0/1; round 0..R-1.
inline void tree_round_spans(uint64_t total_padded, int R, int half, int round, TreeSpan* in, TreeSpan* out) {
    const uint64_t half_out0 = total_padded / 4;                 // outputs of round 0 per half
    auto out_of = [&](int r) {
        TreeSpan s;
        s.count = half_out0 >> r;
        if (r == R - 1) { s.array = 2; s.first = (uint64_t)half * (half_out0 >> (R - 1)); }
        else if (r & 1) { s.array = 1; s.first = (uint64_t)half * (half_out0 / 2); }
        else { s.array = 0; s.first = (uint64_t)half * half_out0; }
        return s;
    };
    *out = out_of(round);
    if (round == 0) { in->array = -1; in->first = (uint64_t)half * 2 * half_out0; in->count = 2 * half_out0; }
    else *in = out_of(round - 1);
}
