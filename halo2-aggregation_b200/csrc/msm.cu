// msm.cu — Pippenger G1 multi-scalar multiplication for sm_100a.
//
// Computes the value of halo2 `arithmetic::best_multiexp(coeffs, bases)` — reached from the
// reference through `commit_lagrange` (examples/simple-example.rs:638-640) and every commitment of
// `create_proof` (:606-613, :702-709) — as a bucket method laid out for the GPU:
//
//   1. digits+histogram : scalars leave Montgomery form, are cut into W = ceil(254/c) signed c-bit
//                         digits; |digit| selects one of 2^(c-1) buckets per window; count per bucket.
//   2. scan             : exclusive prefix over the W * 2^(c-1) counters; in the same pass every bucket is
//                         cut into tasks of at most L entries (a second prefix) so that no thread ever
//                         owns more than L additions, however skewed the scalars are.
//   3. scatter          : (point index | sign) written to its bucket's slice: a counting sort.
//   4. accumulate       : one thread per task sums its points with XYZZ mixed additions; bases are
//                         fetched with 128-bit loads, the next point prefetched during the current add.
//   4b. merge           : buckets that were cut into several tasks are folded by one thread each (up to 32
//                         tasks) or one warp each (lanes stride over the partial sums, then a shuffle tree).
//   5. reduce           : per window, sum_b b * bucket[b] by segmented running sums.
//   6. combine          : W window sums come back to the host; Horner with c doublings per window,
//                         one inversion to canonical affine.
// The arithmetic is 254-bit Montgomery on the 32-bit integer pipe (field.cuh); nothing here is
// a dense contraction, so no tensor cores.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "curve.cuh"
#include "host_bn254.hpp"
#include "tree_layout.hpp"

using namespace h2a;

namespace {

constexpr int SCAN_ITEMS = 8;       // per thread
constexpr int SCAN_THREADS = 1024;  // per block
constexpr int SCAN_CHUNK = SCAN_ITEMS * SCAN_THREADS;

template <int C>
struct Win {
    static constexpr int W = (254 + C - 1) / C;
    static constexpr uint32_t B = 1u << (C - 1);
};

// Signed-digit recoding of a canonical 254-bit scalar: digit_w in (-2^(c-1), 2^(c-1)], carries
// ripple upward; the top window never overflows because 254 - (W-1)c <= c-1 for every c in 3..24.
// f(w, mag, neg) is called for EVERY window (mag == 0: no entry), so a warp stays converged across the calls.
template <int C, class F>
__device__ __forceinline__ void for_each_digit(const uint32_t (&s)[8], F f) {
    constexpr int W = Win<C>::W;
    constexpr uint32_t B = Win<C>::B;
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < W; w++) {
        const int off = w * C, i = off >> 5, sh = off & 31;
        uint32_t v = s[i] >> sh;
        if (sh + C > 32 && i + 1 < 8) v |= s[i + 1] << (32 - sh);
        v &= (1u << C) - 1u;
        v += carry;
        const bool neg = v > B;
        carry = neg ? 1u : 0u;
        const uint32_t mag = neg ? (2u * B - v) : v;
        f(w, mag, neg);
    }
}

// Do the lowest digits of this warp's scalars collide?  Uniformly random scalars never do (32 draws from 2^(c-1)
// buckets); small, repeated or sorted ones — selector and range-checked columns, permuted lookup columns, grand
// products that stay at one — always do, and their plain atomics then serialise on a handful of counters.  Such a
// warp issues ONE atomic per distinct bucket (match.any groups the lanes) instead of one per lane.
template <int C>
__device__ __forceinline__ bool warp_digits_collide(const uint32_t (&s)[8], bool live) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t key = live ? (s[0] & ((1u << (C < 32 ? C : 31)) - 1u)) : (0x80000000u | lane);
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    return __any_sync(0xffffffffu, (peers & (peers - 1u)) != 0u);
}

// PRE = the bases carry precomputed window tables T[w][i] = 2^(c*w) * P_i: every window then feeds the
// same 2^(c-1) buckets and an entry names the table element (w * stride + first + i) instead of i.
template <int C, bool PRE>
__global__ void __launch_bounds__(256) msm_hist_kernel(const uint8_t* __restrict__ scalars, uint32_t n,
                                                       uint32_t* __restrict__ hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const bool live = i < n;     // the grid is whole warps; dead lanes stay for the warp votes
    Fr s = live ? Fr::load(scalars + 32ull * i).from_mont() : Fr::zero();
    if (!warp_digits_collide<C>(s.l, live)) {
        for_each_digit<C>(s.l, [&](int w, uint32_t mag, bool) {
            if (mag) atomicAdd(&hist[(PRE ? 0u : (uint32_t)w * Win<C>::B) + mag - 1u], 1u);
        });
        return;
    }
    for_each_digit<C>(s.l, [&](int w, uint32_t mag, bool) {
        const uint32_t b = (PRE ? 0u : (uint32_t)w * Win<C>::B) + mag - 1u;
        const unsigned peers = __match_any_sync(0xffffffffu, mag ? b : (0x80000000u | lane));
        if (mag && lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(&hist[b], (uint32_t)__popc(peers));
    });
}

template <int C, bool PRE>
__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint8_t* __restrict__ scalars, uint32_t n,
                                                          uint32_t stride, uint32_t first,
                                                          uint32_t* __restrict__ cursor,
                                                          uint32_t* __restrict__ sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const bool live = i < n;
    Fr s = live ? Fr::load(scalars + 32ull * i).from_mont() : Fr::zero();
    if (!warp_digits_collide<C>(s.l, live)) {
        for_each_digit<C>(s.l, [&](int w, uint32_t mag, bool neg) {
            if (!mag) return;
            const uint32_t pos = atomicAdd(&cursor[(PRE ? 0u : (uint32_t)w * Win<C>::B) + mag - 1u], 1u);
            const uint32_t idx = PRE ? (uint32_t)w * stride + first + i : i;
            sorted[pos] = idx | (neg ? 0x80000000u : 0u);
        });
        return;
    }
    for_each_digit<C>(s.l, [&](int w, uint32_t mag, bool neg) {
        const uint32_t b = (PRE ? 0u : (uint32_t)w * Win<C>::B) + mag - 1u;
        const unsigned peers = __match_any_sync(0xffffffffu, mag ? b : (0x80000000u | lane));
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        uint32_t base = 0;
        if (mag && lane == leader) base = atomicAdd(&cursor[b], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mag) {
            const uint32_t idx = PRE ? (uint32_t)w * stride + first + i : i;
            sorted[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = idx | (neg ? 0x80000000u : 0u);
        }
    });
}

// Builds the window tables: thread i walks P_i -> 2^c P_i -> 2^(2c) P_i ... in XYZZ (X, Y parked in the
// destination slots), then normalises all W-1 multiples with ONE field inversion (Montgomery's trick over
// the thread's own chain).  The per-window ZZ, ZZZ and prefix products live in local memory by design.
constexpr int MAX_PRE_WINDOWS = 24;
__global__ void __launch_bounds__(128) msm_precompute_kernel(const uint8_t* __restrict__ bases, uint32_t n, int c,
                                                             int n_windows, uint8_t* __restrict__ table) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine p = Affine::load(bases + 64ull * i);
    p.store(table + 64ull * i);
    if (p.is_identity()) {
        for (int w = 1; w < n_windows; w++) p.store(table + 64ull * ((size_t)w * n + i));
        return;
    }
    Fq zz[MAX_PRE_WINDOWS], zzz[MAX_PRE_WINDOWS], pre[MAX_PRE_WINDOWS];
    XYZZ cur = XYZZ::from_affine(p);
    Fq run = Fq::one();
    for (int w = 1; w < n_windows; w++) {
        for (int k = 0; k < c; k++) cur = cur.dbl();
        uint8_t* slot = table + 64ull * ((size_t)w * n + i);
        cur.x.store(slot);
        cur.y.store(slot + 32);
        zz[w] = cur.zz;
        zzz[w] = cur.zzz;
        run = run * cur.zzz;
        pre[w] = run;
    }
    Fq inv = run.inv();  // 1 / (zzz_1 ... zzz_{W-1}); a point of prime order never doubles to the identity
    for (int w = n_windows - 1; w >= 1; w--) {
        Fq zi = (w > 1) ? inv * pre[w - 1] : inv;  // 1 / zzz_w
        inv = inv * zzz[w];
        Fq zzi = (zi * zz[w]).sqr();               // 1 / zz_w  (zz^3 = zzz^2)
        uint8_t* slot = table + 64ull * ((size_t)w * n + i);
        Fq x = Fq::load(slot) * zzi, y = Fq::load(slot + 32) * zi;
        x.store(slot);
        y.store(slot + 32);
    }
}

// ---- exclusive scan over the bucket counters, in place, carried together with a second prefix:
// tasks(k) = ceil(count(k) / L).  Both ride in one 64-bit value (low word: entries, high word: tasks).
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total) {
    __shared__ uint64_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t ws = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
        uint64_t winc = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        warp_sums[lane] = winc - ws;  // exclusive
        if (lane == 31 && total) *total = winc;
    }
    __syncthreads();
    uint64_t r = inc - v + warp_sums[wid];
    __syncthreads();
    return r;
}
// `pad` rounds a non-empty bucket's slot count up to a multiple of pad (a power of two; 1 = none)
__device__ __forceinline__ uint64_t pack_count(uint32_t cnt, uint32_t task_len, uint32_t pad) {
    uint32_t slots = (cnt + pad - 1u) & ~(pad - 1u);
    return (uint64_t)slots | ((uint64_t)((cnt + task_len - 1) / task_len) << 32);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint32_t* __restrict__ data, uint32_t n,
                                                                       uint32_t task_len, uint32_t pad,
                                                                       uint64_t* __restrict__ block_sums) {
    uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += (base + k < n) ? pack_count(data[base + k], task_len, pad) : 0;
    __shared__ uint64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(uint64_t* __restrict__ block_sums, uint32_t nblocks) {
    // single block; loops if there are more than SCAN_THREADS block sums
    __shared__ uint64_t total;
    uint64_t running = 0;
    for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < nblocks ? block_sums[i] : 0;
        uint64_t ex = block_exclusive_scan(v, &total);
        if (i < nblocks) block_sums[i] = running + ex;
        running += total;
        __syncthreads();
    }
}

// data[k] <- start of bucket k; task_off[k] <- first task of bucket k (task_off[n] = number of tasks);
// buckets cut into more than one task are appended to multi_list.
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint32_t* __restrict__ data, uint32_t n,
                                                                  uint32_t task_len, uint32_t pad,
                                                                  const uint64_t* __restrict__ block_sums,
                                                                  uint32_t* __restrict__ task_off,
                                                                  uint32_t* __restrict__ multi_list, uint32_t multi_cap,
                                                                  uint32_t* __restrict__ n_multi) {
    uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? pack_count(data[base + k], task_len, pad) : 0;
        s += v[k];
    }
    uint64_t ex = block_exclusive_scan(s, nullptr) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) {
            data[base + k] = (uint32_t)ex;
            task_off[base + k] = (uint32_t)(ex >> 32);
            const uint32_t nt = (uint32_t)(v[k] >> 32);
            // buckets cut into 2..32 tasks are merged by one thread each (list grows up from 0),
            // heavier ones by one warp each (list grows down from the end)
            if (nt > 32u) multi_list[multi_cap - 1u - atomicAdd(n_multi + 1, 1u)] = base + k;
            else if (nt > 1u) multi_list[atomicAdd(n_multi, 1u)] = base + k;
        }
        ex += v[k];
        if (base + k == n - 1) {
            task_off[n] = (uint32_t)(ex >> 32);
            n_multi[2] = (uint32_t)ex;   // total slots (padded): everything beyond it in the sorted array is padding
        }
    }
}

// ---- bucket accumulation: one thread per task (a run of at most task_len entries of one bucket).
// The fuller buckets are scheduled first: the top window's (last) when every window has its own buckets,
// the lowest-numbered ones when precomputed tables fold all windows into one bucket set.
// After the scatter, ends[k] is the end of bucket k's slice.  partial[t] receives task t's sum.
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const uint8_t* __restrict__ bases,
                                                             const uint32_t* __restrict__ sorted,
                                                             const uint32_t* __restrict__ ends,
                                                             const uint32_t* __restrict__ task_off, uint32_t n_buckets,
                                                             uint32_t task_len, bool top_down,
                                                             uint8_t* __restrict__ partial) {
    const uint32_t n_tasks = task_off[n_buckets];
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_tasks) return;
    const uint32_t t = top_down ? n_tasks - 1u - tid : tid;
    uint32_t lo = 0, hi = n_buckets;  // task_off[lo] <= t < task_off[hi]
    while (hi - lo > 1u) {
        uint32_t mid = (lo + hi) >> 1;
        if (task_off[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t k = lo;
    uint32_t j = (k ? ends[k - 1] : 0u) + (t - task_off[k]) * task_len;
    const uint32_t end = min(j + task_len, ends[k]);
    XYZZ acc = XYZZ::identity();
    if (j < end) {
        uint32_t e = sorted[j];
        Affine p = Affine::load_gather(bases + 64ull * (e & 0x7fffffffu));
        for (;;) {
            ++j;
            uint32_t e_next = 0;
            Affine p_next;
            const bool more = j < end;
            if (more) {
                e_next = sorted[j];
                p_next = Affine::load_gather(bases + 64ull * (e_next & 0x7fffffffu));
            }
            if (!p.is_identity()) acc.add_affine(p, (e >> 31) != 0);
            if (!more) break;
            e = e_next;
            p = p_next;
        }
    }
    acc.store(partial + 128ull * t);
}

__device__ __noinline__ void xyzz_add_nl(XYZZ& a, const XYZZ& b) { a.add(b); }

// ---- merge: a bucket that was cut into several tasks is folded by one warp; the sum replaces the
// bucket's first partial.
__device__ __forceinline__ XYZZ shfl_down_xyzz(const XYZZ& v, int off) {
    XYZZ r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], off);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], off);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], off);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], off);
    }
    return r;
}
constexpr int MERGE_WARPS = 4;
__global__ void __launch_bounds__(32 * MERGE_WARPS) msm_merge_heavy_kernel(const uint32_t* __restrict__ task_off,
                                                                          const uint32_t* __restrict__ multi_list,
                                                                          uint32_t multi_cap,
                                                                          const uint32_t* __restrict__ n_multi,
                                                                          uint8_t* __restrict__ partial) {
    const uint32_t m = blockIdx.x * MERGE_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= n_multi[1]) return;
    const uint32_t k = multi_list[multi_cap - 1u - m], t0 = task_off[k], cnt = task_off[k + 1] - t0;
    XYZZ acc = XYZZ::identity();
    for (uint32_t i = lane; i < cnt; i += 32) {
        XYZZ b = XYZZ::load(partial + 128ull * (t0 + i));
        xyzz_add_nl(acc, b);
    }
    for (int off = 16; off > 0; off >>= 1) {
        XYZZ other = shfl_down_xyzz(acc, off);
        if ((int)lane < off) xyzz_add_nl(acc, other);
    }
    if (lane == 0) acc.store(partial + 128ull * t0);
}
__global__ void __launch_bounds__(128) msm_merge_light_kernel(const uint32_t* __restrict__ task_off,
                                                              const uint32_t* __restrict__ multi_list,
                                                              const uint32_t* __restrict__ n_multi,
                                                              uint8_t* __restrict__ partial) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_multi[0]) return;
    const uint32_t k = multi_list[m], t0 = task_off[k], cnt = task_off[k + 1] - t0;
    XYZZ acc = XYZZ::load(partial + 128ull * t0);
    for (uint32_t i = 1; i < cnt; i++) {
        XYZZ b = XYZZ::load(partial + 128ull * (t0 + i));
        xyzz_add_nl(acc, b);
    }
    acc.store(partial + 128ull * t0);
}

// =====================================================================================================
// Batched-affine accumulation (ctx->msm_algo == 1).  Every bucket's slice of the sorted entries is padded to a
// multiple of 2^R slots (dummy entries = identity), so R rounds of a perfectly regular pairwise tree
//     out[o] = in[2o] + in[2o+1]
// never cross a bucket boundary and need no per-output lookup: round 0 gathers the bases through the entries,
// later rounds stream the previous round's point array.  Additions are affine — lambda = (y2-y1)/(x2-x1),
// 2M + 1S — and their inversions are shared by a two-level Montgomery trick: a forward kernel leaves per-thread
// prefix products and one total per thread, a small kernel inverts the totals 64 at a time with one field
// inversion each, a backward kernel finishes the additions: about 6 products per addition instead of the 10 of an
// XYZZ mixed addition.  After R rounds a bucket is down to (padded size / 2^R) points, which the XYZZ task kernel
// folds exactly as in the other algorithm (so skewed inputs keep their bounded per-thread work).
constexpr int AFF_B = 32;  // outputs per thread; a warp owns 32 * AFF_B consecutive outputs, lane l takes l, l+32, ...

// kind: 0 result = p1, 4 result = p2, 1 generic addition, 2 doubling, 3 cancellation (identity)
__device__ __forceinline__ int pair_case(const Affine& p1, const Affine& p2, Fq& d) {
    if (p2.is_identity()) { d = Fq::one(); return 0; }
    if (p1.is_identity()) { d = Fq::one(); return 4; }
    Fq dx = p2.x - p1.x;
    if (!dx.is_zero()) { d = dx; return 1; }
    Fq sy = p1.y + p2.y;
    if (sy.is_zero()) { d = Fq::one(); return 3; }
    d = p1.y.dbl();
    return 2;
}

template <bool FIRST>
__device__ __forceinline__ Affine load_input(const uint8_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                             const uint8_t* __restrict__ in_pts, uint32_t i) {
    if (FIRST) {
        const uint32_t e = sorted[i];
        Affine p;
        if (e == 0xffffffffu) { p.x = Fq::zero(); p.y = Fq::zero(); return p; }   // padding slot
        p = Affine::load_gather(bases + 64ull * (e & 0x7fffffffu));
        if ((e >> 31) && !p.is_identity()) p.y = p.y.neg();
        return p;
    }
    return Affine::load(in_pts + 64ull * i);
}

// forward pass: prefix products of this thread's denominators (parked in the scratch, coalesced), chunk total out
template <bool FIRST>
__global__ void __launch_bounds__(128) aff_forward_kernel(const uint8_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                          const uint8_t* __restrict__ in_pts, uint32_t n_out,
                                                          const uint32_t* __restrict__ total_slots, uint32_t o0, int round,
                                                          uint8_t* __restrict__ scratch, uint8_t* __restrict__ totals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t base_o = warp * 32u * AFF_B;
    {   // outputs past the real slots of this round are padding only: clip (o0 = first output of this half)
        const uint32_t live = (*total_slots + (2u << round) - 1u) >> (round + 1);
        n_out = live > o0 ? min(n_out, live - o0) : 0u;
    }
    Fq run = Fq::one();
    uint32_t o = base_o + lane;
    // the denominator of a generic addition needs the two x coordinates only (the y halves are fetched just for the rare
    // cases — an identity / padding slot, equal x — that pair_case has to tell apart); the x pair of the NEXT output is
    // requested before the current one is consumed, so two gathers per thread are in flight
    Fq x1, x2;
    bool pad_slot = false;
    auto fetch = [&](uint32_t oo, Fq& f1, Fq& f2, bool& pad) {
        const uint8_t *a1, *a2;
        pad = false;
        if (FIRST) {
            const uint32_t e1 = sorted[2u * oo], e2 = sorted[2u * oo + 1u];
            pad = e1 == 0xffffffffu || e2 == 0xffffffffu;
            a1 = bases + 64ull * (e1 & 0x7fffffffu);
            a2 = bases + 64ull * (e2 & 0x7fffffffu);
        } else {
            a1 = in_pts + 128ull * oo;
            a2 = a1 + 64;
        }
        if (!pad) {
            f1 = FIRST ? Fq::load_gather(a1) : Fq::load(a1);
            f2 = FIRST ? Fq::load_gather(a2) : Fq::load(a2);
        }
    };
    if (o < n_out) fetch(o, x1, x2, pad_slot);
#pragma unroll 1
    for (int j = 0; j < AFF_B && o < n_out; j++, o += 32u) {
        Fq n1, n2;
        bool npad = false;
        const bool more = j + 1 < (int)AFF_B && o + 32u < n_out;
        if (more) fetch(o + 32u, n1, n2, npad);
        Fq d;
        bool slow = pad_slot;
        if (!slow) {
            d = x2 - x1;
            slow = x1.is_zero() || x2.is_zero() || d.is_zero();
        }
        if (slow) {
            Affine p1 = load_input<FIRST>(bases, sorted, in_pts, 2u * o), p2 = load_input<FIRST>(bases, sorted, in_pts, 2u * o + 1u);
            pair_case(p1, p2, d);
        }
        run.store(scratch + 32ull * (((size_t)warp * AFF_B + j) * 32u + lane));
        run = run * d;
        x1 = n1; x2 = n2; pad_slot = npad;
    }
    run.store(totals + 32ull * t);
}
// in-place inversion of the chunk totals, INV_T per thread with one field inversion (second level of Montgomery's
// trick).  The kernel is a pure latency chain, hence the short batch; Fq::inv is branch-free (division steps).
constexpr uint32_t INV_T = 16;
__global__ void __launch_bounds__(64) aff_invert_totals_kernel(uint8_t* __restrict__ totals, uint32_t count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lo = t * INV_T;
    if (lo >= count) return;
    const uint32_t cnt = min(INV_T, count - lo);
    Fq pre[INV_T];
    Fq run = Fq::one();
    for (uint32_t q = 0; q < cnt; q++) {
        pre[q] = run;
        run = run * Fq::load(totals + 32ull * (lo + q));
    }
    Fq inv = run.inv();
    for (uint32_t q = cnt; q-- > 0;) {
        Fq v = Fq::load(totals + 32ull * (lo + q));
        (inv * pre[q]).store(totals + 32ull * (lo + q));
        inv = inv * v;
    }
}
// backward pass: individual inverses from the inverted chunk total, finish the additions
template <bool FIRST>
__global__ void __launch_bounds__(128, 6) aff_backward_kernel(const uint8_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                           const uint8_t* __restrict__ in_pts, uint32_t n_out,
                                                           const uint32_t* __restrict__ total_slots, uint32_t o0, int round,
                                                           const uint8_t* __restrict__ scratch, const uint8_t* __restrict__ totals,
                                                           uint8_t* __restrict__ out_pts) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t base_o = warp * 32u * AFF_B;
    {
        const uint32_t live = (*total_slots + (2u << round) - 1u) >> (round + 1);
        n_out = live > o0 ? min(n_out, live - o0) : 0u;
    }
    if (base_o + lane >= n_out) return;
    const uint32_t cnt = min((uint32_t)AFF_B, (n_out - base_o - lane + 31u) / 32u);  // outputs of this lane
    Fq inv = Fq::load(totals + 32ull * t);
    uint32_t o = base_o + lane + 32u * (cnt - 1u);
#pragma unroll 1
    for (int j = (int)cnt - 1; j >= 0; j--, o -= 32u) {
        Affine p1 = load_input<FIRST>(bases, sorted, in_pts, 2u * o), p2 = load_input<FIRST>(bases, sorted, in_pts, 2u * o + 1u);
        Fq d;
        const int kind = pair_case(p1, p2, d);
        Affine r = p1;
        if (kind == 4) r = p2;
        else if (kind == 3) { r.x = Fq::zero(); r.y = Fq::zero(); }
        else if (kind != 0) {
            Fq dinv = inv * Fq::load(scratch + 32ull * (((size_t)warp * AFF_B + j) * 32u + lane));
            inv = inv * d;
            Fq num;
            if (kind == 1) num = p2.y - p1.y;
            else { Fq xx = p1.x.sqr(); num = xx.dbl() + xx; p2.x = p1.x; }
            Fq lam = num * dinv;
            r.x = lam.sqr() - p1.x - p2.x;
            r.y = lam * (p1.x - r.x) - p1.y;
        }
        r.store(out_pts + 64ull * o);
    }
}

// counts after the R regular rounds: cnt2[k] = padded slots of bucket k / 2^R  (cursor[k] - starts[k] = real count)
__global__ void aff_counts_after_kernel(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ cursor, uint32_t nb,
                                        int R, uint32_t* __restrict__ cnt2) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    const uint32_t pad = 1u << R, m = cursor[k] - starts[k];
    cnt2[k] = ((m + pad - 1u) & ~(pad - 1u)) >> R;
}

// XYZZ task kernel over a compact point array (second stage of the batched-affine algorithm): same tasks, same
// merge, but the entries are the points themselves.  starts[k] = first point of bucket k.
__global__ void __launch_bounds__(128) msm_accumulate_pts_kernel(const uint8_t* __restrict__ pts, const uint32_t* __restrict__ starts,
                                                                 const uint32_t* __restrict__ task_off, uint32_t n_buckets,
                                                                 const uint32_t* __restrict__ total_pts, uint32_t task_len,
                                                                 bool top_down, uint8_t* __restrict__ partial) {
    const uint32_t n_tasks = task_off[n_buckets];
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_tasks) return;
    const uint32_t t = top_down ? n_tasks - 1u - tid : tid;
    uint32_t lo = 0, hi = n_buckets;
    while (hi - lo > 1u) {
        uint32_t mid = (lo + hi) >> 1;
        if (task_off[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t k = lo;
    uint32_t j = starts[k] + (t - task_off[k]) * task_len;
    const uint32_t bucket_end = (k + 1 < n_buckets) ? starts[k + 1] : *total_pts;   // real total, left by the scan
    const uint32_t end = min(j + task_len, bucket_end);
    XYZZ acc = XYZZ::identity();
    for (; j < end; j++) {
        Affine p = Affine::load(pts + 64ull * j);
        if (!p.is_identity()) acc.add_affine(p, false);
    }
    acc.store(partial + 128ull * t);
}

// ---- bucket reduction: window total = sum_t (t+1) * bucket[t].  A single thread's point addition is a chain of
// dependent Montgomery products (~7 us), so the phase is latency-bound and organised for few additions in series:
//   level 1  one thread per segment j of L = 2^log_l buckets: S_j = plain sum, T_j = sum_i (i+1) * bucket[jL+i]
//            (running sum, 2L additions in series; L = 16 measured best against L = 4, 8, 32 — a variant issuing the
//            two additions of a step interleaved in one thread needed 246 registers and was no faster)
//   level 2  total = sum_j T_j + L * sum_j j * S_j, and sum_j j * S_j = sum_b 2^b * P_b with
//            P_b = sum of the S_j whose index has bit b set: plain tree sums, one block per (window, b, chunk)
//   level 3  tree over the chunks
//   level 4  (several windows) 2^(b + log_l) * P_b by doublings, one thread per bit, and the sum over the bits;
//            with a single window the host finishes the 17-point Horner itself.
// Bucket k's value is partial[task_off[k]] (identity when the bucket has no task).

__device__ __forceinline__ XYZZ load_bucket(const uint8_t* __restrict__ partial, const uint32_t* __restrict__ task_off, uint32_t k) {
    const uint32_t t0 = task_off[k], t1 = task_off[k + 1];
    return t1 > t0 ? XYZZ::load(partial + 128ull * t0) : XYZZ::identity();
}
__global__ void __launch_bounds__(128) msm_reduce_l1_kernel(const uint8_t* __restrict__ partial, const uint32_t* __restrict__ task_off,
                                                            uint32_t total_segs, uint32_t log_l, uint8_t* __restrict__ s_out,
                                                            uint8_t* __restrict__ t_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_segs) return;
    const uint32_t k0 = t << log_l;
    XYZZ run = load_bucket(partial, task_off, k0 + (1u << log_l) - 1u), acc = XYZZ::identity();
    for (int i = (int)(1u << log_l) - 2; i >= 0; i--) {
        xyzz_add_nl(acc, run);
        const XYZZ b = load_bucket(partial, task_off, k0 + (uint32_t)i);
        xyzz_add_nl(run, b);
    }
    xyzz_add_nl(acc, run);
    run.store(s_out + 128ull * t);
    acc.store(t_out + 128ull * t);
}

constexpr int RED_THREADS = 128;
__device__ __forceinline__ XYZZ block_tree_sum(XYZZ acc, uint8_t* sh) {
    acc.store(sh + 128 * threadIdx.x);
    __syncthreads();
    for (int stride = RED_THREADS / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            XYZZ a = XYZZ::load(sh + 128 * threadIdx.x), b = XYZZ::load(sh + 128 * (threadIdx.x + stride));
            xyzz_add_nl(a, b);
            a.store(sh + 128 * threadIdx.x);
        }
        __syncthreads();
    }
    return XYZZ::load(sh);
}
// grid (chunks, windows * (nbits + 1)); q < nbits: P_q restricted to the chunk; q == nbits: the chunk's sum of T
__global__ void __launch_bounds__(RED_THREADS) msm_reduce_l2_kernel(const uint8_t* __restrict__ s_in, const uint8_t* __restrict__ t_in,
                                                                    uint32_t segs, uint32_t nbits, uint32_t chunk_len,
                                                                    uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t sh[RED_THREADS * 128];
    const uint32_t w = blockIdx.y / (nbits + 1u), q = blockIdx.y % (nbits + 1u), ch = blockIdx.x;
    const uint8_t* src = (q == nbits ? t_in : s_in) + 128ull * ((size_t)w * segs);
    XYZZ acc = XYZZ::identity();
    for (uint32_t j = ch * chunk_len + threadIdx.x; j < (ch + 1u) * chunk_len; j += RED_THREADS) {
        if (q == nbits || ((j >> q) & 1u)) {
            XYZZ b = XYZZ::load(src + 128ull * j);
            xyzz_add_nl(acc, b);
        }
    }
    acc = block_tree_sum(acc, sh);
    if (threadIdx.x == 0) acc.store(out + 128ull * ((size_t)blockIdx.y * gridDim.x + ch));
}
// one block per (window, q): sum of the chunk partials
__global__ void __launch_bounds__(RED_THREADS) msm_reduce_l3_kernel(const uint8_t* __restrict__ in, uint32_t chunks,
                                                                    uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t sh[RED_THREADS * 128];
    XYZZ acc = XYZZ::identity();
    for (uint32_t c = threadIdx.x; c < chunks; c += RED_THREADS) {
        XYZZ b = XYZZ::load(in + 128ull * ((size_t)blockIdx.x * chunks + c));
        xyzz_add_nl(acc, b);
    }
    acc = block_tree_sum(acc, sh);
    if (threadIdx.x == 0) acc.store(out + 128ull * blockIdx.x);
}
// one block per window: T + sum_b 2^(b + log_l) * P_b
__global__ void __launch_bounds__(RED_THREADS) msm_reduce_l4_kernel(const uint8_t* __restrict__ in, uint32_t nbits, uint32_t log_l,
                                                                    uint8_t* __restrict__ win_out) {
    __shared__ __align__(16) uint8_t sh[RED_THREADS * 128];
    const uint32_t q = threadIdx.x;
    XYZZ acc = XYZZ::identity();
    if (q <= nbits) {
        acc = XYZZ::load(in + 128ull * ((size_t)blockIdx.x * (nbits + 1u) + q));
        if (q < nbits)
            for (uint32_t d = 0; d < q + log_l; d++) acc = acc.dbl();
    }
    acc = block_tree_sum(acc, sh);
    if (threadIdx.x == 0) acc.store(win_out + 128ull * blockIdx.x);
}

int pick_window(size_t n) {  // from the measured sweep (tools/sweep.py --windows ...), B200
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    if (lg >= 17) return 16;
    if (lg >= 14) return 15;
    return std::max(6, std::min(10, lg - 4));
}

// M columns of n scalars each over the same bases in ONE pass (PRE only for M > 1): column j owns the bucket set
// [j * B, (j + 1) * B) exactly as window w does without tables, so the scan, the tree, the task kernels and the bucket
// reduction see one MSM of M * n points — the launch-bound and latency-bound phases are paid once per group of columns
// instead of once per column — and the reduction leaves one result (nbits + 1 level sums) per column.
template <int C, bool PRE>
int msm_launch_c(h2a_ctx* ctx, const uint8_t* d_bases, uint32_t stride, uint32_t first, const uint8_t* const* d_cols, int M, size_t n_col) {
    constexpr int W = Win<C>::W;
    constexpr uint32_t B = Win<C>::B;
    if (M < 1 || (!PRE && M != 1)) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: %d columns in one pass need precomputed tables", M);
    const int BW = PRE ? M : W;  // sets of buckets: one per column (tables) or one per window
    const uint32_t nb = (uint32_t)BW * B;
    const size_t n = n_col * (size_t)M;   // points of the whole pass
    if ((uint64_t)n * W >= (1ull << 32) || (PRE && (uint64_t)stride * W >= (1ull << 31)))
        H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: n=%zu with %d windows overflows 32-bit positions", n, W);
    // bucket reduction geometry (see msm_reduce_l1_kernel): segments of 2^log_l buckets, chunks of <= 2048 segments
    uint32_t log_l = 0;
    while ((2u << log_l) <= (uint32_t)ctx->msm_seg_len && (2u << log_l) <= B / 4u) log_l++;
    const uint32_t segs = B >> log_l;                 // per window
    uint32_t nbits = 0;
    while ((1u << nbits) < segs) nbits++;
    const uint32_t chunk_len = std::min(segs, (uint32_t)ctx->msm_red_chunk), chunks = segs / chunk_len;
    // points returned to the host: one per window, or (PRE) the nbits + 1 level sums of the single window
    const uint32_t groups = PRE ? (nbits + 1u) * (uint32_t)M : (uint32_t)W;
    const uint32_t scan_blocks = (nb + SCAN_CHUNK - 1) / SCAN_CHUNK;
    // task length: twice the mean bucket load (uniformly distributed digits then give one task per bucket while
    // the fuller buckets fed by a narrow top window are cut into a few equal tasks), shorter when buckets are scarce
    // so that about 2^17 tasks exist to fill 148 SMs
    const uint64_t mean_load = ((uint64_t)n * W) / nb;
    uint64_t task_len64 = std::max<uint64_t>(128, 2 * mean_load);
    if (nb < (1u << 17)) task_len64 = std::max<uint64_t>(32, std::min<uint64_t>(task_len64, ((uint64_t)n * W) >> 17));
    const uint32_t task_len = (uint32_t)task_len64;
    uint32_t max_multi = (uint32_t)std::min<uint64_t>(nb, (uint64_t)n * W / task_len + 1);
    uint32_t max_heavy = (uint32_t)std::min<uint64_t>(nb, (uint64_t)n * W / (32ull * task_len) + 1);
    const uint32_t max_tasks = nb + (uint32_t)((uint64_t)n * W / task_len) + 1;
    cudaStream_t st = ctx->stream;

    H2A_TRY(h2a_reserve(ctx, ctx->offsets, (size_t)nb * 4));
    H2A_TRY(h2a_reserve(ctx, ctx->cursor, ((size_t)nb + 1) * 4));                    // task_off
    H2A_TRY(h2a_reserve(ctx, ctx->misc, (size_t)scan_blocks * 8 + 64));              // block sums + n_multi
    H2A_TRY(h2a_reserve(ctx, ctx->heavy, ((size_t)max_multi + 1) * 4));              // multi_list
    H2A_TRY(h2a_reserve(ctx, ctx->sorted, n * (size_t)W * 4));
    H2A_TRY(h2a_reserve(ctx, ctx->buckets, (size_t)max_tasks * 128));                // per-task partial sums
    H2A_TRY(h2a_reserve(ctx, ctx->segsums, (size_t)BW * segs * 256 + (size_t)BW * (nbits + 1) * (chunks + 1) * 128));   // S, T, chunk partials, level sums
    H2A_TRY(h2a_reserve(ctx, ctx->winsums, (size_t)groups * 128));
    H2A_TRY(h2a_reserve_pinned(ctx, (size_t)groups * 128));
    uint32_t* offsets = (uint32_t*)ctx->offsets.p;
    uint32_t* task_off = (uint32_t*)ctx->cursor.p;
    uint64_t* block_sums = (uint64_t*)ctx->misc.p;
    uint32_t* n_multi = (uint32_t*)((uint8_t*)ctx->misc.p + (size_t)scan_blocks * 8);
    uint32_t* multi_list = (uint32_t*)ctx->heavy.p;
    uint32_t* sorted = (uint32_t*)ctx->sorted.p;
    uint8_t* partial = (uint8_t*)ctx->buckets.p;

    // batched-affine algorithm: R regular rounds over slices padded to 2^R slots, R from the mean bucket load
    int R = 0;
    if (ctx->msm_algo == 1) {
        while (R < 5 && (mean_load >> (R + 3)) >= 1) R++;   // keep >= 8 real points per 2^R of padding
        if (ctx->msm_rounds_bias) R = std::max(0, std::min(5, R + ctx->msm_rounds_bias));
    }
    const uint32_t pad = 1u << R;
    const uint64_t padded_ub = (uint64_t)n * W + (uint64_t)(pad - 1) * std::min<uint64_t>(nb, (uint64_t)n * W);
    if (padded_ub + 2ull * pad >= (1ull << 32)) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: padded entry count overflows 32 bits");
    // upper bound of the padded slot count, itself a multiple of 2^R; slots past the real total stay padding
    const uint32_t total_padded = (uint32_t)((padded_ub + 2ull * pad - 1) & ~(uint64_t)(2ull * pad - 1));   // two equal halves per round
    H2A_TRY(h2a_reserve(ctx, ctx->sorted, (size_t)total_padded * 4));
    sorted = (uint32_t*)ctx->sorted.p;

    h2a_prof_begin(ctx, 0);
    H2A_CUDA(ctx, cudaMemsetAsync(offsets, 0, (size_t)nb * 4, st));
    H2A_CUDA(ctx, cudaMemsetAsync(n_multi, 0, 8, st));
    const uint32_t pt_blocks = (uint32_t)((n_col + 255) / 256);
    for (int col = 0; col < M; col++) {
        msm_hist_kernel<C, PRE><<<pt_blocks, 256, 0, st>>>(d_cols[col], (uint32_t)n_col, offsets + (size_t)col * B);
        H2A_LAUNCH_CHECK(ctx);
    }
    h2a_prof_mark(ctx);
    scan_block_sums_kernel<<<scan_blocks, SCAN_THREADS, 0, st>>>(offsets, nb, task_len, pad, block_sums);
    H2A_LAUNCH_CHECK(ctx);
    scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(block_sums, scan_blocks);
    H2A_LAUNCH_CHECK(ctx);
    scan_apply_kernel<<<scan_blocks, SCAN_THREADS, 0, st>>>(offsets, nb, task_len, pad, block_sums, task_off, multi_list,
                                                            max_multi + 1, n_multi);
    H2A_LAUNCH_CHECK(ctx);
    h2a_prof_mark(ctx);
    uint32_t* starts = nullptr;
    if (R > 0) {   // keep the padded starts (the scatter turns `offsets` into cursors) and pre-fill the padding slots
        H2A_TRY(h2a_reserve(ctx, ctx->aff_u32, ((size_t)nb + 1) * 4 * 2 + 64));
        starts = (uint32_t*)ctx->aff_u32.p;
        H2A_CUDA(ctx, cudaMemcpyAsync(starts, offsets, (size_t)nb * 4, cudaMemcpyDeviceToDevice, st));
        H2A_CUDA(ctx, cudaMemsetAsync(sorted, 0xff, (size_t)total_padded * 4, st));
    }
    for (int col = 0; col < M; col++) {
        msm_scatter_kernel<C, PRE><<<pt_blocks, 256, 0, st>>>(d_cols[col], (uint32_t)n_col, stride, first, offsets + (size_t)col * B, sorted);
        H2A_LAUNCH_CHECK(ctx);
    }
    h2a_prof_mark(ctx);
    if (R == 0) {
        msm_accumulate_kernel<<<(max_tasks + 127) / 128, 128, 0, st>>>(d_bases, sorted, offsets, task_off, nb, task_len, !PRE,
                                                                       partial);
        H2A_LAUNCH_CHECK(ctx);
    } else {
        const uint32_t half_out0 = total_padded / 4;                                // outputs of round 0 per half
        const uint32_t n_final = half_out0 >> (R - 1);                             // outputs of the last round per half
        // Point arrays.  The halves run on two streams with nothing ordering one against the other, so they share no
        // region: half h ping-pongs between its own part of aff_a (half_out0 points) and of aff_b (half_out0 / 2), and
        // only the LAST round writes to the common array aff_c, half 0 in front of half 1, which nobody reads before
        // the join.  (One shared ping-pong pair is a race: a half that gets a round ahead overwrites, with its round
        // r+1 sums, points the other half is still reading in round r.)
        H2A_TRY(h2a_reserve(ctx, ctx->aff_a, (size_t)(total_padded / 2) * 64));
        H2A_TRY(h2a_reserve(ctx, ctx->aff_b, (size_t)(total_padded / 4 + 2) * 64));
        H2A_TRY(h2a_reserve(ctx, ctx->aff_c, ((size_t)2 * n_final + 2) * 64));
        // the two halves of the slot array are independent trees: they run on two streams so that one half's
        // forward / backward kernels fill the latency gap of the other half's totals inversion
        if (!ctx->stream2) H2A_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, ctx->stream_priority));
        if (!ctx->ev_fork) {
            H2A_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            H2A_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        const uint64_t scratch_elems = (uint64_t)half_out0 + 128ull * AFF_B;       // per half
        const uint64_t totals_elems = scratch_elems / AFF_B + 4096;
        H2A_TRY(h2a_reserve(ctx, ctx->aff_scratch, 2 * (scratch_elems + totals_elems) * 32));
        // the second stream starts one phase late (after the first half's round-0 forward pass): the halves then run
        // out of step and each one's latency-bound inversion hides under the other's forward / backward kernel
        // msm_tree_prio: the compute-bound backward passes go to a second, LOWER-priority stream per half, so that whenever a
        // block slot frees up the block scheduler hands it to a pending forward-pass block (memory-latency-bound gathers,
        // little arithmetic) first: the two kinds of blocks then share every SM instead of taking turns on the machine.
        const bool split_prio = ctx->msm_tree_prio != 0;
        if (split_prio && !ctx->stream_lo[0]) {
            for (int h = 0; h < 2; h++) {
                H2A_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream_lo[h], cudaStreamNonBlocking, ctx->stream_priority + 1));
                for (int e = 0; e < 2; e++) H2A_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_tree[h][e], cudaEventDisableTiming));
            }
        }
        for (int half = 0; half < 2; half++) {
            cudaStream_t hs = half ? ctx->stream2 : st;
            cudaStream_t bs = split_prio ? ctx->stream_lo[half] : hs;               // stream of the backward passes
            uint8_t* scratch = (uint8_t*)ctx->aff_scratch.p + (size_t)half * (scratch_elems + totals_elems) * 32;
            uint8_t* totals = scratch + scratch_elems * 32;
            uint8_t* const arrays[3] = {(uint8_t*)ctx->aff_a.p, (uint8_t*)ctx->aff_b.p, (uint8_t*)ctx->aff_c.p};
            uint32_t n_out = half_out0;
            for (int round = 0; round < R; round++) {
                const size_t o0 = (size_t)half * n_out;                            // first output of this half in this round (slot arithmetic only)
                TreeSpan in_span, out_span;
                tree_round_spans(total_padded, R, half, round, &in_span, &out_span);   // tree_layout.hpp: the halves share no region
                const uint8_t* pts_in = round ? arrays[in_span.array] + 64 * in_span.first : nullptr;
                uint8_t* pts_out = arrays[out_span.array] + 64 * out_span.first;
                const uint32_t threads = (uint32_t)((((uint64_t)n_out + 32ull * AFF_B - 1) / (32ull * AFF_B)) * 32);   // whole warps
                const uint32_t blocks = (threads + 127) / 128;
                if (round == 0) aff_forward_kernel<true><<<blocks, 128, 0, hs>>>(d_bases, sorted + 2 * o0, nullptr, n_out, n_multi + 2, (uint32_t)o0, round, scratch, totals);
                else aff_forward_kernel<false><<<blocks, 128, 0, hs>>>(nullptr, nullptr, pts_in, n_out, n_multi + 2, (uint32_t)o0, round, scratch, totals);
                H2A_LAUNCH_CHECK(ctx);
                if (half == 0 && round == 0) {
                    H2A_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
                    H2A_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
                }
                aff_invert_totals_kernel<<<(blocks * 128 / INV_T + 63) / 64, 64, 0, hs>>>(totals, blocks * 128);
                H2A_LAUNCH_CHECK(ctx);
                if (split_prio) {
                    H2A_CUDA(ctx, cudaEventRecord(ctx->ev_tree[half][0], hs));
                    H2A_CUDA(ctx, cudaStreamWaitEvent(bs, ctx->ev_tree[half][0], 0));
                }
                if (round == 0) aff_backward_kernel<true><<<blocks, 128, 0, bs>>>(d_bases, sorted + 2 * o0, nullptr, n_out, n_multi + 2, (uint32_t)o0, round, scratch, totals, pts_out);
                else aff_backward_kernel<false><<<blocks, 128, 0, bs>>>(nullptr, nullptr, pts_in, n_out, n_multi + 2, (uint32_t)o0, round, scratch, totals, pts_out);
                H2A_LAUNCH_CHECK(ctx);
                if (split_prio) {   // the next round's forward pass reads these sums (and reuses the scratch)
                    H2A_CUDA(ctx, cudaEventRecord(ctx->ev_tree[half][1], bs));
                    H2A_CUDA(ctx, cudaStreamWaitEvent(hs, ctx->ev_tree[half][1], 0));
                }
                n_out /= 2;
            }
        }
        uint8_t* const pts_final = (uint8_t*)ctx->aff_c.p;
        H2A_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
        H2A_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
        uint8_t* pts_in = pts_final;
        // second stage: the surviving points (padded count / 2^R per bucket) through the XYZZ task machinery
        const uint32_t total_pts = total_padded >> R;
        uint32_t* starts2 = starts + (nb + 1);
        aff_counts_after_kernel<<<(nb + 255) / 256, 256, 0, st>>>(starts, offsets, nb, R, starts2);
        H2A_LAUNCH_CHECK(ctx);
        const uint32_t task_len2 = std::max<uint32_t>(16, task_len >> R);
        const uint32_t max_multi2 = (uint32_t)std::min<uint64_t>(nb, (uint64_t)total_pts / task_len2 + 1);
        max_heavy = (uint32_t)std::min<uint64_t>(nb, (uint64_t)total_pts / (32ull * task_len2) + 1);
        max_multi = max_multi2;
        H2A_TRY(h2a_reserve(ctx, ctx->heavy, ((size_t)max_multi + 2) * 4));
        multi_list = (uint32_t*)ctx->heavy.p;
        H2A_CUDA(ctx, cudaMemsetAsync(n_multi, 0, 8, st));
        scan_block_sums_kernel<<<scan_blocks, SCAN_THREADS, 0, st>>>(starts2, nb, task_len2, 1, block_sums);
        H2A_LAUNCH_CHECK(ctx);
        scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(block_sums, scan_blocks);
        H2A_LAUNCH_CHECK(ctx);
        scan_apply_kernel<<<scan_blocks, SCAN_THREADS, 0, st>>>(starts2, nb, task_len2, 1, block_sums, task_off, multi_list,
                                                                max_multi + 1, n_multi);
        H2A_LAUNCH_CHECK(ctx);
        const uint32_t max_tasks2 = nb + total_pts / task_len2 + 1;
        H2A_TRY(h2a_reserve(ctx, ctx->buckets, (size_t)max_tasks2 * 128));
        partial = (uint8_t*)ctx->buckets.p;
        msm_accumulate_pts_kernel<<<(max_tasks2 + 127) / 128, 128, 0, st>>>(pts_in, starts2, task_off, nb, n_multi + 2, task_len2, !PRE,
                                                                            partial);
        H2A_LAUNCH_CHECK(ctx);
    }
    msm_merge_heavy_kernel<<<(max_heavy + MERGE_WARPS - 1) / MERGE_WARPS, 32 * MERGE_WARPS, 0, st>>>(
        task_off, multi_list, max_multi + 1, n_multi, partial);
    H2A_LAUNCH_CHECK(ctx);
    msm_merge_light_kernel<<<(max_multi + 127) / 128, 128, 0, st>>>(task_off, multi_list, n_multi, partial);
    H2A_LAUNCH_CHECK(ctx);
    h2a_prof_mark(ctx);
    {
        uint8_t* s_buf = (uint8_t*)ctx->segsums.p;
        uint8_t* t_buf = s_buf + (size_t)BW * segs * 128;
        uint8_t* chunk_buf = t_buf + (size_t)BW * segs * 128;
        uint8_t* level_buf = chunk_buf + (size_t)BW * (nbits + 1) * chunks * 128;
        msm_reduce_l1_kernel<<<(BW * segs + 127) / 128, 128, 0, st>>>(partial, task_off, BW * segs, log_l, s_buf, t_buf);
        H2A_LAUNCH_CHECK(ctx);
        uint8_t* l2_out = chunks > 1 ? chunk_buf : (PRE ? (uint8_t*)ctx->winsums.p : level_buf);
        msm_reduce_l2_kernel<<<dim3(chunks, BW * (nbits + 1)), RED_THREADS, 0, st>>>(s_buf, t_buf, segs, nbits, chunk_len, l2_out);
        H2A_LAUNCH_CHECK(ctx);
        if (chunks > 1) {
            msm_reduce_l3_kernel<<<BW * (nbits + 1), RED_THREADS, 0, st>>>(chunk_buf, chunks, PRE ? (uint8_t*)ctx->winsums.p : level_buf);
            H2A_LAUNCH_CHECK(ctx);
        }
        if (!PRE) {
            msm_reduce_l4_kernel<<<BW, RED_THREADS, 0, st>>>(level_buf, nbits, log_l, (uint8_t*)ctx->winsums.p);
            H2A_LAUNCH_CHECK(ctx);
        }
    }
    h2a_prof_mark(ctx);
    H2A_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->winsums.p, (size_t)groups * 128, cudaMemcpyDeviceToHost, st));
    ctx->msm_pending.active = true;
    ctx->msm_pending.pre = PRE;
    ctx->msm_pending.c = C;
    ctx->msm_pending.groups = PRE ? nbits + 1u : groups;
    ctx->msm_pending.cols = PRE ? (uint32_t)M : 1u;
    ctx->msm_pending.log_l = log_l;
    return H2A_OK;
}

}  // namespace

namespace {
int msm_dispatch(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* const* d_cols, int m, size_t n) {
    if (ctx->msm_pending.active) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: a launch is already pending on this lane");
    if (n == 0) {
        ctx->msm_pending.active = true;
        ctx->msm_pending.groups = 0;
        ctx->msm_pending.cols = (uint32_t)m;
        return H2A_OK;
    }
    if (n * (size_t)m > ((size_t)1 << 27)) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: n=%zu exceeds 2^27 points per call", n * (size_t)m);
    const bool pre = bases->table != nullptr && (ctx->msm_window_override == 0 || ctx->msm_window_override == bases->table_c);
    int c = pre ? bases->table_c : (ctx->msm_window_override ? ctx->msm_window_override : pick_window(n));
    switch (c) {
#define H2A_CASE(C)                                                                                                 \
    case C:                                                                                                         \
        return pre ? msm_launch_c<C, true>(ctx, bases->table, (uint32_t)bases->n, (uint32_t)offset, d_cols, m, n)   \
                   : msm_launch_c<C, false>(ctx, bases->d + 64 * offset, 0, 0, d_cols, m, n);
        H2A_CASE(6) H2A_CASE(7) H2A_CASE(8) H2A_CASE(9) H2A_CASE(10) H2A_CASE(11) H2A_CASE(12) H2A_CASE(13)
        H2A_CASE(14) H2A_CASE(15) H2A_CASE(16) H2A_CASE(17) H2A_CASE(18) H2A_CASE(19) H2A_CASE(20)
#undef H2A_CASE
        default:
            H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: window width %d not in 6..20", c);
    }
}
}  // namespace

int h2a_msm_launch(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* d_scalars, size_t n) {
    return msm_dispatch(ctx, bases, offset, &d_scalars, 1, n);
}
// m columns of n scalars each over bases [0, n) in one pass; needs the precomputed tables.  h2a_msm_finish then writes m points.
int h2a_msm_launch_cols(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* const* d_cols, int m, size_t n) {
    if (m < 1) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: no columns");
    return msm_dispatch(ctx, bases, 0, d_cols, m, n);
}

int h2a_msm_finish(h2a_ctx* ctx, uint8_t* out_affine) {
    if (!ctx->msm_pending.active) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: nothing pending on this lane");
    ctx->msm_pending.active = false;
    const uint32_t cols = std::max(1u, ctx->msm_pending.cols);
    if (ctx->msm_pending.groups == 0) {
        memset(out_affine, 0, 64 * (size_t)cols);
        return H2A_OK;
    }
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    using namespace h2a_host;
    const uint8_t* ws = (const uint8_t*)ctx->pinned;
    if (ctx->msm_pending.pre) {  // per column a single window (the tables carry the 2^(c*w) factors): T + 2^log_l * sum_b 2^b P_b
        const uint32_t nbits = ctx->msm_pending.groups - 1;
        for (uint32_t col = 0; col < cols; col++, ws += 128 * (size_t)(nbits + 1)) {
            PointX acc = px_identity();
            for (int b = (int)nbits - 1; b >= 0; b--) acc = px_add(px_dbl(acc), px_load(ws + 128 * b));
            for (uint32_t i = 0; i < ctx->msm_pending.log_l; i++) acc = px_dbl(acc);
            acc = px_add(acc, px_load(ws + 128 * nbits));
            affine_store(out_affine + 64 * (size_t)col, px_to_affine(acc));
        }
    } else {                     // Horner over the window sums: acc = acc * 2^c + S_w, top window first
        PointX acc = px_identity();
        for (int w = (int)ctx->msm_pending.groups - 1; w >= 0; w--) {
            for (int i = 0; i < ctx->msm_pending.c; i++) acc = px_dbl(acc);
            acc = px_add(acc, px_load(ws + 128 * w));
        }
        affine_store(out_affine, px_to_affine(acc));
    }
    h2a_prof_mark(ctx);
    h2a_prof_end(ctx);
    return H2A_OK;
}

int h2a_msm_run(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* d_scalars, size_t n,
                uint8_t out_affine[64]) {
    H2A_TRY(h2a_msm_launch(ctx, bases, offset, d_scalars, n));
    return h2a_msm_finish(ctx, out_affine);
}

// Two lanes alternate: while lane A's latency-bound tail (bucket reduction, window sums, copy back) drains, lane B's
// histogram / scatter / accumulation kernels already occupy the SMs.  Lane B first waits for everything queued on the
// main stream, so scalars produced there by earlier kernels are complete.
// With precomputed tables and columns of equal length, a lane takes a GROUP of columns per pass (msm_launch_c with
// M > 1, at most ctx->msm_group_cols of them and 2^23 points): the whole group shares one scan, one tree and one
// bucket reduction.
// h_src (optional): column j is first copied from host memory h_src[j] into d_scalars[j] on its lane's stream, so the
// copy of one pass overlaps the computation of the previous one (groups stay small there: the first pass's copies are exposed).
int h2a_msm_batch_dev(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* const* d_scalars, const size_t* n, int m,
                      uint8_t* out_affine, const uint8_t* const* h_src) {
    if (m <= 0) return H2A_OK;
    if (m == 1) {
        if (h_src && h_src[0] && n[0])
            H2A_CUDA(ctx, cudaMemcpyAsync((void*)d_scalars[0], h_src[0], 32 * n[0], cudaMemcpyHostToDevice, ctx->stream));
        return h2a_msm_run(ctx, bases, 0, d_scalars[0], n[0], out_affine);
    }
    h2a_ctx* alt = nullptr;
    H2A_TRY(h2a_get_alt(ctx, &alt));
    alt->msm_window_override = ctx->msm_window_override;
    alt->msm_algo = ctx->msm_algo;
    // pass sizes
    bool same_n = n[0] > 0;
    for (int j = 1; j < m; j++) same_n = same_n && n[j] == n[0];
    const bool pre = bases->table != nullptr && (ctx->msm_window_override == 0 || ctx->msm_window_override == bases->table_c);
    int gmax = 1;
    if (pre && same_n && ctx->msm_algo == 1) {
        gmax = h_src ? ctx->msm_group_cols_host : ctx->msm_group_cols;
        while (gmax > 1 && n[0] * (size_t)gmax > ((size_t)1 << 23)) gmax--;
    }
    int passes = (m + gmax - 1) / gmax;
    if (passes == 1 && m >= 4) passes = 2;   // both lanes get work
    std::vector<int> first(passes + 1, 0);
    for (int q = 0; q < passes; q++) first[q + 1] = first[q] + m / passes + (q < m % passes ? 1 : 0);

    struct Event {   // released on every exit path
        cudaEvent_t e = nullptr;
        ~Event() { if (e) cudaEventDestroy(e); }
    } ready;
    H2A_CUDA(ctx, cudaEventCreateWithFlags(&ready.e, cudaEventDisableTiming));
    H2A_CUDA(ctx, cudaEventRecord(ready.e, ctx->stream));
    H2A_CUDA(ctx, cudaStreamWaitEvent(alt->stream, ready.e, 0));
    h2a_ctx* lanes[2] = {ctx, alt};
    int pending_pass[2] = {-1, -1};
    int rc = H2A_OK;
    const bool prof = ctx->profiling;
    ctx->profiling = false;  // per-phase events of interleaved launches would be meaningless
    for (int q = 0; q < passes && rc == H2A_OK; q++) {
        h2a_ctx* lane = lanes[q & 1];
        if (pending_pass[q & 1] >= 0) {
            rc = h2a_msm_finish(lane, out_affine + 64 * (size_t)first[pending_pass[q & 1]]);
            pending_pass[q & 1] = -1;
            if (rc != H2A_OK) break;
        }
        const int j0 = first[q], cnt = first[q + 1] - j0;
        if (h_src)
            for (int j = j0; j < j0 + cnt; j++) {
                if (!h_src[j] || !n[j]) continue;
                cudaError_t e = cudaMemcpyAsync((void*)d_scalars[j], h_src[j], 32 * n[j], cudaMemcpyHostToDevice, lane->stream);
                if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = H2A_ERR_CUDA; break; }
            }
        if (rc != H2A_OK) break;
        rc = cnt == 1 ? h2a_msm_launch(lane, bases, 0, d_scalars[j0], n[j0]) : h2a_msm_launch_cols(lane, bases, d_scalars + j0, cnt, n[j0]);
        if (rc == H2A_OK) pending_pass[q & 1] = q;
        else if (lane != ctx) ctx->err = lane->err;
    }
    for (int l = 0; l < 2; l++) {
        if (pending_pass[l] >= 0) {
            int r2 = h2a_msm_finish(lanes[l], out_affine + 64 * (size_t)first[pending_pass[l]]);
            if (rc == H2A_OK) rc = r2;
        }
    }
    ctx->profiling = prof;
    ctx->launches += alt->launches;
    alt->launches = 0;
    return rc;
}

// Builds T[w][i] = 2^(c*w) * P_i for w < ceil(254/c) next to the bases (W * n * 64 bytes of HBM).
int h2a_msm_precompute(h2a_ctx* ctx, h2a_bases* bases, int c) {
    if (c < 0) c = bases->n >= (1u << 16) ? 20 : 16;  // automatic choice from the measured sweep
    if (c < 11 || c > 20) H2A_FAIL(ctx, H2A_ERR_INVALID, "precompute: window width %d not in 11..20", c);
    const int W = (254 + c - 1) / c;
    if ((uint64_t)bases->n * W >= (1ull << 31)) H2A_FAIL(ctx, H2A_ERR_INVALID, "precompute: %zu bases x %d windows too large", bases->n, W);
    if (bases->table) {
        H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        H2A_CUDA(ctx, cudaFree(bases->table));
        bases->table = nullptr;
    }
    if (bases->n == 0) return H2A_OK;
    void* t = nullptr;
    H2A_CUDA(ctx, cudaMalloc(&t, (size_t)W * bases->n * 64));
    msm_precompute_kernel<<<(unsigned)((bases->n + 127) / 128), 128, 0, ctx->stream>>>(bases->d, (uint32_t)bases->n, c, W, (uint8_t*)t);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(t);
        H2A_FAIL(ctx, H2A_ERR_CUDA, "precompute: %s", cudaGetErrorString(e));
    }
    bases->table = (uint8_t*)t;
    bases->table_c = c;
    return H2A_OK;
}

const char* h2a_msm_phase_name(int i) {
    static const char* names[] = {"digits+histogram", "scan", "scatter", "accumulate+merge", "reduce", "d2h+combine"};
    return (i >= 0 && i < 6) ? names[i] : "";
}
