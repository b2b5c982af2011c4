// Host-side training-batch sampler, stream-identical to the reference's `sample_function` (sampler.py:16-81) for a
// given seed: the reference draws from numpy's legacy `RandomState` -- MT19937 plus masked-rejection
// bounded integers -- one or more `randint(1, usernum+1)` for the user (until one with > 1 training events), then,
// newest position first, one or more `randint(1, itemnum+1)` per filled position (rejecting the user's items).
// This file restates exactly that consumption order in C++ (no numpy, no Python loop): `genrand_int32`, the legacy
// seeding `init_genrand(seed)`, and `do v = next32() & mask while (v > rng)`.  Plain host code; it lives in the same
// library so that the sampler, the time-feature rules and the kernels ship as one artifact.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/cast_b200.h"

namespace {

struct MT19937 {
  uint32_t mt[624];
  int pos;
  void seed(uint32_t s) {  // numpy _legacy_seeding(int) -> mt19937_seed == init_genrand
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    pos = 624;
  }
  void gen() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    int kk = 0;
    uint32_t y;
    for (; kk < 624 - 397; ++kk) {
      y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    for (; kk < 623; ++kk) {
      y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    pos = 0;
  }
  uint32_t next32() {
    if (pos >= 624) gen();
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // RandomState.randint(low, high) for high - low - 1 <= 0xFFFFFFFE: masked rejection on 32-bit outputs
  long bounded(long low, long high) {
    const uint64_t rng = (uint64_t)(high - low - 1);
    if (rng == 0) return low;
    uint64_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint32_t v;
    do { v = next32() & (uint32_t)mask; } while (v > rng);
    return low + (long)v;
  }
};

struct Sampler {
  int usernum, itemnum, T;
  std::vector<long> uptr;  // [usernum + 2]: events of user u are [uptr[u], uptr[u+1])
  std::vector<int> items, ratings, hours, days;
  std::vector<long long> ts;
  std::vector<long long> edges;
  std::vector<uint8_t> member;
  MT19937 rs;
};

}  // namespace

extern "C" void* cast_sampler_create(int usernum, int itemnum, const long* user_ptr, const int* items,
                                     const int* ratings, const int* hours, const int* days, const long long* ts,
                                     int maxlen, unsigned seed, const long long* edges, int n_edges) {
  if (usernum <= 0 || itemnum <= 0 || !user_ptr || !items || maxlen <= 0) return nullptr;
  Sampler* s = new Sampler();
  s->usernum = usernum;
  s->itemnum = itemnum;
  s->T = maxlen;
  s->uptr.assign(user_ptr, user_ptr + usernum + 2);
  const long n = s->uptr[usernum + 1];
  s->items.assign(items, items + n);
  if (ratings) s->ratings.assign(ratings, ratings + n);
  if (hours) s->hours.assign(hours, hours + n);
  if (days) s->days.assign(days, days + n);
  if (ts) s->ts.assign(ts, ts + n);
  if (edges && n_edges > 0) s->edges.assign(edges, edges + n_edges);
  s->member.assign((size_t)itemnum + 2, 0);
  s->rs.seed(seed);
  return s;
}

extern "C" void cast_sampler_destroy(void* h) { delete static_cast<Sampler*>(h); }

// One batch of B samples, every output [B, T] int32 (left-padded with zeros), `user` [B].  Optional outputs may be null.
static int sampler_next(void* h, int B, int* user, int* seq, int* pos, int* neg, int* timeseq, int* ratings,
                        int* hours, int* days, long long* ts_raw) {
  Sampler* s = static_cast<Sampler*>(h);
  if (!s || B <= 0 || !user || !seq || !pos || !neg) return CAST_ERR_BAD_ARG;
  const int T = s->T;
  for (int b = 0; b < B; ++b) {
    long u = s->rs.bounded(1, (long)s->usernum + 1);
    while (s->uptr[u + 1] - s->uptr[u] <= 1) u = s->rs.bounded(1, (long)s->usernum + 1);
    const long beg = s->uptr[u], n = s->uptr[u + 1] - beg;
    const int k = (int)((n - 1) < T ? (n - 1) : T);  // filled positions: events [n-1-k, n-1) as inputs
    const long lo = beg + n - 1 - k;
    int* sq = seq + (long)b * T;
    int* ps = pos + (long)b * T;
    int* ng = neg + (long)b * T;
    memset(sq, 0, sizeof(int) * T);
    memset(ps, 0, sizeof(int) * T);
    memset(ng, 0, sizeof(int) * T);
    user[b] = (int)u;
    for (int j = 0; j < k; ++j) {
      sq[T - k + j] = s->items[lo + j];
      ps[T - k + j] = s->items[lo + j + 1];
    }
    // negatives are drawn newest position first (sampler.py:44-58), rejecting every training item of the user
    for (long e = beg; e < beg + n; ++e) s->member[s->items[e]] = 1;
    for (int j = k - 1; j >= 0; --j) {
      long t = s->rs.bounded(1, (long)s->itemnum + 1);
      while (s->member[t]) t = s->rs.bounded(1, (long)s->itemnum + 1);
      ng[T - k + j] = (int)t;
    }
    for (long e = beg; e < beg + n; ++e) s->member[s->items[e]] = 0;
    if (ratings) {
      int* o = ratings + (long)b * T;
      memset(o, 0, sizeof(int) * T);
      if (!s->ratings.empty()) for (int j = 0; j < k; ++j) o[T - k + j] = s->ratings[lo + j];
    }
    if (hours) {
      int* o = hours + (long)b * T;
      memset(o, 0, sizeof(int) * T);
      if (!s->hours.empty()) for (int j = 0; j < k; ++j) o[T - k + j] = s->hours[lo + j];
    }
    if (days) {
      int* o = days + (long)b * T;
      memset(o, 0, sizeof(int) * T);
      if (!s->days.empty()) for (int j = 0; j < k; ++j) o[T - k + j] = s->days[lo + j];
    }
    if (ts_raw) {   // raw timestamps of the window: the device computes bins / hours / weekdays (cast_time_features)
      long long* o = ts_raw + (long)b * T;
      memset(o, 0, sizeof(long long) * T);
      if (!s->ts.empty()) for (int j = 0; j < k; ++j) o[T - k + j] = s->ts[lo + j];
    }
    if (timeseq) {
      int* o = timeseq + (long)b * T;
      memset(o, 0, sizeof(int) * T);
      if (!s->ts.empty() && !s->edges.empty() && k > 0) {
        const long long ref = s->ts[lo + k - 1];  // newest input event of the window (sampler.py:61-72)
        const int ne = (int)s->edges.size();
        for (int j = 0; j < k; ++j) {
          const long long delta = ref - s->ts[lo + j];
          int a = 0, c = ne;  // number of edges <= delta
          while (a < c) {
            const int mid = (a + c) >> 1;
            if (s->edges[mid] <= delta) a = mid + 1; else c = mid;
          }
          o[T - k + j] = a;
        }
      }
    }
  }
  return CAST_OK;
}

extern "C" int cast_sampler_next(void* h, int B, int* user, int* seq, int* pos, int* neg, int* timeseq, int* ratings,
                                 int* hours, int* days) {
  return sampler_next(h, B, user, seq, pos, neg, timeseq, ratings, hours, days, nullptr);
}

/* the same stream, handing out the RAW int64 timestamps [B, T] of the window instead of host-computed time features */
extern "C" int cast_sampler_next_raw(void* h, int B, int* user, int* seq, int* pos, int* neg, long long* ts) {
  if (!ts) return CAST_ERR_BAD_ARG;
  return sampler_next(h, B, user, seq, pos, neg, nullptr, nullptr, nullptr, nullptr, ts);
}
