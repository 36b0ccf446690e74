/*
 * shim.h — TEST INFRASTRUCTURE.  Force-included (`g++ -include oracle/shim.h`) in front of the
 * UNMODIFIED reference sources so that every source of randomness the reference rollouts use
 * is replaced by one pre-drawn / counter-based u32 stream:
 *   std::default_random_engine          (Quoridor.h:50, Quoridor.cpp:14,831)
 *   std::uniform_real_distribution      (Quoridor.cpp:814, used at :696,:721,:748,:792)
 *   std::shuffle                        (Quoridor.cpp:720)
 *   rand() / srand()                    (Quoridor.cpp:796, TicTacToe.cpp:86,93)
 * and so that the harness can reach the reference classes' members (`#define class struct`;
 * Quoridor_state has no access specifier before its data, Quoridor.h:32-34).
 * The draw conventions are documented in oracle/philox_ref.h.  No reference source is edited
 * or copied; this header only renames std identifiers the reference spells.
 */
#ifndef ORACLE_SHIM_H
#define ORACLE_SHIM_H

/* every standard header the reference includes, BEFORE any macro games */
#include <iostream>
#include <iomanip>
#include <algorithm>
#include <cmath>
#include <ctime>
#include <cstdlib>
#include <cstdio>
#include <cstdint>
#include <cassert>
#include <random>
#include <queue>
#include <deque>
#include <vector>
#include <forward_list>
#include <unordered_map>
#include <string>
#include <stdexcept>
#include <chrono>
#include <atomic>
#include <thread>
#include <mutex>
#include <utility>
#include <iterator>
#include <pthread.h>
#include <stdlib.h>

#include "philox_ref.h"

namespace refshim {
struct DrawSource {
    const uint32_t *data;    /* explicit stream (read modulo len) or NULL for Philox          */
    uint64_t len;
    uint64_t seed;           /* Philox key                                                    */
    uint64_t rollout_id;     /* Philox stream id                                              */
    uint64_t pos;            /* draws consumed so far                                         */
};
extern thread_local DrawSource g_src;
inline uint32_t next_u32() {
    DrawSource &s = g_src;
    uint32_t v = s.data ? s.data[s.pos % s.len] : philox_draw(s.seed, s.rollout_id, (uint32_t)s.pos);
    s.pos++;
    return v;
}
}  // namespace refshim

#ifndef ORACLE_SHIM_NATIVE_RNG
namespace std {
struct stream_engine {
    typedef uint32_t result_type;
    stream_engine() {}
    template <typename T> explicit stream_engine(T) {}
    result_type operator()() { return refshim::next_u32(); }
};
template <typename T> struct stream_real_dist {
    stream_real_dist(T, T) {}
    T operator()(stream_engine &) { return (T)refshim::next_u32() * (T)(1.0 / 4294967296.0); }
};
template <typename It, typename G> void stream_shuffle(It first, It last, G &) {
    size_t n = (size_t)(last - first);
    if (n < 2) return;
    size_t j0 = mulhi32(refshim::next_u32(), (uint32_t)n);
    std::vector<std::pair<uint32_t, uint32_t> > keys;            /* (key, input index) of the others */
    for (size_t i = 0; i < n; i++) if (i != j0) keys.push_back(std::make_pair(refshim::next_u32(), (uint32_t)i));
    std::sort(keys.begin(), keys.end());
    std::vector<typename std::iterator_traits<It>::value_type> tmp(first, last);
    first[0] = tmp[j0];
    for (size_t i = 0; i + 1 < n; i++) first[i + 1] = tmp[keys[i].second];
}
}  // namespace std
inline int stream_rand() { return (int)(refshim::next_u32() & 0x7fffffffu); }
inline void stream_srand(unsigned) {}

#define default_random_engine stream_engine
#define uniform_real_distribution stream_real_dist
#define shuffle stream_shuffle
#define rand stream_rand
#define srand stream_srand
#endif /* !ORACLE_SHIM_NATIVE_RNG */

/* member access for the harness; must come after all std headers */
#define class struct

#endif
