// std::unordered_set<std::string> behind a C ABI (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).
//
// The reference's C++ tiers keep every agg_hit bucket in an unordered_set<string>
// (mixed_precs_caching/evlfu_32.hpp:49) and evict `lists_C1[min].begin()` (evlfu_32.cpp:230), so
// its victim order is libstdc++'s iteration order, a function of the insert/erase history only.
// oracle/tiers.py replays that history on the very same container through this shim, which makes
// the sequential restatement comparable request-by-request with the compiled reference
// (oracle/_ref/lib*.so, built by the same g++ against the same libstdc++).  Our own code.
#include <cstring>
#include <string>
#include <unordered_set>

using Set = std::unordered_set<std::string>;

extern "C" {

void *uset_new(void) { return new Set(); }
void uset_free(void *s) { delete static_cast<Set *>(s); }
long uset_size(void *s) { return static_cast<long>(static_cast<Set *>(s)->size()); }
void uset_insert(void *s, const char *k) { static_cast<Set *>(s)->insert(std::string(k)); }
int uset_erase(void *s, const char *k) { return static_cast<int>(static_cast<Set *>(s)->erase(std::string(k))); }
int uset_contains(void *s, const char *k) { return static_cast<Set *>(s)->count(std::string(k)) ? 1 : 0; }

// Copies *begin() into out (cap bytes) and erases it; returns 0 when the set is empty.
int uset_pop_begin(void *s, char *out, int cap) {
    Set *set = static_cast<Set *>(s);
    if (set->empty()) return 0;
    auto it = set->begin();
    strncpy(out, it->c_str(), static_cast<size_t>(cap) - 1);
    out[cap - 1] = 0;
    set->erase(it);
    return 1;
}

}  // extern "C"
