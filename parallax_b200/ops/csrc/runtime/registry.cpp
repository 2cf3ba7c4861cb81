// Op registry: name-keyed LRU cache of validated collective signatures with a
// stable bit position per entry.
//
// Parity: horovod/common/response_cache.{h,cc} — `ResponseCache` (LRU list +
// bit positions, capacity 1024: global_state.h:135, eviction warning
// response_cache.cc:107-115) and `CacheCoordinator` (bit-vector agreement with
// one all-reduce(AND) and an OR pass for invalidations, :303-432).  A hit means
// "same name, dtype, shape, device as last time" and lets the op skip the
// cross-rank validation exchange (Horovod: skip the gather/bcast negotiation,
// operations.cc:1403-1409).  The cross-rank agreement itself is one tiny
// all-reduce of the bit vector issued by the Python layer.
#include <cstdint>
#include <cstring>
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {
struct Entry { std::string name; uint64_t sig; int bit; };
struct Registry {
  size_t capacity = 1024;
  std::list<Entry> lru;                                      // front = most recent
  std::unordered_map<std::string, std::list<Entry>::iterator> map;
  std::vector<int> free_bits;
  int next_bit = 0;
  long hits = 0, misses = 0, invalid = 0, evictions = 0;
  std::mutex mu;
};
Registry R;
}  // namespace

extern "C" {

void px_registry_reset(int capacity) {
  std::lock_guard<std::mutex> lk(R.mu);
  R.lru.clear(); R.map.clear(); R.free_bits.clear(); R.next_bit = 0;
  R.capacity = capacity > 0 ? capacity : 1024;
  R.hits = R.misses = R.invalid = R.evictions = 0;
}

// 0 = MISS (unknown name), 1 = HIT (same signature), 2 = INVALID (signature changed)
int px_registry_lookup(const char* name, uint64_t sig) {
  std::lock_guard<std::mutex> lk(R.mu);
  auto it = R.map.find(name);
  if (it == R.map.end()) { R.misses++; return 0; }
  if (it->second->sig != sig) { R.invalid++; return 2; }
  R.lru.splice(R.lru.begin(), R.lru, it->second);
  R.hits++;
  return 1;
}

// insert/update after a successful cross-rank validation; returns the bit position
int px_registry_put(const char* name, uint64_t sig) {
  std::lock_guard<std::mutex> lk(R.mu);
  auto it = R.map.find(name);
  if (it != R.map.end()) {
    it->second->sig = sig;
    R.lru.splice(R.lru.begin(), R.lru, it->second);
    return it->second->bit;
  }
  if (R.lru.size() >= R.capacity) {
    Entry& victim = R.lru.back();
    R.free_bits.push_back(victim.bit);
    R.map.erase(victim.name);
    R.lru.pop_back();
    R.evictions++;
  }
  int bit;
  if (!R.free_bits.empty()) { bit = R.free_bits.back(); R.free_bits.pop_back(); }
  else bit = R.next_bit++;
  R.lru.push_front(Entry{name, sig, bit});
  R.map[name] = R.lru.begin();
  return bit;
}

int px_registry_erase(const char* name) {
  std::lock_guard<std::mutex> lk(R.mu);
  auto it = R.map.find(name);
  if (it == R.map.end()) return 0;
  R.free_bits.push_back(it->second->bit);
  R.lru.erase(it->second);
  R.map.erase(it);
  return 1;
}

// bit vector of cached entries (words of 64 bits); returns number of words needed
int px_registry_bits(uint64_t* out, int max_words) {
  std::lock_guard<std::mutex> lk(R.mu);
  const int words = (R.next_bit + 63) / 64;
  if (out) {
    memset(out, 0, sizeof(uint64_t) * max_words);
    for (auto& e : R.lru)
      if (e.bit / 64 < max_words) out[e.bit / 64] |= (1ull << (e.bit % 64));
  }
  return words;
}

// order-independent digest of the whole cache (names + signatures): equal digests on all ranks
// <=> every rank will take the same HIT / MISS decisions from here on
uint64_t px_registry_digest() {
  std::lock_guard<std::mutex> lk(R.mu);
  uint64_t d = 0x9E3779B97F4A7C15ull * (R.lru.size() + 1);
  for (auto& e : R.lru) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : e.name) { h ^= c; h *= 1099511628211ull; }
    h ^= e.sig + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    d += h * 0xD6E8FEB86659FD93ull;
  }
  return d;
}

void px_registry_stats(long* hits, long* misses, long* invalid, long* evictions, long* size) {
  std::lock_guard<std::mutex> lk(R.mu);
  *hits = R.hits; *misses = R.misses; *invalid = R.invalid; *evictions = R.evictions;
  *size = (long)R.lru.size();
}

}  // extern "C"
