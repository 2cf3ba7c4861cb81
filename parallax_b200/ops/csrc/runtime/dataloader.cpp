// Native input pipeline: multi-threaded record readers feeding a bounded
// (optionally shuffling) pool, plus a vocabulary encoder.
//
// Parity: the reference's input pipelines run inside TensorFlow's C++ runtime —
// `RecordInput` / `TFRecordReader` + `string_input_producer` +
// `RandomShuffleQueue(capacity, min_after_dequeue)` fed by queue-runner threads
// (tf_cnn_benchmarks `preprocessing.py:501-560`, skip_thoughts
// `ops/input_ops.py:63-131`), `TextLineDataset` + `lookup_ops.index_table_from_file`
// (nmt `utils/iterator_utils.py:72-160`, `utils/vocab_utils.py:104-118`), and
// `Dataset.shard(num_shards, index)` (`common/shard.py:69-87`).  This file is the
// native counterpart: file list -> N reader threads -> pool -> `next`.
//
//   kind 0 = text lines, kind 1 = TFRecord (u64 length, u32 masked-crc32c(length),
//   payload, u32 masked-crc32c(payload)).
//   sharding: by file (files i with i % S == k ... contiguous is decided by the
//   caller, which passes only its files) or by record (record j kept iff
//   j % S == k, j counted over the whole file list in order).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ------------------------------------------------------------------ crc32c
uint32_t g_crc_table[8][256];
std::once_flag g_crc_once;

void crc_init() {
  const uint32_t poly = 0x82F63B78u;           // Castagnoli, reflected
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ poly : (c >> 1);
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t)
      g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
}

uint32_t crc32c(const uint8_t* p, size_t n) {
  std::call_once(g_crc_once, crc_init);
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {                              // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^
        g_crc_table[5][(lo >> 16) & 0xFF] ^ g_crc_table[4][lo >> 24] ^
        g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^
        g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) c = g_crc_table[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

inline uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u; }

// ------------------------------------------------------------------ loader
struct Loader {
  std::vector<std::string> files;
  int kind = 0, threads = 1, epochs = 1;
  size_t capacity = 1024, min_after = 0;
  bool shuffle = false, verify = true, by_record = false;
  int num_shards = 1, shard_id = 0;
  uint64_t seed = 0;

  std::mutex mu;
  std::condition_variable not_full, not_empty;
  std::vector<std::string> pool;               // shuffle mode: random element out
  std::deque<std::string> fifo;                // FIFO mode: file order
  std::mt19937_64 rng;
  int live_readers = 0;
  std::atomic<bool> closing{false};
  std::string error;

  // work distribution: (epoch, file index) pairs handed out in order
  std::mutex wmu;
  int cur_epoch = 0;
  size_t cur_file = 0;
  std::vector<size_t> order;

  std::atomic<long> records{0}, bytes{0}, crc_errors{0};
  std::vector<std::thread> workers;
  // a record that did not fit the caller's buffer waits here for the retry (one
  // consumer thread per loader)
  std::string pending;
  bool has_pending = false;
  bool strict = false;

  size_t size_locked() const { return shuffle ? pool.size() : fifo.size(); }

  bool next_file(std::string* out, size_t* index) {
    std::lock_guard<std::mutex> lk(wmu);
    while (true) {
      if (epochs > 0 && cur_epoch >= epochs) return false;
      if (cur_file == 0) {
        order.resize(files.size());
        for (size_t i = 0; i < files.size(); ++i) order[i] = i;
        if (shuffle && !by_record) {
          std::mt19937_64 r(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(cur_epoch + 1));
          for (size_t i = order.size(); i > 1; --i) std::swap(order[i - 1], order[r() % i]);
        }
      }
      if (cur_file < order.size()) {
        *index = order[cur_file++];
        *out = files[*index];
        return true;
      }
      cur_file = 0;
      ++cur_epoch;
    }
  }

  bool push(std::string&& rec) {
    std::unique_lock<std::mutex> lk(mu);
    not_full.wait(lk, [&] { return closing || size_locked() < capacity; });
    if (closing) return false;
    bytes += (long)rec.size();
    ++records;
    if (shuffle) pool.emplace_back(std::move(rec)); else fifo.emplace_back(std::move(rec));
    not_empty.notify_one();
    return true;
  }

  void fail(const std::string& msg) {
    std::lock_guard<std::mutex> lk(mu);
    if (error.empty()) error = msg;
    closing = true;
    not_empty.notify_all(); not_full.notify_all();
  }

  // record-sharding needs the global record index: count records of earlier files
  // lazily (only in by_record mode, single reader thread enforced by the opener)
  long global_index = 0;

  bool keep(long j) const { return !by_record || num_shards <= 1 || (j % num_shards) == shard_id; }

  void read_text(const std::string& fn) {
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) { fail("cannot open " + fn); return; }
    std::string line;
    char buf[1 << 16];
    bool have = false;
    auto emit = [&]() -> bool {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      long j = global_index++;
      bool ok = true;
      if (keep(j)) ok = push(std::move(line));
      line.clear();
      return ok;
    };
    while (fgets(buf, sizeof(buf), f)) {
      size_t n = strlen(buf);
      have = true;
      if (n && buf[n - 1] == '\n') {
        line.append(buf, n - 1);
        if (!emit()) { fclose(f); return; }
        have = false;
      } else {
        line.append(buf, n);
      }
    }
    if (have) emit();
    fclose(f);
  }

  void read_tfrecord(const std::string& fn) {
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) { fail("cannot open " + fn); return; }
    while (true) {
      uint8_t hdr[12];
      size_t got = fread(hdr, 1, 12, f);
      if (got == 0) break;
      if (got != 12) { fail("truncated record header in " + fn); break; }
      uint64_t len; uint32_t lcrc;
      memcpy(&len, hdr, 8); memcpy(&lcrc, hdr + 8, 4);
      if (verify && mask_crc(crc32c(hdr, 8)) != lcrc) {
        ++crc_errors; fail("corrupted record length in " + fn); break;
      }
      if (len > (1ull << 31)) { fail("implausible record length in " + fn); break; }
      std::string rec(len, '\0');
      uint32_t dcrc;
      if (fread(&rec[0], 1, len, f) != len || fread(&dcrc, 1, 4, f) != 4) {
        fail("truncated record in " + fn); break;
      }
      if (verify && mask_crc(crc32c((const uint8_t*)rec.data(), len)) != dcrc) {
        ++crc_errors; fail("corrupted record data in " + fn); break;
      }
      long j = global_index++;
      if (keep(j) && !push(std::move(rec))) break;
    }
    fclose(f);
  }

  void worker() {
    std::string fn; size_t idx;
    while (!closing && next_file(&fn, &idx)) {
      if (kind == 0) read_text(fn); else read_tfrecord(fn);
    }
    std::lock_guard<std::mutex> lk(mu);
    --live_readers;
    not_empty.notify_all();
  }

  // 0 = record, 1 = end of data, -1 = error
  int pop(std::string* out) {
    std::unique_lock<std::mutex> lk(mu);
    not_empty.wait(lk, [&] {
      if (!error.empty()) return true;
      size_t n = size_locked();
      if (live_readers == 0) return true;
      // one reader: pop only from a FULL pool, so the stream is a function of the
      // seed alone; several readers: RandomShuffleQueue's `min_after_dequeue` rule
      if (shuffle) return strict ? n >= capacity : n > min_after;
      return n > 0;
    });
    if (!error.empty()) return -1;
    size_t n = size_locked();
    if (n == 0) return 1;
    if (shuffle) {
      size_t j = rng() % n;
      std::swap(pool[j], pool[n - 1]);
      *out = std::move(pool.back());
      pool.pop_back();
    } else {
      *out = std::move(fifo.front());
      fifo.pop_front();
    }
    not_full.notify_one();
    return 0;
  }

  void start() {
    rng.seed(seed ^ 0xD1B54A32D192ED03ull);
    int n = by_record ? 1 : threads;           // record order must be deterministic
    if (!shuffle) n = 1;                       // FIFO mode keeps file order
    live_readers = n;
    strict = shuffle && n == 1;
    for (int i = 0; i < n; ++i) workers.emplace_back([this] { worker(); });
  }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu);
      closing = true;
      not_full.notify_all(); not_empty.notify_all();
    }
    for (auto& t : workers) if (t.joinable()) t.join();
    workers.clear();
  }
};

std::mutex g_mu;
std::unordered_map<int, std::unique_ptr<Loader>> g_loaders;
int g_next = 1;

Loader* find(int h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_loaders.find(h);
  return it == g_loaders.end() ? nullptr : it->second.get();
}

// --------------------------------------------------------------- vocabulary
struct Vocab {
  std::unordered_map<std::string, int> index;
  int unk = 0;
};
std::unordered_map<int, std::unique_ptr<Vocab>> g_vocabs;

}  // namespace

extern "C" {

uint32_t px_crc32c(const void* data, size_t n) { return crc32c((const uint8_t*)data, n); }
uint32_t px_masked_crc32c(const void* data, size_t n) { return mask_crc(crc32c((const uint8_t*)data, n)); }

int px_loader_open(const char** files, int nfiles, int kind, int threads, long capacity,
                   long min_after_dequeue, int shuffle, unsigned long long seed, int epochs,
                   int num_shards, int shard_id, int shard_by_record, int verify_crc) {
  if (nfiles <= 0 || kind < 0 || kind > 1 || num_shards < 1 || shard_id < 0 ||
      shard_id >= num_shards) return -1;
  std::unique_ptr<Loader> L(new Loader());
  for (int i = 0; i < nfiles; ++i) L->files.emplace_back(files[i]);
  L->kind = kind; L->threads = threads > 0 ? threads : 1;
  L->capacity = capacity > 0 ? (size_t)capacity : 1024;
  L->min_after = shuffle ? (size_t)(min_after_dequeue > 0 ? min_after_dequeue : 0) : 0;
  if (L->min_after >= L->capacity) L->min_after = L->capacity - 1;
  L->shuffle = shuffle != 0; L->seed = seed; L->epochs = epochs;
  L->num_shards = num_shards; L->shard_id = shard_id;
  L->by_record = shard_by_record != 0; L->verify = verify_crc != 0;
  L->start();
  std::lock_guard<std::mutex> lk(g_mu);
  int h = g_next++;
  g_loaders[h] = std::move(L);
  return h;
}

// 0 = ok (len set), 1 = end of data, 2 = buffer too small (len = needed; the record
// is kept and returned by the next call), -1 = error (see px_loader_error)
int px_loader_next(int h, void* buf, size_t cap, size_t* len) {
  Loader* L = find(h);
  if (!L) return -1;
  if (!L->has_pending) {
    int rc = L->pop(&L->pending);
    if (rc != 0) return rc;
    L->has_pending = true;
  }
  *len = L->pending.size();
  if (L->pending.size() > cap) return 2;
  memcpy(buf, L->pending.data(), L->pending.size());
  L->has_pending = false;
  L->pending.clear();
  return 0;
}

// up to n records packed back to back; offsets[0..count]; returns count (0 at end),
// -1 on error, -2 if the first record alone does not fit (offsets[1] = needed)
int px_loader_next_batch(int h, int n, void* buf, size_t cap, size_t* offsets) {
  Loader* L = find(h);
  if (!L) return -1;
  size_t used = 0; int count = 0;
  offsets[0] = 0;
  while (count < n) {
    if (!L->has_pending) {
      int rc = L->pop(&L->pending);
      if (rc == 1) break;
      if (rc != 0) return -1;
      L->has_pending = true;
    }
    if (used + L->pending.size() > cap) {
      if (count == 0) { offsets[1] = L->pending.size(); return -2; }
      break;
    }
    memcpy((char*)buf + used, L->pending.data(), L->pending.size());
    used += L->pending.size();
    offsets[++count] = used;
    L->has_pending = false;
    L->pending.clear();
  }
  return count;
}

int px_loader_stats(int h, long* records, long* bytes, long* crc_errors) {
  Loader* L = find(h);
  if (!L) return -1;
  *records = L->records.load(); *bytes = L->bytes.load(); *crc_errors = L->crc_errors.load();
  return 0;
}

int px_loader_error(int h, char* out, size_t cap) {
  Loader* L = find(h);
  if (!L) return -1;
  std::lock_guard<std::mutex> lk(L->mu);
  snprintf(out, cap, "%s", L->error.c_str());
  return (int)L->error.size();
}

int px_loader_close(int h) {
  std::unique_ptr<Loader> L;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_loaders.find(h);
    if (it == g_loaders.end()) return -1;
    L = std::move(it->second);
    g_loaders.erase(it);
  }
  L->stop();
  return 0;
}

// ---- vocabulary: newline separated words -> ids; whitespace tokenisation -------
int px_vocab_create(const char* words, size_t nbytes, int unk_id) {
  std::unique_ptr<Vocab> v(new Vocab());
  v->unk = unk_id;
  const char* p = words; const char* end = words + nbytes;
  int id = 0;
  while (p < end) {
    const char* q = (const char*)memchr(p, '\n', end - p);
    if (!q) q = end;
    v->index.emplace(std::string(p, q - p), id);   // first occurrence wins
    ++id;
    p = q + 1;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  int h = g_next++;
  g_vocabs[h] = std::move(v);
  return h;
}

// encode one line; returns the number of tokens (may exceed cap: only cap are written)
int px_vocab_encode(int h, const char* line, size_t n, long long* out, int cap) {
  Vocab* v;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_vocabs.find(h);
    if (it == g_vocabs.end()) return -1;
    v = it->second.get();
  }
  int count = 0;
  size_t i = 0;
  while (i < n) {
    while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\n' || line[i] == '\r')) ++i;
    size_t s = i;
    while (i < n && !(line[i] == ' ' || line[i] == '\t' || line[i] == '\n' || line[i] == '\r')) ++i;
    if (i > s) {
      if (count < cap) {
        auto it = v->index.find(std::string(line + s, i - s));
        out[count] = it == v->index.end() ? v->unk : it->second;
      }
      ++count;
    }
  }
  return count;
}

int px_vocab_free(int h) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_vocabs.erase(h) ? 0 : -1;
}

}  // extern "C"
