// Timeline: Chrome-tracing JSON writer fed through a lock-free bounded queue
// and drained by a writer thread.
//
// Parity: horovod/common/timeline.{h,cc} — TimelineWriter (lock-free queue of
// 1 Mi records + writer thread, timeline.h:46-74), per-tensor rows with
// NEGOTIATE / top-level op / ACTIVITY states (timeline.cc:184-298), cycle
// markers (timeline.cc:300-307); GPU activities are timed with CUDA events
// replayed off the critical path (horovod/common/ops/cuda_operations.cc:77-93).
// Here there is no negotiation phase; rows are buckets / tables / user ops and
// activities are e.g. WAIT_FOR_DATA, DENSE_STEP, SPARSE_PUSH, SPARSE_APPLY.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Rec {
  char name[48];
  char cat[24];
  char args[72];
  char ph;          // 'B','E','X','i','M'
  int32_t tid;
  int64_t ts_us;
  int64_t dur_us;
};

// Vyukov bounded MPMC queue
template <typename T, size_t N>
class Queue {
 public:
  Queue() { for (size_t i = 0; i < N; ++i) cells_[i].seq.store(i, std::memory_order_relaxed); }
  bool push(const T& v) {
    size_t pos = head_.load(std::memory_order_relaxed);
    for (;;) {
      Cell& c = cells_[pos & (N - 1)];
      size_t seq = c.seq.load(std::memory_order_acquire);
      intptr_t d = (intptr_t)seq - (intptr_t)pos;
      if (d == 0) {
        if (head_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) {
          c.v = v; c.seq.store(pos + 1, std::memory_order_release); return true;
        }
      } else if (d < 0) return false;   // full: drop (never block the training thread)
      else pos = head_.load(std::memory_order_relaxed);
    }
  }
  bool pop(T& v) {
    size_t pos = tail_.load(std::memory_order_relaxed);
    for (;;) {
      Cell& c = cells_[pos & (N - 1)];
      size_t seq = c.seq.load(std::memory_order_acquire);
      intptr_t d = (intptr_t)seq - (intptr_t)(pos + 1);
      if (d == 0) {
        if (tail_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) {
          v = c.v; c.seq.store(pos + N, std::memory_order_release); return true;
        }
      } else if (d < 0) return false;
      else pos = tail_.load(std::memory_order_relaxed);
    }
  }
 private:
  struct Cell { std::atomic<size_t> seq; T v; };
  std::vector<Cell> cells_{N};
  std::atomic<size_t> head_{0}, tail_{0};
};

struct GpuRange { int id; Rec rec; cudaEvent_t start, end; bool ended; };

struct Timeline {
  std::atomic<bool> on{false};
  FILE* f = nullptr;
  std::thread writer;
  Queue<Rec, 1 << 16>* q = nullptr;
  std::mutex gpu_mu;
  std::vector<GpuRange> gpu;
  int next_id = 1;
  cudaEvent_t base_ev = nullptr;
  int64_t base_us = 0;
  bool have_cuda = false;
  bool first = true;
  std::atomic<long> dropped{0}, written{0};
  int rank = 0;
};
Timeline T;

int64_t now_us() {
  return std::chrono::duration_cast<std::chrono::microseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

void json_escape(const char* s, std::string& out) {
  for (; *s; ++s) {
    if (*s == '"' || *s == '\\') { out.push_back('\\'); out.push_back(*s); }
    else if ((unsigned char)*s < 0x20) out.push_back(' ');
    else out.push_back(*s);
  }
}

void write_rec(const Rec& r) {
  std::string s;
  s += T.first ? "\n" : ",\n";
  T.first = false;
  s += "{\"name\": \""; json_escape(r.name, s);
  s += "\", \"cat\": \""; json_escape(r.cat, s);
  s += "\", \"ph\": \""; s.push_back(r.ph);
  s += "\", \"pid\": " + std::to_string(T.rank) + ", \"tid\": " + std::to_string(r.tid) +
       ", \"ts\": " + std::to_string(r.ts_us);
  if (r.ph == 'X') s += ", \"dur\": " + std::to_string(r.dur_us);
  if (r.ph == 'i') s += ", \"s\": \"g\"";
  if (r.args[0]) { s += ", \"args\": {\"info\": \""; json_escape(r.args, s); s += "\"}"; }
  s += "}";
  fputs(s.c_str(), T.f);
  T.written++;
}

void drain_gpu(bool final_pass) {
  std::lock_guard<std::mutex> lk(T.gpu_mu);
  for (size_t i = 0; i < T.gpu.size();) {
    GpuRange& g = T.gpu[i];
    bool done = g.ended && cudaEventQuery(g.end) == cudaSuccess;
    if (!done && final_pass && g.ended) done = cudaEventSynchronize(g.end) == cudaSuccess;
    if (done) {
      float ms0 = 0, ms1 = 0;
      cudaEventElapsedTime(&ms0, T.base_ev, g.start);
      cudaEventElapsedTime(&ms1, g.start, g.end);
      g.rec.ph = 'X';
      g.rec.ts_us = T.base_us + (int64_t)(ms0 * 1000.0);
      g.rec.dur_us = (int64_t)(ms1 * 1000.0);
      write_rec(g.rec);
      cudaEventDestroy(g.start); cudaEventDestroy(g.end);
      T.gpu[i] = T.gpu.back(); T.gpu.pop_back();
    } else if (final_pass) {
      cudaEventDestroy(g.start); if (g.ended) cudaEventDestroy(g.end);
      T.gpu[i] = T.gpu.back(); T.gpu.pop_back();
    } else ++i;
  }
  cudaGetLastError();
}

void writer_loop() {
  Rec r;
  while (T.on.load(std::memory_order_acquire)) {
    bool any = false;
    while (T.q->pop(r)) { write_rec(r); any = true; }
    if (T.have_cuda) drain_gpu(false);
    if (!any) std::this_thread::sleep_for(std::chrono::milliseconds(2));
  }
  while (T.q->pop(r)) write_rec(r);
  if (T.have_cuda) drain_gpu(true);
}

void fill(Rec& r, const char* name, const char* cat, const char* args, char ph, int tid) {
  memset(&r, 0, sizeof(r));
  strncpy(r.name, name ? name : "", sizeof(r.name) - 1);
  strncpy(r.cat, cat ? cat : "", sizeof(r.cat) - 1);
  strncpy(r.args, args ? args : "", sizeof(r.args) - 1);
  r.ph = ph; r.tid = tid;
}

}  // namespace

extern "C" {

int px_timeline_start(const char* path, int rank, int use_cuda) {
  if (T.on.load()) return 1;
  T.f = fopen(path, "w");
  if (!T.f) return -1;
  fputs("[", T.f);
  T.first = true; T.rank = rank; T.dropped = 0; T.written = 0;
  if (!T.q) T.q = new Queue<Rec, 1 << 16>();
  T.have_cuda = false;
  if (use_cuda) {
    if (cudaEventCreate(&T.base_ev) == cudaSuccess) {
      cudaEventRecord(T.base_ev, 0);
      cudaEventSynchronize(T.base_ev);
      T.have_cuda = true;
    }
    cudaGetLastError();
  }
  T.base_us = now_us();
  T.on.store(true, std::memory_order_release);
  T.writer = std::thread(writer_loop);
  Rec r; fill(r, "process_name", "__metadata", "", 'M', 0);
  snprintf(r.args, sizeof(r.args), "parallax rank %d", rank);
  r.ts_us = T.base_us; T.q->push(r);
  return 0;
}

int px_timeline_enabled() { return T.on.load() ? 1 : 0; }

// ph: 'B' begin, 'E' end, 'i' instant, 'X' complete (dur_us used)
int px_timeline_event(const char* name, const char* cat, char ph, int tid, long long dur_us,
                      const char* args) {
  if (!T.on.load(std::memory_order_acquire)) return 0;
  Rec r; fill(r, name, cat, args, ph, tid);
  r.ts_us = now_us(); r.dur_us = dur_us;
  if (ph == 'X') r.ts_us -= dur_us;
  if (!T.q->push(r)) { T.dropped++; return -1; }
  return 0;
}

// GPU activity: begin records an event on `stream`; end records the closing
// event; the writer thread emits the range once both have completed.
int px_timeline_gpu_begin(const char* name, const char* cat, int tid, const char* args,
                          cudaStream_t stream) {
  if (!T.on.load(std::memory_order_acquire) || !T.have_cuda) return 0;
  GpuRange g; g.ended = false;
  fill(g.rec, name, cat, args, 'X', tid);
  if (cudaEventCreate(&g.start) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaEventCreate(&g.end) != cudaSuccess) { cudaEventDestroy(g.start); cudaGetLastError(); return 0; }
  cudaEventRecord(g.start, stream);
  std::lock_guard<std::mutex> lk(T.gpu_mu);
  g.id = T.next_id++;
  T.gpu.push_back(g);
  return g.id;
}

int px_timeline_gpu_end(int id, cudaStream_t stream) {
  if (id <= 0 || !T.on.load(std::memory_order_acquire)) return 0;
  std::lock_guard<std::mutex> lk(T.gpu_mu);
  for (auto& g : T.gpu)
    if (g.id == id) { cudaEventRecord(g.end, stream); g.ended = true; return 0; }
  return -1;
}

long px_timeline_written() { return T.written.load(); }
long px_timeline_dropped() { return T.dropped.load(); }

int px_timeline_stop() {
  if (!T.on.load()) return 0;
  T.on.store(false, std::memory_order_release);
  if (T.writer.joinable()) T.writer.join();
  fputs("\n]\n", T.f);
  fclose(T.f); T.f = nullptr;
  if (T.have_cuda) { cudaEventDestroy(T.base_ev); T.have_cuda = false; }
  return 0;
}

}  // extern "C"
