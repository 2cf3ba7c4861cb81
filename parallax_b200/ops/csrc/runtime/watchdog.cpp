// Stall watchdog: detects a training step (or collective) that stopped making
// progress — typically a peer that died or never entered a collective, which
// would leave our device-side spin barriers waiting forever.
//
// Parity: Horovod's stall detector on the coordinator,
// horovod/common/operations.cc:703-784 (`CheckForStalledTensors`: warn after
// HOROVOD_STALL_CHECK_TIME_SECONDS listing the missing ranks, optional hard
// shutdown after HOROVOD_STALL_SHUTDOWN_TIME_SECONDS, knobs :1023-1027) and
// horovod/test/test_stall.py.  There is no coordinator here: every rank runs
// the check locally.  Progress = host heartbeat (`beat`) or completion of the
// CUDA event armed at the end of each step.  On a stall the signal pad is
// copied to the host on a private non-blocking stream (works while a kernel
// spins) and the peers whose barrier epochs lag are reported as missing.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unistd.h>

namespace {
struct Watchdog {
  std::atomic<bool> on{false};
  std::thread th;
  std::mutex mu;
  int rank = 0, world = 1;
  double warn_s = 60, shutdown_s = 0;
  std::atomic<long long> last_beat_us{0};
  std::atomic<long long> step{0};
  cudaEvent_t ev = nullptr;       // armed event (owned by the caller)
  bool armed = false;
  const uint32_t* pad_dev = nullptr;
  size_t pad_words = 0;
  std::atomic<int> stalls{0};
  bool warned = false;
  int exit_code = 17;
  // stall-path resources are created up front: cudaMallocHost / cudaFreeHost synchronise the
  // device, which never completes while a kernel is spinning on a dead peer's flag
  cudaStream_t rd_stream = nullptr;
  uint32_t* rd_host = nullptr;
};
Watchdog W;

long long now_us() {
  return std::chrono::duration_cast<std::chrono::microseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

void report_missing() {
  if (!W.pad_dev || W.pad_words == 0 || !W.rd_stream || !W.rd_host) return;
  cudaStream_t s = W.rd_stream;
  uint32_t* host = W.rd_host;
  // never block here: poll the copy for at most one second (a wedged device must not keep the
  // watchdog from reaching its shutdown limit)
  bool have = false;
  if (cudaMemcpyAsync(host, W.pad_dev, W.pad_words * 4, cudaMemcpyDeviceToHost, s) == cudaSuccess) {
    for (int i = 0; i < 100 && !have; ++i) {
      if (cudaStreamQuery(s) == cudaSuccess) have = true;
      else { cudaGetLastError(); std::this_thread::sleep_for(std::chrono::milliseconds(10)); }
    }
  }
  if (!have)
    fprintf(stderr, "[parallax watchdog] rank %d: device did not return the signal pad within 1 s\n",
            W.rank);
  if (have) {
    // layout [channel][block][src]: for block 0 of every channel report lagging peers
    const int MAXR = 16, MAXB = 128;
    for (int ch = 0; ch < 8; ++ch) {
      const uint32_t* row = host + (size_t)(ch * MAXB) * MAXR;
      uint32_t mx = 0;
      for (int r = 0; r < W.world; ++r) mx = row[r] > mx ? row[r] : mx;
      if (mx == 0) continue;
      char buf[256]; int n = 0; bool any = false;
      for (int r = 0; r < W.world && n < 240; ++r)
        if (row[r] < mx) { n += snprintf(buf + n, sizeof(buf) - n, "%s%d", any ? "," : "", r); any = true; }
      if (any)
        fprintf(stderr, "[parallax watchdog] rank %d: channel %d epoch %u — missing ranks: %s\n",
                W.rank, ch, mx, buf);
    }
  }
  cudaGetLastError();
}

void loop() {
  while (W.on.load(std::memory_order_acquire)) {
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    bool progressed = false;
    {
      std::lock_guard<std::mutex> lk(W.mu);
      if (W.armed && W.ev) {
        cudaError_t e = cudaEventQuery(W.ev);
        if (e == cudaSuccess) { W.armed = false; progressed = true; }
        else cudaGetLastError();
      }
    }
    if (progressed) { W.last_beat_us = now_us(); W.warned = false; }
    const double idle = (now_us() - W.last_beat_us.load()) / 1e6;
    bool waiting;
    { std::lock_guard<std::mutex> lk(W.mu); waiting = W.armed; }
    if (!waiting && W.pad_dev) continue;           // nothing outstanding on the device
    if (idle > W.warn_s && !W.warned) {
      W.warned = true;
      W.stalls++;
      fprintf(stderr,
              "[parallax watchdog] rank %d: no progress for %.1f s at step %lld. One or more "
              "ranks may have died or are not calling the same collectives in the same order.\n",
              W.rank, idle, W.step.load());
      fflush(stderr);
      // best-effort diagnosis on its own thread: whatever the device does, this loop goes on
      // to enforce the shutdown limit
      std::thread(report_missing).detach();
    }
    if (W.shutdown_s > 0 && idle > W.shutdown_s) {
      fprintf(stderr, "[parallax watchdog] rank %d: stalled for %.1f s > shutdown limit %.1f s; "
                      "terminating.\n", W.rank, idle, W.shutdown_s);
      fflush(stderr);
      _exit(W.exit_code);
    }
  }
}
}  // namespace

extern "C" {

int px_watchdog_start(int rank, int world, double warn_s, double shutdown_s, const void* pad_dev,
                      size_t pad_words) {
  if (W.on.load()) return 1;
  W.rank = rank; W.world = world; W.warn_s = warn_s; W.shutdown_s = shutdown_s;
  W.pad_dev = (const uint32_t*)pad_dev; W.pad_words = pad_words;
  W.last_beat_us = now_us(); W.warned = false; W.stalls = 0; W.armed = false;
  if (W.pad_dev && pad_words > 0 && !W.rd_host) {
    if (cudaStreamCreateWithFlags(&W.rd_stream, cudaStreamNonBlocking) != cudaSuccess) W.rd_stream = nullptr;
    if (cudaMallocHost(&W.rd_host, pad_words * 4) != cudaSuccess) W.rd_host = nullptr;
    cudaGetLastError();
  }
  W.on.store(true, std::memory_order_release);
  W.th = std::thread(loop);
  return 0;
}

// host heartbeat (CPU fabric, or "a step was enqueued")
void px_watchdog_beat(long long step) {
  W.step = step; W.last_beat_us = now_us(); W.warned = false;
}

// progress is the completion of `ev` (recorded by the caller at step end)
void px_watchdog_arm(cudaEvent_t ev, long long step) {
  std::lock_guard<std::mutex> lk(W.mu);
  W.ev = ev; W.armed = true; W.step = step;
}

void px_watchdog_disarm() { std::lock_guard<std::mutex> lk(W.mu); W.armed = false; }
int px_watchdog_stalls() { return W.stalls.load(); }

int px_watchdog_stop() {
  if (!W.on.load()) return 0;
  W.on.store(false, std::memory_order_release);
  if (W.th.joinable()) W.th.join();
  return 0;
}

}  // extern "C"
