// host_stage.h — uploads from host columns that are NOT page-locked.
//
// plonky2's prover hands `PolynomialBatch::from_values` a `Vec<PolynomialValues<F>>` it allocated
// itself ([P2] plonk/prover.rs: `wires_values`, reached from
// /root/reference/src/vtfhe/ivc_based_vpbs.rs:302/:333/:364): ordinary pageable memory.
// cudaMemcpyAsync from such memory is staged by the driver through its own bounce buffer ON THE
// CALLING THREAD, so every upload finishes before the first kernel is even enqueued and nothing
// overlaps.  Here the library does the staging itself: a ring of pinned slots, a few copy threads
// that fill a slot in parallel, and an uploader thread that sends each slot on the H2D stream while
// the calling thread is already enqueueing the kernels of the previous column chunk.  The compute
// stream waits for chunk k's event, and the calling thread waits (on the host) until that event has
// actually been recorded before it enqueues the wait.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace hoststage {

struct Segment {
  char* dst;
  const char* src;
  size_t bytes;
};

// T - 1 persistent worker threads + the caller copy a list of segments, split into <= 256 KiB tasks.
class CopyPool {
 public:
  explicit CopyPool(unsigned threads) {
    for (unsigned t = 1; t < threads; t++) workers_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  unsigned threads() const { return (unsigned)workers_.size() + 1; }
  void copy(const std::vector<Segment>& segs) {
    tasks_.clear();
    constexpr size_t TASK = 256u << 10;
    for (const Segment& s : segs)
      for (size_t o = 0; o < s.bytes; o += TASK)
        tasks_.push_back(Segment{s.dst + o, s.src + o, s.bytes - o < TASK ? s.bytes - o : TASK});
    if (workers_.empty() || tasks_.size() == 1) {
      for (const Segment& t : tasks_) memcpy(t.dst, t.src, t.bytes);
      return;
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      next_.store(0);
      pending_ = (unsigned)workers_.size();
      gen_++;
    }
    cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void drain() {
    for (;;) {
      const size_t i = next_.fetch_add(1);
      if (i >= tasks_.size()) return;
      memcpy(tasks_[i].dst, tasks_[i].src, tasks_[i].bytes);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      drain();
      {
        std::lock_guard<std::mutex> g(mu_);
        pending_--;
      }
      done_cv_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<Segment> tasks_;
  std::atomic<size_t> next_{0};
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  uint64_t gen_ = 0;
  unsigned pending_ = 0;
  bool stop_ = false;
};

// The pinned ring of one context.
struct Ring {
  static constexpr size_t SLOT_BYTES = 4u << 20;
  static constexpr int SLOTS = 4;
  char* slot[SLOTS] = {};
  cudaEvent_t free_ev[SLOTS] = {};
  bool used[SLOTS] = {};
  CopyPool* pool = nullptr;
  cudaError_t ensure(unsigned threads) {
    for (int s = 0; s < SLOTS; s++) {
      if (!slot[s]) {
        cudaError_t e = cudaHostAlloc((void**)&slot[s], SLOT_BYTES, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&free_ev[s], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
      }
    }
    if (pool && pool->threads() != threads) {
      delete pool;
      pool = nullptr;
    }
    if (!pool) pool = new CopyPool(threads);
    return cudaSuccess;
  }
  void release() {
    delete pool;
    pool = nullptr;
    for (int s = 0; s < SLOTS; s++) {
      if (slot[s]) cudaFreeHost(slot[s]);
      if (free_ev[s]) cudaEventDestroy(free_ev[s]);
      slot[s] = nullptr;
      free_ev[s] = nullptr;
      used[s] = false;
    }
  }
};

// One staged upload: column chunks [c0, c1) of per-column host pointers into a column-major device
// buffer (column c at dev + c * n), chunk k's event recorded on `stream` once its bytes are queued.
struct Upload {
  int device = 0;
  Ring* ring = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t* dev = nullptr;
  const uint64_t* const* cols = nullptr;
  uint64_t n = 0;
  std::vector<std::pair<uint32_t, uint32_t>> chunks;
  std::vector<cudaEvent_t> ready;  // one per chunk (may be shorter: no event for that chunk)
  cudaEvent_t done_ev = nullptr;   // optional (timing): recorded after the last chunk

  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  unsigned recorded = 0;  // chunks whose copies AND event record have been enqueued
  bool finished = false;
  cudaError_t err = cudaSuccess;

  void start() { th = std::thread([this] { run(); }); }
  // Host-side wait: chunk k's event has been recorded (or the upload failed / ended).
  void wait_recorded(unsigned k) {
    std::unique_lock<std::mutex> l(mu);
    cv.wait(l, [&] { return recorded > k || finished; });
  }
  cudaError_t join() {
    if (th.joinable()) th.join();
    return err;
  }
  ~Upload() { join(); }

 private:
  void run() {
    cudaError_t e = cudaSetDevice(device);
    int s = 0;
    std::vector<Segment> segs;
    for (unsigned k = 0; k < chunks.size() && e == cudaSuccess; k++) {
      // the chunk as (host, device offset, bytes) pieces of at most one slot
      uint32_t c = chunks[k].first;
      uint64_t off = 0;  // bytes already taken from column c
      const uint64_t col_bytes = n * sizeof(uint64_t);
      while (c < chunks[k].second && e == cudaSuccess) {
        if (ring->used[s]) e = cudaEventSynchronize(ring->free_ev[s]);
        if (e != cudaSuccess) break;
        segs.clear();
        size_t fill = 0;
        uint64_t* dst_dev = dev + (uint64_t)c * n + off / sizeof(uint64_t);
        while (c < chunks[k].second && fill < Ring::SLOT_BYTES) {
          const size_t take = (size_t)((col_bytes - off) < (Ring::SLOT_BYTES - fill) ? (col_bytes - off)
                                                                                      : (Ring::SLOT_BYTES - fill));
          segs.push_back(Segment{ring->slot[s] + fill, reinterpret_cast<const char*>(cols[c]) + off, take});
          fill += take;
          off += take;
          if (off == col_bytes) {
            c++;
            off = 0;
          }
        }
        ring->pool->copy(segs);
        e = cudaMemcpyAsync(dst_dev, ring->slot[s], fill, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaEventRecord(ring->free_ev[s], stream);
        ring->used[s] = true;
        s = (s + 1) % Ring::SLOTS;
      }
      if (e == cudaSuccess && k < ready.size()) e = cudaEventRecord(ready[k], stream);
      {
        std::lock_guard<std::mutex> g(mu);
        if (e == cudaSuccess) recorded = k + 1;
      }
      cv.notify_all();
    }
    if (e == cudaSuccess && done_ev) e = cudaEventRecord(done_ev, stream);
    {
      std::lock_guard<std::mutex> g(mu);
      err = e;
      finished = true;
    }
    cv.notify_all();
  }
};

// true when `p` is ordinary host memory the driver would have to stage (not cudaHostAlloc'ed /
// cudaHostRegister'ed, not managed, not device)
inline bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace hoststage
