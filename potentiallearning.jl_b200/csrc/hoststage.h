// hoststage.h — copies between PAGEABLE host memory and the device at (close to) the PCIe rate.
//
// cudaMemcpy from/to pageable memory is staged by the driver through a small bounce buffer on the calling thread: config 2's 1.6 GB
// result takes 105 ms instead of the 28 ms the link needs (profiles/r01_pageable_probe.json), and a Julia Array is always pageable.
// Here the staging is ours: a ring of pinned buffers, DMA on a CUDA stream, and a pool of host threads doing the pageable <-> pinned
// memcpy of several chunks concurrently (one thread moves ~10 GB/s, the link 55 GB/s). Pinned / registered host pointers bypass all of
// this (cudaPointerGetAttributes) and are copied directly.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

namespace plb {

class HostStager {
 public:
  static constexpr size_t kSlotBytes = 16u << 20;
  static constexpr int kSlots = 16;     // ring capacity (only nthreads_ + 2 of them are allocated)
  static constexpr int kDefaultThreads = 8;

  ~HostStager() { shutdown(); }

  // Start the ring and the worker threads now (otherwise at first use) with at most `max_threads` workers. A library driving several
  // GPUs from one process starts every device's stager up front: pinned allocations may synchronise all devices, which must not
  // happen while another device already waits inside a collective.
  cudaError_t start(int max_threads) {
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (!started_ && max_threads > 0) thread_cap_ = max_threads;
    }
    return ensure();
  }

  static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
  }

  // dev (pitched) -> pageable host (pitched), rows x width bytes, in stream order on `copy_stream` (make it wait on the producer's
  // event first). wait = false: returns once every chunk is queued; call finish() before touching dst.
  cudaError_t d2h_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t copy_stream, bool wait = true) {
    cudaError_t e = ensure();
    if (e != cudaSuccess) return e;
    if (width == 0 || rows == 0) return cudaSuccess;
    if (width > kSlotBytes) {  // one row does not fit a slot: let the driver stage it
      finish();
      e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, copy_stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(copy_stream);
      return e;
    }
    const size_t rows_per = kSlotBytes / width;
    for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
      const size_t nr = rows - r0 < rows_per ? rows - r0 : rows_per;
      const int s = next_slot();  // the job that last used this slot is done (its host memcpy AND the DMA it waited for / issued)
      if (h2d_used_[s]) {
        cudaEventSynchronize(ev_[s]);
        h2d_used_[s] = false;
      }
      e = cudaMemcpy2DAsync(slot_[s], width, (const char*)src + r0 * spitch, spitch, width, nr, cudaMemcpyDeviceToHost, copy_stream);
      if (e != cudaSuccess) break;
      e = cudaEventRecord(ev_[s], copy_stream);
      if (e != cudaSuccess) break;
      char* d = (char*)dst + r0 * dpitch;
      submit(s, [this, s, d, dpitch, width, nr]() {
        cudaEventSynchronize(ev_[s]);
        const char* b = (const char*)slot_[s];
        if (dpitch == width) memcpy(d, b, width * nr);
        else
          for (size_t r = 0; r < nr; ++r) memcpy(d + r * dpitch, b + r * width, width);
      });
    }
    if (wait || e != cudaSuccess) finish();
    return e;
  }

  // pageable host (pitched) -> dev (pitched). The worker that has copied a chunk into its pinned slot enqueues the DMA on `copy_stream`
  // itself. wait = false: returns once every chunk is queued; after finish() every DMA has been ENQUEUED on copy_stream (record your
  // event then) and `src` is no longer read.
  cudaError_t h2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t copy_stream, bool wait = true) {
    cudaError_t e = ensure();
    if (e != cudaSuccess) return e;
    if (width == 0 || rows == 0) return cudaSuccess;
    if (width > kSlotBytes) {
      finish();
      return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, copy_stream);
    }
    const size_t rows_per = kSlotBytes / width;
    for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
      const size_t nr = rows - r0 < rows_per ? rows - r0 : rows_per;
      const int s = next_slot();
      const bool used = h2d_used_[s];
      h2d_used_[s] = true;
      const char* sp = (const char*)src + r0 * spitch;
      char* dp = (char*)dst + r0 * dpitch;
      submit(s, [this, s, used, sp, spitch, dp, dpitch, width, nr, copy_stream]() {
        if (used) cudaEventSynchronize(ev_[s]);  // the previous DMA out of this slot has drained it
        char* b = (char*)slot_[s];
        if (spitch == width) memcpy(b, sp, width * nr);
        else
          for (size_t r = 0; r < nr; ++r) memcpy(b + r * width, sp + r * spitch, width);
        cudaError_t ee = cudaMemcpy2DAsync(dp, dpitch, b, width, width, nr, cudaMemcpyHostToDevice, copy_stream);
        if (ee == cudaSuccess) ee = cudaEventRecord(ev_[s], copy_stream);
        if (ee != cudaSuccess) {
          std::unique_lock<std::mutex> lk(mu_);
          async_error_ = ee;
        }
      });
    }
    if (wait) return finish();
    return cudaSuccess;
  }

  // all queued chunks are done (D2H: landed in the destination; H2D: their DMAs are enqueued on the copy stream)
  cudaError_t finish() {
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this]() { return pending_ == 0; });
    const cudaError_t e = async_error_;
    async_error_ = cudaSuccess;
    return e;
  }

  void shutdown() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (!started_) return;
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
    threads_.clear();
    for (int s = 0; s < kSlots; ++s) {
      if (slot_[s]) cudaFreeHost(slot_[s]);
      if (ev_[s]) cudaEventDestroy(ev_[s]);
      slot_[s] = nullptr;
      ev_[s] = nullptr;
    }
    started_ = false;
    stop_ = false;
  }

 private:
  cudaError_t ensure() {
    std::unique_lock<std::mutex> lk(mu_);
    if (started_) return cudaSuccess;
    {
      const char* env = getenv("PLB200_STAGE_THREADS");
      int t = env ? atoi(env) : kDefaultThreads;
      const int hw = (int)std::thread::hardware_concurrency();
      if (hw > 0 && t > hw) t = hw;
      if (thread_cap_ > 0 && t > thread_cap_) t = thread_cap_;
      nthreads_ = t < 1 ? 1 : (t > kSlots - 2 ? kSlots - 2 : t);
      nslots_ = nthreads_ + 2;
    }
    for (int s = 0; s < nslots_; ++s) {
      cudaError_t e = cudaHostAlloc(&slot_[s], kSlotBytes, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
      e = cudaEventCreateWithFlags(&ev_[s], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
      busy_[s] = false;
      h2d_used_[s] = false;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    for (int t = 0; t < nthreads_; ++t)
      threads_.emplace_back([this, dev]() {
        cudaSetDevice(dev);
        for (;;) {
          std::pair<int, std::function<void()>> job;
          {
            std::unique_lock<std::mutex> lk2(mu_);
            cv_.wait(lk2, [this]() { return stop_ || !jobs_.empty(); });
            if (stop_ && jobs_.empty()) return;
            job = std::move(jobs_.front());
            jobs_.pop();
          }
          job.second();
          {
            std::unique_lock<std::mutex> lk2(mu_);
            busy_[job.first] = false;
            --pending_;
          }
          done_cv_.notify_all();
        }
      });
    started_ = true;
    return cudaSuccess;
  }
  void submit(int s, std::function<void()> fn) {
    {
      std::unique_lock<std::mutex> lk(mu_);
      busy_[s] = true;
      ++pending_;
      jobs_.emplace(s, std::move(fn));
    }
    cv_.notify_one();
  }
  // round-robin slot whose previous job has completed
  int next_slot() {
    const int s = cursor_;
    cursor_ = (cursor_ + 1) % nslots_;
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this, s]() { return !busy_[s]; });
    return s;
  }

  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::queue<std::pair<int, std::function<void()>>> jobs_;
  std::vector<std::thread> threads_;
  void* slot_[kSlots] = {};
  cudaEvent_t ev_[kSlots] = {};
  bool busy_[kSlots] = {};
  bool h2d_used_[kSlots] = {};  // the slot's event marks an H2D DMA that may still be reading it
  int cursor_ = 0;
  int nthreads_ = kDefaultThreads, nslots_ = kDefaultThreads + 2;
  int thread_cap_ = 0;
  cudaError_t async_error_ = cudaSuccess;
  int pending_ = 0;
  bool started_ = false, stop_ = false;
};

}  // namespace plb
