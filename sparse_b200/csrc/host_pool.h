// host_pool.h -- a small persistent fork-join pool of host threads for the staging work of the HOST-buffer entry points
// (narrowing int64 indices into pinned staging buffers while the previous chunk is on the PCIe bus).  The workers are
// created on first use, sleep on a condition variable between calls and are never joined (the pool is leaked on
// purpose: static destruction order against the CUDA runtime at interpreter exit is not worth fighting).
#pragma once
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace b2s {

class HostPool {
  public:
    explicit HostPool(int workers) : n_(workers < 0 ? 0 : workers) {
        for (int w = 0; w < n_; ++w) std::thread([this, w] { loop(w); }).detach();
    }
    int size() const { return n_ + 1; }  // workers + the calling thread

    // fn(begin, end) over a partition of [0, n) into size() contiguous ranges; returns when all ranges are done
    void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)> &fn) {
        const int parts = size();
        if (n_ == 0 || n < (int64_t)parts * 4096) {
            fn(0, n);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            total_ = n;
            pending_ = n_;
            ++gen_;
        }
        cv_.notify_all();
        run_part(parts - 1, parts, n, fn);  // the caller takes the last range
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    static void run_part(int p, int parts, int64_t n, const std::function<void(int64_t, int64_t)> &fn) {
        const int64_t per = (n + parts - 1) / parts;
        const int64_t b = (int64_t)p * per, e = b + per < n ? b + per : n;
        if (b < e) fn(b, e);
    }
    void loop(int w) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int64_t, int64_t)> *fn;
            int64_t n;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                fn = fn_;
                n = total_;
            }
            run_part(w, n_ + 1, n, *fn);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    const int n_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int64_t, int64_t)> *fn_ = nullptr;
    int64_t total_ = 0;
    int pending_ = 0;
    uint64_t gen_ = 0;
};

}  // namespace b2s
