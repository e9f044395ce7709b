// pp_stage.hpp -- host-side staging copies for callers that hold ordinary (pageable) memory.
//
// A Rust caller passes Vec<f64> storage, numpy passes malloc'ed arrays: cudaMemcpyAsync on such memory is staged by
// the driver through its own bounce buffers on ONE host thread (about 12 GB/s on the B200 boxes, a fifth of what the
// link carries from pinned memory).  The batch entry points therefore do the staging themselves: a small fork-join
// pool copies each chunk into / out of the context's pinned ring with several threads while the previous chunks are
// on the wire.  Plain C++ (no CUDA), so the pool is unit-tested on the CPU (tests/test_stage_pool.py).
#pragma once

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

struct pp_copy_piece {
    void *dst;
    const void *src;
    size_t bytes;
};

class pp_stage_pool {
   public:
    // `threads` counts the calling thread: T - 1 workers are spawned, the caller copies the first share itself
    explicit pp_stage_pool(int threads) : T_(threads < 1 ? 1 : (threads > 64 ? 64 : threads)) {
        for (int i = 1; i < T_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~pp_stage_pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_job_.notify_all();
        for (std::thread &t : workers_)
            if (t.joinable()) t.join();
    }
    pp_stage_pool(const pp_stage_pool &) = delete;
    pp_stage_pool &operator=(const pp_stage_pool &) = delete;

    int threads() const { return T_; }

    // copies every piece, each split into T contiguous shares (share boundaries on 64-byte multiples of the offset);
    // returns when all bytes are in place.  One run at a time (the callers hold the context mutex).
    void run(const pp_copy_piece *pieces, size_t n) {
        if (n == 0) return;
        if (T_ == 1) {
            share(pieces, n, 0);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            pieces_ = pieces;
            n_ = n;
            pending_ = T_ - 1;
            ++generation_;
        }
        cv_job_.notify_all();
        share(pieces, n, 0);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
    }

   private:
    void share(const pp_copy_piece *pieces, size_t n, int tid) const {
        for (size_t k = 0; k < n; ++k) {
            const pp_copy_piece &p = pieces[k];
            if (p.bytes == 0) continue;
            const size_t blocks = (p.bytes + 63) / 64;  // 64-byte blocks, the last one possibly short
            const size_t b0 = blocks * (size_t)tid / (size_t)T_, b1 = blocks * (size_t)(tid + 1) / (size_t)T_;
            if (b1 <= b0) continue;
            const size_t lo = b0 * 64, hi = (b1 * 64 < p.bytes) ? b1 * 64 : p.bytes;
            memcpy((char *)p.dst + lo, (const char *)p.src + lo, hi - lo);
        }
    }
    void loop(int tid) {
        uint64_t seen = 0;
        for (;;) {
            const pp_copy_piece *pieces;
            size_t n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                pieces = pieces_;
                n = n_;
            }
            share(pieces, n, tid);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) cv_done_.notify_all();
            }
        }
    }

    const int T_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_;
    const pp_copy_piece *pieces_ = nullptr;
    size_t n_ = 0;
    uint64_t generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};
