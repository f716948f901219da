// Shared host thread pool of libvacmap_b200 (pure C++17, no CUDA).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace vmp {

// Process-wide pool of host threads shared by every Driver / backend (the pipelined workers all draw from
// it, so the host is never oversubscribed).  parallel_for hands out blocks of indices; the caller works too.
class HostPool {
public:
    static HostPool &get()
    {
        // VM_HOST_THREADS: share of the host this process may use (one process per GPU on a multi-GPU box)
        static HostPool pool([] {
            const char *e = getenv("VM_HOST_THREADS");
            const int n = e ? atoi(e) : 0;
            return n > 0 ? n : (int)std::max(1u, std::thread::hardware_concurrency());
        }());
        return pool;
    }
    int size() const { return (int)threads_.size() + 1; }

    void run(int64_t n, int max_threads, int64_t grain, const std::function<void(int64_t)> &fn)
    {
        if (n <= 0) return;
        if (max_threads <= 1 || n <= grain || threads_.empty()) {
            for (int64_t i = 0; i < n; ++i) fn(i);
            return;
        }
        auto job = std::make_shared<Job>();
        job->n = n;
        job->grain = grain;
        job->fn = &fn;
        job->helpers_wanted = (int)std::min<int64_t>(std::min<int64_t>(max_threads - 1, (int64_t)threads_.size()), (n + grain - 1) / grain - 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            queue_.push_back(job);
        }
        cv_.notify_all();
        work(*job);
        std::unique_lock<std::mutex> lk(job->mu);
        job->cv.wait(lk, [&] { return job->active == 0 && job->next.load() >= job->n; });
        if (job->failed) std::rethrow_exception(job->error);
    }

private:
    struct Job {
        int64_t n = 0, grain = 1;
        const std::function<void(int64_t)> *fn = nullptr;
        std::atomic<int64_t> next{0};
        int helpers_wanted = 0, helpers = 0;   // guarded by HostPool::mu_
        int active = 0;                        // guarded by mu
        bool failed = false;
        std::exception_ptr error;
        std::mutex mu;
        std::condition_variable cv;
    };

    explicit HostPool(int n)
    {
        for (int t = 1; t < n; ++t) threads_.emplace_back([this] { loop(); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
    }

    void work(Job &job)
    {
        {
            std::lock_guard<std::mutex> lk(job.mu);
            ++job.active;
        }
        try {
            for (;;) {
                const int64_t i0 = job.next.fetch_add(job.grain);
                if (i0 >= job.n) break;
                const int64_t i1 = std::min(job.n, i0 + job.grain);
                for (int64_t i = i0; i < i1; ++i) (*job.fn)(i);
            }
        } catch (...) {
            std::lock_guard<std::mutex> lk(job.mu);
            if (!job.failed) { job.failed = true; job.error = std::current_exception(); }
            job.next.store(job.n);
        }
        {
            std::lock_guard<std::mutex> lk(job.mu);
            --job.active;
        }
        job.cv.notify_all();
    }

    void loop()
    {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                for (;;) {
                    while (!queue_.empty() && (queue_.front()->next.load() >= queue_.front()->n ||
                                               queue_.front()->helpers >= queue_.front()->helpers_wanted))
                        queue_.pop_front();
                    if (stop_ || !queue_.empty()) break;
                    cv_.wait(lk);
                }
                if (stop_) return;
                job = queue_.front();
                ++job->helpers;
                // spread the pool over the jobs in flight: the next idle thread looks at the next job first
                if (queue_.size() > 1) { queue_.pop_front(); queue_.push_back(job); }
            }
            work(*job);
        }
    }

    std::vector<std::thread> threads_;
    std::deque<std::shared_ptr<Job>> queue_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

static inline void parallel_for(int64_t n, int threads, const std::function<void(int64_t)> &fn, int64_t grain = 16)
{
    HostPool::get().run(n, threads, grain, fn);
}

// out = concatenation of parts[0..m) (moved), start[t] = offset of parts[t] in out; parallel over the parts
template <typename T>
static inline void parallel_concat(std::vector<std::vector<T>> &parts, int threads, std::vector<T> &out, std::vector<int64_t> &start)
{
    const int64_t m = (int64_t)parts.size();
    start.assign((size_t)m + 1, 0);
    for (int64_t t = 0; t < m; ++t) start[t + 1] = start[t] + (int64_t)parts[t].size();
    out.resize((size_t)start[m]);
    parallel_for(m, threads, [&](int64_t t) {
        std::move(parts[t].begin(), parts[t].end(), out.begin() + start[t]);
    }, 64);
}

} // namespace vmp
