// ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Stand-in for the un-vendored third-party module bshoshany/thread-pool v4.1.0
// (pinned at /root/reference/CMakeLists.txt:16-21). Surface used by the reference
// (/root/reference/src/simulation.cpp:230,244-250): BS::thread_pool(n),
// detach_loop<T>(first, last, fn(index)), wait(). Results in the reference are
// index-addressed, so any correct pool gives identical output.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

namespace BS
{
    class thread_pool
    {
    public:
        explicit thread_pool(std::size_t n_threads = 0)
        {
            if (n_threads == 0)
            {
                n_threads = std::thread::hardware_concurrency();
                if (n_threads == 0)
                    n_threads = 1;
            }
            for (std::size_t i = 0; i < n_threads; ++i)
                workers_.emplace_back([this] { worker(); });
        }
        ~thread_pool()
        {
            wait();
            {
                std::lock_guard<std::mutex> lk(mu_);
                stop_ = true;
            }
            cv_task_.notify_all();
            for (auto &t : workers_)
                t.join();
        }
        std::size_t get_thread_count() const { return workers_.size(); }

        // Splits [first, last) into blocks (one per worker by default) and runs fn(i) for every index.
        template <typename T, typename F>
        void detach_loop(T first, T last, F &&fn, std::size_t num_blocks = 0)
        {
            if (last <= first)
                return;
            const std::size_t total = static_cast<std::size_t>(last - first);
            if (num_blocks == 0)
                num_blocks = workers_.size() * 8; // finer than one block per worker: trials have uneven cost
            if (num_blocks > total)
                num_blocks = total;
            const std::size_t base = total / num_blocks, rem = total % num_blocks;
            T begin = first;
            for (std::size_t b = 0; b < num_blocks; ++b)
            {
                const T end = begin + static_cast<T>(base + (b < rem ? 1 : 0));
                push([fn, begin, end] {
                    for (T i = begin; i < end; ++i)
                        fn(i);
                });
                begin = end;
            }
        }
        void wait()
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [this] { return tasks_.empty() && running_ == 0; });
        }

    private:
        void push(std::function<void()> task)
        {
            {
                std::lock_guard<std::mutex> lk(mu_);
                tasks_.push(std::move(task));
            }
            cv_task_.notify_one();
        }
        void worker()
        {
            for (;;)
            {
                std::function<void()> task;
                {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_task_.wait(lk, [this] { return stop_ || !tasks_.empty(); });
                    if (stop_ && tasks_.empty())
                        return;
                    task = std::move(tasks_.front());
                    tasks_.pop();
                    ++running_;
                }
                task();
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    --running_;
                    if (tasks_.empty() && running_ == 0)
                        cv_done_.notify_all();
                }
            }
        }
        std::vector<std::thread> workers_;
        std::queue<std::function<void()>> tasks_;
        std::mutex mu_;
        std::condition_variable cv_task_, cv_done_;
        std::size_t running_ = 0;
        bool stop_ = false;
    };
}
