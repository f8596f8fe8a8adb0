// thread_pool.h -- persistent helper threads shared by the host-side batch entry points.
#pragma once

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dg {

// Persistent helper threads: every round hands the same job to all of them (each pulls game indices from an atomic
// counter); creating threads per round would cost more than the round itself.
class Helpers {
  public:
    explicit Helpers(int n) {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~Helpers() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++generation_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void run(const std::function<void()>& job) {          // the caller works too; returns when everybody is done
        if (threads_.empty()) { job(); return; }
        { std::lock_guard<std::mutex> g(m_); job_ = &job; remaining_ = (int)threads_.size(); ++generation_; }
        cv_.notify_all();
        job();
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [this] { return remaining_ == 0; });
        job_ = nullptr;
    }

  private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void()>* job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                job = job_;
            }
            (*job)();
            { std::lock_guard<std::mutex> g(m_); if (--remaining_ == 0) done_cv_.notify_one(); }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void()>* job_ = nullptr;
    uint64_t generation_ = 0;
    int remaining_ = 0;
    bool stop_ = false;
};

}  // namespace dg
