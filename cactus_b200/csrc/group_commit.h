// group_commit.h -- merge the device batches of concurrent callers ("group commit").
//
// The reference enters the BAR code from OpenMP teams (bar/impl/bar.c:90-94: one thread per flower; SURVEY.md 8b
// "Threading"), so several host threads submit device batches at the same time. A device batch is only as efficient as it
// is wide, therefore callers do not queue up behind a mutex: every caller enqueues its request; whoever finds the device
// idle becomes the leader, takes the request at the head of the queue plus every queued request that may share a launch
// with it, runs them as ONE batch, marks them done and steps down. While a batch runs, later callers pile up and form the
// next, wider batch. A single caller degenerates to "run my own request" with no copying and no waiting.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

namespace barb200 {

template <class Req>
class GroupCommit {
public:
    // Req needs `bool done` (false on entry) and `int rc`. can_merge(head, other) -> bool; exec(std::vector<Req *> &batch) runs the
    // batch and stores each request's outcome in the request; if it throws (out of memory), every request of the batch gets
    // rc = -2 (BARB200_ENOMEM) and nothing propagates -- the callers sit behind a C ABI. Returns when r->done.
    template <class CanMerge, class Exec>
    void submit(Req *r, CanMerge can_merge, Exec exec) {
        std::unique_lock<std::mutex> lk(mu_);
        queue_.push_back(r);
        while (!r->done) {
            if (leader_active_) { cv_.wait(lk); continue; }
            leader_active_ = true;
            std::vector<Req *> batch;
            batch.push_back(queue_.front());
            queue_.pop_front();
            for (auto it = queue_.begin(); it != queue_.end();) {
                if (can_merge(*batch.front(), **it)) { batch.push_back(*it); it = queue_.erase(it); } else ++it;
            }
            lk.unlock();
            try { exec(batch); } catch (...) { for (Req *b : batch) b->rc = -2; }
            lk.lock();
            for (Req *b : batch) b->done = true;
            leader_active_ = false;
            cv_.notify_all();
        }
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Req *> queue_;
    bool leader_active_ = false;
};

}  // namespace barb200
