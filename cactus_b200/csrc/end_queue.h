// end_queue.h -- the global end queue: every end of every flower a caller has submitted, drained by one worker per device
// lane into device batches ("funnel all POA jobs into one async batching queue").
//
// The reference aligns the ends of one flower per OpenMP thread, synchronously (bar/impl/bar.c:90-164 ->
// make_flower_alignment_poa -> one abpoa_msa per window, poaBarAligner.c:609). A device batch is only as efficient as it is wide,
// so here a caller SUBMITS a flower's ends (a ticket) and collects the alignments later; the shim's bar() submits every leaf
// flower before it waits for the first one, and CAF runs on the caller's threads while the lanes work (shim/cactus_bar_shim.c).
// Synchronous callers are submit + wait: whatever several threads have submitted by the time a lane becomes free runs as ONE
// batch (what the group commit of round 1 did for POA is a special case of this queue).
//
// A lane worker takes tickets off the queue up to a job / cost limit, builds the next window of every unfinished end
// (bar_windows.h), runs the batch through `exec` (barb200.cu: run_jobs_on_lane; tests/hosttest: a CPU stand-in), trims the
// windows, and puts tickets with unfinished ends back at the FRONT of the queue. A batch that fails is re-run ticket by ticket,
// so only the offending caller sees the error. Stitching and the cross-end trimming run in the waiting thread.
// No CUDA in this file.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "bar_windows.h"

namespace barb200 {

struct Ticket {
    int64_t n_ends = 0;
    std::vector<barwin::EndState> ends;
    int64_t window_size = 0, max_prog_rows = 0; double max_prog_length_diff = 0; int default_progressive = 1;
    bool consistent = false;                                    // cross-end trimming requested (make_consistent_partial_order_alignments)
    std::vector<std::vector<int64_t>> right_end_indexes, right_end_row_indexes, overlaps;
    // queue state (under EndQueue::mu_)
    int64_t active = 0;                                         // ends that still have windows to align
    bool done = false;
    int rc = 0; std::string err;
    double cost = 0;                                            // rough cells of the next round (batch sizing)
};

class EndQueue {
public:
    // exec(lane, jobs, results, err) -> 0 or a BARB200_E* code (err = message)
    typedef std::function<int(int, const std::vector<HostJob> &, std::vector<JobResult> &, std::string &)> Exec;

    EndQueue(int n_lanes, Exec exec, int64_t max_jobs = 6144, double max_cost = 1.6e11) : exec_(exec), max_jobs_(max_jobs), max_cost_(max_cost) {
        for (int l = 0; l < n_lanes; ++l) workers_.emplace_back([this, l] { lane_loop(l); });
    }
    ~EndQueue() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        work_cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void submit(Ticket *t) {
        t->active = 0; t->cost = 0;
        for (auto &E : t->ends) if (!E.done) { ++t->active; t->cost += end_cost(E, t->window_size); }
        std::lock_guard<std::mutex> lk(mu_);
        if (t->active == 0 || t->rc) { t->done = true; return; }
        pending_.push_back(t);
        pending_jobs_ += t->active;
        last_submit_ = std::chrono::steady_clock::now();
        work_cv_.notify_one();
    }
    void wait(Ticket *t) {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [t] { return t->done; });
    }
    // batches run so far / jobs in them (reports, tests)
    void stats(int64_t *batches, int64_t *jobs) { std::lock_guard<std::mutex> lk(mu_); if (batches) *batches = n_batches_; if (jobs) *jobs = n_jobs_; }

private:
    static double end_cost(const barwin::EndState &E, int64_t window) {
        if (E.seq_no < 2) return 0;
        const double L = (double)std::min<int64_t>(window, E.seq_lens[0] > 0 ? E.seq_lens[0] : 1);
        return (double)(E.seq_no - 1) * L * std::min(L + 1.0, 2.0 * (1000.0 + 0.1 * L));
    }

    void lane_loop(int lane) {
        while (true) {
            std::vector<Ticket *> batch;
            {
                std::unique_lock<std::mutex> lk(mu_);
                work_cv_.wait(lk, [this] { return stop_ || !pending_.empty(); });
                if (pending_.empty()) return;               // stop requested and nothing left
                // a burst of submissions is still arriving (the shim's bar() submits thousands of flowers in a row): let it pile up
                // instead of launching a batch of the first few -- only while tickets keep coming (the last one less than 600 us ago),
                // until a full batch is waiting, and for 20 ms at most
                {
                    const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(20);
                    while (!stop_ && pending_jobs_ < max_jobs_) {
                        const auto now = std::chrono::steady_clock::now();
                        if (now >= t_end || now - last_submit_ >= std::chrono::microseconds(600)) break;
                        work_cv_.wait_for(lk, std::chrono::microseconds(300));
                    }
                }
                if (pending_.empty()) { if (stop_) return; continue; }
                int64_t jobs = 0; double cost = 0;
                while (!pending_.empty()) {
                    Ticket *t = pending_.front();
                    if (!batch.empty() && (jobs + t->active > max_jobs_ || cost + t->cost > max_cost_)) break;
                    pending_.pop_front(); batch.push_back(t); jobs += t->active; cost += t->cost; pending_jobs_ -= t->active;
                }
                if (!pending_.empty()) work_cv_.notify_one();   // more work than one batch: wake another lane
            }
            process(lane, batch);
            {
                std::lock_guard<std::mutex> lk(mu_);
                for (auto it = batch.rbegin(); it != batch.rend(); ++it) {
                    Ticket *t = *it;
                    if (t->rc || t->active == 0) t->done = true; else { pending_.push_front(t); pending_jobs_ += t->active; }
                }
                if (!pending_.empty()) work_cv_.notify_one();
            }
            done_cv_.notify_all();
        }
    }

    // one window round of the unfinished ends of `batch`
    void process(int lane, std::vector<Ticket *> &batch) {
        std::vector<HostJob> jobs; std::vector<std::pair<Ticket *, barwin::EndState *>> owner;
        for (Ticket *t : batch) {
            if (t->rc) continue;
            for (auto &E : t->ends) {
                if (E.done) continue;
                HostJob job;
                const std::string e = barwin::end_prepare_window(E, t->window_size, t->default_progressive, t->max_prog_rows, t->max_prog_length_diff, job);
                if (!e.empty()) { t->rc = BARB200_EINVAL; t->err = e; break; }
                jobs.push_back(job); owner.emplace_back(t, &E);
            }
        }
        // (jobs of tickets that just failed are still in the list; they are run and dropped -- simpler than compacting, and rare)
        if (jobs.empty()) return;
        static const bool trace = getenv("BARB200_TIMING") != nullptr;
        auto ms_now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double tq0 = trace ? ms_now() : 0;
        std::vector<JobResult> res; std::string err;
        int rc = 0;
        for (size_t at = 0; at < jobs.size() && !rc; at += (size_t)max_chunk_) {           // a single huge ticket is cut into chunks
            const size_t end = std::min(jobs.size(), at + (size_t)max_chunk_);
            if (at == 0 && end == jobs.size()) { rc = exec_(lane, jobs, res, err); break; }
            std::vector<HostJob> part(jobs.begin() + at, jobs.begin() + end); std::vector<JobResult> pres;
            rc = exec_(lane, part, pres, err);
            if (!rc) for (auto &r : pres) res.push_back(std::move(r));
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            ++n_batches_; n_jobs_ += (int64_t)jobs.size();
        }
        if (rc) {
            if (batch.size() > 1) {
                // isolate the failure: the windows were prepared already (offsets moved), so the tickets cannot simply be re-queued;
                // run each ticket's prepared jobs on its own
                size_t at = 0;
                for (Ticket *t : batch) {
                    size_t n = 0;
                    while (at + n < owner.size() && owner[at + n].first == t) ++n;
                    if (n && !t->rc) {
                        std::vector<HostJob> part(jobs.begin() + at, jobs.begin() + at + n); std::vector<JobResult> pres; std::string perr;
                        const int prc = exec_(lane, part, pres, perr);
                        if (prc) { t->rc = prc; t->err = perr; }
                        else consume(owner, at, n, pres);
                    }
                    at += n;
                }
            } else if (!batch.empty() && !batch[0]->rc) { batch[0]->rc = rc; batch[0]->err = err; }
            finish_counts(batch);
            return;
        }
        const double tq1 = trace ? ms_now() : 0;
        consume(owner, 0, owner.size(), res);
        finish_counts(batch);
        if (trace) fprintf(stderr, "barb200 timing: queue lane %d: %zu tickets, %zu jobs, start %.1f ms, device batch %.1f ms, trim %.1f ms\n", lane, batch.size(),
                           jobs.size(), fmod(tq0, 100000.0), tq1 - tq0, ms_now() - tq1);
    }

    void consume(std::vector<std::pair<Ticket *, barwin::EndState *>> &owner, size_t at, size_t n, std::vector<JobResult> &res) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t k = 0; k < (int64_t)n; ++k) {
            Ticket *t = owner[at + k].first;
            const std::string e = barwin::end_consume_window(*owner[at + k].second, res[k]);
            if (!e.empty()) {
#pragma omp critical(barb200_end_queue_err)
                { t->rc = BARB200_EINVAL; t->err = e; }
            }
        }
    }
    static void finish_counts(std::vector<Ticket *> &batch) {
        for (Ticket *t : batch) {
            t->active = 0; t->cost = 0;
            for (auto &E : t->ends) if (!E.done) { ++t->active; t->cost += end_cost(E, t->window_size); }
        }
    }

    Exec exec_;
    int64_t max_jobs_, max_chunk_ = 1 << 15, pending_jobs_ = 0; double max_cost_;
    std::chrono::steady_clock::time_point last_submit_ = std::chrono::steady_clock::now();
    std::mutex mu_;
    std::condition_variable work_cv_, done_cv_;
    std::deque<Ticket *> pending_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
    int64_t n_batches_ = 0, n_jobs_ = 0;
};

}  // namespace barb200
