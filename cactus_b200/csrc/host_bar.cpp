// host_bar.cpp -- the MSA-level entry points of the C ABI (include/barb200.h) on top of the end queue:
//
//   barb200_flower_submit / barb200_flower_wait      asynchronous: one ticket = the ends of one flower (or any set of ends)
//   barb200_msa_make_partial_order_alignment[_batch] msa_make_partial_order_alignment, bar/impl/poaBarAligner.c:463-749
//   barb200_make_consistent_partial_order_alignments make_consistent_partial_order_alignments, :751-801
//
// The synchronous calls are submit + wait. The window / trimming / stitching logic lives in bar_windows.h, the queue and its
// lane workers in end_queue.h (both free of CUDA, so that tests/hosttest runs them on the CPU with a stand-in device); this file
// binds them to the device lanes of barb200.cu.
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <mutex>
#include <vector>
#include "host_api.h"
#include "end_queue.h"

namespace barb200 {

static std::mutex g_queue_create_mu;

static EndQueue *queue_of(barb200_ctx *ctx) {
    void **slot = dispatcher_slot(ctx);
    std::lock_guard<std::mutex> lk(g_queue_create_mu);
    if (!*slot) {
        mark_lanes_shared(ctx);                 // from now on every lane plans with its share of the device memory
        int64_t max_jobs = 6144; double max_cost = 8e10;
        if (getenv("BARB200_QUEUE_MAX_JOBS")) max_jobs = atoll(getenv("BARB200_QUEUE_MAX_JOBS"));
        if (getenv("BARB200_QUEUE_MAX_COST")) max_cost = atof(getenv("BARB200_QUEUE_MAX_COST"));
        *slot = new EndQueue(total_lanes(ctx), [ctx](int lane, const std::vector<HostJob> &jobs, std::vector<JobResult> &res, std::string &err) {
            const int rc = run_jobs_on_lane(ctx, lane, jobs, res);
            if (rc) err = get_error(ctx);
            return rc;
        }, max_jobs, max_cost);
    }
    return (EndQueue *)*slot;
}

void dispatcher_destroy(barb200_ctx *ctx) {
    void **slot = dispatcher_slot(ctx);
    std::lock_guard<std::mutex> lk(g_queue_create_mu);
    if (*slot) { delete (EndQueue *)*slot; *slot = nullptr; }
}

}  // namespace barb200

using namespace barb200;

struct barb200_ticket { Ticket t; };

extern "C" void barb200_msa_destruct(barb200_msa *m) { if (m) { free(m->seq_lens); free(m->msa); free(m); } }

extern "C" barb200_ticket *barb200_flower_submit(barb200_ctx *ctx, int64_t end_no, const int64_t *end_lengths, char ***end_strings,
        int **end_string_lengths, int64_t **right_end_indexes, int64_t **right_end_row_indexes, int64_t **overlaps,
        int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff) {
    if (!ctx) return nullptr;
    if (end_no < 0 || (end_no > 0 && (!end_lengths || !end_strings || !end_string_lengths))) { set_error(ctx, "bad arguments"); return nullptr; }
    if (window_size <= 0) { set_error(ctx, "window_size must be positive"); return nullptr; }
    std::unique_ptr<barb200_ticket> bt(new barb200_ticket());
    Ticket &t = bt->t;
    t.n_ends = end_no; t.window_size = window_size; t.max_prog_rows = max_prog_rows; t.max_prog_length_diff = max_prog_length_diff;
    t.default_progressive = default_progressive(ctx);
    t.ends.resize((size_t)end_no);
    for (int64_t e = 0; e < end_no; ++e) {
        const std::string err = barwin::end_init(t.ends[e], end_lengths[e], end_strings[e], end_string_lengths[e]);
        if (!err.empty()) { set_error(ctx, err); return nullptr; }
    }
    t.consistent = right_end_indexes != nullptr;
    if (t.consistent) {
        if (!right_end_row_indexes || !overlaps) { set_error(ctx, "bad arguments"); return nullptr; }
        t.right_end_indexes.resize(end_no); t.right_end_row_indexes.resize(end_no); t.overlaps.resize(end_no);
        for (int64_t e = 0; e < end_no; ++e) {
            t.right_end_indexes[e].assign(right_end_indexes[e], right_end_indexes[e] + end_lengths[e]);
            t.right_end_row_indexes[e].assign(right_end_row_indexes[e], right_end_row_indexes[e] + end_lengths[e]);
            t.overlaps[e].assign(overlaps[e], overlaps[e] + end_lengths[e]);
        }
    }
    queue_of(ctx)->submit(&t);
    return bt.release();
}

extern "C" barb200_msa **barb200_flower_wait(barb200_ctx *ctx, barb200_ticket *bt) {
    if (!ctx || !bt) return nullptr;
    std::unique_ptr<barb200_ticket> own(bt);
    Ticket &t = bt->t;
    queue_of(ctx)->wait(&t);
    if (t.rc) { set_error(ctx, t.err.empty() ? "device batch failed" : t.err); return nullptr; }
    barb200_msa **msas = (barb200_msa **)calloc((size_t)(t.n_ends > 0 ? t.n_ends : 1), sizeof(barb200_msa *));
    if (!msas) { set_error(ctx, "host allocation failed"); return nullptr; }
    bool ok = true;
    for (int64_t e = 0; e < t.n_ends && ok; ++e) { msas[e] = barwin::end_stitch(t.ends[e]); ok = msas[e] != nullptr; }
    if (!ok) set_error(ctx, "host allocation failed");
    if (ok && t.consistent && !barwin::consistency_trim(t.n_ends, msas, t.right_end_indexes, t.right_end_row_indexes, t.overlaps)) {
        set_error(ctx, "inconsistent end / row indexes or overlaps"); ok = false;
    }
    if (!ok) { for (int64_t e = 0; e < t.n_ends; ++e) barb200_msa_destruct(msas[e]); free(msas); return nullptr; }
    return msas;
}

extern "C" int barb200_msa_make_partial_order_alignment_batch(barb200_ctx *ctx, int64_t n_ends, const int64_t *seq_no,
        char ***seqs, int **seq_lens, int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff,
        barb200_msa **out) {
    if (!ctx || n_ends < 0 || (n_ends > 0 && (!seq_no || !seqs || !seq_lens || !out))) return BARB200_EINVAL;
    for (int64_t e = 0; e < n_ends; ++e) out[e] = nullptr;
    barb200_ticket *t = barb200_flower_submit(ctx, n_ends, seq_no, seqs, seq_lens, nullptr, nullptr, nullptr, window_size, max_prog_rows, max_prog_length_diff);
    if (!t) return BARB200_EINVAL;
    barb200_msa **ms = barb200_flower_wait(ctx, t);
    if (!ms) return BARB200_EJOB;
    for (int64_t e = 0; e < n_ends; ++e) out[e] = ms[e];
    free(ms);
    return BARB200_OK;
}

extern "C" barb200_msa *barb200_msa_make_partial_order_alignment(barb200_ctx *ctx, char **seqs, int *seq_lens, int64_t seq_no,
        int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff) {
    barb200_msa *m = nullptr;
    char **s1[1] = {seqs}; int *l1[1] = {seq_lens};
    if (barb200_msa_make_partial_order_alignment_batch(ctx, 1, &seq_no, s1, l1, window_size, max_prog_rows, max_prog_length_diff, &m)) return nullptr;
    return m;
}

extern "C" barb200_msa **barb200_make_consistent_partial_order_alignments(barb200_ctx *ctx, int64_t end_no, int64_t *end_lengths,
        char ***end_strings, int **end_string_lengths, int64_t **right_end_indexes, int64_t **right_end_row_indexes,
        int64_t **overlaps, int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff) {
    if (!ctx || end_no < 0) return nullptr;
    if (end_no > 0 && (!right_end_indexes || !right_end_row_indexes || !overlaps)) { set_error(ctx, "bad arguments"); return nullptr; }
    barb200_ticket *t = barb200_flower_submit(ctx, end_no, end_lengths, end_strings, end_string_lengths, right_end_indexes, right_end_row_indexes,
                                              overlaps, window_size, max_prog_rows, max_prog_length_diff);
    if (!t) return nullptr;
    return barb200_flower_wait(ctx, t);
}

extern "C" int barb200_queue_stats(barb200_ctx *ctx, int64_t *batches, int64_t *jobs) {
    if (!ctx) return BARB200_EINVAL;
    queue_of(ctx)->stats(batches, jobs);
    return BARB200_OK;
}
