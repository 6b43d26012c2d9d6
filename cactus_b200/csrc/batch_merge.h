// batch_merge.h -- the pair-HMM request that goes through GroupCommit (group_commit.h) and how a group of them becomes ONE
// device batch: inputs concatenated (views, nothing is copied but pointers), outputs handed back to their callers.
// (POA callers go through the end queue instead, end_queue.h.)
// The device batch itself is a callable (`impl`), so that tests/hosttest can drive this plumbing with a stand-in on the CPU.
#pragma once
#include <stdint.h>
#include <string.h>
#include <iterator>
#include <vector>
#include "group_commit.h"
#include "host_api.h"

namespace barb200 {

// ---- cPecan: barb200_pecan_aligned_pairs_batch of concurrent callers ---------------------------------------------------------
struct PecanRequest {
    bool done = false;
    barb200_pecan_params p;
    int64_t n = 0;
    const char *const *sx = nullptr; const int64_t *lx = nullptr; const char *const *sy = nullptr; const int64_t *ly = nullptr;
    const int64_t *const *anchors = nullptr; const int64_t *n_anchor = nullptr;
    const uint8_t *ragged_left = nullptr, *ragged_right = nullptr;
    int64_t **triples_out = nullptr; int64_t *n_out = nullptr; double **posteriors_out = nullptr; int64_t *cells_out = nullptr;
    int rc = 0;
};

inline bool pecan_can_merge(const PecanRequest &a, const PecanRequest &b) {
    return a.p.threshold == b.p.threshold && a.p.min_diags_between_traceback == b.p.min_diags_between_traceback &&
           a.p.traceback_diagonals == b.p.traceback_diagonals && a.p.diagonal_expansion == b.p.diagonal_expansion &&
           a.p.split_matrix_bigger_than_this == b.p.split_matrix_bigger_than_this && a.p.dynamic_anchor_expansion == b.p.dynamic_anchor_expansion &&
           (a.posteriors_out != nullptr) == (b.posteriors_out != nullptr);
}

// impl(p, n, sx, lx, sy, ly, anchors, n_anchor, rl, rr, triples_out, n_out, posteriors_out, cells_out) -> int
template <class Impl>
void run_pecan_group(std::vector<PecanRequest *> &batch, Impl impl) {
    if (batch.size() == 1) {
        PecanRequest *q = batch[0];
        q->rc = impl(&q->p, q->n, q->sx, q->lx, q->sy, q->ly, q->anchors, q->n_anchor, q->ragged_left, q->ragged_right, q->triples_out, q->n_out,
                     q->posteriors_out, q->cells_out);
        return;
    }
    int64_t total = 0;
    for (PecanRequest *q : batch) total += q->n;
    const size_t m = (size_t)(total > 0 ? total : 1);
    std::vector<const char *> sx(m), sy(m);
    std::vector<int64_t> lx(m), ly(m), na(m, 0), n_out(m, 0), cells(m, 0);
    std::vector<const int64_t *> an(m, nullptr);
    std::vector<uint8_t> rl(m, 0), rr(m, 0);
    std::vector<int64_t *> trip(m, nullptr);
    std::vector<double *> post(m, nullptr);
    const bool want_post = batch[0]->posteriors_out != nullptr;
    int64_t o = 0;
    for (PecanRequest *q : batch) {
        for (int64_t i = 0; i < q->n; ++i, ++o) {
            sx[o] = q->sx[i]; sy[o] = q->sy[i]; lx[o] = q->lx[i]; ly[o] = q->ly[i];
            na[o] = q->n_anchor ? q->n_anchor[i] : 0;
            an[o] = (q->anchors && na[o]) ? q->anchors[i] : nullptr;
            rl[o] = q->ragged_left ? q->ragged_left[i] : 0; rr[o] = q->ragged_right ? q->ragged_right[i] : 0;
        }
    }
    const int rc = impl(&batch[0]->p, total, sx.data(), lx.data(), sy.data(), ly.data(), an.data(), na.data(), rl.data(), rr.data(), trip.data(),
                        n_out.data(), want_post ? post.data() : nullptr, cells.data());
    if (rc != 0) {       // as above: isolate the failure to the request that caused it
        for (PecanRequest *q : batch)
            q->rc = impl(&q->p, q->n, q->sx, q->lx, q->sy, q->ly, q->anchors, q->n_anchor, q->ragged_left, q->ragged_right, q->triples_out, q->n_out,
                         q->posteriors_out, q->cells_out);
        return;
    }
    o = 0;
    for (PecanRequest *q : batch) {
        q->rc = rc;
        for (int64_t i = 0; i < q->n; ++i, ++o) {
            if (rc != 0) continue;
            q->triples_out[i] = trip[o]; q->n_out[i] = n_out[o];
            if (want_post) q->posteriors_out[i] = post[o];
            if (q->cells_out) q->cells_out[i] = cells[o];
        }
    }
}

}  // namespace barb200
