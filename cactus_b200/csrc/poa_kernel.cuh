// poa_kernel.cuh -- device-side arguments of the fused batched POA kernel (see poa_kernel.cu).
#pragma once
#include "poa_types.h"

namespace barb200 {

// Byte offsets of every per-slot array from the slot base (computed once per batch on the host).
struct SlotLayout {
    int64_t slot_bytes;
    int node_cap, in_pool, out_pool, W, cigar_cap, fc_cap;
    int64_t plane_cap;   // ints per slot
    int64_t o_base, o_aln_n, o_aln_id, o_in_off, o_in_n, o_in_cap, o_out_off, o_out_n, o_out_cap;
    int64_t o_in_id, o_in_w, o_out_id, o_out_w, o_out_rid;
    int64_t o_index_to_node, o_node_to_index, o_remain, o_msa_rank, o_tmp0, o_tmp1;
    int64_t o_row_rec, o_pre_row;
    int64_t o_row_off, o_row_info, o_cigar, o_fc;
};

enum { PH_DP = 0, PH_BACKTRACK = 1, PH_FUSE = 2, PH_TOPO = 3, PH_MSA = 4, PH_TOTAL = 5, PH_N = 8 };

struct BatchArgs {
    const JobDesc *jobs;        // all jobs of the stage (internal order)
    int job_base, n_jobs;       // this launch's class: jobs [job_base, job_base + n_jobs)
    const uint8_t *seqs;        // packed 0..4 codes of all jobs
    const int *lens;            // per sequence
    const int64_t *soff;        // per sequence: offset from the job's seq_off
    const int *order;           // per sequence slot a: which read is aligned a-th (guide tree, guide_tree_kernel)
    const int *gt_status;       // per job: nonzero = the guide tree ran out of key space (the job is retried with more)
    uint8_t *msa; int *msa_len; int *status; long long *cells;
    uint8_t *slots; int *planes;
    int *next_job;              // the class's work counter
    unsigned long long *phase_clk;   // [gridDim.x * PH_N] clock64 per phase, or nullptr
    int serial_phases;          // debugging aid: 1 = run the graph phases in their serial reference form
    int bfs_order;              // debugging aid: 1 = recompute abPOA's BFS order after every fusion instead of splicing
    int scratch_bytes;          // dynamic shared memory per CTA (topological sort scratch)
    SlotLayout lay;
    PoaParams P;
};

// ---- K0: guide_tree_kernel (guide_tree.cu): the read order of every job of a stage, one persistent CTA per scratch slot ----
struct GuideTreeArgs {
    const JobDesc *jobs; int n_jobs;
    const uint8_t *seqs; const int *lens; const int64_t *soff;
    int *order;                 // out: per sequence slot, indexed like lens
    int *gt_status;             // out: per job, 0 or JOB_ERR_GT_CAP
    int *next_job;              // work counter
    uint8_t *scratch; int64_t slot_bytes;       // per-CTA scratch
    int64_t o_keys, o_gx, o_hit, o_jac, o_score; int key_cap;
    int k, w;
};
constexpr int kGuideTreeThreads = 256;
constexpr int kGuideTreeTileKeys = 2048;        // shared-memory tile of the sort (16 KB)
void launch_guide_tree(const GuideTreeArgs &A, int ctas, void *stream);

}  // namespace barb200

