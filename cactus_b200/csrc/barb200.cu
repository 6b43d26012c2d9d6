// barb200.cu -- host orchestration + C ABI of libbarb200.so (see include/barb200.h).
//
// A *stage* packs a set of POA jobs (one job = one abpoa_msa call of the reference), computes their guide-tree
// orders on host threads, sizes the per-CTA device slots for the set, uploads everything once and then launches the
// persistent fused kernel (poa_kernel.cu) with one CTA per slot; CTAs pull jobs from a device-side counter.
// Jobs that outgrow the optimistic slot sizing (DP planes / MSA columns) come back flagged and are re-run in a
// second, worst-case-sized launch. No CPU fallback exists: if CUDA is unavailable, creation fails.
#include <cuda_runtime.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#include "host_api.h"
#include "batch_merge.h"
#include "poa_kernel.cuh"

namespace barb200 {
extern "C" __global__ void poa_msa_kernel_t32(const BatchArgs A);
extern "C" __global__ void poa_msa_kernel_t64(const BatchArgs A);
extern "C" __global__ void poa_msa_kernel_t128(const BatchArgs A);
extern "C" __global__ void poa_msa_kernel_t256(const BatchArgs A);
extern "C" __global__ void poa_msa_kernel_t640(const BatchArgs A);
extern "C" __global__ void poa_msa_kernel_t1024(const BatchArgs A);
}
typedef void (*poa_kernel_fn)(const barb200::BatchArgs);
// CTA-size classes: a CTA of T threads sweeps rows of up to 16*T columns (query length + 1)
// scratch = dynamic shared memory per CTA for the topological sort, sized so that the class's CTAs per SM still fit
static const struct { int T; poa_kernel_fn fn; int scratch; } kKernels[] = {
    {32, barb200::poa_msa_kernel_t32, 10 * 1024}, {64, barb200::poa_msa_kernel_t64, 24 * 1024}, {128, barb200::poa_msa_kernel_t128, 48 * 1024},
    {256, barb200::poa_msa_kernel_t256, 96 * 1024}, {640, barb200::poa_msa_kernel_t640, 200 * 1024}, {1024, barb200::poa_msa_kernel_t1024, 200 * 1024}};
static const int kNumKernels = 6;
using namespace barb200;

struct barb200_ctx {
    barb200_params p;
    PoaParams P;
    HostParams hp;
    int device = 0, sm_count = 0;
    size_t smem_optin = 0;
    std::mutex mu;                      // serialises device batches (one kernel owns the slot arena at a time)
    std::mutex err_mu, cache_mu;
    std::string err;
    cudaStream_t copy_stream = nullptr; // uploads of the NEXT chunk while a kernel runs
    // two slot arenas with a stream each: chunk k+1's kernel is queued (and its CTAs move in as chunk k's retire) while
    // chunk k's results are downloaded and unpacked
    struct Arena { uint8_t *d_slots = nullptr; size_t slots_bytes = 0; int *d_planes = nullptr; size_t planes_bytes = 0; cudaStream_t stream = nullptr; } ar[2];
    // grow-only cache of device blocks for the per-stage buffers (cudaMalloc / cudaFree per call cost milliseconds
    // and cudaFree synchronises the device, which would stall the upload / kernel overlap)
    std::vector<std::pair<void *, size_t>> free_blocks; size_t cached_bytes = 0;
    uint8_t *h_pinned = nullptr; size_t h_pinned_bytes = 0;   // pinned staging buffer for the MSA download
    int *h_ready = nullptr; unsigned ready_slot = 0;          // pinned ring of "jobs released" values (copied to the device by the copy engine)
    unsigned long long *d_clk = nullptr; size_t clk_entries = 0;
    GroupCommit<PoaRequest> poa_group;      // concurrent run_jobs callers share device batches (group_commit.h)
    GroupCommit<PecanRequest> pecan_group;  // likewise for barb200_pecan_aligned_pairs_batch
    void *pecan_scratch = nullptr; size_t pecan_scratch_bytes = 0;   // pecan.cu's batch call (grow-only)
    void *pecan_pinned[2] = {nullptr, nullptr}; size_t pecan_pinned_bytes[2] = {0, 0};   // pinned staging: 0 upload, 1 download
};

namespace barb200 {
// the message is kept twice: in the context (for callers whose request ran inside another thread's merged batch) and per
// thread (concurrent callers do not overwrite each other's text)
static thread_local std::string tls_err;
static thread_local const barb200_ctx *tls_err_ctx = nullptr;
void set_error(barb200_ctx *ctx, const std::string &msg) {
    if (!ctx) return;
    tls_err = msg; tls_err_ctx = ctx;
    std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = msg;
}
// threads for host-side work: the OpenMP default capped by the cgroup CPU quota (containers on big hosts often see
// all logical CPUs but may only use a few; oversubscribing them slows the packing / guide-tree loops down)
static int usable_host_threads() {
    int n = omp_get_max_threads();
    FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char q[64]; long long period = 0;
        if (fscanf(f, "%63s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            const long long quota = atoll(q);
            const int lim = (int)((quota + period / 2) / period);
            if (lim >= 1 && lim < n) n = lim;
        }
        fclose(f);
    }
    return n;
}
int host_threads(barb200_ctx *ctx) {
    static const int dflt = usable_host_threads();
    return ctx->p.host_threads > 0 ? ctx->p.host_threads : dflt;
}
int default_progressive(barb200_ctx *ctx) { return ctx->p.progressive_poa; }
std::mutex &device_mutex(barb200_ctx *ctx) { return ctx->mu; }
int ctx_device(barb200_ctx *ctx) { return ctx->device; }
int ctx_sm_count(barb200_ctx *ctx) { return ctx->sm_count; }
double ctx_mem_fraction(barb200_ctx *ctx) { return ctx->p.mem_fraction > 0 ? ctx->p.mem_fraction : 0.85; }
}

static cudaError_t ctx_alloc(barb200_ctx *ctx, void **p, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
    {
        std::lock_guard<std::mutex> lk(ctx->cache_mu);
        int best = -1;
        for (size_t i = 0; i < ctx->free_blocks.size(); ++i)
            if (ctx->free_blocks[i].second >= bytes && ctx->free_blocks[i].second <= 2 * bytes + (1 << 20) &&
                (best < 0 || ctx->free_blocks[i].second < ctx->free_blocks[best].second)) best = (int)i;
        if (best >= 0) { *p = ctx->free_blocks[best].first; ctx->cached_bytes -= ctx->free_blocks[best].second; ctx->free_blocks.erase(ctx->free_blocks.begin() + best); return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {       // give the cache back and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(ctx->cache_mu);
        for (auto &b : ctx->free_blocks) cudaFree(b.first);
        ctx->free_blocks.clear(); ctx->cached_bytes = 0;
        e = cudaMalloc(p, bytes);
    }
    return e;
}
static size_t block_size_of(size_t bytes) { return (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255; }
static void ctx_free(barb200_ctx *ctx, void *p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(ctx->cache_mu);
    if (ctx->cached_bytes + block_size_of(bytes) > ((size_t)4 << 30) || ctx->free_blocks.size() >= 64) { cudaFree(p); return; }
    ctx->free_blocks.emplace_back(p, block_size_of(bytes)); ctx->cached_bytes += block_size_of(bytes);
}

namespace barb200 {
int device_alloc(barb200_ctx *ctx, void **p, size_t bytes) { return ctx_alloc(ctx, p, bytes) == cudaSuccess ? 0 : -1; }
void device_free(barb200_ctx *ctx, void *p, size_t bytes) { ctx_free(ctx, p, bytes); }
// grow-only scratch of the pair-HMM batch call (tens of GB of rings: cudaMalloc of that size costs ~0.2 s per call)
void *pecan_scratch(barb200_ctx *ctx, size_t bytes) {
    if (ctx->pecan_scratch_bytes >= bytes && ctx->pecan_scratch) return ctx->pecan_scratch;
    if (ctx->pecan_scratch) { cudaFree(ctx->pecan_scratch); ctx->pecan_scratch = nullptr; ctx->pecan_scratch_bytes = 0; }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->pecan_scratch = p; ctx->pecan_scratch_bytes = bytes;
    return p;
}
// grow-only pinned staging buffers of the pair-HMM batch call (which: 0 upload, 1 download); nullptr on failure
void *pecan_pinned(barb200_ctx *ctx, int which, size_t bytes) {
    if (ctx->pecan_pinned_bytes[which] >= bytes && ctx->pecan_pinned[which]) return ctx->pecan_pinned[which];
    if (ctx->pecan_pinned[which]) { cudaFreeHost(ctx->pecan_pinned[which]); ctx->pecan_pinned[which] = nullptr; ctx->pecan_pinned_bytes[which] = 0; }
    void *p = nullptr;
    bytes += bytes / 4;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->pecan_pinned[which] = p; ctx->pecan_pinned_bytes[which] = bytes;
    return p;
}
}  // namespace barb200

#define CUDA_TRY(ctx, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { \
    set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(_e)); return BARB200_ECUDA; } } while (0)

extern "C" void barb200_params_default(barb200_params *p) {
    static const int mat[25] = {91, -114, -61, -123, -100, -114, 100, -125, -61, -100, -61, -125, 100, -114, -100,
                                -123, -61, -114, 91, -100, -100, -100, -100, -100, 100};
    memset(p, 0, sizeof(*p));
    memcpy(p->mat, mat, sizeof(mat));
    p->gap_open1 = 400; p->gap_ext1 = 30; p->gap_open2 = 1200; p->gap_ext2 = 1;
    p->wb = 1000; p->wf = 0.1f;
    p->k = 15; p->w = 5; p->min_w = 500;
    p->progressive_poa = 1; p->disable_seeding = 1;
    p->device = 0; p->threads_per_block = 0; p->ctas_per_sm = 0; p->mem_fraction = 0.0; p->host_threads = 0;
    p->collect_phase_clocks = 0;
}

static void fail(char *errbuf, int n, const std::string &m) { if (errbuf && n > 0) { snprintf(errbuf, n, "%s", m.c_str()); } }

extern "C" barb200_ctx *barb200_create(const barb200_params *p, char *errbuf, int errbuf_len) {
    if (!p) { fail(errbuf, errbuf_len, "null params"); return nullptr; }
    if (p->gap_open1 <= 0 || p->gap_open2 <= 0 || p->gap_ext1 < 0 || p->gap_ext2 < 0 || (p->gap_ext1 == 0 && p->gap_ext2 == 0)) {
        fail(errbuf, errbuf_len, "only the convex gap model (both gap opens > 0) is supported; that is what Cactus configures"); return nullptr; }
    if (p->wb < 0) { fail(errbuf, errbuf_len, "partialOrderAlignmentBandConstant must be >= 0 (adaptive band)"); return nullptr; }
    if ((int64_t)p->gap_open1 + p->gap_ext1 >= 65535 || (int64_t)p->gap_open2 + p->gap_ext2 >= 65535) {
        fail(errbuf, errbuf_len, "gap open + extension must be below 65535 (the traceback planes keep E as a 16-bit distance below H)"); return nullptr; }
    if (!p->disable_seeding) { fail(errbuf, errbuf_len, "minimizer seeding (partialOrderAlignmentDisableSeeding=0) is not supported"); return nullptr; }
    if (p->k <= 0 || p->k > 28 || p->w <= 0 || p->w >= 256) { fail(errbuf, errbuf_len, "minimizer k must be in 1..28 and w in 1..255"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { fail(errbuf, errbuf_len, std::string("no CUDA device: ") + cudaGetErrorString(e)); return nullptr; }
    if (p->device < 0 || p->device >= ndev) { fail(errbuf, errbuf_len, "device ordinal out of range"); return nullptr; }
    barb200_ctx *ctx = new barb200_ctx();
    ctx->p = *p; ctx->device = p->device;
    if (cudaSetDevice(p->device) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaSetDevice failed"); delete ctx; return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaGetDeviceProperties failed"); delete ctx; return nullptr; }
    ctx->sm_count = prop.multiProcessorCount; ctx->smem_optin = prop.sharedMemPerBlockOptin;
    PoaParams &P = ctx->P;
    memcpy(P.mat, p->mat, sizeof(P.mat));
    P.o1 = p->gap_open1; P.e1 = p->gap_ext1; P.o2 = p->gap_open2; P.e2 = p->gap_ext2; P.wb = p->wb; P.wf = p->wf;
    P.max_mat = 0; P.min_mis = 0;
    for (int i = 0; i < 25; ++i) { P.max_mat = std::max(P.max_mat, P.mat[i]); P.min_mis = std::max(P.min_mis, -P.mat[i]); }
    // the reference's int32 minus infinity, abpoa_align_simd.c:1299
    const int oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    P.inf_min = std::max(std::max(INT32_MIN + P.min_mis, INT32_MIN + oe1), INT32_MIN + oe2) + 512 * std::max(P.e1, P.e2);
    ctx->hp = HostParams{p->k, p->w, p->min_w, p->progressive_poa};
    for (int i = 0; i < kNumKernels; ++i) cudaFuncSetAttribute(kKernels[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kKernels[i].scratch);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaStreamCreate failed"); delete ctx; return nullptr; }
    if (cudaMallocHost((void **)&ctx->h_ready, 1024 * sizeof(int)) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaMallocHost failed"); delete ctx; return nullptr; }
    for (int a = 0; a < 2; ++a)
        if (cudaStreamCreateWithFlags(&ctx->ar[a].stream, cudaStreamNonBlocking) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaStreamCreate failed"); delete ctx; return nullptr; }
    return ctx;
}

extern "C" void barb200_destroy(barb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int a = 0; a < 2; ++a) {
        if (ctx->ar[a].d_slots) cudaFree(ctx->ar[a].d_slots);
        if (ctx->ar[a].d_planes) cudaFree(ctx->ar[a].d_planes);
        if (ctx->ar[a].stream) cudaStreamDestroy(ctx->ar[a].stream);
    }
    if (ctx->d_clk) cudaFree(ctx->d_clk);
    if (ctx->pecan_scratch) cudaFree(ctx->pecan_scratch);
    for (int i = 0; i < 2; ++i) if (ctx->pecan_pinned[i]) cudaFreeHost(ctx->pecan_pinned[i]);
    for (auto &b : ctx->free_blocks) cudaFree(b.first);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_ready) cudaFreeHost(ctx->h_ready);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

// The returned pointer stays valid until the calling thread's next engine call (thread-local copy, taken under the lock).
extern "C" const char *barb200_last_error(barb200_ctx *ctx) {
    if (!ctx) return "null context";
    static thread_local std::string out;
    if (barb200::tls_err_ctx == ctx && !barb200::tls_err.empty()) out = barb200::tls_err;
    else { std::lock_guard<std::mutex> lk(ctx->err_mu); out = ctx->err; }
    return out.c_str();
}
extern "C" void barb200_free(void *p) { free(p); }

extern "C" int barb200_device_info(barb200_ctx *ctx, int *sm_count, int64_t *mem_total, int64_t *mem_free, char *name, int name_len) {
    if (!ctx) return BARB200_EINVAL;
    cudaSetDevice(ctx->device);
    cudaDeviceProp prop; CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    size_t f = 0, t = 0; CUDA_TRY(ctx, cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (mem_total) *mem_total = (int64_t)t;
    if (mem_free) *mem_free = (int64_t)f;
    if (name && name_len > 0) snprintf(name, name_len, "%s", prop.name);
    return BARB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stage
// ---------------------------------------------------------------------------------------------------------
struct barb200_stage {
    barb200_ctx *ctx = nullptr;
    int64_t n_jobs = 0, n_seqs = 0, n_bases = 0;
    std::vector<int> n_seq, lens, order, progressive;
    std::vector<int64_t> soff, job_len_off, job_seq_off, job_sum_len;
    std::vector<int> job_max_len;
    std::vector<JobDesc> desc;
    int64_t msa_bytes = 0;
    bool worst_case = false;
    // device: one block from the context's cache holds all per-stage arrays
    void *d_block = nullptr; size_t d_block_bytes = 0;
    uint8_t *d_seqs = nullptr, *d_msa = nullptr; int *d_lens = nullptr, *d_order = nullptr; int64_t *d_soff = nullptr;
    JobDesc *d_desc = nullptr; int *d_msa_len = nullptr, *d_status = nullptr, *d_next = nullptr; long long *d_cells = nullptr;
    // sizing
    SlotLayout lay; int T = 0, slots = 0, kernel_class = 0; size_t dyn_smem = 0;
    // results of the last run
    std::vector<int> status, msa_len; std::vector<long long> cells;
    barb200_stage *retry = nullptr; std::vector<int64_t> retry_jobs;
    int64_t launches = 0; bool ran = false;
    int arena = 0; double mem_share = 1.0; cudaEvent_t e0 = nullptr, e1 = nullptr; bool launched = false;
    int *d_ready = nullptr; const uint8_t *host_seqs = nullptr; int64_t orders_done = 0;   // streamed guide trees (see stage_stream_orders)
    uint64_t clk[6] = {0, 0, 0, 0, 0, 0};
};

static void stage_free_device(barb200_stage *st) {
    ctx_free(st->ctx, st->d_block, st->d_block_bytes);
    st->d_block = nullptr; st->d_block_bytes = 0;
    st->d_seqs = st->d_msa = nullptr; st->d_lens = st->d_order = nullptr; st->d_soff = nullptr; st->d_desc = nullptr;
    st->d_msa_len = st->d_status = st->d_next = nullptr; st->d_cells = nullptr;
}

extern "C" void barb200_stage_destroy(barb200_stage *st) {
    if (!st) return;
    cudaSetDevice(st->ctx->device);
    if (st->retry) barb200_stage_destroy(st->retry);
    if (st->e0) { cudaEventDestroy(st->e0); cudaEventDestroy(st->e1); }
    stage_free_device(st);
    delete st;
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Decide slot sizes, threads per block, shared memory and the number of resident CTAs for this set of jobs.
static int plan_stage(barb200_stage *st) {
    barb200_ctx *ctx = st->ctx;
    int64_t max_nodes = 4, max_edges = 4, max_len = 1, max_k = 1;
    int64_t plane_need = 0;
    for (int64_t j = 0; j < st->n_jobs; ++j) {
        const int64_t sum = st->job_sum_len[j], ml = st->job_max_len[j], K = st->n_seq[j];
        max_nodes = std::max(max_nodes, sum + 2); max_edges = std::max(max_edges, sum + K); max_len = std::max(max_len, ml);
        max_k = std::max(max_k, K);
        // rows the graph can reach: worst case every base a new node; optimistic: the longest read plus a share of the rest
        int64_t rows = st->worst_case ? sum + 2 : std::min<int64_t>(sum + 2, ml + (sum - ml) / 8 + 256);
        plane_need = std::max(plane_need, rows * (TB / CPT) * (align_up(ml + 1, CPT) + CPT));
    }
    SlotLayout &Y = st->lay;
    memset(&Y, 0, sizeof(Y));
    Y.node_cap = (int)max_nodes; Y.in_pool = (int)(4 * max_edges + 64); Y.out_pool = Y.in_pool;
    Y.W = (int)(1 + ((max_k - 1) >> 6)); Y.cigar_cap = (int)(max_len + max_nodes + 16);
    Y.plane_cap = align_up(plane_need, 4);
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 16); return r; };
    const int64_t N = Y.node_cap;
    Y.o_base = take(N); Y.o_aln_n = take(N); Y.o_aln_id = take(N * 16);
    Y.o_in_off = take(N * 4); Y.o_in_n = take(N * 4); Y.o_in_cap = take(N * 4);
    Y.o_out_off = take(N * 4); Y.o_out_n = take(N * 4); Y.o_out_cap = take(N * 4);
    Y.o_in_id = take((int64_t)Y.in_pool * 4); Y.o_in_w = take((int64_t)Y.in_pool * 4);
    Y.o_out_id = take((int64_t)Y.out_pool * 4); Y.o_out_w = take((int64_t)Y.out_pool * 4);
    Y.o_out_rid = take((int64_t)Y.out_pool * 8 * Y.W);
    Y.o_index_to_node = take(N * 4); Y.o_node_to_index = take(N * 4); Y.o_remain = take(N * 4); Y.o_msa_rank = take(N * 4);
    Y.o_tmp0 = take(N * 4); Y.o_tmp1 = take(N * 4);
    Y.o_row_rec = take(N * 16); Y.o_pre_row = take((int64_t)Y.in_pool * 4);
    Y.o_row_off = take(N * 8); Y.o_row_info = take(N * 16);
    Y.o_cigar = take((int64_t)Y.cigar_cap * 8);
    Y.fc_cap = (int)(max_len + 2); Y.o_fc = take((int64_t)Y.fc_cap * 8);
    Y.slot_bytes = align_up(o, 256);

    // smallest CTA-size class whose 16 columns per thread cover the longest query (+ column 0)
    int cls = -1;
    for (int i = 0; i < kNumKernels; ++i) if ((int64_t)kKernels[i].T * CPT >= max_len + 1 && kKernels[i].T >= ctx->p.threads_per_block) { cls = i; break; }
    if (cls < 0) { set_error(ctx, "a sequence is longer than the device engine's row limit (16383 bases per window)"); return BARB200_EINVAL; }
    st->T = kKernels[cls].T; st->kernel_class = cls; st->dyn_smem = kKernels[cls].scratch;
    if (getenv("BARB200_SCRATCH_KB")) st->dyn_smem = (size_t)atoi(getenv("BARB200_SCRATCH_KB")) * 1024;   // tuning aid
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kKernels[cls].fn, st->T, st->dyn_smem));
    if (per_sm < 1) { set_error(ctx, "kernel does not fit on an SM with the requested configuration"); return BARB200_EINVAL; }
    if (ctx->p.ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->p.ctas_per_sm);
    int64_t slots = std::min<int64_t>(st->n_jobs, (int64_t)per_sm * ctx->sm_count);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
    const double frac = ctx->p.mem_fraction > 0 ? ctx->p.mem_fraction : 0.85;
    const barb200_ctx::Arena &AR = ctx->ar[st->arena];
    const double budget = (double)(free_b + AR.slots_bytes + AR.planes_bytes) * frac * st->mem_share;
    const double per_slot = (double)Y.slot_bytes + (double)Y.plane_cap * 4.0;
    if (per_slot > budget) { set_error(ctx, "a single job needs more device memory than is available"); return BARB200_ENOMEM; }
    slots = std::max<int64_t>(1, std::min<int64_t>(slots, (int64_t)(budget / per_slot)));
    st->slots = (int)slots;
    return BARB200_OK;
}

static int ensure_arena(barb200_ctx *ctx, int a, size_t slots_bytes, size_t planes_bytes, size_t clk_entries) {
    barb200_ctx::Arena &AR = ctx->ar[a];
    if (slots_bytes > AR.slots_bytes) {
        if (AR.d_slots) cudaFree(AR.d_slots);
        AR.d_slots = nullptr; AR.slots_bytes = 0;
        if (cudaMalloc(&AR.d_slots, slots_bytes) != cudaSuccess) { cudaGetLastError(); set_error(ctx, "cudaMalloc(slots) failed"); return BARB200_ENOMEM; }
        AR.slots_bytes = slots_bytes;
    }
    if (planes_bytes > AR.planes_bytes) {
        if (AR.d_planes) cudaFree(AR.d_planes);
        AR.d_planes = nullptr; AR.planes_bytes = 0;
        if (cudaMalloc(&AR.d_planes, planes_bytes) != cudaSuccess) { cudaGetLastError(); set_error(ctx, "cudaMalloc(planes) failed"); return BARB200_ENOMEM; }
        AR.planes_bytes = planes_bytes;
    }
    if (clk_entries > ctx->clk_entries) {
        if (ctx->d_clk) cudaFree(ctx->d_clk);
        ctx->d_clk = nullptr; ctx->clk_entries = 0;
        if (cudaMalloc(&ctx->d_clk, clk_entries * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); set_error(ctx, "cudaMalloc(clk) failed"); return BARB200_ENOMEM; }
        ctx->clk_entries = clk_entries;
    }
    return BARB200_OK;
}

// tell the device that jobs [0, n) may start: a 4-byte copy from a pinned ring by the COPY ENGINE (a kernel could not
// be used for this: the persistent POA kernel owns every SM while its CTAs wait for the value)
static cudaError_t post_ready(barb200_stage *st, int64_t n) {
    barb200_ctx *ctx = st->ctx;
    if (ctx->ready_slot && (ctx->ready_slot & 1023) == 0) { cudaError_t e = cudaStreamSynchronize(ctx->copy_stream); if (e != cudaSuccess) return e; }
    int *slot = ctx->h_ready + (ctx->ready_slot++ & 1023);
    *slot = (int)n;
    return cudaMemcpyAsync(st->d_ready, slot, 4, cudaMemcpyHostToDevice, ctx->copy_stream);
}

// guide-tree orders (abpoa_seed.c:85-156, 232-325 via guide_tree.cpp) of jobs [j0, j1) into st->order
static void host_orders(barb200_stage *st, int64_t j0, int64_t j1) {
    barb200_ctx *ctx = st->ctx;
    const int nthreads = host_threads(ctx);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t j = j0; j < j1; ++j) {
        const int K = st->n_seq[j];
        std::vector<const uint8_t *> ptr(K);
        const uint8_t *base = st->host_seqs + st->job_seq_off[j];
        for (int i = 0; i < K; ++i) ptr[i] = base + st->soff[st->job_len_off[j] + i];
        guide_tree_order(ctx->hp, st->progressive[j], K, ptr.data(), st->lens.data() + st->job_len_off[j], st->order.data() + st->job_len_off[j]);
    }
}

static int stage_build(barb200_ctx *ctx, int64_t n_jobs, const int *n_seq, const int *seq_lens, const uint8_t *seqs,
                       const int *progressive, bool worst_case, barb200_stage **out, int arena = 0, double mem_share = 1.0,
                       bool defer_orders = false) {
    if (!ctx || n_jobs < 0 || (n_jobs > 0 && (!n_seq || !seq_lens || !seqs))) { set_error(ctx, "bad arguments"); return BARB200_EINVAL; }
    cudaSetDevice(ctx->device);
    barb200_stage *st = new barb200_stage();
    st->ctx = ctx; st->n_jobs = n_jobs; st->worst_case = worst_case; st->arena = arena; st->mem_share = mem_share;
    st->n_seq.assign(n_seq, n_seq + n_jobs);
    st->job_len_off.resize(n_jobs + 1); st->job_seq_off.resize(n_jobs + 1);
    st->job_sum_len.resize(n_jobs); st->job_max_len.resize(n_jobs);
    int64_t ns = 0, nb = 0;
    for (int64_t j = 0; j < n_jobs; ++j) {
        if (n_seq[j] <= 0) { set_error(ctx, "job without sequences"); delete st; return BARB200_EINVAL; }
        st->job_len_off[j] = ns; st->job_seq_off[j] = nb;
        int64_t sum = 0; int ml = 0;
        for (int i = 0; i < n_seq[j]; ++i) {
            const int l = seq_lens[ns + i];
            if (l <= 0) { set_error(ctx, "empty sequence in a POA job (the shim substitutes 'N', poaBarAligner.c:551-562)"); delete st; return BARB200_EINVAL; }
            sum += l; ml = std::max(ml, l);
        }
        st->job_sum_len[j] = sum; st->job_max_len[j] = ml;
        ns += n_seq[j]; nb += sum;
    }
    st->job_len_off[n_jobs] = ns; st->job_seq_off[n_jobs] = nb;
    st->n_seqs = ns; st->n_bases = nb;
    st->lens.assign(seq_lens, seq_lens + ns);
    st->soff.resize(ns); st->order.resize(ns); st->progressive.resize(n_jobs);
    st->desc.resize(n_jobs);
    int64_t msa_off = 0; int bad = 0;
    for (int64_t j = 0; j < n_jobs; ++j) {
        int64_t o = 0;
        for (int i = 0; i < n_seq[j]; ++i) { st->soff[st->job_len_off[j] + i] = o; o += seq_lens[st->job_len_off[j] + i]; }
        st->progressive[j] = progressive ? progressive[j] : ctx->hp.progressive_poa;
        const int64_t sum = st->job_sum_len[j], ml = st->job_max_len[j];
        int64_t stride = worst_case ? sum : std::min<int64_t>(sum, ml + ml / 2 + 64);
        stride = align_up(stride, 16);
        JobDesc &d = st->desc[j];
        d.n_seq = n_seq[j]; d.seq_off = st->job_seq_off[j]; d.len_off = st->job_len_off[j]; d.msa_off = msa_off; d.msa_stride = (int)stride;
        msa_off += stride * n_seq[j];
    }
    st->msa_bytes = msa_off;
    // input validation (codes 0..4), then the guide-tree orders on host threads -- unless the caller streams them in
    // behind the running kernel (stage_stream_orders)
    const int nthreads = host_threads(ctx);
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : bad)
    for (int64_t b = 0; b < (nb + 65535) / 65536; ++b) {
        const int64_t e = std::min<int64_t>(nb, (b + 1) * 65536);
        uint8_t m = 0;
        for (int64_t t = b * 65536; t < e; ++t) m |= seqs[t] > 4;
        bad |= m;
    }
    if (bad) { set_error(ctx, "sequence code > 4"); delete st; return BARB200_EINVAL; }
    st->host_seqs = seqs;
    if (!defer_orders) {
        host_orders(st, 0, n_jobs);
        st->orders_done = n_jobs;
    }
    if (n_jobs == 0) { *out = st; return BARB200_OK; }
    int rc = plan_stage(st);
    if (rc) { delete st; return rc; }
    // device buffers (one cached block) + upload on the copy stream, so that it overlaps a running kernel
    size_t off = 0;
    auto sub = [&](size_t bytes) { size_t r = off; off = (off + std::max<size_t>(bytes, 16) + 255) & ~(size_t)255; return r; };
    const size_t o_seqs = sub(nb), o_lens = sub(ns * 4), o_order = sub(ns * 4), o_soff = sub(ns * 8), o_desc = sub(n_jobs * sizeof(JobDesc)),
                 o_msa = sub(st->msa_bytes), o_msa_len = sub(n_jobs * 4), o_status = sub(n_jobs * 4), o_cells = sub(n_jobs * 8), o_next = sub(4), o_ready = sub(4);
    cudaError_t e = ctx_alloc(ctx, &st->d_block, off);
    if (e != cudaSuccess) {
        cudaGetLastError(); set_error(ctx, std::string("cudaMalloc(stage) failed: ") + cudaGetErrorString(e));
        delete st; return BARB200_ENOMEM;
    }
    st->d_block_bytes = off;
    uint8_t *blk = (uint8_t *)st->d_block;
    st->d_seqs = blk + o_seqs; st->d_lens = (int *)(blk + o_lens); st->d_order = (int *)(blk + o_order); st->d_soff = (int64_t *)(blk + o_soff);
    st->d_desc = (JobDesc *)(blk + o_desc); st->d_msa = blk + o_msa; st->d_msa_len = (int *)(blk + o_msa_len); st->d_status = (int *)(blk + o_status);
    st->d_cells = (long long *)(blk + o_cells); st->d_next = (int *)(blk + o_next); st->d_ready = (int *)(blk + o_ready);
    cudaStream_t s = ctx->copy_stream;
    if ((e = cudaMemcpyAsync(st->d_seqs, seqs, nb, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_lens, st->lens.data(), ns * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (st->orders_done && (e = cudaMemcpyAsync(st->d_order, st->order.data(), ns * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) ||
        (e = post_ready(st, st->orders_done)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_soff, st->soff.data(), ns * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_desc, st->desc.data(), n_jobs * sizeof(JobDesc), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess) {
        set_error(ctx, std::string("H2D failed: ") + cudaGetErrorString(e)); stage_free_device(st); delete st; return BARB200_ECUDA;
    }
    *out = st;
    return BARB200_OK;
}

extern "C" int barb200_stage_create(barb200_ctx *ctx, int64_t n_jobs, const int *n_seq, const int *seq_lens,
                                    const uint8_t *seqs, const int *progressive, barb200_stage **out) {
    if (!ctx || !out) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return stage_build(ctx, n_jobs, n_seq, seq_lens, seqs, progressive, false, out);
}

static int stage_run_locked(barb200_stage *st, float *kernel_ms);

// Guide trees behind the running kernel: the stage was built with defer_orders, its kernel is already queued and its
// CTAs wait for `ready`. Jobs are released chunk by chunk as soon as their read orders have been computed and uploaded.
static int stage_stream_orders(barb200_stage *st) {
    barb200_ctx *ctx = st->ctx;
    const int64_t chunk = std::max<int64_t>(256, st->slots);
    while (st->orders_done < st->n_jobs) {
        const int64_t j0 = st->orders_done, j1 = std::min<int64_t>(st->n_jobs, j0 + chunk);
        host_orders(st, j0, j1);
        const int64_t a = st->job_len_off[j0], b = st->job_len_off[j1];
        cudaError_t e = cudaMemcpyAsync(st->d_order + a, st->order.data() + a, (b - a) * 4, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = post_ready(st, j1);
        if (e != cudaSuccess) {
            // release everything so that the kernel can drain, then report
            post_ready(st, st->n_jobs); cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->ar[st->arena].stream);
            set_error(ctx, std::string("streaming guide trees: ") + cudaGetErrorString(e)); return BARB200_ECUDA;
        }
        st->orders_done = j1;
    }
    return BARB200_OK;
}

// queue the stage's kernel (+ the download of the job statuses) on its arena's stream; returns without waiting
static int stage_launch(barb200_stage *st) {
    barb200_ctx *ctx = st->ctx;
    cudaSetDevice(ctx->device);
    st->launches = 0; st->ran = false; st->launched = false;
    if (st->n_jobs == 0) return BARB200_OK;
    if (st->retry) { barb200_stage_destroy(st->retry); st->retry = nullptr; st->retry_jobs.clear(); }
    const size_t clk_n = ctx->p.collect_phase_clocks ? (size_t)st->slots * PH_N : 0;
    int rc = ensure_arena(ctx, st->arena, (size_t)st->lay.slot_bytes * st->slots, (size_t)st->lay.plane_cap * 4 * st->slots, clk_n);
    if (rc) return rc;
    barb200_ctx::Arena &AR = ctx->ar[st->arena];
    cudaStream_t s = AR.stream;
    CUDA_TRY(ctx, cudaMemsetAsync(st->d_next, 0, 4, s));
    if (clk_n) CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_clk, 0, clk_n * sizeof(unsigned long long), s));
    BatchArgs A;
    A.jobs = st->d_desc; A.n_jobs = (int)st->n_jobs; A.seqs = st->d_seqs; A.lens = st->d_lens; A.soff = st->d_soff; A.order = st->d_order;
    A.msa = st->d_msa; A.msa_len = st->d_msa_len; A.status = st->d_status; A.cells = st->d_cells;
    A.slots = AR.d_slots; A.planes = AR.d_planes; A.next_job = st->d_next; A.ready = st->d_ready;
    A.phase_clk = clk_n ? ctx->d_clk : nullptr;
    A.serial_phases = getenv("BARB200_DEBUG_SERIAL") ? 1 : 0;
    A.bfs_order = getenv("BARB200_DEBUG_BFS") ? 1 : 0;
    A.scratch_bytes = (int)st->dyn_smem; A.lay = st->lay; A.P = ctx->P;
    if (!st->e0) { CUDA_TRY(ctx, cudaEventCreate(&st->e0)); CUDA_TRY(ctx, cudaEventCreate(&st->e1)); }
    CUDA_TRY(ctx, cudaEventRecord(st->e0, s));
    kKernels[st->kernel_class].fn<<<st->slots, st->T, st->dyn_smem, s>>>(A);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { set_error(ctx, std::string("kernel launch: ") + cudaGetErrorString(le)); return BARB200_ECUDA; }
    CUDA_TRY(ctx, cudaEventRecord(st->e1, s));
    // (no device-to-host copy here: into pageable memory it would block the host until the kernel is done and
    // serialise the chunk pipeline; stage_finish fetches the statuses after its synchronisation)
    st->launched = true;
    return BARB200_OK;
}

// wait for the stage's kernel, collect its device time, re-run capacity misses in a worst-case-sized stage
static int stage_finish(barb200_stage *st, float *kernel_ms) {
    barb200_ctx *ctx = st->ctx;
    if (kernel_ms) *kernel_ms = 0.f;
    if (st->n_jobs == 0) { st->ran = true; return BARB200_OK; }
    if (!st->launched) { set_error(ctx, "stage_finish without stage_launch"); return BARB200_EINVAL; }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->ar[st->arena].stream;
    cudaError_t se = cudaStreamSynchronize(s);
    if (se != cudaSuccess) { set_error(ctx, std::string("kernel execution: ") + cudaGetErrorString(se)); return BARB200_ECUDA; }
    float ms = 0.f; cudaEventElapsedTime(&ms, st->e0, st->e1);
    st->launches = 1; st->launched = false;
    st->status.resize(st->n_jobs);
    CUDA_TRY(ctx, cudaMemcpyAsync(st->status.data(), st->d_status, st->n_jobs * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const size_t clk_n = ctx->p.collect_phase_clocks ? (size_t)st->slots * PH_N : 0;
    if (clk_n) {
        std::vector<unsigned long long> h(clk_n);
        CUDA_TRY(ctx, cudaMemcpy(h.data(), ctx->d_clk, clk_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 6; ++k) st->clk[k] = 0;
        for (int b = 0; b < st->slots; ++b) for (int k = 0; k < 6; ++k) st->clk[k] += h[(size_t)b * PH_N + k];
    }
    // capacity misses -> one worst-case-sized retry launch for just those jobs
    std::vector<int64_t> redo;
    for (int64_t j = 0; j < st->n_jobs; ++j) {
        const int sc = st->status[j];
        if (sc == JOB_OK) continue;
        if (!st->worst_case && (sc == JOB_ERR_PLANE_CAP || sc == JOB_ERR_MSA_CAP)) redo.push_back(j);
        else {
            char buf[160]; snprintf(buf, sizeof(buf), "job %lld failed on the device with status %d", (long long)j, sc);
            set_error(ctx, buf); return BARB200_EJOB;
        }
    }
    if (!redo.empty()) {
        std::vector<int> r_nseq, r_lens, r_prog; std::vector<uint8_t> r_seqs;
        // the retry stage needs host copies of the inputs: fetch them back from the device buffers we own
        std::vector<uint8_t> h_seqs(st->n_bases);
        CUDA_TRY(ctx, cudaMemcpy(h_seqs.data(), st->d_seqs, st->n_bases, cudaMemcpyDeviceToHost));
        for (int64_t j : redo) {
            r_nseq.push_back(st->n_seq[j]); r_prog.push_back(st->progressive[j]);
            for (int i = 0; i < st->n_seq[j]; ++i) r_lens.push_back(st->lens[st->job_len_off[j] + i]);
            r_seqs.insert(r_seqs.end(), h_seqs.begin() + st->job_seq_off[j], h_seqs.begin() + st->job_seq_off[j + 1]);
        }
        barb200_stage *rs = nullptr;
        int rc = stage_build(ctx, (int64_t)redo.size(), r_nseq.data(), r_lens.data(), r_seqs.data(), r_prog.data(), true, &rs, st->arena, st->mem_share);
        if (rc) return rc;
        float rms = 0.f;
        rc = stage_run_locked(rs, &rms);
        if (rc) { barb200_stage_destroy(rs); return rc; }
        st->retry = rs; st->retry_jobs = redo; st->launches += rs->launches; ms += rms;
        for (int k = 0; k < 6; ++k) st->clk[k] += rs->clk[k];
    }
    if (kernel_ms) *kernel_ms = ms;
    st->ran = true;
    return BARB200_OK;
}

static int stage_run_locked(barb200_stage *st, float *kernel_ms) {
    int rc = stage_launch(st);
    if (rc) return rc;
    return stage_finish(st, kernel_ms);
}

extern "C" int barb200_stage_run(barb200_stage *st, float *kernel_ms) {
    if (!st) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(st->ctx->mu);
    return stage_run_locked(st, kernel_ms);
}

extern "C" int64_t barb200_stage_launches(barb200_stage *st) { return st ? st->launches : 0; }

extern "C" int barb200_stage_phase_clocks(barb200_stage *st, uint64_t out[6]) {
    if (!st || !out) return BARB200_EINVAL;
    for (int k = 0; k < 6; ++k) out[k] = st->clk[k];
    return BARB200_OK;
}

static int stage_fetch_locked(barb200_stage *st, uint8_t **msa_out, int *msa_len, int64_t *cells) {
    barb200_ctx *ctx = st->ctx;
    if (!st->ran) { set_error(ctx, "stage_fetch before stage_run"); return BARB200_EINVAL; }
    if (st->n_jobs == 0) return BARB200_OK;
    cudaSetDevice(ctx->device);
    // the retry stage first: it shares the context's pinned staging buffer
    std::vector<uint8_t *> r_out; std::vector<int> r_len; std::vector<int64_t> r_cells;
    if (st->retry) {
        const size_t n = st->retry_jobs.size();
        r_out.assign(n, nullptr); r_len.assign(n, 0); r_cells.assign(n, 0);
        int rc = stage_fetch_locked(st->retry, r_out.data(), r_len.data(), r_cells.data());
        if (rc) return rc;
    }
    st->msa_len.resize(st->n_jobs); st->cells.resize(st->n_jobs);
    if ((size_t)st->msa_bytes > ctx->h_pinned_bytes) {
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        ctx->h_pinned = nullptr; ctx->h_pinned_bytes = 0;
        const size_t want = (size_t)st->msa_bytes + ((size_t)st->msa_bytes >> 2);
        if (cudaMallocHost((void **)&ctx->h_pinned, want) != cudaSuccess) { cudaGetLastError(); set_error(ctx, "cudaMallocHost failed"); return BARB200_ENOMEM; }
        ctx->h_pinned_bytes = want;
    }
    uint8_t *h_msa = ctx->h_pinned;
    cudaStream_t s = ctx->ar[st->arena].stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(st->msa_len.data(), st->d_msa_len, st->n_jobs * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(st->cells.data(), st->d_cells, st->n_jobs * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_msa, st->d_msa, st->msa_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    std::vector<int64_t> redo_pos(st->n_jobs, -1);
    for (size_t i = 0; i < st->retry_jobs.size(); ++i) redo_pos[st->retry_jobs[i]] = (int64_t)i;
    int oom = 0;
    const int nthreads = host_threads(ctx);
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : oom)
    for (int64_t j = 0; j < st->n_jobs; ++j) {
        if (redo_pos[j] >= 0) {
            const int64_t i = redo_pos[j];
            if (msa_out) msa_out[j] = r_out[i]; else free(r_out[i]);
            if (msa_len) msa_len[j] = r_len[i];
            if (cells) cells[j] = r_cells[i];
            continue;
        }
        const int K = st->n_seq[j], ml = st->msa_len[j];
        if (msa_len) msa_len[j] = ml;
        if (cells) cells[j] = st->cells[j];
        if (msa_out) {
            uint8_t *o = (uint8_t *)malloc((size_t)K * (ml > 0 ? ml : 1));
            if (!o) { oom = 1; msa_out[j] = nullptr; continue; }
            const uint8_t *src = h_msa + st->desc[j].msa_off;
            for (int i = 0; i < K; ++i) memcpy(o + (size_t)i * ml, src + (size_t)i * st->desc[j].msa_stride, ml);
            msa_out[j] = o;
        }
    }
    if (oom) { set_error(ctx, "host allocation failed"); return BARB200_ENOMEM; }
    return BARB200_OK;
}

extern "C" int barb200_stage_fetch(barb200_stage *st, uint8_t **msa_out, int *msa_len, int64_t *cells) {
    if (!st) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(st->ctx->mu);
    return stage_fetch_locked(st, msa_out, msa_len, cells);
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool timing_on() { static const bool on = getenv("BARB200_TIMING") != nullptr; return on; }

extern "C" int barb200_poa_msa_batch(barb200_ctx *ctx, int64_t n_jobs, const int *n_seq, const int *seq_lens,
                                     const uint8_t *seqs, const int *progressive, uint8_t **msa_out, int *msa_len,
                                     int64_t *cells) {
    if (!ctx) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (msa_out) for (int64_t j = 0; j < n_jobs; ++j) msa_out[j] = nullptr;
    barb200_stage *st = nullptr;
    const double t0 = now_ms();
    // one launch; the guide trees are computed behind it and jobs are released to the kernel as their orders arrive
    const bool stream = n_jobs >= 512 && !getenv("BARB200_NO_PIPELINE");
    int rc = stage_build(ctx, n_jobs, n_seq, seq_lens, seqs, progressive, false, &st, 0, 1.0, stream);
    if (rc) return rc;
    const double t1 = now_ms(); float kms = 0.f;
    rc = stage_launch(st);
    if (!rc && stream) rc = stage_stream_orders(st);
    if (!rc) rc = stage_finish(st, &kms);
    const double t2 = now_ms();
    if (!rc) rc = stage_fetch_locked(st, msa_out, msa_len, cells);
    const double t3 = now_ms();
    barb200_stage_destroy(st);
    if (timing_on()) fprintf(stderr, "barb200 timing: build %.1f ms, run %.1f ms (%.1f on the device), fetch %.1f ms, destroy %.1f ms\n", t1 - t0, t2 - t1, kms, t3 - t2, now_ms() - t3);
    return rc;
}

namespace barb200 {
static int run_jobs_now(barb200_ctx *ctx, const std::vector<HostJob> &jobs, std::vector<JobResult> &results);

// Callers on different host threads (the reference enters the BAR code from OpenMP teams, bar/impl/bar.c:90-94) do not queue up
// behind the device one by one: whatever is waiting when the device becomes free runs as ONE batch (group_commit.h).
int run_jobs(barb200_ctx *ctx, const std::vector<HostJob> &jobs, std::vector<JobResult> &results) {
    PoaRequest r;
    r.jobs = &jobs; r.results = &results;
    ctx->poa_group.submit(&r, [](const PoaRequest &, const PoaRequest &) { return true; },
                          [ctx](std::vector<PoaRequest *> &batch) {
                              run_poa_group(batch, [ctx](const std::vector<HostJob> &j, std::vector<JobResult> &res) { return run_jobs_now(ctx, j, res); });
                          });
    return r.rc;
}
GroupCommit<PecanRequest> &pecan_group(barb200_ctx *ctx) { return ctx->pecan_group; }

static int run_jobs_now(barb200_ctx *ctx, const std::vector<HostJob> &jobs, std::vector<JobResult> &results) {
    const int64_t n = (int64_t)jobs.size();
    results.assign(n, JobResult());
    if (n == 0) return BARB200_OK;
    std::vector<int> n_seq(n), prog(n), lens; std::vector<uint8_t> seqs;
    for (int64_t j = 0; j < n; ++j) {
        n_seq[j] = jobs[j].n_seq; prog[j] = jobs[j].progressive;
        int64_t sum = 0;
        for (int i = 0; i < jobs[j].n_seq; ++i) { lens.push_back(jobs[j].lens[i]); sum += jobs[j].lens[i]; }
        seqs.insert(seqs.end(), jobs[j].seqs, jobs[j].seqs + sum);
    }
    std::vector<uint8_t *> out(n, nullptr); std::vector<int> ml(n, 0); std::vector<int64_t> cells(n, 0);
    int rc = barb200_poa_msa_batch(ctx, n, n_seq.data(), lens.data(), seqs.data(), prog.data(), out.data(), ml.data(), cells.data());
    if (rc) { for (auto p : out) free(p); return rc; }
    for (int64_t j = 0; j < n; ++j) {
        results[j].msa_len = ml[j]; results[j].cells = cells[j];
        results[j].msa.assign(out[j], out[j] + (size_t)jobs[j].n_seq * ml[j]);
        free(out[j]);
    }
    return BARB200_OK;
}
}  // namespace barb200
