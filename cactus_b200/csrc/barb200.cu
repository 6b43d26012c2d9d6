// barb200.cu -- device orchestration + C ABI of libbarb200.so (see include/barb200.h).
//
// A *stage* is one device batch of POA jobs (one job = one abpoa_msa call of the reference):
//   * jobs are BUCKETED by the CTA-size class their longest sequence needs (32 .. 1024 threads, 16 columns per thread) and,
//     inside a class, ordered by estimated cost, largest first. Every class gets its own slot layout sized from ITS largest
//     job, its own persistent launch on its own stream (largest class first; the launches run concurrently, the block
//     scheduler fills whatever a class leaves free) and its own work counter -- a batch of thousands of short adjacencies
//     and a few 10 kbp windows no longer runs everything in the CTA shape and the slot size of the largest job;
//   * the host does no per-job work besides packing: the guide-tree read orders are computed by the CTA that owns the job
//     (guide_tree.cuh);
//   * jobs that outgrow the optimistic plane / MSA sizing come back flagged and are re-run with geometrically larger
//     slots (x4, then worst case).
// A context drives one or more devices; every device has two *lanes* (slot arena + streams + pinned staging) so that one
// batch's upload / download / unpacking overlaps the other's kernels. host_bar.cpp's dispatcher feeds the lanes of all
// devices from one queue of ends. No CPU fallback exists: if CUDA is unavailable, creation fails.
#include <cuda_runtime.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
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
    {32, barb200::poa_msa_kernel_t32, 10 * 1024}, {64, barb200::poa_msa_kernel_t64, 24 * 1024}, {128, barb200::poa_msa_kernel_t128, 40 * 1024},
    {256, barb200::poa_msa_kernel_t256, 96 * 1024}, {640, barb200::poa_msa_kernel_t640, 200 * 1024}, {1024, barb200::poa_msa_kernel_t1024, 200 * 1024}};
static const int kNumKernels = 6;
static const int kMaxDevices = 8;
using namespace barb200;

namespace barb200 {
// one in-flight device batch: slot arena, streams, pinned staging
struct Lane {
    int index = 0;                      // global lane index (device * lanes_per_device + lane)
    std::mutex busy;                    // held while a batch (or a staged run) owns the arena
    uint8_t *d_slots = nullptr; size_t slots_bytes = 0;
    int *d_planes = nullptr; size_t planes_bytes = 0;
    cudaStream_t main = nullptr, copy = nullptr, cls[kNumKernels] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t cls_done[kNumKernels] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long *d_clk = nullptr; size_t clk_entries = 0;
};
struct Device {
    int ordinal = 0, sm_count = 0; size_t mem_total = 0;
    std::vector<std::unique_ptr<Lane>> lanes;
    // grow-only cache of device blocks for the per-stage buffers (cudaMalloc / cudaFree per call cost milliseconds
    // and cudaFree synchronises the device, which would stall the upload / kernel overlap)
    std::mutex cache_mu; std::vector<std::pair<void *, size_t>> free_blocks; size_t cached_bytes = 0;
    // pinned staging blocks (packed inputs up, MSA bytes down), shared by the device's lanes: cudaMallocHost of a 60 MB block takes
    // ~25 ms, so a lane that meets its first large batch borrows the block the other lane has grown already
    std::mutex pin_mu; std::vector<std::pair<void *, size_t>> free_pinned;
};
}  // namespace barb200

struct barb200_ctx {
    barb200_params p;
    PoaParams P;
    HostParams hp;
    std::vector<std::unique_ptr<Device>> devs;
    int lanes_per_device = 2;
    bool lanes_shared = false;          // set once the dispatcher runs: every lane plans with its share of the device memory
    std::mutex mu;                      // serialises the pair-HMM batches (device 0)
    std::mutex err_mu;
    std::string err;
    GroupCommit<PecanRequest> pecan_group;  // concurrent barb200_pecan_aligned_pairs_batch callers share device batches
    void *pecan_scratch = nullptr; size_t pecan_scratch_bytes = 0;   // pecan.cu's batch call (grow-only)
    void *pecan_pinned[2] = {nullptr, nullptr}; size_t pecan_pinned_bytes[2] = {0, 0};   // pinned staging: 0 upload, 1 download
    void *dispatcher = nullptr;         // host_bar.cpp's end queue (created on first use, destroyed with the context)
    double last_timing[6] = {0, 0, 0, 0, 0, 0};   // of the most recent device batch (barb200_last_batch_timing), under err_mu
};

namespace barb200 {
// the message is kept twice: in the context (for callers whose request ran inside another thread's batch) and per
// thread (concurrent callers do not overwrite each other's text)
static thread_local std::string tls_err;
static thread_local const barb200_ctx *tls_err_ctx = nullptr;
void set_error(barb200_ctx *ctx, const std::string &msg) {
    if (!ctx) return;
    tls_err = msg; tls_err_ctx = ctx;
    std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = msg;
}
std::string get_error(barb200_ctx *ctx) {
    if (tls_err_ctx == ctx && !tls_err.empty()) return tls_err;
    std::lock_guard<std::mutex> lk(ctx->err_mu); return ctx->err;
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
// a batch over several devices runs one host thread per device: each takes its share of the host threads for its parallel loops
static thread_local int tl_thread_divisor = 1;
int host_threads(barb200_ctx *ctx) {
    static const int dflt = usable_host_threads();
    const int n = ctx->p.host_threads > 0 ? ctx->p.host_threads : dflt;
    return std::max(1, n / std::max(1, tl_thread_divisor));
}
int default_progressive(barb200_ctx *ctx) { return ctx->p.progressive_poa; }
std::mutex &device_mutex(barb200_ctx *ctx) { return ctx->mu; }
int ctx_device(barb200_ctx *ctx) { return ctx->devs[0]->ordinal; }
int ctx_sm_count(barb200_ctx *ctx) { return ctx->devs[0]->sm_count; }
double ctx_mem_fraction(barb200_ctx *ctx) { return ctx->p.mem_fraction > 0 ? ctx->p.mem_fraction : 0.85; }
int total_lanes(barb200_ctx *ctx) { return (int)ctx->devs.size() * ctx->lanes_per_device; }
void **dispatcher_slot(barb200_ctx *ctx) { return &ctx->dispatcher; }
void mark_lanes_shared(barb200_ctx *ctx) { ctx->lanes_shared = true; }
static Device &dev_of_lane(barb200_ctx *ctx, int lane) { return *ctx->devs[lane / ctx->lanes_per_device]; }
static Lane &lane_of(barb200_ctx *ctx, int lane) { return *ctx->devs[lane / ctx->lanes_per_device]->lanes[lane % ctx->lanes_per_device]; }
}  // namespace barb200

static size_t block_size_of(size_t bytes) { return (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255; }
static cudaError_t dev_alloc(Device &D, void **p, size_t bytes) {
    bytes = block_size_of(bytes);
    {
        std::lock_guard<std::mutex> lk(D.cache_mu);
        int best = -1;
        for (size_t i = 0; i < D.free_blocks.size(); ++i)
            if (D.free_blocks[i].second >= bytes && D.free_blocks[i].second <= 2 * bytes + (1 << 20) &&
                (best < 0 || D.free_blocks[i].second < D.free_blocks[best].second)) best = (int)i;
        if (best >= 0) { *p = D.free_blocks[best].first; D.cached_bytes -= D.free_blocks[best].second; D.free_blocks.erase(D.free_blocks.begin() + best); return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {       // give the cache back and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(D.cache_mu);
        for (auto &b : D.free_blocks) cudaFree(b.first);
        D.free_blocks.clear(); D.cached_bytes = 0;
        e = cudaMalloc(p, bytes);
    }
    return e;
}
static void *pinned_take(Device &D, size_t bytes, size_t *got) {
    {
        std::lock_guard<std::mutex> lk(D.pin_mu);
        int best = -1;
        for (size_t i = 0; i < D.free_pinned.size(); ++i)
            if (D.free_pinned[i].second >= bytes && (best < 0 || D.free_pinned[i].second < D.free_pinned[best].second)) best = (int)i;
        if (best >= 0) { void *p = D.free_pinned[best].first; *got = D.free_pinned[best].second; D.free_pinned.erase(D.free_pinned.begin() + best); return p; }
        // nothing fits: retire the smallest block (the pool stays at a handful of blocks, each grown to the largest batch seen)
        if (D.free_pinned.size() >= 4) {
            size_t m = 0;
            for (size_t i = 1; i < D.free_pinned.size(); ++i) if (D.free_pinned[i].second < D.free_pinned[m].second) m = i;
            cudaFreeHost(D.free_pinned[m].first); D.free_pinned.erase(D.free_pinned.begin() + m);
        }
    }
    void *p = nullptr;
    const size_t want = bytes + (bytes >> 2) + 4096;
    if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got = want;
    return p;
}
static void pinned_give(Device &D, void *p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(D.pin_mu);
    D.free_pinned.emplace_back(p, bytes);
}
struct PinnedBlock {       // scope guard
    Device &D; void *p = nullptr; size_t bytes = 0;
    PinnedBlock(Device &d, size_t want) : D(d) { p = pinned_take(D, want, &bytes); }
    ~PinnedBlock() { pinned_give(D, p, bytes); }
    PinnedBlock(const PinnedBlock &) = delete; PinnedBlock &operator=(const PinnedBlock &) = delete;
};

static void dev_free(Device &D, void *p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(D.cache_mu);
    if (D.cached_bytes + block_size_of(bytes) > ((size_t)4 << 30) || D.free_blocks.size() >= 64) { cudaFree(p); return; }
    D.free_blocks.emplace_back(p, block_size_of(bytes)); D.cached_bytes += block_size_of(bytes);
}

namespace barb200 {
int device_alloc(barb200_ctx *ctx, void **p, size_t bytes) { return dev_alloc(*ctx->devs[0], p, bytes) == cudaSuccess ? 0 : -1; }
void device_free(barb200_ctx *ctx, void *p, size_t bytes) { dev_free(*ctx->devs[0], p, bytes); }
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
GroupCommit<PecanRequest> &pecan_group(barb200_ctx *ctx) { return ctx->pecan_group; }
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
    p->collect_phase_clocks = 0; p->n_devices = 0; p->lanes = 0;
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
    if (p->k <= 0 || p->k > 20 || p->w <= 0 || p->w >= 256) { fail(errbuf, errbuf_len, "minimizer k must be in 1..20 and w in 1..255 (guide-tree keys are hash << 24 | span << 16 | read)"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { fail(errbuf, errbuf_len, std::string("no CUDA device: ") + cudaGetErrorString(e)); return nullptr; }
    // device list: params.devices[0 .. n_devices) (n_devices = -1: every visible device), else the single params.device
    std::vector<int> ords;
    if (p->n_devices < 0) for (int d = 0; d < std::min(ndev, kMaxDevices); ++d) ords.push_back(d);
    else if (p->n_devices > 0) { for (int d = 0; d < std::min(p->n_devices, kMaxDevices); ++d) ords.push_back(p->devices[d]); }
    else ords.push_back(p->device);
    for (size_t a = 0; a < ords.size(); ++a) {
        if (ords[a] < 0 || ords[a] >= ndev) { fail(errbuf, errbuf_len, "device ordinal out of range"); return nullptr; }
        for (size_t b = 0; b < a; ++b) if (ords[a] == ords[b]) { fail(errbuf, errbuf_len, "device listed twice"); return nullptr; }
    }
    std::unique_ptr<barb200_ctx> ctx(new barb200_ctx());
    ctx->p = *p;
    ctx->lanes_per_device = p->lanes > 0 ? std::min(p->lanes, 2) : 2;
    if (getenv("BARB200_LANES")) ctx->lanes_per_device = std::max(1, std::min(2, atoi(getenv("BARB200_LANES"))));
    PoaParams &P = ctx->P;
    memcpy(P.mat, p->mat, sizeof(P.mat));
    P.o1 = p->gap_open1; P.e1 = p->gap_ext1; P.o2 = p->gap_open2; P.e2 = p->gap_ext2; P.wb = p->wb; P.wf = p->wf;
    P.max_mat = 0; P.min_mis = 0;
    for (int i = 0; i < 25; ++i) { P.max_mat = std::max(P.max_mat, P.mat[i]); P.min_mis = std::max(P.min_mis, -P.mat[i]); }
    // the reference's int32 minus infinity, abpoa_align_simd.c:1299
    const int oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    P.inf_min = std::max(std::max(INT32_MIN + P.min_mis, INT32_MIN + oe1), INT32_MIN + oe2) + 512 * std::max(P.e1, P.e2);
    ctx->hp = HostParams{p->k, p->w, p->min_w, p->progressive_poa};
    int lane_index = 0;
    for (int ord : ords) {
        if (cudaSetDevice(ord) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaSetDevice failed"); barb200_destroy(ctx.release()); return nullptr; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ord) != cudaSuccess) { fail(errbuf, errbuf_len, "cudaGetDeviceProperties failed"); barb200_destroy(ctx.release()); return nullptr; }
        std::unique_ptr<Device> D(new Device());
        D->ordinal = ord; D->sm_count = prop.multiProcessorCount; D->mem_total = prop.totalGlobalMem;
        for (int i = 0; i < kNumKernels; ++i) cudaFuncSetAttribute(kKernels[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kKernels[i].scratch);
        bool ok = true;
        for (int l = 0; l < ctx->lanes_per_device && ok; ++l) {
            std::unique_ptr<Lane> L(new Lane());
            L->index = lane_index++;
            ok = cudaStreamCreateWithFlags(&L->main, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&L->copy, cudaStreamNonBlocking) == cudaSuccess;
            for (int c = 0; c < kNumKernels && ok; ++c)
                ok = cudaStreamCreateWithFlags(&L->cls[c], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&L->cls_done[c], cudaEventDisableTiming) == cudaSuccess;
            D->lanes.push_back(std::move(L));
        }
        ctx->devs.push_back(std::move(D));
        if (!ok) { fail(errbuf, errbuf_len, "creating streams / pinned memory failed"); barb200_destroy(ctx.release()); return nullptr; }
    }
    cudaSetDevice(ords[0]);
    return ctx.release();
}

namespace barb200 { void dispatcher_destroy(barb200_ctx *ctx); }

extern "C" void barb200_destroy(barb200_ctx *ctx) {
    if (!ctx) return;
    dispatcher_destroy(ctx);            // joins the lane threads (host_bar.cpp)
    for (auto &D : ctx->devs) {
        cudaSetDevice(D->ordinal);
        for (auto &L : D->lanes) {
            if (L->d_slots) cudaFree(L->d_slots);
            if (L->d_planes) cudaFree(L->d_planes);
            if (L->d_clk) cudaFree(L->d_clk);
            for (int c = 0; c < kNumKernels; ++c) { if (L->cls[c]) cudaStreamDestroy(L->cls[c]); if (L->cls_done[c]) cudaEventDestroy(L->cls_done[c]); }
            if (L->main) cudaStreamDestroy(L->main);
            if (L->copy) cudaStreamDestroy(L->copy);
        }
        for (auto &b : D->free_blocks) cudaFree(b.first);
        for (auto &b : D->free_pinned) cudaFreeHost(b.first);
    }
    if (!ctx->devs.empty()) cudaSetDevice(ctx->devs[0]->ordinal);
    if (ctx->pecan_scratch) cudaFree(ctx->pecan_scratch);
    for (int i = 0; i < 2; ++i) if (ctx->pecan_pinned[i]) cudaFreeHost(ctx->pecan_pinned[i]);
    delete ctx;
}

// The returned pointer stays valid until the calling thread's next call of this function (thread-local copy).
extern "C" const char *barb200_last_error(barb200_ctx *ctx) {
    if (!ctx) return "null context";
    static thread_local std::string out;
    out = get_error(ctx);
    return out.c_str();
}
extern "C" void barb200_free(void *p) { free(p); }
extern "C" void barb200_free_many(void *const *p, int64_t n) {
    if (!p) return;
    for (int64_t i = 0; i < n; ++i) free(p[i]);
}
// rows[i] (bytes[i] bytes each, e.g. the MSAs of a batch) -> dst back to back; parallel copy
extern "C" void barb200_pack_rows(void *const *rows, const int64_t *bytes, int64_t n, uint8_t *dst) {
    if (!rows || !bytes || !dst || n <= 0) return;
    std::vector<int64_t> off((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) off[i + 1] = off[i] + bytes[i];
    static const int nt = std::min(usable_host_threads(), 16);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n; ++i) if (rows[i] && bytes[i] > 0) memcpy(dst + off[i], rows[i], (size_t)bytes[i]);
}

extern "C" int barb200_device_count(barb200_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

extern "C" int barb200_device_info(barb200_ctx *ctx, int *sm_count, int64_t *mem_total, int64_t *mem_free, char *name, int name_len) {
    if (!ctx) return BARB200_EINVAL;
    cudaSetDevice(ctx->devs[0]->ordinal);
    cudaDeviceProp prop; CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->devs[0]->ordinal));
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
// one CTA-size class of a stage: jobs [job_base, job_base + n_jobs) in the stage's internal order
struct Bucket {
    int cls = 0, T = 0; size_t dyn_smem = 0;
    int64_t job_base = 0, n_jobs = 0;
    SlotLayout lay; int slots = 0;
    size_t slot_off = 0; int64_t plane_off = 0; size_t clk_off = 0;   // offsets into the lane's arena (bytes / ints / entries)
};

struct barb200_stage {
    barb200_ctx *ctx = nullptr;
    int lane = 0;
    int64_t n_jobs = 0, n_seqs = 0, n_bases = 0;
    // everything below is in the stage's INTERNAL job order: class-major (largest class first), cost-descending inside a
    // class; perm[internal] = the caller's job index
    std::vector<int64_t> perm;
    std::vector<int> n_seq, lens, progressive;
    std::vector<int64_t> soff, job_len_off, job_seq_off, job_sum_len;
    std::vector<int> job_max_len;
    std::vector<JobDesc> desc;
    std::vector<Bucket> buckets;
    int64_t msa_bytes = 0;
    double grow = 1.0; bool worst_case = false;
    // device: one block from the device's cache holds all per-stage arrays
    void *d_block = nullptr; size_t d_block_bytes = 0;
    uint8_t *d_seqs = nullptr, *d_msa = nullptr; int *d_lens = nullptr; int64_t *d_soff = nullptr;
    JobDesc *d_desc = nullptr; int *d_msa_len = nullptr, *d_status = nullptr, *d_next = nullptr; long long *d_cells = nullptr;
    int *d_order = nullptr, *d_gt_status = nullptr; uint8_t *d_gt_scratch = nullptr;
    GuideTreeArgs gt; int gt_ctas = 0; float gt_ms = 0.f; cudaEvent_t e_gt = nullptr;   // K0: the guide trees of the stage (guide_tree.cu)
    // results of the last run
    std::vector<int> status, msa_len; std::vector<long long> cells;
    barb200_stage *retry = nullptr; std::vector<int64_t> retry_jobs;     // internal ids
    int64_t launches = 0; bool ran = false;
    cudaEvent_t e0 = nullptr, e1 = nullptr; bool launched = false;
    const uint8_t *host_seqs = nullptr;      // the caller's buffer while it is valid (capacity-miss retries re-upload from it)
    uint64_t clk[7] = {0, 0, 0, 0, 0, 0, 0};
};

static void stage_free_device(barb200_stage *st) {
    dev_free(dev_of_lane(st->ctx, st->lane), st->d_block, st->d_block_bytes);
    st->d_block = nullptr;
    st->d_seqs = st->d_msa = nullptr; st->d_lens = nullptr; st->d_soff = nullptr; st->d_desc = nullptr;
    st->d_msa_len = st->d_status = st->d_next = nullptr; st->d_cells = nullptr; st->d_order = st->d_gt_status = nullptr; st->d_gt_scratch = nullptr;
}

extern "C" void barb200_stage_destroy(barb200_stage *st) {
    if (!st) return;
    cudaSetDevice(dev_of_lane(st->ctx, st->lane).ordinal);
    if (st->retry) barb200_stage_destroy(st->retry);
    if (st->e0) { cudaEventDestroy(st->e0); cudaEventDestroy(st->e1); }
    if (st->e_gt) cudaEventDestroy(st->e_gt);
    stage_free_device(st);
    delete st;
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int class_of_len(barb200_ctx *ctx, int64_t max_len) {
    for (int i = 0; i < kNumKernels; ++i) if ((int64_t)kKernels[i].T * CPT >= max_len + 1 && kKernels[i].T >= ctx->p.threads_per_block) return i;
    return -1;
}
// banded cells the job will sweep, roughly: (K-1) alignments x ~mean length rows x band width (SURVEY.md 8e's cost)
static double job_cost(const barb200_ctx *ctx, int K, int64_t sum, int64_t ml) {
    const double w = 2.0 * (ctx->P.wb + ctx->P.wf * (double)ml) + 1.0;
    return (double)(K - 1) * (double)(sum / std::max(1, K) + 1) * std::min<double>((double)ml + 1.0, w) + 2000.0 * K;
}

// Slot sizes, shared memory and resident CTAs of every bucket; carve the lane's arena.
static int plan_stage(barb200_stage *st) {
    barb200_ctx *ctx = st->ctx;
    Device &D = dev_of_lane(ctx, st->lane);
    Lane &LN = lane_of(ctx, st->lane);
    for (Bucket &B : st->buckets) {
        int64_t max_nodes = 4, max_edges = 4, max_len = 1, max_k = 1, plane_need = 0;
        for (int64_t j = B.job_base; j < B.job_base + B.n_jobs; ++j) {
            const int64_t sum = st->job_sum_len[j], ml = st->job_max_len[j], K = st->n_seq[j];
            max_nodes = std::max(max_nodes, sum + 2); max_edges = std::max(max_edges, sum + K); max_len = std::max(max_len, ml);
            max_k = std::max(max_k, K);
            // rows the graph can reach: worst case every base a new node; optimistic: the longest read plus a share of the rest
            int64_t rows = sum + 2;
            if (!st->worst_case) rows = std::min<int64_t>(rows, (int64_t)(st->grow * (double)(ml + (sum - ml) / 8 + 256)));
            // columns a row stores: the adaptive band is [min(maxL, d) - w, max(maxR, d) + w] (abpoa_align_simd.c:946-960), i.e. 2w plus
            // however far the best columns of the predecessors have drifted from the row's diagonal d; a quarter of 2w is allowed for
            // that before the geometric retries take over. Only windows longer than ~2.5 kbp are narrower than the whole row
            // (w = 1000 + 0.1 L): a 10 kbp window plans 0.77 GB of planes instead of 1.53 GB, so every SM gets a resident CTA.
            int64_t width = ml + 1;
            if (!st->worst_case) {
                const int64_t w = (int64_t)ctx->P.wb + (int64_t)(ctx->P.wf * (double)ml);
                width = std::min<int64_t>(width, (int64_t)(st->grow * (2.5 * (double)w + 64.0)));
            }
            plane_need = std::max(plane_need, rows * (TB / CPT) * (align_up(width, CPT) + CPT));
        }
        SlotLayout &Y = B.lay;
        memset(&Y, 0, sizeof(Y));
        Y.node_cap = (int)max_nodes; Y.in_pool = (int)(4 * max_edges + 64); Y.out_pool = Y.in_pool;
        Y.W = (int)(1 + ((max_k - 1) >> 6)); Y.cigar_cap = (int)(max_len + max_nodes + 16);
        Y.plane_cap = align_up(plane_need, 8);
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
        B.T = kKernels[B.cls].T; B.dyn_smem = kKernels[B.cls].scratch;
        if (getenv("BARB200_SCRATCH_KB")) B.dyn_smem = (size_t)atoi(getenv("BARB200_SCRATCH_KB")) * 1024;   // tuning aid
        int per_sm = 0;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kKernels[B.cls].fn, B.T, B.dyn_smem));
        if (per_sm < 1) { set_error(ctx, "kernel does not fit on an SM with the requested configuration"); return BARB200_EINVAL; }
        if (ctx->p.ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->p.ctas_per_sm);
        B.slots = (int)std::min<int64_t>(B.n_jobs, (int64_t)per_sm * D.sm_count);
    }
    // memory: what is free now + what this lane's arena already holds, minus the stage's own block; a lane of a context
    // whose dispatcher runs plans with its share so that the other lane of the device can do the same
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
    const double frac = ctx->p.mem_fraction > 0 ? ctx->p.mem_fraction : 0.85;
    double budget = (double)(free_b + LN.slots_bytes + LN.planes_bytes) * frac - (double)st->d_block_bytes;
    if (ctx->lanes_shared && ctx->lanes_per_device > 1) budget = std::min(budget, (double)total_b * frac / ctx->lanes_per_device - (double)st->d_block_bytes);
    auto need_of = [&](const Bucket &B) { return (double)B.slots * ((double)B.lay.slot_bytes + (double)B.lay.plane_cap * 4.0); };
    double need = 0;
    for (const Bucket &B : st->buckets) need += need_of(B);
    for (int pass = 0; pass < 8 && need > budget; ++pass) {
        const double scale = budget / need;
        need = 0;
        for (Bucket &B : st->buckets) { B.slots = std::max(1, (int)((double)B.slots * scale)); need += need_of(B); }
    }
    if (need > budget) { set_error(ctx, "a single job needs more device memory than is available"); return BARB200_ENOMEM; }
    size_t so = 0, co = 0; int64_t po = 0;
    for (Bucket &B : st->buckets) {
        B.slot_off = so; B.plane_off = po; B.clk_off = co;
        so += (size_t)B.lay.slot_bytes * B.slots; po += B.lay.plane_cap * B.slots; co += (size_t)B.slots * PH_N;
    }
    return BARB200_OK;
}

static int ensure_arena(barb200_ctx *ctx, Device &D, Lane &LN, size_t slots_bytes, size_t planes_bytes, size_t clk_entries) {
    // a lane that has to grow goes straight to what its sibling already needed (same workload, same budget: plan_stage): which lane
    // meets the first large batch is a matter of timing, and a re-allocation costs tens of milliseconds plus a device synchronisation.
    // If that much is not free (the sibling was sized while it had the device to itself) the lane takes what the stage needs.
    auto grow = [&](void **p, size_t *have, size_t need, size_t sibling_max, const char *what) {
        if (need <= *have) return BARB200_OK;
        if (*p) cudaFree(*p);
        *p = nullptr; *have = 0;
        size_t want = std::max(need, sibling_max);
        if (cudaMalloc(p, want) != cudaSuccess) {
            cudaGetLastError(); *p = nullptr;
            if (want == need || cudaMalloc(p, need) != cudaSuccess) { cudaGetLastError(); *p = nullptr; set_error(ctx, std::string("cudaMalloc(") + what + ") failed"); return BARB200_ENOMEM; }
            want = need;
        }
        *have = want;
        return BARB200_OK;
    };
    size_t sib_slots = 0, sib_planes = 0;
    for (auto &other : D.lanes) if (other.get() != &LN) { sib_slots = std::max(sib_slots, other->slots_bytes); sib_planes = std::max(sib_planes, other->planes_bytes); }
    int rc = grow((void **)&LN.d_slots, &LN.slots_bytes, slots_bytes, sib_slots, "slots");
    if (!rc) rc = grow((void **)&LN.d_planes, &LN.planes_bytes, planes_bytes, sib_planes, "planes");
    if (rc) return rc;
    if (clk_entries > LN.clk_entries) {
        if (LN.d_clk) cudaFree(LN.d_clk);
        LN.d_clk = nullptr; LN.clk_entries = 0;
        if (cudaMalloc(&LN.d_clk, clk_entries * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); set_error(ctx, "cudaMalloc(clk) failed"); return BARB200_ENOMEM; }
        LN.clk_entries = clk_entries;
    }
    return BARB200_OK;
}

// n_seq / seq_lens / seqs / progressive are in the CALLER's job order; seq_off[j] (may be null = consecutive) is the offset of
// job j's first base in `seqs`, n_bases_total the size of `seqs`. grow / worst_case size the slots (capacity-miss retries).
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool timing_on() { static const bool on = getenv("BARB200_TIMING") != nullptr; return on; }

static int stage_build(barb200_ctx *ctx, int lane, int64_t n_jobs, const int *n_seq, const int *seq_lens, const uint8_t *seqs, const int64_t *seq_off,
                       int64_t n_bases_total, const int *progressive, double grow, bool worst_case, barb200_stage **out) {
    if (!ctx || n_jobs < 0 || (n_jobs > 0 && (!n_seq || !seq_lens || !seqs))) { set_error(ctx, "bad arguments"); return BARB200_EINVAL; }
    if (n_jobs > 0x7ffffff0) { set_error(ctx, "too many jobs in one stage"); return BARB200_EINVAL; }
    Device &D = dev_of_lane(ctx, lane);
    Lane &LN = lane_of(ctx, lane);
    cudaSetDevice(D.ordinal);
    std::unique_ptr<barb200_stage> st(new barb200_stage());
    st->ctx = ctx; st->lane = lane; st->n_jobs = n_jobs; st->grow = grow; st->worst_case = worst_case;
    const double tb0 = now_ms();
    // ---- caller-order facts ----
    std::vector<int64_t> c_len_off(n_jobs + 1), c_seq_off(n_jobs + 1), c_sum(n_jobs);
    std::vector<int> c_ml(n_jobs), c_cls(n_jobs);
    std::vector<double> c_cost(n_jobs);
    int64_t ns = 0, nb = 0;
    for (int64_t j = 0; j < n_jobs; ++j) {
        if (n_seq[j] <= 0) { set_error(ctx, "job without sequences"); return BARB200_EINVAL; }
        c_len_off[j] = ns; c_seq_off[j] = seq_off ? seq_off[j] : nb;
        int64_t sum = 0; int ml = 0;
        for (int i = 0; i < n_seq[j]; ++i) {
            const int l = seq_lens[ns + i];
            if (l <= 0) { set_error(ctx, "empty sequence in a POA job (the shim substitutes 'N', poaBarAligner.c:551-562)"); return BARB200_EINVAL; }
            sum += l; ml = std::max(ml, l);
        }
        c_sum[j] = sum; c_ml[j] = ml; c_cls[j] = class_of_len(ctx, ml); c_cost[j] = job_cost(ctx, n_seq[j], sum, ml);
        if (c_cls[j] < 0) { set_error(ctx, "a sequence is longer than the device engine's row limit (16383 bases per window)"); return BARB200_EINVAL; }
        ns += n_seq[j]; nb += sum;
    }
    c_len_off[n_jobs] = ns; c_seq_off[n_jobs] = nb;
    if (!seq_off) n_bases_total = nb;
    st->n_seqs = ns; st->n_bases = n_bases_total;
    // input validation (codes 0..4)
    {
        int bad = 0;
        const int nthreads = host_threads(ctx);
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : bad)
        for (int64_t b = 0; b < (n_bases_total + 65535) / 65536; ++b) {
            const int64_t e = std::min<int64_t>(n_bases_total, (b + 1) * 65536);
            uint8_t m = 0;
            for (int64_t t = b * 65536; t < e; ++t) m |= seqs[t] > 4;
            bad |= m;
        }
        if (bad) { set_error(ctx, "sequence code > 4"); return BARB200_EINVAL; }
    }
    const double tb1 = now_ms();
    // ---- internal order: largest class first, inside a class by estimated cost, largest first (stable) ----
    st->perm.resize(n_jobs);
    std::iota(st->perm.begin(), st->perm.end(), (int64_t)0);
    std::stable_sort(st->perm.begin(), st->perm.end(), [&](int64_t a, int64_t b) {
        if (c_cls[a] != c_cls[b]) return c_cls[a] > c_cls[b];
        return c_cost[a] > c_cost[b];
    });
    st->n_seq.resize(n_jobs); st->progressive.resize(n_jobs);
    st->job_len_off.resize(n_jobs + 1); st->job_seq_off.resize(n_jobs + 1); st->job_sum_len.resize(n_jobs); st->job_max_len.resize(n_jobs);
    st->lens.resize(ns); st->soff.resize(ns); st->desc.resize(n_jobs);
    int64_t lo = 0, msa_off = 0;
    for (int64_t j = 0; j < n_jobs; ++j) {
        const int64_t c = st->perm[j];
        const int K = n_seq[c];
        st->n_seq[j] = K; st->progressive[j] = progressive ? progressive[c] : ctx->hp.progressive_poa;
        st->job_len_off[j] = lo; st->job_seq_off[j] = c_seq_off[c]; st->job_sum_len[j] = c_sum[c]; st->job_max_len[j] = c_ml[c];
        int64_t o = 0;
        for (int i = 0; i < K; ++i) { st->lens[lo + i] = seq_lens[c_len_off[c] + i]; st->soff[lo + i] = o; o += st->lens[lo + i]; }
        const int64_t sum = c_sum[c], ml = c_ml[c];
        int64_t stride = sum;
        if (!worst_case) stride = std::min<int64_t>(sum, (int64_t)(grow * (double)(ml + ml / 2 + 64)));
        stride = align_up(stride, 16);
        JobDesc &d = st->desc[j];
        d.n_seq = K; d.seq_off = c_seq_off[c]; d.len_off = lo; d.msa_off = msa_off; d.msa_stride = (int)stride; d.progressive = st->progressive[j];
        if (d.progressive && K > 65535) { set_error(ctx, "progressive mode with more than 65535 sequences in one window is not supported"); return BARB200_EINVAL; }
        msa_off += stride * K;
        lo += K;
        if (st->buckets.empty() || st->buckets.back().cls != c_cls[c]) { Bucket B; B.cls = c_cls[c]; B.job_base = j; st->buckets.push_back(B); }
        st->buckets.back().n_jobs++;
    }
    st->job_len_off[n_jobs] = lo; st->job_seq_off[n_jobs] = n_bases_total;
    st->msa_bytes = msa_off;
    st->host_seqs = seqs;
    if (n_jobs == 0) { *out = st.release(); return BARB200_OK; }
    // device buffers (one cached block) + upload on the lane's copy stream, so that it overlaps a running kernel
    size_t off = 0;
    auto sub = [&](size_t bytes) { size_t r = off; off = (off + std::max<size_t>(bytes, 16) + 255) & ~(size_t)255; return r; };
    // K0's scratch: one slot per resident guide-tree CTA, sized from the stage's largest job (keys: ~2 / (w + 1) per base; ties in
    // repeats can push up to 2 w per base -- capacity misses are retried like the planes')
    GuideTreeArgs &GA = st->gt;
    memset(&GA, 0, sizeof(GA));
    {
        int64_t key_need = 64, gx_need = 8, mk = 1;
        for (int64_t j = 0; j < n_jobs; ++j) {
            if (!(st->progressive[j] && st->n_seq[j] > 2)) continue;
            const int64_t sum = st->job_sum_len[j], worst = 2 * (int64_t)ctx->p.w * sum + st->n_seq[j];
            key_need = std::max(key_need, worst_case ? worst : std::min<int64_t>(worst, (int64_t)(grow * (double)(sum / 2 + 64))));
            gx_need = std::max(gx_need, sum); mk = std::max<int64_t>(mk, st->n_seq[j]);
        }
        int64_t kc = 64; while (kc < key_need) kc <<= 1;            // the sort pads to a power of two
        int64_t o = 0;
        auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 16); return r; };
        GA.key_cap = (int)kc;
        GA.o_keys = take(kc * 8); GA.o_gx = take(gx_need * 8); GA.o_hit = take(mk * (mk + 1) / 2 * 4); GA.o_jac = take(mk * (mk - 1) / 2 * 8 + 8); GA.o_score = take(mk * 8);
        GA.slot_bytes = align_up(o, 256);
        st->gt_ctas = (int)std::min<int64_t>(std::min<int64_t>(n_jobs, (int64_t)4 * D.sm_count), std::max<int64_t>(1, ((int64_t)2 << 30) / GA.slot_bytes));
    }
    const size_t o_seqs = sub(n_bases_total), o_lens = sub(ns * 4), o_soff = sub(ns * 8), o_desc = sub(n_jobs * sizeof(JobDesc)),
                 o_msa = sub(st->msa_bytes), o_msa_len = sub(n_jobs * 4), o_status = sub(n_jobs * 4), o_cells = sub(n_jobs * 8), o_next = sub(4 * (kNumKernels + 1)),
                 o_order = sub(ns * 4), o_gts = sub(n_jobs * 4), o_gtscr = sub((size_t)GA.slot_bytes * st->gt_ctas);
    st->d_block_bytes = off;
    const double tb2 = now_ms();
    int rc = plan_stage(st.get());
    if (rc) return rc;
    const double tb3 = now_ms();
    cudaError_t e = dev_alloc(D, &st->d_block, off);
    if (e != cudaSuccess) {
        cudaGetLastError(); set_error(ctx, std::string("cudaMalloc(stage) failed: ") + cudaGetErrorString(e));
        st->d_block = nullptr; return BARB200_ENOMEM;
    }
    uint8_t *blk = (uint8_t *)st->d_block;
    st->d_seqs = blk + o_seqs; st->d_lens = (int *)(blk + o_lens); st->d_soff = (int64_t *)(blk + o_soff);
    st->d_desc = (JobDesc *)(blk + o_desc); st->d_msa = blk + o_msa; st->d_msa_len = (int *)(blk + o_msa_len); st->d_status = (int *)(blk + o_status);
    st->d_cells = (long long *)(blk + o_cells); st->d_next = (int *)(blk + o_next);
    st->d_order = (int *)(blk + o_order); st->d_gt_status = (int *)(blk + o_gts); st->d_gt_scratch = blk + o_gtscr;
    GA.jobs = st->d_desc; GA.n_jobs = (int)n_jobs; GA.seqs = st->d_seqs; GA.lens = st->d_lens; GA.soff = st->d_soff; GA.order = st->d_order; GA.gt_status = st->d_gt_status;
    GA.next_job = st->d_next + kNumKernels; GA.scratch = st->d_gt_scratch; GA.k = ctx->p.k; GA.w = ctx->p.w;
    cudaStream_t s = LN.copy;
    const double tb4 = now_ms();
    if ((e = cudaMemcpyAsync(st->d_seqs, seqs, n_bases_total, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_lens, st->lens.data(), ns * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_soff, st->soff.data(), ns * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_desc, st->desc.data(), n_jobs * sizeof(JobDesc), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess) {
        set_error(ctx, std::string("H2D failed: ") + cudaGetErrorString(e)); stage_free_device(st.get()); return BARB200_ECUDA;
    }
    if (timing_on()) fprintf(stderr, "barb200 timing: stage_build: scan + validate %.2f ms, order + descriptors %.2f, plan %.2f, alloc %.2f, upload %.2f\n",
                             tb1 - tb0, tb2 - tb1, tb3 - tb2, tb4 - tb3, now_ms() - tb4);
    *out = st.release();
    return BARB200_OK;
}

extern "C" int barb200_stage_create(barb200_ctx *ctx, int64_t n_jobs, const int *n_seq, const int *seq_lens,
                                    const uint8_t *seqs, const int *progressive, barb200_stage **out) {
    if (!ctx || !out) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(lane_of(ctx, 0).busy);
    int rc = stage_build(ctx, 0, n_jobs, n_seq, seq_lens, seqs, nullptr, 0, progressive, 1.0, false, out);
    if (!rc) (*out)->host_seqs = nullptr;       // the caller's buffer is only guaranteed during this call
    return rc;
}

static int stage_run_locked(barb200_stage *st, float *kernel_ms);

// queue the stage's kernels on its lane's streams; returns without waiting
static int stage_launch(barb200_stage *st) {
    barb200_ctx *ctx = st->ctx;
    Device &D = dev_of_lane(ctx, st->lane);
    Lane &LN = lane_of(ctx, st->lane);
    cudaSetDevice(D.ordinal);
    st->launches = 0; st->ran = false; st->launched = false;
    if (st->n_jobs == 0) return BARB200_OK;
    if (st->retry) { barb200_stage_destroy(st->retry); st->retry = nullptr; st->retry_jobs.clear(); }
    size_t slots_bytes = 0, plane_ints = 0, clk_n = 0;
    for (const Bucket &B : st->buckets) { slots_bytes += (size_t)B.lay.slot_bytes * B.slots; plane_ints += (size_t)B.lay.plane_cap * B.slots; clk_n += (size_t)B.slots * PH_N; }
    if (!ctx->p.collect_phase_clocks) clk_n = 0;
    int rc = ensure_arena(ctx, dev_of_lane(ctx, st->lane), LN, slots_bytes, plane_ints * 4, clk_n);
    if (rc) return rc;
    cudaStream_t s = LN.main;
    CUDA_TRY(ctx, cudaMemsetAsync(st->d_next, 0, 4 * (kNumKernels + 1), s));
    CUDA_TRY(ctx, cudaMemsetAsync(st->d_msa, 0, st->msa_bytes, s));     // the rows' padding travels back with the MSAs (one D2H copy): defined bytes
    if (clk_n) CUDA_TRY(ctx, cudaMemsetAsync(LN.d_clk, 0, clk_n * sizeof(unsigned long long), s));
    if (!st->e0) { CUDA_TRY(ctx, cudaEventCreate(&st->e0)); CUDA_TRY(ctx, cudaEventCreate(&st->e1)); }
    if (!st->e_gt) CUDA_TRY(ctx, cudaEventCreate(&st->e_gt));
    CUDA_TRY(ctx, cudaEventRecord(st->e0, s));
    // K0: the guide trees of every job (guide_tree.cu); the POA launches below wait for it through the stream / e_gt
    launch_guide_tree(st->gt, st->gt_ctas, (void *)s);
    { cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) { set_error(ctx, std::string("guide tree launch: ") + cudaGetErrorString(le)); return BARB200_ECUDA; } }
    CUDA_TRY(ctx, cudaEventRecord(st->e_gt, s));
    st->launches++;
    const bool single = st->buckets.size() == 1;
    for (size_t b = 0; b < st->buckets.size(); ++b) {          // largest class first
        const Bucket &B = st->buckets[b];
        BatchArgs A;
        A.jobs = st->d_desc; A.job_base = (int)B.job_base; A.n_jobs = (int)B.n_jobs; A.seqs = st->d_seqs; A.lens = st->d_lens; A.soff = st->d_soff;
        A.order = st->d_order; A.gt_status = st->d_gt_status;
        A.msa = st->d_msa; A.msa_len = st->d_msa_len; A.status = st->d_status; A.cells = st->d_cells;
        A.slots = LN.d_slots + B.slot_off; A.planes = LN.d_planes + B.plane_off; A.next_job = st->d_next + b;
        A.phase_clk = clk_n ? LN.d_clk + B.clk_off : nullptr;
        A.serial_phases = getenv("BARB200_DEBUG_SERIAL") ? 1 : 0;
        A.bfs_order = getenv("BARB200_DEBUG_BFS") ? 1 : 0;
        A.scratch_bytes = (int)B.dyn_smem; A.lay = B.lay; A.P = ctx->P;
        cudaStream_t cs = single ? s : LN.cls[B.cls];
        if (!single) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, st->e_gt, 0));
        kKernels[B.cls].fn<<<B.slots, B.T, B.dyn_smem, cs>>>(A);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) { set_error(ctx, std::string("kernel launch: ") + cudaGetErrorString(le)); return BARB200_ECUDA; }
        if (!single) { CUDA_TRY(ctx, cudaEventRecord(LN.cls_done[B.cls], cs)); CUDA_TRY(ctx, cudaStreamWaitEvent(s, LN.cls_done[B.cls], 0)); }
        st->launches++;
    }
    CUDA_TRY(ctx, cudaEventRecord(st->e1, s));
    // (no device-to-host copy here: into pageable memory it would block the host until the kernels are done;
    // stage_finish fetches the statuses after its synchronisation)
    st->launched = true;
    return BARB200_OK;
}

// wait for the stage's kernels, collect their device time, re-run capacity misses with larger slots
static int stage_finish(barb200_stage *st, float *kernel_ms) {
    barb200_ctx *ctx = st->ctx;
    Device &D = dev_of_lane(ctx, st->lane);
    Lane &LN = lane_of(ctx, st->lane);
    if (kernel_ms) *kernel_ms = 0.f;
    if (st->n_jobs == 0) { st->ran = true; return BARB200_OK; }
    if (!st->launched) { set_error(ctx, "stage_finish without stage_launch"); return BARB200_EINVAL; }
    cudaSetDevice(D.ordinal);
    cudaStream_t s = LN.main;
    cudaError_t se = cudaStreamSynchronize(s);
    if (se != cudaSuccess) { set_error(ctx, std::string("kernel execution: ") + cudaGetErrorString(se)); return BARB200_ECUDA; }
    float ms = 0.f; cudaEventElapsedTime(&ms, st->e0, st->e1);
    st->gt_ms = 0.f; cudaEventElapsedTime(&st->gt_ms, st->e0, st->e_gt);
    st->launched = false;
    st->status.resize(st->n_jobs);
    CUDA_TRY(ctx, cudaMemcpyAsync(st->status.data(), st->d_status, st->n_jobs * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    for (int k = 0; k < 7; ++k) st->clk[k] = 0;
    st->clk[6] = (uint64_t)(st->gt_ms * 1e6f);          // K0 is a kernel of its own: its device time, in nanoseconds
    if (ctx->p.collect_phase_clocks) {
        size_t clk_n = 0;
        for (const Bucket &B : st->buckets) clk_n += (size_t)B.slots * PH_N;
        std::vector<unsigned long long> h(clk_n);
        CUDA_TRY(ctx, cudaMemcpy(h.data(), LN.d_clk, clk_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (size_t b = 0; b < clk_n / PH_N; ++b) for (int k = 0; k < 6; ++k) st->clk[k] += h[b * PH_N + k];
    }
    // capacity misses -> retry launch for just those jobs with larger slots (x4 first, then worst case)
    std::vector<int64_t> redo;
    for (int64_t j = 0; j < st->n_jobs; ++j) {
        const int sc = st->status[j];
        if (sc == JOB_OK) continue;
        if (!st->worst_case && (sc == JOB_ERR_PLANE_CAP || sc == JOB_ERR_MSA_CAP || sc == JOB_ERR_GT_CAP)) redo.push_back(j);
        else {
            char buf[160]; snprintf(buf, sizeof(buf), "job %lld failed on the device with status %d", (long long)st->perm[j], sc);
            set_error(ctx, buf); return BARB200_EJOB;
        }
    }
    if (!redo.empty()) {
        std::vector<int> r_nseq, r_lens, r_prog; std::vector<int64_t> r_off;
        // the retry stage needs host copies of the inputs: use the caller's buffer while it is valid, else fetch them back
        std::vector<uint8_t> h_seqs;
        const uint8_t *src = st->host_seqs;
        if (!src) {
            h_seqs.resize(st->n_bases);
            CUDA_TRY(ctx, cudaMemcpy(h_seqs.data(), st->d_seqs, st->n_bases, cudaMemcpyDeviceToHost));
            src = h_seqs.data();
        }
        for (int64_t j : redo) {
            r_nseq.push_back(st->n_seq[j]); r_prog.push_back(st->progressive[j]); r_off.push_back(st->job_seq_off[j]);
            for (int i = 0; i < st->n_seq[j]; ++i) r_lens.push_back(st->lens[st->job_len_off[j] + i]);
        }
        barb200_stage *rs = nullptr;
        const bool go_worst = st->grow >= 4.0;
        int rc = stage_build(ctx, st->lane, (int64_t)redo.size(), r_nseq.data(), r_lens.data(), src, r_off.data(), st->n_bases, r_prog.data(),
                             go_worst ? 1.0 : st->grow * 4.0, go_worst, &rs);
        if (rc) return rc;
        float rms = 0.f;
        rc = stage_run_locked(rs, &rms);
        rs->host_seqs = nullptr;
        if (rc) { barb200_stage_destroy(rs); return rc; }
        st->retry = rs; st->retry_jobs = redo; st->launches += rs->launches; ms += rms;
        for (int k = 0; k < 7; ++k) st->clk[k] += rs->clk[k];
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
    std::lock_guard<std::mutex> lk(lane_of(st->ctx, st->lane).busy);
    return stage_run_locked(st, kernel_ms);
}

extern "C" int64_t barb200_stage_launches(barb200_stage *st) { return st ? st->launches : 0; }

extern "C" int barb200_stage_phase_clocks(barb200_stage *st, uint64_t out[7]) {
    if (!st || !out) return BARB200_EINVAL;
    for (int k = 0; k < 7; ++k) out[k] = st->clk[k];
    return BARB200_OK;
}

// per-bucket facts of a stage for reports: for bucket b, out[4b..4b+3] = threads per CTA, jobs, resident CTAs, plane ints per slot
extern "C" int barb200_stage_buckets(barb200_stage *st, int64_t *out, int max_buckets) {
    if (!st) return 0;
    int n = 0;
    for (const Bucket &B : st->buckets) {
        if (n >= max_buckets) break;
        if (out) { out[4 * n] = B.T; out[4 * n + 1] = B.n_jobs; out[4 * n + 2] = B.slots; out[4 * n + 3] = B.lay.plane_cap; }
        ++n;
    }
    return n;
}

// dest(caller job index, K, msa_len) -> where the K x msa_len bytes go (nullptr: allocation failure)
typedef std::function<uint8_t *(int64_t, int, int)> MsaDest;

static int stage_fetch_locked(barb200_stage *st, const MsaDest &dest, int *msa_len, int64_t *cells, const std::vector<int64_t> *outer = nullptr) {
    barb200_ctx *ctx = st->ctx;
    Device &D = dev_of_lane(ctx, st->lane);
    Lane &LN = lane_of(ctx, st->lane);
    if (!st->ran) { set_error(ctx, "stage_fetch before stage_run"); return BARB200_EINVAL; }
    if (st->n_jobs == 0) return BARB200_OK;
    cudaSetDevice(D.ordinal);
    // caller index of internal job j (a retry stage's "caller" is the parent stage: `outer` maps its job list to the real caller)
    auto caller_of = [&](int64_t j) { const int64_t c = st->perm[j]; return outer ? (*outer)[c] : c; };
    // the retry stage first: it shares the lane's pinned staging buffer
    if (st->retry) {
        std::vector<int64_t> op(st->retry_jobs.size());
        for (size_t i = 0; i < op.size(); ++i) op[i] = caller_of(st->retry_jobs[i]);
        int rc = stage_fetch_locked(st->retry, dest, msa_len, cells, &op);
        if (rc) return rc;
    }
    st->msa_len.resize(st->n_jobs); st->cells.resize(st->n_jobs);
    PinnedBlock down(D, (size_t)st->msa_bytes);
    if (!down.p) { set_error(ctx, "cudaMallocHost failed"); return BARB200_ENOMEM; }
    uint8_t *h_msa = (uint8_t *)down.p;
    cudaStream_t s = LN.main;
    CUDA_TRY(ctx, cudaMemcpyAsync(st->msa_len.data(), st->d_msa_len, st->n_jobs * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(st->cells.data(), st->d_cells, st->n_jobs * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_msa, st->d_msa, st->msa_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    std::vector<char> redone(st->n_jobs, 0);
    for (int64_t j : st->retry_jobs) redone[j] = 1;
    int oom = 0;
    const int nthreads = host_threads(ctx);
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : oom)
    for (int64_t j = 0; j < st->n_jobs; ++j) {
        if (redone[j]) continue;
        const int64_t c = caller_of(j);
        const int K = st->n_seq[j], ml = st->msa_len[j];
        if (msa_len) msa_len[c] = ml;
        if (cells) cells[c] = st->cells[j];
        if (dest) {
            uint8_t *o = dest(c, K, ml);
            if (!o) { oom = 1; continue; }
            const uint8_t *src = h_msa + st->desc[j].msa_off;
            for (int i = 0; i < K; ++i) memcpy(o + (size_t)i * ml, src + (size_t)i * st->desc[j].msa_stride, ml);
        }
    }
    if (oom) { set_error(ctx, "host allocation failed"); return BARB200_ENOMEM; }
    return BARB200_OK;
}

static MsaDest malloc_dest(uint8_t **msa_out) {
    if (!msa_out) return MsaDest();
    return [msa_out](int64_t c, int K, int ml) { uint8_t *o = (uint8_t *)malloc((size_t)K * (ml > 0 ? ml : 1)); msa_out[c] = o; return o; };
}

extern "C" int barb200_stage_fetch(barb200_stage *st, uint8_t **msa_out, int *msa_len, int64_t *cells) {
    if (!st) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(lane_of(st->ctx, st->lane).busy);
    return stage_fetch_locked(st, malloc_dest(msa_out), msa_len, cells);
}


// one device batch on one lane (the caller holds the lane): build + launch + streamed guide trees + finish + fetch
static int batch_on_lane(barb200_ctx *ctx, int lane, int64_t n_jobs, const int *n_seq, const int *seq_lens, const uint8_t *seqs, const int64_t *seq_off,
                         int64_t n_bases, const int *progressive, const MsaDest &dest, int *msa_len, int64_t *cells, float *device_ms) {
    barb200_stage *st = nullptr;
    const double t0 = now_ms();
    int rc = stage_build(ctx, lane, n_jobs, n_seq, seq_lens, seqs, seq_off, n_bases, progressive, 1.0, false, &st);
    if (rc) return rc;
    const double t1 = now_ms(); float kms = 0.f;
    rc = stage_launch(st);
    if (!rc) rc = stage_finish(st, &kms);
    const double t2 = now_ms();
    if (!rc) rc = stage_fetch_locked(st, dest, msa_len, cells);
    const double t3 = now_ms();
    barb200_stage_destroy(st);
    if (device_ms) *device_ms = kms;
    {
        std::lock_guard<std::mutex> lk(ctx->err_mu);
        ctx->last_timing[0] = t1 - t0; ctx->last_timing[1] = t2 - t1; ctx->last_timing[2] = kms; ctx->last_timing[3] = t3 - t2;
        ctx->last_timing[4] = now_ms() - t0; ctx->last_timing[5] = (double)n_jobs;
    }
    if (timing_on()) fprintf(stderr, "barb200 timing: lane %d, %lld jobs: build %.1f ms, run %.1f ms (%.1f on the device), fetch %.1f ms, destroy %.1f ms\n",
                             lane, (long long)n_jobs, t1 - t0, t2 - t1, kms, t3 - t2, now_ms() - t3);
    return rc;
}

// job-count / byte limits of ONE device batch (larger requests are cut into chunks)
static const int64_t kMaxJobsPerBatch = 1 << 15;
static const int64_t kMaxBasesPerBatch = (int64_t)768 << 20;

extern "C" int barb200_poa_msa_batch(barb200_ctx *ctx, int64_t n_jobs, const int *n_seq, const int *seq_lens,
                                     const uint8_t *seqs, const int *progressive, uint8_t **msa_out, int *msa_len,
                                     int64_t *cells) {
    if (!ctx || n_jobs < 0 || (n_jobs > 0 && (!n_seq || !seq_lens || !seqs))) { if (ctx) set_error(ctx, "bad arguments"); return BARB200_EINVAL; }
    if (msa_out) for (int64_t j = 0; j < n_jobs; ++j) msa_out[j] = nullptr;
    if (n_jobs == 0) return BARB200_OK;
    // offsets + estimated cost in the caller's order
    std::vector<int64_t> len_off(n_jobs + 1), seq_off(n_jobs + 1);
    std::vector<double> cost(n_jobs);
    int64_t ns = 0, nb = 0;
    for (int64_t j = 0; j < n_jobs; ++j) {
        if (n_seq[j] <= 0) { set_error(ctx, "job without sequences"); return BARB200_EINVAL; }
        len_off[j] = ns; seq_off[j] = nb;
        int64_t sum = 0, ml = 0;
        for (int i = 0; i < n_seq[j]; ++i) {
            const int l = seq_lens[ns + i];
            if (l <= 0) { set_error(ctx, "empty sequence in a POA job (the shim substitutes 'N', poaBarAligner.c:551-562)"); return BARB200_EINVAL; }
            sum += l; ml = std::max<int64_t>(ml, l);
        }
        cost[j] = job_cost(ctx, n_seq[j], sum, ml);
        ns += n_seq[j]; nb += sum;
    }
    len_off[n_jobs] = ns; seq_off[n_jobs] = nb;
    const int ndev = (int)ctx->devs.size();
    // ---- deal: one device -> everything; several -> cost-sorted, each job to the device with the least work so far (LPT) ----
    std::vector<std::vector<int64_t>> share(ndev);
    if (ndev == 1) { share[0].resize(n_jobs); std::iota(share[0].begin(), share[0].end(), (int64_t)0); }
    else {
        std::vector<int64_t> idx(n_jobs);
        std::iota(idx.begin(), idx.end(), (int64_t)0);
        std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
        std::vector<double> load(ndev, 0.0);
        for (int64_t j : idx) { const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin()); share[d].push_back(j); load[d] += cost[j]; }
        for (auto &s : share) std::sort(s.begin(), s.end());
    }
    std::vector<int> rcs(ndev, BARB200_OK);
    std::vector<std::string> errs(ndev);
    int active_devs = 0;
    for (int d = 0; d < ndev; ++d) active_devs += !share[d].empty();
    auto run_device = [&](int d) {
        tl_thread_divisor = std::max(1, active_devs);
        const std::vector<int64_t> &mine = share[d];
        const int lane = d * ctx->lanes_per_device;
        std::lock_guard<std::mutex> lk(lane_of(ctx, lane).busy);
        size_t at = 0;
        while (at < mine.size() && rcs[d] == BARB200_OK) {
            // one chunk: a bounded number of jobs and bases
            size_t end = at; int64_t bases = 0;
            while (end < mine.size() && (int64_t)(end - at) < kMaxJobsPerBatch &&
                   (end == at || bases + (seq_off[mine[end] + 1] - seq_off[mine[end]]) <= kMaxBasesPerBatch)) {
                bases += seq_off[mine[end] + 1] - seq_off[mine[end]]; ++end;
            }
            const int64_t m = (int64_t)(end - at);
            std::vector<int> c_nseq(m), c_lens, c_prog(m);
            std::vector<int64_t> c_off(m);
            for (int64_t k = 0; k < m; ++k) {
                const int64_t j = mine[at + k];
                c_nseq[k] = n_seq[j]; c_prog[k] = progressive ? progressive[j] : ctx->hp.progressive_poa; c_off[k] = seq_off[j];
                c_lens.insert(c_lens.end(), seq_lens + len_off[j], seq_lens + len_off[j + 1]);
            }
            std::vector<int> c_ml(m); std::vector<int64_t> c_cells(m);
            const int64_t *jobs_of = mine.data() + at;
            MsaDest dest;
            if (msa_out) dest = [msa_out, jobs_of](int64_t c, int K, int ml) { uint8_t *o = (uint8_t *)malloc((size_t)K * (ml > 0 ? ml : 1)); msa_out[jobs_of[c]] = o; return o; };
            const int rc = batch_on_lane(ctx, lane, m, c_nseq.data(), c_lens.data(), seqs, c_off.data(), nb, c_prog.data(), dest, c_ml.data(), c_cells.data(), nullptr);
            if (rc) { rcs[d] = rc; errs[d] = get_error(ctx); break; }
            for (int64_t k = 0; k < m; ++k) { if (msa_len) msa_len[jobs_of[k]] = c_ml[k]; if (cells) cells[jobs_of[k]] = c_cells[k]; }
            at = end;
        }
    };
    if (ndev == 1) run_device(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; ++d) if (!share[d].empty()) th.emplace_back(run_device, d);
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < ndev; ++d) if (rcs[d]) {
        if (msa_out) for (int64_t j = 0; j < n_jobs; ++j) { free(msa_out[j]); msa_out[j] = nullptr; }
        set_error(ctx, errs[d]); return rcs[d];
    }
    return BARB200_OK;
}

// host-side phases of the most recent device batch, milliseconds: out[0] build (pack, validation, planning, H2D), out[1] launch +
// streamed guide trees + wait, out[2] device time of the kernels, out[3] fetch (D2H + unpack), out[4] total, out[5] jobs
extern "C" int barb200_last_batch_timing(barb200_ctx *ctx, double out[6]) {
    if (!ctx || !out) return BARB200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    for (int k = 0; k < 6; ++k) out[k] = ctx->last_timing[k];
    return BARB200_OK;
}

namespace barb200 {
// One device batch of host jobs on lane `lane` (host_bar.cpp's dispatcher): the jobs' sequences are packed once into the
// lane's pinned upload buffer, the MSAs unpacked straight into the results.
int run_jobs_on_lane(barb200_ctx *ctx, int lane, const std::vector<HostJob> &jobs, std::vector<JobResult> &results) {
    const int64_t n = (int64_t)jobs.size();
    results.assign(n, JobResult());
    if (n == 0) return BARB200_OK;
    Lane &LN = lane_of(ctx, lane);
    std::lock_guard<std::mutex> lk(LN.busy);
    cudaSetDevice(dev_of_lane(ctx, lane).ordinal);
    std::vector<int> n_seq(n), prog(n), lens; std::vector<int64_t> off(n + 1);
    int64_t nb = 0;
    for (int64_t j = 0; j < n; ++j) {
        n_seq[j] = jobs[j].n_seq; prog[j] = jobs[j].progressive; off[j] = nb;
        for (int i = 0; i < jobs[j].n_seq; ++i) { lens.push_back(jobs[j].lens[i]); nb += jobs[j].lens[i]; }
    }
    off[n] = nb;
    PinnedBlock up(dev_of_lane(ctx, lane), (size_t)nb);
    if (!up.p) { set_error(ctx, "cudaMallocHost failed"); return BARB200_ENOMEM; }
    uint8_t *const h_up = (uint8_t *)up.p;
    const int nthreads = host_threads(ctx);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t j = 0; j < n; ++j) memcpy(h_up + off[j], jobs[j].seqs, (size_t)(off[j + 1] - off[j]));
    std::vector<int> ml(n, 0); std::vector<int64_t> cells(n, 0);
    JobResult *res = results.data();
    static uint8_t empty_sink[1];
    MsaDest dest = [res](int64_t c, int K, int m) { res[c].msa.resize((size_t)K * m); return res[c].msa.empty() ? empty_sink : res[c].msa.data(); };
    const int rc = batch_on_lane(ctx, lane, n, n_seq.data(), lens.data(), h_up, nullptr, nb, prog.data(), dest, ml.data(), cells.data(), nullptr);
    if (rc) return rc;
    for (int64_t j = 0; j < n; ++j) { results[j].msa_len = ml[j]; results[j].cells = cells[j]; }
    return BARB200_OK;
}
}  // namespace barb200
