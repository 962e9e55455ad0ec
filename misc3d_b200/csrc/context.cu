/*
 * context.cu -- m3d_ctx life cycle, the hypothesis-count exchange (NCCL via dlopen, or a caller
 * callback) and the host-only helpers of the C-ABI (sample table, ordered scan).
 */
#include "context.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <limits>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "scan.h"

namespace m3d {

/* ---------------------------------------------------------------- NCCL through dlopen ------
 * libnccl is not linked: the library must load on boxes without it (and on the CPU-only build
 * box).  When the process already holds torch's libnccl.so.2 the same soname resolves to it. */
struct Id128 {
    char b[128];
};
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, /* ncclUniqueId by value = 128 bytes */ Id128, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load(std::string *why) {
    if (g_nccl.h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        if (why) *why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return false;
    }
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather =
        (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        if (why) *why = "libnccl is missing a required symbol";
        dlclose(h);
        return false;
    }
    g_nccl.h = h;
    return true;
}

bool exchange_has_nccl(const m3d_ctx *ctx) { return ctx->world > 1 && ctx->nccl_comm != nullptr; }

int exchange_allgather_nccl(m3d_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank, cudaStream_t stream) {
    if (!exchange_has_nccl(ctx)) return ctx->fail(M3D_ERR_INTERNAL, "no NCCL communicator");
    const int ncclChar = 0;
    const int rc = g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclChar, ctx->nccl_comm, stream);
    if (rc != 0)
        return ctx->fail(M3D_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return M3D_OK;
}

int exchange_allgather(m3d_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank) {
    if (ctx->world <= 1) {
        if (d_send != d_recv)
            M3D_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return M3D_OK;
    }
    if (ctx->nccl_comm) {
        const int ncclChar = 0; /* ncclInt8 / ncclChar */
        const int rc = g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclChar, ctx->nccl_comm, ctx->stream);
        if (rc != 0)
            return ctx->fail(M3D_ERR_NCCL, "ncclAllGather: %s",
                             g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
        return M3D_OK;
    }
    if (ctx->xfn) {
        if (ctx->x_on_device) {
            const int rc = ctx->xfn(ctx->xuser, d_send, d_recv, bytes_per_rank, 1);
            if (rc != 0) return ctx->fail(M3D_ERR_NCCL, "exchange callback failed (%d)", rc);
            return M3D_OK;
        }
        /* host callback: stage through pinned memory */
        const size_t tot = bytes_per_rank * (size_t)ctx->world;
        M3D_CUDA(ctx, ctx->h_stage.reserve(tot + bytes_per_rank));
        char *hs = ctx->h_stage.as<char>();
        M3D_CUDA(ctx, cudaMemcpyAsync(hs, d_send, bytes_per_rank, cudaMemcpyDeviceToHost, ctx->stream));
        M3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const int rc = ctx->xfn(ctx->xuser, hs, hs + bytes_per_rank, bytes_per_rank, 0);
        if (rc != 0) return ctx->fail(M3D_ERR_NCCL, "exchange callback failed (%d)", rc);
        M3D_CUDA(ctx, cudaMemcpyAsync(d_recv, hs + bytes_per_rank, tot, cudaMemcpyHostToDevice, ctx->stream));
        return M3D_OK;
    }
    return ctx->fail(M3D_ERR_NCCL, "world size %d but no exchange configured", ctx->world);
}

}  // namespace m3d

/* ---------------------------------------------------------------- staged upload of pageable buffers */
struct CopyPool {
    static constexpr size_t kPiece = 2u << 20;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;
    uint64_t job_id = 0; /* bumped per job */
    const char *src = nullptr;
    char *dst = nullptr;
    size_t bytes = 0, pieces = 0;
    std::atomic<size_t> next{0};
    std::atomic<int> active{0};
    std::vector<std::atomic<uint8_t>> done;
    CopyPool(int n) : done(4096) {
        for (int i = 0; i < n; ++i) workers.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : workers) t.join();
    }
    bool work_one() {
        const size_t i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= pieces) return false;
        const size_t off = i * kPiece, len = std::min(kPiece, bytes - off);
        memcpy(dst + off, src + off, len);
        done[i].store(1, std::memory_order_release);
        return true;
    }
    void work() {
        while (work_one()) {
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || job_id != seen; });
                if (stop) return;
                seen = job_id;
                active.fetch_add(1);
            }
            work();
            active.fetch_sub(1);
        }
    }
    void start(const void *s, void *d, size_t n) {
        while (active.load() != 0) std::this_thread::yield(); /* stragglers of the previous job */
        {
            std::lock_guard<std::mutex> lk(mu);
            src = static_cast<const char *>(s);
            dst = static_cast<char *>(d);
            bytes = n;
            pieces = (n + kPiece - 1) / kPiece;
            for (size_t i = 0; i < pieces; ++i) done[i].store(0, std::memory_order_relaxed);
            next.store(0);
            ++job_id;
        }
        cv.notify_all();
    }
};

namespace m3d {
int host_to_device(m3d_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    static const bool staged_ok = !(getenv("M3D_STAGED_UPLOAD") && atoi(getenv("M3D_STAGED_UPLOAD")) == 0);
    bool pageable = false;
    if (staged_ok && bytes >= (4u << 20) && bytes <= CopyPool::kPiece * 4096) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src) != cudaSuccess) cudaGetLastError();
        else pageable = at.type == cudaMemoryTypeUnregistered;
    }
    if (!pageable) {
        M3D_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return M3D_OK;
    }
    /* the previous staged upload may still be in flight on the copy engine (another stream): it must have drained before
     * the staging buffer is overwritten */
    M3D_CUDA(ctx, cudaStreamSynchronize(stream));
    M3D_CUDA(ctx, ctx->h_upload.reserve(bytes));
    if (!ctx->pool) {
        static const int nthreads = getenv("M3D_UPLOAD_THREADS") ? std::max(1, std::min(32, atoi(getenv("M3D_UPLOAD_THREADS")))) : 1;
        ctx->pool = new CopyPool(nthreads);
    }
    CopyPool &P = *ctx->pool;
    char *stage = ctx->h_upload.as<char>();
    P.start(src, stage, bytes);
    for (size_t i = 0; i < P.pieces; ++i) {
        while (!P.done[i].load(std::memory_order_acquire)) /* help staging (one piece at a time) until piece i is there */
            if (!P.work_one()) std::this_thread::yield();
        const size_t off = i * CopyPool::kPiece, len = std::min(CopyPool::kPiece, bytes - off);
        M3D_CUDA(ctx, cudaMemcpyAsync(static_cast<char *>(dst) + off, stage + off, len, cudaMemcpyHostToDevice, stream));
    }
    return M3D_OK;
}
}  // namespace m3d

using namespace m3d;

struct m3d_feat_scratch;
extern "C" void m3d_feat_scratch_free(m3d_feat_scratch *f);

/* fp32 FFMA throughput probe: the denominator of the scoring kernel's ALU roofline
 * (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks) */
__global__ void __launch_bounds__(256) ffma_probe_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b);
            x1 = fmaf(x1, a, b);
            x2 = fmaf(x2, a, b);
            x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b);
            x5 = fmaf(x5, a, b);
            x6 = fmaf(x6, a, b);
            x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" {

int m3d_probe_fp64_dfma(m3d_ctx *c, double *dfma_per_s) {
    if (!c || !dfma_per_s) return M3D_ERR_INVALID_ARG;
    M3D_CUDA(c, cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256, iters = 256;
    M3D_CUDA(c, c->d_tmp5.reserve(sizeof(double) * (size_t)blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        M3D_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
        dfma_probe_kernel<<<blocks, threads, 0, c->stream>>>(c->d_tmp5.as<double>(), iters, 0.999, 0.001);
        M3D_LAUNCHED(c);
        M3D_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
        M3D_CUDA(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        const double r = (double)blocks * threads * (double)iters * 64.0 / (ms * 1e-3);
        if (rep > 0 && r > best) best = r;
    }
    *dfma_per_s = best;
    return M3D_OK;
}

int m3d_probe_fp32_ffma(m3d_ctx *c, double *ffma_per_s) {
    if (!c || !ffma_per_s) return M3D_ERR_INVALID_ARG;
    M3D_CUDA(c, cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
    M3D_CUDA(c, c->d_tmp5.reserve(sizeof(float) * (size_t)blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        M3D_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
        ffma_probe_kernel<<<blocks, threads, 0, c->stream>>>(c->d_tmp5.as<float>(), iters, 0.999f, 0.001f);
        M3D_LAUNCHED(c);
        M3D_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
        M3D_CUDA(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        const double r = (double)blocks * threads * (double)iters * 64.0 / (ms * 1e-3);
        if (rep > 0 && r > best) best = r;
    }
    *ffma_per_s = best;
    return M3D_OK;
}

int m3d_abi_version(void) { return M3D_ABI_VERSION; }

int m3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int ctx_create_impl(int device, void *stream, bool own, m3d_ctx **out) {
    if (!out) return M3D_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return M3D_ERR_CUDA; /* no CPU fallback: the product path needs the GPU */
    }
    m3d_ctx *c = new m3d_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete c;
        return M3D_ERR_CUDA;
    }
    if (own) {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return M3D_ERR_CUDA;
        }
        c->own_stream = true;
    } else {
        c->stream = (cudaStream_t)stream;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    for (auto &e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            delete c;
            return M3D_ERR_CUDA;
        }
    *out = c;
    return M3D_OK;
}

int m3d_ctx_create(int device, m3d_ctx **out) { return ctx_create_impl(device, nullptr, true, out); }
int m3d_ctx_create_on_stream(int device, void *cuda_stream, m3d_ctx **out) {
    return ctx_create_impl(device, cuda_stream, false, out);
}

void m3d_host_unregister_all(m3d_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (auto &r : c->registered)
        if (cudaHostUnregister(const_cast<void *>(r.first)) != cudaSuccess) cudaGetLastError();
    c->registered.clear();
}

void m3d_ctx_destroy(m3d_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    m3d_host_unregister_all(c);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    if (c->scratch_cloud) m3d_cloud_free(c->scratch_cloud);
    DevBuf *db[] = {&c->d_samples, &c->d_counts, &c->d_counts_all, &c->d_blk, &c->d_part, &c->d_small,
                    &c->d_inl,     &c->d_models, &c->d_valid,      &c->d_tmp0, &c->d_tmp1, &c->d_tmp2,
                    &c->d_tmp3,    &c->d_tmp4,   &c->d_tmp5,      &c->d_queue,    &c->d_tiles, &c->d_rownrm, &c->d_rowmap, &c->d_recs, &c->d_draw, &c->d_mtjump,
                    &c->d_metas,   &c->d_models_all, &c->d_valid_all};
    for (auto *b : db) b->release();
    PinBuf *pb[] = {&c->h_samples, &c->h_counts, &c->h_small, &c->h_stage, &c->h_rownrm, &c->h_metas, &c->h_upload};
    for (auto *b : pb) b->release();
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c->pool;
    m3d_feat_scratch_free(c->feat);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *m3d_last_error(const m3d_ctx *c) { return c ? c->err.c_str() : "null context"; }
void *m3d_ctx_stream(const m3d_ctx *c) { return c ? (void *)c->stream : nullptr; }
uint64_t m3d_ctx_launch_count(const m3d_ctx *c) { return c ? c->launches : 0; }

int m3d_nccl_unique_id(char id[M3D_NCCL_ID_BYTES]) {
    std::string why;
    if (!nccl_load(&why)) return M3D_ERR_NCCL;
    return g_nccl.GetUniqueId(id) == 0 ? M3D_OK : M3D_ERR_NCCL;
}

int m3d_ctx_init_nccl(m3d_ctx *c, const char id[M3D_NCCL_ID_BYTES], int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return M3D_ERR_INVALID_ARG;
    std::string why;
    if (!nccl_load(&why)) return c->fail(M3D_ERR_NCCL, "%s", why.c_str());
    M3D_CUDA(c, cudaSetDevice(c->device));
    Id128 u;
    memcpy(u.b, id, 128);
    void *comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world, u, rank);
    if (rc != 0)
        return c->fail(M3D_ERR_NCCL, "ncclCommInitRank: %s",
                       g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    c->nccl_comm = comm;
    c->rank = rank;
    c->world = world;
    return M3D_OK;
}

int m3d_ctx_set_exchange(m3d_ctx *c, m3d_allgather_fn fn, void *user, int on_device, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return M3D_ERR_INVALID_ARG;
    c->xfn = fn;
    c->xuser = user;
    c->x_on_device = on_device;
    c->rank = rank;
    c->world = world;
    return M3D_OK;
}

/* RandomSampler<size_t>::operator() (utils.h:81-97): idx = rng() % size, keep if not yet drawn */
void m3d_sample_table(uint32_t seed, size_t n, int k, size_t rows, uint32_t *out) {
    SampleStream s(seed, n);
    s.draw_rows(k, rows, out);
}

void m3d_shard_rows(size_t rows, int rank, int world, uint32_t *out, size_t *n_local, size_t *padded) {
    const ShardMap sm{(uint32_t)rows, (uint32_t)std::max(world, 1), (uint32_t)std::max(rank, 0)};
    const uint32_t mine = sm.local_rows();
    if (out)
        for (uint32_t l = 0; l < mine; ++l) out[l] = sm.wave_row(l);
    if (n_local) *n_local = mine;
    if (padded) *padded = sm.padded();
}

int m3d_ordered_scan(const uint64_t *counts, const uint8_t *valid, const double *err, size_t rows,
                     size_t n_points, int k, double probability, uint64_t max_iteration,
                     m3d_ransac_stats *st) {
    if (!counts || !valid || !st) return M3D_ERR_INVALID_ARG;
    OrderedScan scan(n_points, k, probability, max_iteration);
    for (size_t i = 0; i < rows; ++i) {
        const uint64_t cnt = counts[i];
        scan.step(i, valid[i] != 0, cnt, [&](uint64_t j, bool /*exact*/, double *rmse) {
            const double e = err ? err[j] : 0.0;
            *rmse = counts[j] ? e / std::sqrt((double)counts[j]) : 1e10;
            return 0;
        });
        if (scan.stopped) break;
    }
    scan.fill(st);
    return M3D_OK;
}

} /* extern "C" */
