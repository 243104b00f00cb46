// mss_ood_metrics_dist -- the multi-GPU exact metric behind ONE C entry point, for callers that bind
// include/mss_b200.h directly and bring their own NCCL communicator (one process or thread per GPU).  It is the
// "nccl" exchange of evaluator.StreamingEvaluator restated in C++ on top of the stage-level entry points:
//   1. all-reduce of (count, positives, nan, inf)                         -> None / ValueError decisions
//   2. all-reduce of the sampled top-16-bit key histogram                 -> identical splitters on every rank
//   3. local stable partition of both streams + grouped ncclSend/ncclRecv -> rank r owns key range r
//   4. local sort + merge-path counts with the global prefixes
//   5. all-gather of the per-threshold (tps, fps)                         -> the same float64 tail on every rank
// NCCL is not linked: the symbols are resolved at run time from the NCCL the caller's process has loaded (the
// communicator was created by it), so the library itself keeps depending on the CUDA runtime only.
// Temporaries come from the stream-ordered allocator (cudaMallocAsync) and are freed before returning: received sizes
// are only known after the exchange, so a caller-sized workspace would have to assume the worst case.
#include <dlfcn.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace mss {

typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_sendrecv_fn)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_group_fn)(void);
constexpr int NCCL_INT64 = 4, NCCL_UINT32 = 3, NCCL_SUM = 0;      // ncclDataType_t / ncclRedOp_t values (stable ABI)

struct Nccl {
    nccl_allreduce_fn all_reduce = nullptr;
    nccl_allgather_fn all_gather = nullptr;
    nccl_sendrecv_fn send = nullptr, recv = nullptr;
    nccl_group_fn group_start = nullptr, group_end = nullptr;
    bool ok() const { return all_reduce && all_gather && send && recv && group_start && group_end; }
};

static const Nccl &nccl() {
    static const Nccl n = [] {
        Nccl x;
        void *h = RTLD_DEFAULT;
        if (!dlsym(h, "ncclAllReduce")) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            x.all_reduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
            x.all_gather = (nccl_allgather_fn)dlsym(h, "ncclAllGather");
            x.send = (nccl_sendrecv_fn)dlsym(h, "ncclSend");
            x.recv = (nccl_sendrecv_fn)dlsym(h, "ncclRecv");
            x.group_start = (nccl_group_fn)dlsym(h, "ncclGroupStart");
            x.group_end = (nccl_group_fn)dlsym(h, "ncclGroupEnd");
        }
        return x;
    }();
    return n;
}

#define MSS_CHECK_NCCL(expr)                                                                   \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != 0) {                                                                         \
            mss::set_error("%s failed with ncclResult %d (%s:%d)", #expr, _r, __FILE__, __LINE__); \
            return MSS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// device temporaries of one call, released on every exit path
struct Temps {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Temps(cudaStream_t s) : st(s) {}
    template <typename T>
    T *get(size_t n) {
        void *p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T *)p;
    }
    ~Temps() {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
};

// Key-range splitters from the GLOBAL top-16-bit histogram: world - 1 ascending uint32 keys that balance the key counts as
// evenly as bin granularity allows (the integer arithmetic of evaluator.choose_splitters: identical inputs give identical
// splitters on every rank).
static void choose_splitters(const long long *hist, int bins_log2, int world, std::vector<uint32_t> &out) {
    const int bins = 1 << bins_log2;
    long long total = 0;
    for (int b = 0; b < bins; b++) total += hist[b];
    out.clear();
    long long cum = 0;
    int b = 0;
    for (int j = 1; j < world; j++) {
        const long long target = (total * j + world - 1) / world;          // ceil(total * j / world)
        while (b < bins && cum + hist[b] < target) cum += hist[b++];       // first bin whose cumulative count reaches target
        const long long boundary = std::min<long long>(b + 1, bins);       // the splitter sits behind that bin
        const unsigned long long key = (unsigned long long)boundary << (32 - bins_log2);
        out.push_back((uint32_t)std::min<unsigned long long>(key, 0xFFFFFFFFull));
    }
}

// Bins a splitter would have to fall INSIDE of: the bin in front of a chosen boundary when it holds more than half of one
// rank's share (evaluator.heavy_bins).
static std::vector<int> heavy_bins(const long long *hist, int bins_log2, int world, const std::vector<uint32_t> &spl) {
    const int bins = 1 << bins_log2;
    long long total = 0;
    for (int b = 0; b < bins; b++) total += hist[b];
    const long long share = std::max<long long>(total / std::max(world, 1), 1);
    std::vector<int> out;
    for (uint32_t s : spl) {
        const int b = (int)std::min<long long>((long long)(s >> (32 - bins_log2)), bins) - 1;
        if (b >= 0 && b < bins && hist[b] * 2 > share && std::find(out.begin(), out.end(), b) == out.end()) out.push_back(b);
    }
    std::sort(out.begin(), out.end());
    return out;
}

// Splitters with a second level inside the heavy bins (evaluator.refine_splitters, same integer arithmetic).
// fine[i * 65536 ...] = global histogram of the low 16 key bits inside heavy[i].
static void refine_splitters(const long long *hist, int world, const std::vector<int> &heavy, const std::vector<long long> &fine,
                             std::vector<uint32_t> &out) {
    const int bins = 1 << 16;
    long long total = 0;
    for (int b = 0; b < bins; b++) total += hist[b];
    out.clear();
    long long cum = 0;                                                     // keys in bins < b
    int b = 0;
    for (int j = 1; j < world; j++) {
        const long long target = (total * j + world - 1) / world;
        while (b < bins - 1 && cum + hist[b] < target) cum += hist[b++];
        unsigned long long key = (unsigned long long)(b + 1) << 16;
        const auto it = std::find(heavy.begin(), heavy.end(), b);
        if (it != heavy.end()) {
            const long long *f = fine.data() + (size_t)(it - heavy.begin()) * bins;
            long long tot_f = 0;
            for (int i = 0; i < bins; i++) tot_f += f[i];
            if (tot_f > 0) {
                const long long inside = target - cum, hb = std::max<long long>(hist[b], 1);
                const long long want = std::max<long long>((inside * tot_f + hist[b] - 1) / hb, 1);
                long long fc = 0;
                int i = 0;
                while (i < bins && fc + f[i] < want) fc += f[i++];         // first low value whose cumulative count reaches want
                key = ((unsigned long long)b << 16) + (unsigned long long)std::min(i + 1, bins);
            }
        }
        uint32_t k32 = (uint32_t)std::min<unsigned long long>(key, 0xFFFFFFFFull);
        if (!out.empty()) k32 = std::max(k32, out.back());
        out.push_back(k32);
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_ood_metrics_dist(const mss_eval_buffers *ev, void *nccl_comm, int rank, int world, double out_host[3],
                                    int64_t counts_host[4], void *stream) {
    MSS_REQUIRE(ev && ev->keys && ev->state && nccl_comm && out_host && world >= 1 && rank >= 0 && rank < world && world <= 256,
                "mss_ood_metrics_dist: bad arguments");
    const Nccl &nc = nccl();
    if (!nc.ok()) {
        set_error("mss_ood_metrics_dist: NCCL symbols not found in this process (load the NCCL that created the communicator first)");
        return MSS_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Temps tmp(st);

    // 1. global emptiness / finiteness decisions
    int64_t s4[4];
    int rc = mss_eval_state_host(ev, s4, stream);
    if (rc) return rc;
    const int64_t m = s4[0], n_pos = s4[1], n_neg = m - n_pos;
    long long *d_small = tmp.get<long long>(4 + 2 * world + 2 * world * world + 2 * world);
    MSS_REQUIRE(d_small, "mss_ood_metrics_dist: cudaMallocAsync failed");
    long long h4[4] = {(long long)s4[0], (long long)s4[1], (long long)s4[2], (long long)s4[3]};
    MSS_CHECK_CUDA(cudaMemcpyAsync(d_small, h4, sizeof(h4), cudaMemcpyHostToDevice, st));
    MSS_CHECK_NCCL(nc.all_reduce(d_small, d_small, 4, NCCL_INT64, NCCL_SUM, nccl_comm, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(h4, d_small, sizeof(h4), cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    const long long M = h4[0], P = h4[1];
    if (P == 0 || P == M) return MSS_EMPTY_CLASS;
    if (h4[2]) { set_error("Input contains NaN."); return MSS_ERR_NAN; }
    if (h4[3]) { set_error("Input contains infinity or a value too large for dtype('float32')."); return MSS_ERR_INF; }

    // 2. sampled histogram of both streams -> all-reduce -> splitters
    const uint32_t *neg = ev->keys, *pos = ev->keys + (ev->capacity - n_pos);
    constexpr int BITS = 16;
    long long *d_hist = tmp.get<long long>(2 << BITS);
    MSS_REQUIRE(d_hist, "mss_ood_metrics_dist: cudaMallocAsync failed");
    const int every = (int)std::max<long long>(1, M / ((long long)world << 24));
    rc = mss_keys_histogram_sampled(neg, n_neg, BITS, every, (int64_t *)d_hist, stream);
    if (rc) return rc;
    rc = mss_keys_histogram_sampled(pos, n_pos, BITS, every, (int64_t *)d_hist + (1 << BITS), stream);
    if (rc) return rc;
    MSS_CHECK_NCCL(nc.all_reduce(d_hist, d_hist, (size_t)2 << BITS, NCCL_INT64, NCCL_SUM, nccl_comm, st));
    std::vector<long long> h_hist((size_t)2 << BITS);
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_hist.data(), d_hist, h_hist.size() * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < (1 << BITS); b++) h_hist[b] += h_hist[(1 << BITS) + b];
    std::vector<uint32_t> spl;
    choose_splitters(h_hist.data(), BITS, world, spl);
    // saturated / narrow-range scores: second-level histogram inside the bins a splitter should cut through
    const std::vector<int> heavy = heavy_bins(h_hist.data(), BITS, world, spl);
    if (!heavy.empty()) {
        std::vector<long long> fine(heavy.size() << BITS), h2((size_t)2 << BITS);
        for (size_t i = 0; i < heavy.size(); i++) {
            rc = mss_keys_histogram_refine(neg, n_neg, (unsigned)heavy[i], every, (int64_t *)d_hist, stream);
            if (rc) return rc;
            rc = mss_keys_histogram_refine(pos, n_pos, (unsigned)heavy[i], every, (int64_t *)d_hist + (1 << BITS), stream);
            if (rc) return rc;
            MSS_CHECK_NCCL(nc.all_reduce(d_hist, d_hist, (size_t)2 << BITS, NCCL_INT64, NCCL_SUM, nccl_comm, st));
            MSS_CHECK_CUDA(cudaMemcpyAsync(h2.data(), d_hist, h2.size() * 8, cudaMemcpyDeviceToHost, st));
            MSS_CHECK_CUDA(cudaStreamSynchronize(st));
            for (int b = 0; b < (1 << BITS); b++) fine[(i << BITS) + b] = h2[b] + h2[(1 << BITS) + b];
        }
        refine_splitters(h_hist.data(), world, heavy, fine, spl);
    }
    uint32_t *d_spl = tmp.get<uint32_t>(world);
    MSS_REQUIRE(d_spl, "mss_ood_metrics_dist: cudaMallocAsync failed");
    if (world > 1) MSS_CHECK_CUDA(cudaMemcpyAsync(d_spl, spl.data(), (size_t)(world - 1) * 4, cudaMemcpyHostToDevice, st));

    // 3. local partition of both streams by destination rank
    uint32_t *pk_neg = tmp.get<uint32_t>((size_t)n_neg), *pk_pos = tmp.get<uint32_t>((size_t)n_pos);
    const size_t pws_b = std::max(mss_partition_workspace_bytes(n_neg, world), mss_partition_workspace_bytes(n_pos, world));
    char *pws = tmp.get<char>(pws_b);
    MSS_REQUIRE(pk_neg && pk_pos && pws, "mss_ood_metrics_dist: cudaMallocAsync failed");
    std::vector<int64_t> send(2 * (size_t)world, 0);
    rc = mss_partition_keys(neg, n_neg, d_spl, world, pk_neg, send.data(), pws, pws_b, stream);
    if (rc) return rc;
    rc = mss_partition_keys(pos, n_pos, d_spl, world, pk_pos, send.data() + world, pws, pws_b, stream);
    if (rc) return rc;
    // all-gather of the bucket sizes: counts[src][stream][dst]
    long long *d_cnt = d_small + 4, *d_all = d_cnt + 2 * world;
    std::vector<long long> h_send(send.begin(), send.end()), h_all((size_t)2 * world * world);
    MSS_CHECK_CUDA(cudaMemcpyAsync(d_cnt, h_send.data(), h_send.size() * 8, cudaMemcpyHostToDevice, st));
    MSS_CHECK_NCCL(nc.all_gather(d_cnt, d_all, (size_t)2 * world, NCCL_INT64, nccl_comm, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(h_all.data(), d_all, h_all.size() * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    auto cnt = [&](int src, int stream_id, int dst) { return h_all[((size_t)src * 2 + stream_id) * world + dst]; };
    long long m2_neg = 0, m2_pos = 0, neg_before = 0, pos_before = 0;
    for (int s = 0; s < world; s++) { m2_neg += cnt(s, 0, rank); m2_pos += cnt(s, 1, rank); }
    for (int d = 0; d < rank; d++)
        for (int s = 0; s < world; s++) { neg_before += cnt(s, 0, d); pos_before += cnt(s, 1, d); }

    // exchange: grouped send / recv (an all-to-all with per-peer sizes), one group per stream
    uint32_t *r_neg = tmp.get<uint32_t>((size_t)m2_neg), *r_pos = tmp.get<uint32_t>((size_t)m2_pos);
    MSS_REQUIRE(r_neg && r_pos, "mss_ood_metrics_dist: cudaMallocAsync failed");
    for (int sid = 0; sid < 2; sid++) {
        uint32_t *src_buf = sid ? pk_pos : pk_neg, *dst_buf = sid ? r_pos : r_neg;
        MSS_CHECK_NCCL(nc.group_start());
        long long so = 0, ro = 0;
        for (int p = 0; p < world; p++) {
            const long long sc = cnt(rank, sid, p), rcv = cnt(p, sid, rank);
            if (sc) MSS_CHECK_NCCL(nc.send(src_buf + so, (size_t)sc, NCCL_UINT32, p, nccl_comm, st));
            if (rcv) MSS_CHECK_NCCL(nc.recv(dst_buf + ro, (size_t)rcv, NCCL_UINT32, p, nccl_comm, st));
            so += sc;
            ro += rcv;
        }
        MSS_CHECK_NCCL(nc.group_end());
    }

    // 4. local sort + counts with the global prefixes
    const size_t sws_b = mss_sort_keys_workspace_bytes(m2_neg + m2_pos);
    char *sws = tmp.get<char>(sws_b);
    MSS_REQUIRE(sws, "mss_ood_metrics_dist: cudaMallocAsync failed");
    rc = mss_sort_keys(r_neg, m2_neg, r_pos, m2_pos, sws, sws_b, stream);
    if (rc) return rc;
    const long long m2 = m2_neg + m2_pos;
    long long *tps = tmp.get<long long>((size_t)m2), *fps = tmp.get<long long>((size_t)m2);
    const size_t cws_b = mss_counts_workspace_bytes(m2);
    char *cws = tmp.get<char>(cws_b);
    MSS_REQUIRE(tps && fps && cws, "mss_ood_metrics_dist: cudaMallocAsync failed");
    int64_t T_local = 0;
    rc = mss_counts_from_sorted(r_neg, m2_neg, r_pos, m2_pos, pos_before, neg_before, (int64_t *)tps, (int64_t *)fps, &T_local, cws,
                                cws_b, stream);
    if (rc) return rc;

    // 5. all-gather the thresholds (padded to the longest slice), the identical tail on every rank
    long long *d_T = d_all + 2 * world * world;
    long long hT = T_local;
    std::vector<long long> Ts(world);
    MSS_CHECK_CUDA(cudaMemcpyAsync(d_T, &hT, 8, cudaMemcpyHostToDevice, st));
    MSS_CHECK_NCCL(nc.all_gather(d_T, d_T + world, 1, NCCL_INT64, nccl_comm, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(Ts.data(), d_T + world, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
    MSS_CHECK_CUDA(cudaStreamSynchronize(st));
    long long Tmax = 1, T = 0;
    for (int r = 0; r < world; r++) { Tmax = std::max(Tmax, Ts[r]); T += Ts[r]; }
    long long *pad = tmp.get<long long>((size_t)2 * Tmax), *gathered = tmp.get<long long>((size_t)2 * Tmax * world);
    long long *tps_all = tmp.get<long long>((size_t)T), *fps_all = tmp.get<long long>((size_t)T);
    MSS_REQUIRE(pad && gathered && tps_all && fps_all, "mss_ood_metrics_dist: cudaMallocAsync failed");
    MSS_CHECK_CUDA(cudaMemsetAsync(pad, 0, (size_t)2 * Tmax * 8, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(pad, tps, (size_t)T_local * 8, cudaMemcpyDeviceToDevice, st));
    MSS_CHECK_CUDA(cudaMemcpyAsync(pad + Tmax, fps, (size_t)T_local * 8, cudaMemcpyDeviceToDevice, st));
    MSS_CHECK_NCCL(nc.all_gather(pad, gathered, (size_t)2 * Tmax, NCCL_INT64, nccl_comm, st));
    long long off = 0;
    for (int r = 0; r < world; r++) {
        if (Ts[r]) {
            MSS_CHECK_CUDA(cudaMemcpyAsync(tps_all + off, gathered + (size_t)r * 2 * Tmax, (size_t)Ts[r] * 8, cudaMemcpyDeviceToDevice, st));
            MSS_CHECK_CUDA(cudaMemcpyAsync(fps_all + off, gathered + (size_t)r * 2 * Tmax + Tmax, (size_t)Ts[r] * 8, cudaMemcpyDeviceToDevice, st));
        }
        off += Ts[r];
    }
    const size_t tws_b = mss_tail_workspace_bytes(T);
    char *tws = tmp.get<char>(tws_b);
    MSS_REQUIRE(tws, "mss_ood_metrics_dist: cudaMallocAsync failed");
    int64_t T_roc = 0;
    rc = mss_metrics_tail((const int64_t *)tps_all, (const int64_t *)fps_all, T, 0.95, tws, tws_b, out_host, &T_roc, stream);
    if (rc) return rc;
    if (counts_host) { counts_host[0] = P; counts_host[1] = M - P; counts_host[2] = T; counts_host[3] = T_roc; }
    return MSS_OK;
}
