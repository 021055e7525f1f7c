// Compact G rows and their expansion on the host.
//
// 71 of the 104 entries of a window's Jacobian record never depend on x: the zeros of the reference's
// tabG initialisation and its +-1 entries (src/problem.cpp:1038, 1084, 1098, 1112, 1170, 1182, 1204), and
// two more equal -dt.  On the device they cost nothing extra (they sit in the record slots once), but on
// the host-pointer path they are two thirds of what crosses PCIe.  The kernels can therefore write a
// COMPACT row per trajectory,
//
//     [0, R0)                      objective-row block, as in G
//     [R0 + 31k, R0 + 31k + 31)    the 31 x-dependent entries of window k, in record order
//     [R0 + 31ts, + nbG)           boundary block, as in G
//     [R0 + 31ts + nbG]            -dt
//
// and this file turns such rows back into rows in SNOPT coordinate order: every value the device computed
// is placed where reference computeG (src/problem.cpp:782-806) puts it and the structural constants are
// written as literals.  No arithmetic on x happens here.
//
// Writes are non-temporal (the destination, 11 GB per 65,536-trajectory step, is not read back by this
// library), spread over a small pool of host threads by trajectory.
#include <emmintrin.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <utility>

#include "tolcuda_internal.h"

namespace tolcuda {

namespace {

constexpr int REC = TOLCUDA_REC;
constexpr int NVAR = TOLCUDA_NVAR;

// record position -> source: >= 0 index into the window's 31 values, -1: 0.0, -2: +1.0, -3: -1.0, -4: -dt
struct RecMap {
    int8_t m[REC];
};
constexpr RecMap make_map() {
    RecMap r{};
    for (int j = 0; j < REC; j++) r.m[j] = -1;
    // same positions as record_store / record_init in fg_kernels.cu
    constexpr int pos[NVAR] = {0,  4,  5,  6,  13, 17, 18, 19, 26, 30, 31, 39, 43, 44, 45, 47,
                               50, 52, 56, 57, 58, 59, 60, 65, 69, 70, 71, 72, 73, 78, 91};
    for (int i = 0; i < NVAR; i++) r.m[pos[i]] = (int8_t)i;
    r.m[1] = r.m[15] = r.m[29] = r.m[85] = r.m[99] = -3;
    for (int s = 0; s < TOLCUDA_PF; s++) r.m[13 * s + 12] = -2;
    r.m[87] = r.m[101] = -4;
    return r;
}
constexpr RecMap kMap = make_map();

template <int J>
inline double rec_val(const double *v, const double mdt) {
    constexpr int m = kMap.m[J];
    if constexpr (m >= 0) return v[m];
    else if constexpr (m == -1) return 0.0;
    else if constexpr (m == -2) return 1.0;
    else if constexpr (m == -3) return -1.0;
    else return mdt;
}

inline void nt1(double *p, double v) {
    long long bits;
    __builtin_memcpy(&bits, &v, 8);
    _mm_stream_si64(reinterpret_cast<long long *>(p), bits);
}

// one record, destination 16-byte aligned: 52 streaming 16-byte stores of straight-line code
template <int... P>
inline void record_aligned(double *dst, const double *v, const double mdt, std::integer_sequence<int, P...>) {
    (_mm_stream_pd(dst + 2 * P, _mm_set_pd(rec_val<2 * P + 1>(v, mdt), rec_val<2 * P>(v, mdt))), ...);
}
// destination 8 mod 16: one 8-byte store, 51 pairs shifted by one, one 8-byte store
template <int... P>
inline void record_shifted(double *dst, const double *v, const double mdt, std::integer_sequence<int, P...>) {
    nt1(dst, rec_val<0>(v, mdt));
    (_mm_stream_pd(dst + 1 + 2 * P, _mm_set_pd(rec_val<2 * P + 2>(v, mdt), rec_val<2 * P + 1>(v, mdt))), ...);
    nt1(dst + REC - 1, rec_val<REC - 1>(v, mdt));
}

// streaming copy of cnt doubles to any 8-byte aligned destination
inline void nt_copy(double *dst, const double *src, long cnt) {
    long i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) && cnt > 0) nt1(dst, src[0]), i = 1;
    for (; i + 1 < cnt; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
    if (i < cnt) nt1(dst + i, src[i]);
}

}  // namespace

long compact_len(int form, int ts) {
    int R0, nbG;
    pattern_dims(form, ts, nullptr, nullptr, nullptr, &R0, &nbG);
    return (long)R0 + (long)NVAR * ts + nbG + 1;
}

void expand_row(int form, int ts, const double *src, double *dst) {
    int R0, nbG;
    pattern_dims(form, ts, nullptr, nullptr, nullptr, &R0, &nbG);
    const double *bsrc = src + R0 + (long)NVAR * ts;
    const double mdt = bsrc[nbG];
    nt_copy(dst, src, R0);
    double *rec = dst + R0;
    const double *v = src + R0;
    if ((reinterpret_cast<uintptr_t>(rec) & 15) == 0) {
        for (int k = 0; k < ts; k++, rec += REC, v += NVAR)
            record_aligned(rec, v, mdt, std::make_integer_sequence<int, REC / 2>());
    } else {
        for (int k = 0; k < ts; k++, rec += REC, v += NVAR)
            record_shifted(rec, v, mdt, std::make_integer_sequence<int, REC / 2 - 1>());
    }
    nt_copy(rec, bsrc, nbG);
}

// ---- host thread pool ----------------------------------------------------------------------------------

struct HostPool::Impl {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    const std::function<void(long)> *fn = nullptr;
    std::atomic<long> next{0};
    long nitems = 0;
    long gen = 0;
    int busy = 0;
    bool stop = false;

    void drain() {
        for (;;) {
            const long i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= nitems) break;
            (*fn)(i);
        }
    }
    void worker() {
        long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_go.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            lk.unlock();
            drain();
            lk.lock();
            if (--busy == 0) cv_done.notify_one();
        }
    }
};

int HostPool::default_threads() {
    if (const char *env = std::getenv("TOLCUDA_HOST_THREADS")) {
        const int v = std::atoi(env);
        if (v > 0) return v;
    }
    int cores = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) cores = CPU_COUNT(&set);
    if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
    if (cores <= 0) cores = 1;
    // one process per GPU: leave the other local ranks their share of the cores
    if (const char *env = std::getenv("LOCAL_WORLD_SIZE")) {
        const int w = std::atoi(env);
        if (w > 1) cores = std::max(1, cores / w);
    }
    return std::min(cores, 64);
}

HostPool::HostPool(int threads) : impl_(new Impl), threads_(std::max(1, threads)) {
    for (int t = 1; t < threads_; t++) impl_->th.emplace_back([this] { impl_->worker(); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv_go.notify_all();
    for (auto &t : impl_->th) t.join();
    delete impl_;
}

void HostPool::parallel_for(long n, const std::function<void(long)> &fn) {
    if (n <= 0) return;
    Impl &p = *impl_;
    {
        std::lock_guard<std::mutex> lk(p.mu);
        p.fn = &fn;
        p.nitems = n;
        p.next.store(0, std::memory_order_relaxed);
        p.busy = (int)p.th.size();
        p.gen++;
    }
    p.cv_go.notify_all();
    p.drain();  // the calling thread works too
    std::unique_lock<std::mutex> lk(p.mu);
    p.cv_done.wait(lk, [&] { return p.busy == 0; });
    p.fn = nullptr;
}

void expand_rows(HostPool &pool, int form, int ts, long B, const double *Gc, long ldGc, double *G, long ldG) {
    constexpr long BLK = 4;  // trajectories per work item
    const std::function<void(long)> job = [&](long item) {
        const long b1 = std::min(B, (item + 1) * BLK);
        for (long b = item * BLK; b < b1; b++) expand_row(form, ts, Gc + b * ldGc, G + b * ldG);
        _mm_sfence();  // streaming stores are weakly ordered: make them visible before the item counts as done
    };
    pool.parallel_for((B + BLK - 1) / BLK, job);
}

}  // namespace tolcuda
