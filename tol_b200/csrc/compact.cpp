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
#include <immintrin.h>
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
    for (int j = 0; j < REC; j++) r.m[j] = (int8_t)tolcuda_rec_kind(j);  // the one table, fg_const.h
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

// NT = true: non-temporal stores (batch path: the rows are not read back by this library and are far larger than
// the caches); NT = false: ordinary stores (single-trajectory callback: SNOPT reads G right away)
template <bool NT>
inline void put1(double *p, double v) {
    if (NT) {
        long long bits;
        __builtin_memcpy(&bits, &v, 8);
        _mm_stream_si64(reinterpret_cast<long long *>(p), bits);
    } else {
        *p = v;
    }
}
template <bool NT>
inline void put2(double *p, __m128d v) {
    if (NT) _mm_stream_pd(p, v);
    else _mm_storeu_pd(p, v);
}

// one record, destination 16-byte aligned: 52 streaming 16-byte stores of straight-line code
template <bool NT, int... P>
inline void record_aligned(double *dst, const double *v, const double mdt, std::integer_sequence<int, P...>) {
    (put2<NT>(dst + 2 * P, _mm_set_pd(rec_val<2 * P + 1>(v, mdt), rec_val<2 * P>(v, mdt))), ...);
}
// destination 8 mod 16: one 8-byte store, 51 pairs shifted by one, one 8-byte store
template <bool NT, int... P>
inline void record_shifted(double *dst, const double *v, const double mdt, std::integer_sequence<int, P...>) {
    put1<NT>(dst, rec_val<0>(v, mdt));
    (put2<NT>(dst + 1 + 2 * P, _mm_set_pd(rec_val<2 * P + 2>(v, mdt), rec_val<2 * P + 1>(v, mdt))), ...);
    put1<NT>(dst + REC - 1, rec_val<REC - 1>(v, mdt));
}

// copy of cnt doubles to any 8-byte aligned destination
template <bool NT>
inline void copy_out(double *dst, const double *src, long cnt) {
    long i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) && cnt > 0) put1<NT>(dst, src[0]), i = 1;
    for (; i + 1 < cnt; i += 2) put2<NT>(dst + i, _mm_loadu_pd(src + i));
    if (i < cnt) put1<NT>(dst + i, src[i]);
}

// ---- AVX-512 path ---------------------------------------------------------------------------------------
//
// The record region of a row is a periodic stream: output double j (counted from the first record) holds
// record position j mod 104, and 104 doubles are exactly 13 cache lines.  Starting at the first 64-byte
// aligned address inside the region (Q doubles in), line l of every 13-line period therefore has a fixed
// 8-lane layout: which lanes take the next values of the compact stream (mask) and which constants sit in
// the others (template).  One masked expand-load (the next popcount(mask) compact values into the masked
// lanes, the template elsewhere) and one streaming 64-byte store produce a line.
struct LineTab {
    uint8_t var[13];   // lanes fed from the compact stream
    uint8_t mdt[13];   // lanes holding -dt
    double tmpl[13][8];  // constants (0 in the var and -dt lanes)
    int8_t cnt[13];    // popcount(var)
};
constexpr LineTab make_lines(int q) {
    LineTab t{};
    for (int l = 0; l < 13; l++) {
        for (int e = 0; e < 8; e++) {
            const int m = kMap.m[(q + 8 * l + e) % REC];
            if (m >= 0) t.var[l] |= (uint8_t)(1u << e), t.cnt[l]++;
            else if (m == -4) t.mdt[l] |= (uint8_t)(1u << e);
            t.tmpl[l][e] = m == -2 ? 1.0 : (m == -3 ? -1.0 : 0.0);
        }
    }
    return t;
}
template <int Q>
struct Lines {
    static constexpr LineTab tab = make_lines(Q);
};

inline double rec_val_rt(int pos, const double *&v, const double mdt) {
    const int m = kMap.m[pos];
    if (m >= 0) return *v++;
    return m == -1 ? 0.0 : (m == -2 ? 1.0 : (m == -3 ? -1.0 : mdt));
}

// records of one row: dst = first record (8-byte aligned, Q doubles before the next 64-byte boundary),
// v = the row's compact window values
template <bool NT>
__attribute__((target("avx512f"))) inline void put8(double *p, __m512d v) {
    if (NT) _mm512_stream_pd(p, v);
    else _mm512_store_pd(p, v);
}

template <int Q, bool NT>
__attribute__((target("avx512f"))) void records_avx512(double *dst, const double *v, const double mdt, const int ts) {
    constexpr const LineTab &T = Lines<Q>::tab;
    const long total = (long)REC * ts;
    long j = 0;
    for (; j < Q && j < total; j++) put1<NT>(dst + j, rec_val_rt((int)(j % REC), v, mdt));
    if (j == total) return;
    __m512d tm[13];
    const __m512d vm = _mm512_set1_pd(mdt);
#pragma GCC unroll 13
    for (int l = 0; l < 13; l++) tm[l] = _mm512_mask_blend_pd((__mmask8)T.mdt[l], _mm512_loadu_pd(T.tmpl[l]), vm);
    double *out = dst + Q;
    const long nlines = (total - Q) / 8;
    long i = 0;
    for (; i + 13 <= nlines; i += 13) {
#pragma GCC unroll 13
        for (int l = 0; l < 13; l++) {
            put8<NT>(out + 8 * l, _mm512_mask_expandloadu_pd(tm[l], (__mmask8)T.var[l], v));
            v += T.cnt[l];
        }
        out += REC;
    }
    for (int l = 0; i < nlines; i++, l++, out += 8) {
        put8<NT>(out, _mm512_mask_expandloadu_pd(tm[l], (__mmask8)T.var[l], v));
        v += T.cnt[l];
    }
    for (j = Q + 8 * nlines; j < total; j++) put1<NT>(dst + j, rec_val_rt((int)(j % REC), v, mdt));
}

// AVX-512 available and not switched off (TOLCUDA_NO_AVX512, for tests of the SSE2 path)
bool have_avx512() { return __builtin_cpu_supports("avx512f") && !std::getenv("TOLCUDA_NO_AVX512"); }

template <bool NT>
void records_dispatch_avx512(double *dst, const double *v, const double mdt, const int ts) {
    switch ((8 - (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 7)) & 7) {
        case 0: return records_avx512<0, NT>(dst, v, mdt, ts);
        case 1: return records_avx512<1, NT>(dst, v, mdt, ts);
        case 2: return records_avx512<2, NT>(dst, v, mdt, ts);
        case 3: return records_avx512<3, NT>(dst, v, mdt, ts);
        case 4: return records_avx512<4, NT>(dst, v, mdt, ts);
        case 5: return records_avx512<5, NT>(dst, v, mdt, ts);
        case 6: return records_avx512<6, NT>(dst, v, mdt, ts);
        default: return records_avx512<7, NT>(dst, v, mdt, ts);
    }
}

template <bool NT>
void expand_row_as(int form, int ts, const double *src, double *dst, bool wide) {
    int R0, nbG;
    pattern_dims(form, ts, nullptr, nullptr, nullptr, &R0, &nbG);
    const double *bsrc = src + R0 + (long)NVAR * ts;
    const double mdt = bsrc[nbG];
    copy_out<NT>(dst, src, R0);
    double *rec = dst + R0;
    const double *v = src + R0;
    if (wide) {
        records_dispatch_avx512<NT>(rec, v, mdt, ts);
        rec += (long)REC * ts;
    } else if ((reinterpret_cast<uintptr_t>(rec) & 15) == 0) {
        for (int k = 0; k < ts; k++, rec += REC, v += NVAR)
            record_aligned<NT>(rec, v, mdt, std::make_integer_sequence<int, REC / 2>());
    } else {
        for (int k = 0; k < ts; k++, rec += REC, v += NVAR)
            record_shifted<NT>(rec, v, mdt, std::make_integer_sequence<int, REC / 2 - 1>());
    }
    copy_out<NT>(rec, bsrc, nbG);
}

}  // namespace

long compact_len(int form, int ts) {
    int R0, nbG;
    pattern_dims(form, ts, nullptr, nullptr, nullptr, &R0, &nbG);
    return (long)R0 + (long)NVAR * ts + nbG + 1;
}

void expand_row(int form, int ts, const double *src, double *dst) { expand_row(form, ts, src, dst, have_avx512()); }

void expand_row(int form, int ts, const double *src, double *dst, bool wide) {
    expand_row_as<true>(form, ts, src, dst, wide);
}

void expand_row_cached(int form, int ts, const double *src, double *dst) {
    expand_row_as<false>(form, ts, src, dst, have_avx512());
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
    // the expansion is bound by host memory bandwidth: 8 threads reach it on the B200 host (tools/expandbw.py)
    return std::min(cores, 8);
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
    const bool wide = have_avx512();
    const std::function<void(long)> job = [&](long item) {
        const long b1 = std::min(B, (item + 1) * BLK);
        for (long b = item * BLK; b < b1; b++) expand_row(form, ts, Gc + b * ldGc, G + b * ldG, wide);
        _mm_sfence();  // streaming stores are weakly ordered: make them visible before the item counts as done
    };
    pool.parallel_for((B + BLK - 1) / BLK, job);
}

}  // namespace tolcuda
