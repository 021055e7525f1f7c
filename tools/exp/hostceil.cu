// What the HOST side of the box can sustain, measured with nothing of libtolcuda in the way: the ceiling bench.py's
// end-to-end number (`e2e.ceiling`) is judged against.
//
//   d2h / h2d / bidir   plain cudaMemcpyAsync of `--mb` MiB pinned blocks on G GPUs CONCURRENTLY (one stream per
//                       GPU, one copy command per block), aggregate GB/s from first enqueue to last completion
//   fill                T host threads writing their own slabs with non-temporal 64-byte stores (the expansion's
//                       store pattern), GB/s
//   copy                T host threads memcpy-ing slab to slab (read + write counted)
//   d2h+fill            both at once: DMA ingest while the threads fill -- the compact-rows path's mix
//
//   ./hostceil [--gpus G] [--mb 1024] [--threads T] [--reps 3]      prints one JSON object
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                       \
            std::exit(2);                                                                          \
        }                                                                                          \
    } while (0)

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Dev {
    int id;
    cudaStream_t s_out, s_in;
    char *d_out, *d_in, *h_out, *h_in;
};

__attribute__((target("avx2"))) static void nt_fill(char *p, size_t bytes, long long v) {
    const __m256i w = _mm256_set1_epi64x(v);
    for (size_t i = 0; i + 64 <= bytes; i += 64) {
        _mm256_stream_si256(reinterpret_cast<__m256i *>(p + i), w);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(p + i + 32), w);
    }
    _mm_sfence();
}

template <class Fn>
static double on_threads(int T, Fn fn) {
    std::vector<std::thread> th;
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            ready++;
            while (!go.load(std::memory_order_acquire)) {
            }
            fn(t);
        });
    while (ready.load() < T) {
    }
    const double t0 = now();
    go.store(true, std::memory_order_release);
    for (auto &t : th) t.join();
    return now() - t0;
}

int main(int argc, char **argv) {
    int G = 1, T = 0, reps = 3;
    size_t mb = 1024;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--gpus")) G = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--mb")) mb = std::atol(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--threads")) T = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--reps")) reps = std::atoi(argv[i + 1]);
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (G > ndev) G = ndev;
    if (T <= 0) T = (int)std::thread::hardware_concurrency();
    const size_t bytes = mb << 20;
    std::vector<Dev> dv(G);
    for (int g = 0; g < G; g++) {
        Dev &d = dv[g];
        d.id = g;
        CK(cudaSetDevice(g));
        CK(cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking));
        CK(cudaMalloc(&d.d_out, bytes));
        CK(cudaMalloc(&d.d_in, bytes));
        CK(cudaMallocHost(&d.h_out, bytes));
        CK(cudaMallocHost(&d.h_in, bytes));
        CK(cudaMemset(d.d_out, 1, bytes));
        std::memset(d.h_in, 2, bytes);
        std::memset(d.h_out, 3, bytes);
    }
    auto sync_all = [&] {
        for (Dev &d : dv) {
            CK(cudaSetDevice(d.id));
            CK(cudaStreamSynchronize(d.s_out));
            CK(cudaStreamSynchronize(d.s_in));
        }
    };
    auto dma = [&](bool out, bool in) {  // best of reps, aggregate GB/s
        double best = 1e30;
        for (int r = 0; r < reps + 1; r++) {
            sync_all();
            const double t0 = now();
            for (Dev &d : dv) {
                CK(cudaSetDevice(d.id));
                if (out) CK(cudaMemcpyAsync(d.h_out, d.d_out, bytes, cudaMemcpyDeviceToHost, d.s_out));
                if (in) CK(cudaMemcpyAsync(d.d_in, d.h_in, bytes, cudaMemcpyHostToDevice, d.s_in));
            }
            sync_all();
            const double dt = now() - t0;
            if (r > 0 && dt < best) best = dt;
        }
        return (double)bytes * G * ((out ? 1 : 0) + (in ? 1 : 0)) / best / 1e9;
    };
    const double d2h = dma(true, false), h2d = dma(false, true), bidir = dma(true, true);

    // host threads: slabs of 256 MiB each (pageable, first-touched by their thread)
    const size_t slab = (size_t)256 << 20;
    std::vector<char *> a(T), b(T);
    on_threads(T, [&](int t) {
        a[t] = (char *)aligned_alloc(4096, slab);
        b[t] = (char *)aligned_alloc(4096, slab);
        std::memset(a[t], 1, slab);
        std::memset(b[t], 2, slab);
    });
    double fill = 0, copy = 0;
    for (int r = 0; r < reps; r++) {
        double dt = on_threads(T, [&](int t) { nt_fill(a[t], slab, r + 5); });
        fill = std::max(fill, (double)slab * T / dt / 1e9);
        dt = on_threads(T, [&](int t) { std::memcpy(b[t], a[t], slab); });
        copy = std::max(copy, 2.0 * slab * T / dt / 1e9);
    }
    // DMA ingest while the threads fill: the DMA engines loop until the threads are done
    double mix_dma = 0, mix_fill = 0;
    for (int r = 0; r < reps; r++) {
        std::atomic<bool> stop{false};
        std::atomic<long> copies{0};
        std::thread pump([&] {
            while (!stop.load()) {
                for (Dev &d : dv) {
                    cudaSetDevice(d.id);
                    cudaMemcpyAsync(d.h_out, d.d_out, bytes, cudaMemcpyDeviceToHost, d.s_out);
                }
                for (Dev &d : dv) {
                    cudaSetDevice(d.id);
                    cudaStreamSynchronize(d.s_out);
                }
                copies += G;
            }
        });
        while (copies.load() < G) {
        }  // the DMA stream is running
        const long c0 = copies.load();
        const double t0 = now();
        const int loops = 4;
        const double dt = on_threads(T, [&](int t) {
            for (int l = 0; l < loops; l++) nt_fill(a[t], slab, l + r);
        });
        const double t1 = now();
        const long c1 = copies.load();
        stop.store(true);
        pump.join();
        mix_fill = std::max(mix_fill, (double)slab * T * loops / dt / 1e9);
        mix_dma = std::max(mix_dma, (double)bytes * (c1 - c0) / (t1 - t0) / 1e9);
    }
    std::printf("{\"gpus\": %d, \"threads\": %d, \"block_mb\": %zu, \"d2h_GBps\": %.1f, \"h2d_GBps\": %.1f, \"bidir_GBps\": %.1f, "
                "\"fill_nt_GBps\": %.1f, \"memcpy_rw_GBps\": %.1f, \"mix_d2h_GBps\": %.1f, \"mix_fill_GBps\": %.1f}\n",
                G, T, mb, d2h, h2d, bidir, fill, copy, mix_dma, mix_fill);
    return 0;
}
