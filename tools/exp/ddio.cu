// Does a device->host DMA into a SMALL, constantly reused pinned ring stay in the host's last-level cache (Intel DDIO
// write-update), i.e. off the DRAM bus?  Probe: one thread streams D2H copies of 1 MB pieces into a ring of R bytes
// and reads every piece back; H "hog" threads fill a 2 GB buffer with non-temporal stores (the expansion's write
// stream).  If the ring is LLC-resident the hogs and the copies should both run faster with a small ring than with
// a multi-GB one.  Usage: ddio [hog_threads=8] [seconds=1.5]
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int H = argc > 1 ? std::atoi(argv[1]) : 8;
    const double T = argc > 2 ? std::atof(argv[2]) : 1.5;
    const size_t PIECE = 1 << 20, DEV = 256 << 20, HOG = (size_t)2 << 30;
    char *d;
    CK(cudaMalloc(&d, DEV));
    CK(cudaMemset(d, 1, DEV));
    char *ring, *hog;
    const size_t RMAX = (size_t)4 << 30;
    CK(cudaMallocHost(&ring, RMAX));
    CK(cudaMallocHost(&hog, HOG));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const size_t rings[] = {(size_t)2 << 20, (size_t)4 << 20, (size_t)8 << 20, (size_t)16 << 20, (size_t)64 << 20, (size_t)512 << 20, RMAX};
    for (int with_hog = 0; with_hog < 2; with_hog++)
        for (size_t R : rings) {
            std::atomic<bool> stop{false};
            std::atomic<long long> hog_bytes{0};
            std::vector<std::thread> th;
            if (with_hog)
                for (int t = 0; t < H; t++)
                    th.emplace_back([&, t]() {
                        const size_t per = HOG / H;
                        double *p = (double *)(hog + t * per);
                        const __m256d v = _mm256_set1_pd(1.5);
                        long long done = 0;
                        const size_t blk = (32u << 20) / 8, nblk = per / 8 / blk;
                        for (size_t bi = 0; !stop.load(std::memory_order_relaxed); bi = (bi + 1) % nblk) {
                            double *q = p + bi * blk;
                            for (size_t i = 0; i < blk; i += 4) _mm256_stream_pd(q + i, v);
                            done += 32 << 20;
                        }
                        _mm_sfence();
                        hog_bytes += done;
                    });
            const double t0 = now();
            long long copied = 0;
            double sink = 0;
            size_t off = 0, doff = 0;
            while (now() - t0 < T) {
                for (int q = 0; q < 2; q++) {  // two pieces in flight per synchronisation
                    cudaMemcpyAsync(ring + (off + q * PIECE) % R, d + doff, PIECE, cudaMemcpyDeviceToHost, st);
                    doff = (doff + PIECE) % DEV;
                }
                cudaStreamSynchronize(st);
                for (int q = 0; q < 2; q++) {  // the consumer reads what has just landed
                    const double *p = (const double *)(ring + (off + q * PIECE) % R);
                    __m256d a = _mm256_setzero_pd();
                    for (size_t i = 0; i < PIECE / 8; i += 4) a = _mm256_add_pd(a, _mm256_load_pd(p + i));
                    sink += ((double *)&a)[0];
                }
                off = (off + 2 * PIECE) % R;
                copied += 2 * PIECE;
            }
            const double dt = now() - t0;
            stop = true;
            for (auto &t : th) t.join();
            std::printf("ring %5zu MB  hogs %d: D2H+read %6.1f GB/s   hog stores %6.1f GB/s   (%g)\n", R >> 20, with_hog ? H : 0,
                        copied / dt / 1e9, hog_bytes.load() / dt / 1e9, sink);
            std::fflush(stdout);
        }
    return 0;
}
