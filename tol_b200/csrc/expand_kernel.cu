// Device-side counterpart of compact.cpp: compact G rows (objective-row block | 31 x-dependent values per window |
// boundary block | -dt) -> rows in SNOPT coordinate order, on the GPU.  For callers that keep or move G in compact
// form on the device (e.g. gather the shards of several GPUs on one of them over NVLink at a third of the bytes)
// and need the coordinate-order rows there.  Every value is placed where reference computeG (src/problem.cpp:
// 782-806) puts it; the structural constants of the reference's tabG (src/problem.cpp:1038, 1084, 1098, 1112,
// 1170, 1182, 1204) are written as literals.  Bound by its write stream (104 doubles out per 31 in).
#include <cuda_runtime.h>
#include <stdint.h>

#include "fg_const.h"
#include "fg_launch.h"

namespace {

constexpr int REC = TOLCUDA_REC, NVAR = TOLCUDA_NVAR;
constexpr int WPB = 64;      // windows per block
constexpr int THREADS = 208; // the pairs of a 4-record group: 4 * 104 / 2

// record position -> source: >= 0 index into the window's 31 values, -1: 0.0, -2: +1.0, -3: -1.0, -4: -dt
// (tolcuda_rec_kind, fg_const.h)
struct RecMap {
    signed char m[REC];
};
constexpr RecMap make_map() {
    RecMap r{};
    for (int j = 0; j < REC; j++) r.m[j] = (signed char)tolcuda_rec_kind(j);
    return r;
}
__constant__ RecMap c_map = make_map();

__device__ __forceinline__ double rec_value(const int pos, const double *v, const double mdt) {
    const int m = c_map.m[pos];
    return m >= 0 ? v[m] : (m == -1 ? 0.0 : (m == -2 ? 1.0 : (m == -3 ? -1.0 : mdt)));
}

// grid (ceil(ts / WPB), B); block (x, b) expands windows [WPB*x, WPB*x + WPB) of trajectory b; block x = 0 also
// copies the objective-row block, the last block the boundary block
__global__ void __launch_bounds__(THREADS)
expand_kernel(const int ts, const int R0, const int nbG, const double *__restrict__ Gc, const long ldGc,
              double *__restrict__ G, const long ldG) {
    __shared__ double sv[WPB * NVAR];
    const double *src = Gc + (size_t)blockIdx.y * ldGc;
    double *dst = G + (size_t)blockIdx.y * ldG;
    const int k0 = WPB * blockIdx.x, nk = min(WPB, ts - k0);
    const int t = threadIdx.x;
    for (int i = t; i < nk * NVAR; i += THREADS) sv[i] = __ldg(src + R0 + (size_t)NVAR * k0 + i);
    const double mdt = __ldg(src + R0 + (size_t)NVAR * ts + nbG);
    if (blockIdx.x == 0)
        for (int i = t; i < R0; i += THREADS) dst[i] = __ldg(src + i);
    if (blockIdx.x == gridDim.x - 1)
        for (int i = t; i < nbG; i += THREADS) dst[R0 + (size_t)REC * ts + i] = __ldg(src + R0 + (size_t)NVAR * ts + i);
    __syncthreads();
    double *out = dst + R0 + (size_t)REC * k0;
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // 16-byte stores.  Thread t owns pair t of every 4-record group: its two record positions (p, p + 1)
        // are the same in every group, so the table is read once per thread (a per-iteration lookup would be a
        // divergent constant-memory access: measured 2.4x slower)
        const int w = t / (REC / 2), p = 2 * (t - w * (REC / 2));
        const int m0 = c_map.m[p], m1 = c_map.m[p + 1];
        const double c0 = m0 == -1 ? 0.0 : (m0 == -2 ? 1.0 : (m0 == -3 ? -1.0 : mdt));
        const double c1 = m1 == -1 ? 0.0 : (m1 == -2 ? 1.0 : (m1 == -3 ? -1.0 : mdt));
        for (int r = w; r < nk; r += 4) {
            const double *v = sv + r * NVAR;
            *reinterpret_cast<double2 *>(out + (size_t)r * REC + p) = make_double2(m0 >= 0 ? v[m0] : c0, m1 >= 0 ? v[m1] : c1);
        }
    } else {
        for (int j = t; j < nk * REC; j += THREADS) {
            const int w = j / REC;
            out[j] = rec_value(j - w * REC, sv + w * NVAR, mdt);
        }
    }
}

// coordinate order -> column-compressed order: out[b][p] = G[b][perm[p]].  Writes are coalesced; the reads of a
// block stay within a few neighbouring window records (a column of node k gathers from the records of windows
// k-1 and k), i.e. in L1/L2.
__global__ void __launch_bounds__(256)
repack_kernel(const int neG, const int *__restrict__ perm, const double *__restrict__ G, const long ldG,
              double *__restrict__ out, const long ldo) {
    const double *src = G + (size_t)blockIdx.y * ldG;
    double *dst = out + (size_t)blockIdx.y * ldo;
    const int p0 = blockIdx.x * 2048;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int p = p0 + threadIdx.x + 256 * i;
        if (p < neG) dst[p] = __ldg(src + __ldg(perm + p));
    }
}

}  // namespace

cudaError_t repack_launch(int neG, const int *perm, long B, const double *G, long ldG, double *out, long ldo,
                          cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    for (long b0 = 0; b0 < B; b0 += 65535) {  // gridDim.y limit
        const long nb = B - b0 < 65535 ? B - b0 : 65535;
        const dim3 grid((neG + 2047) / 2048, (unsigned)nb);
        repack_kernel<<<grid, 256, 0, stream>>>(neG, perm, G + b0 * ldG, ldG, out + b0 * ldo, ldo);
    }
    return cudaGetLastError();
}

cudaError_t expand_launch(int form, int ts, int R0, int nbG, long B, const double *Gc, long ldGc, double *G,
                          long ldG, cudaStream_t stream) {
    (void)form;
    if (B <= 0) return cudaSuccess;
    for (long b0 = 0; b0 < B; b0 += 65535) {  // gridDim.y limit
        const long nb = B - b0 < 65535 ? B - b0 : 65535;
        const dim3 grid((ts + WPB - 1) / WPB, (unsigned)nb);
        expand_kernel<<<grid, THREADS, 0, stream>>>(ts, R0, nbG, Gc + b0 * ldGc, ldGc, G + b0 * ldG, ldG);
    }
    return cudaGetLastError();
}
