// Ceiling of the G write stream on B200, without any arithmetic: every CTA writes one 171,520-byte row
// (the padded S10 ts=200 G row) as contiguous pieces of S bytes from shared memory, issued either as TMA
// bulk copies (one lane per warp, D copies in flight per warp) or as coalesced 16-byte lane stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o storebw storebw.cu && ./storebw
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void bulk_store(void *g, const void *s, int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
                 "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// mode 0: TMA bulk copies; mode 1: lane stores (16 B per lane, coalesced)
template <int D>
__global__ void store_kernel(char *out, long row_bytes, int S, int mode, int smem_bytes) {
    extern __shared__ __align__(128) char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x * 16; i < smem_bytes; i += blockDim.x * 16) *reinterpret_cast<double2 *>(sm + i) = make_double2(1.0, 2.0);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    char *row = out + (long)blockIdx.x * row_bytes;
    // pieces of the row are dealt to the warps round robin in groups of 8 (like 32 windows = 8 groups of 4 records)
    const long npieces = row_bytes / S;
    const int per = (smem_bytes / nw) & ~127;
    const char *src = sm + (long)warp * per;
    const int slots = per / S;  // distinct smem pieces a warp cycles through
    int it = 0;
    for (long p0 = (long)warp * 8; p0 < npieces; p0 += (long)nw * 8) {
        for (int q = 0; q < 8 && p0 + q < npieces; q++, it++) {
            char *dst = row + (p0 + q) * S;
            const char *s = src + (long)(it % (slots > 0 ? slots : 1)) * S;
            if (mode == 0) {
                if (lane == 0) {
                    bulk_store(dst, s, S);
                    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(D - 1) : "memory");
                }
                __syncwarp();
            } else {
                for (int i = lane * 16; i < S; i += 512) *reinterpret_cast<double2 *>(dst + i) = *reinterpret_cast<const double2 *>(s + i);
            }
        }
    }
    if (mode == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// the real traffic mix of one S10 ts=200 trajectory per CTA: read the x row (cp.async, waited for), write the F
// row (8-byte lane stores), write the G row as TMA bulk copies of S bytes; no arithmetic
__global__ void mix_kernel(char *out, long row_bytes, const double *x, long ldx, int nx, double *F, long ldF, int nF, int S,
                           int smem_bytes, int wait_x, int do_x, int fmode, int rows_per_cta, int B) {
    extern __shared__ __align__(128) char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int per = (smem_bytes / nw) & ~127;
    char *mine = sm + (long)warp * per;
    for (int rr = 0; rr < rows_per_cta; rr++) {
        const long rowi = (long)blockIdx.x * rows_per_cta + rr;
        if (rowi >= B) break;
        const double *xs = x + rowi * ldx + 352 * warp;
        const int cnt = min(364, nx - 352 * warp);
        if (do_x) for (int i = lane * 2; i + 1 < cnt; i += 64)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(mine + 8 * i)), "l"(xs + i));
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        double acc = reinterpret_cast<double *>(mine)[lane];
        double *Fb = F + rowi * ldF + 1 + 256 * warp;
        const int nf = min(256, nF - 1 - 256 * warp);
        if (fmode) for (int i = lane; i < nf; i += 32) Fb[i] = acc;
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        char *row = out + rowi * row_bytes;
        const long npieces = row_bytes / S;
        for (long p0 = (long)warp * 4; p0 < npieces; p0 += (long)nw * 4)
            for (int q = 0; q < 4 && p0 + q < npieces; q++) {
                if (lane == 0) {
                    bulk_store(row + (p0 + q) * S, mine + 4096, S);
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncwarp();
            }
    }
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 16384;
    const long row = 171520;
    char *out;
    CK(cudaMalloc(&out, row * B));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("rows %d x %ld B = %.2f GB\n", B, row, row * (double)B / 1e9);
    const int Ss[] = {1664, 3328};
    for (int mode = 0; mode < 2; mode++)
        for (int S : Ss)
            for (int nw : {7})
                for (int smem_kb : {64, 100}) {
                    const int smem = smem_kb * 1024;
                    if (((smem / nw) & ~127) < S) continue;
                    if (mode == 1 && (S != 3328 || smem_kb != 64)) continue;
                    auto run = [&](auto kern, const char *nm) {
                        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                        for (int i = 0; i < 2; i++) kern<<<B, nw * 32, smem>>>(out, row, S, mode, smem);
                        CK(cudaEventRecord(e0));
                        for (int i = 0; i < 5; i++) kern<<<B, nw * 32, smem>>>(out, row, S, mode, smem);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        const double bytes = (double)(row / S) * S * B;
                        printf("%s S=%5d warps/CTA=%2d smem=%3dKB %s: %.3f ms  %.0f GB/s\n", mode ? "lane" : "tma ", S, nw, smem_kb, nm, ms / 5, bytes / (ms / 5 * 1e-3) / 1e9);
                    };
                    if (mode == 0) {
                        run(store_kernel<1>, "D=1");
                        run(store_kernel<2>, "D=2");
                        run(store_kernel<4>, "D=4");
                    } else {
                        run(store_kernel<1>, "   ");
                    }
                }
    {
        const long ldx = 2224, ldF = 1616;
        const int nx = 2212, nF = 1612;
        double *x, *F;
        CK(cudaMalloc(&x, ldx * 8 * B));
        CK(cudaMalloc(&F, ldF * 8 * B));
        CK(cudaMemset(x, 0, ldx * 8 * B));
        for (int smem_kb : {88})
            for (int S : {3328, 6656})
                for (int rpc = 1; rpc <= 2; rpc++) {
                    const int wait_x = 1, do_x = 1, fmode = 1;
                    const int smem = smem_kb * 1024;
                    CK(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                    for (int i = 0; i < 2; i++) mix_kernel<<<(B + rpc - 1) / rpc, 224, smem>>>(out, row, x, ldx, nx, F, ldF, nF, S, smem, wait_x, do_x, fmode, rpc, B);
                    CK(cudaEventRecord(e0));
                    for (int i = 0; i < 5; i++) mix_kernel<<<(B + rpc - 1) / rpc, 224, smem>>>(out, row, x, ldx, nx, F, ldF, nF, S, smem, wait_x, do_x, fmode, rpc, B);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double bytes = ((double)(row / S) * S + 8.0 * ((do_x ? nx : 0) + (fmode ? nF : 0))) * B;
                    printf("mix  S=%5d warps/CTA= 7 smem=%3dKB rows/CTA=%d: %.3f ms  %.0f GB/s (x in, F out, G out)\n", S, smem_kb, rpc, ms / 5,
                           bytes / (ms / 5 * 1e-3) / 1e9);
                }
    }
    CK(cudaGetLastError());
    return 0;
}
