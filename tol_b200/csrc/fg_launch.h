// Host-side launch interface of fg_kernels.cu (internal to libtolcuda).
#ifndef TOLCUDA_FG_LAUNCH_H_
#define TOLCUDA_FG_LAUNCH_H_

#include <cuda_runtime.h>

#include "fg_const.h"

#define TOLCUDA_FORM_G7 7
#define TOLCUDA_FORM_S10 10

struct FgLaunch {
    int slot;  // index into the __constant__ FgConst table
    int form, wind, ts, n, neF, R0;
    int B;
    const double *x;
    long ldx;
    double *F;
    long ldF;
    double *G;
    long ldG;
    int needF, needG;
    int minb; // tuning variant: minimum resident CTAs per SM the kernel is compiled for (0 = default)
    cudaStream_t stream;
};

// copy one context's constants into its __constant__ slot (synchronises `stream`)
cudaError_t fg_upload_const(int slot, const FgConst &c, cudaStream_t stream);

// enqueue one batched F/G evaluation; device pointers
cudaError_t fg_launch(const FgLaunch &L);

#endif
