// sm_100a kernels for tol's SNOPT user function: objective, defect + boundary constraints F and the
// sparse Jacobian values G in SNOPT coordinate order, for B independent trajectories per launch.
//
// Replaces (reference file:line, lingaqing/tol):
//   problem::modelWind            src/problem.cpp:475-531   (cases 0, 1)
//   problem::computeF             src/problem.cpp:765-774
//   problem::dynamicConstraints   src/problem.cpp:929-1021
//   problem::computeG             src/problem.cpp:782-806
//   problem::dynamicsGradients    src/problem.cpp:1035-1208
//   problemS10::{cost,boundaryConstraints,costGradient,boundaryGradients}  src/problemS10.cpp:227-415
//   problemG7::{cost,boundaryConstraints,costGradient,boundaryGradients}   src/problemG7.cpp:225-513
//
// Work decomposition.  One CTA owns one trajectory; one thread owns one collocation window k (node k
// and the 8 states of node k+1).  The reference instead walks the neG coordinate entries and, for
// every single entry, re-evaluates the whole 12-column expression table of that entry's row
// (src/problem.cpp:785-802, 1074-1199); here every parenthesised sub-expression is evaluated once
// per node.
//
// Numerics.  Compiled with -fmad=false: every product and sum is rounded separately, in the C
// left-to-right association of the cited reference line, exactly like the reference's x86-64 -O2
// build (no FMA contraction).  Sub-expressions the reference multiplies by a wind-gradient
// component that is identically zero under the selected wind model are dropped: x + 0*y == x
// for finite y, so this changes no value (only, possibly, the sign of a zero).  What is left to
// differ from the reference is the last-ulp behaviour of sin/cos (CUDA libdevice vs glibc).
//
// Memory.  The trajectory's decision vector (n doubles, contiguous) is staged in shared memory with
// coalesced 16-byte loads; every output leaves through shared memory as full coalesced 16-byte
// stores: the 104-value Jacobian record of each window is assembled in a per-warp tile whose
// structural constants (0, +-1, -dt) are written once per CTA, and F plus the objective row are
// staged in the (by then dead) x buffer.  Cost sums use warp-shuffle reductions.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fg_const.h"
#include "fg_launch.h"

__constant__ FgConst c_fg[TOLCUDA_MAX_CTX];

namespace {

constexpr int PX = TOLCUDA_PX;
constexpr int PF = TOLCUDA_PF;
constexpr int REC = TOLCUDA_REC;
constexpr int REC_LD = 106;  // smem stride of a record: 16-byte aligned, and with 8 bytes x (2*106)
                             // words the 8/16 lanes of a pass hit distinct banks
constexpr int NVAR = 31;     // x-dependent entries of a record

// ---- per-window arithmetic ----------------------------------------------------------------------

struct WindowOut {
    double f[PF];    // defects F[1+8k .. 8+8k]
    double v[NVAR];  // x-dependent Jacobian entries, order = kVarIdx
};

// record positions of WindowOut::v (row s starts at 13*s: [d/d dt, d/d c0..c10 @k, d/d c_s @k+1])
__device__ constexpr int kVarIdx[NVAR] = {
    0,  4,  5,  6,           // F1: dt, Va, gam, chi        src/problem.cpp:1084-1088
    13, 17, 18, 19,          // F2                          :1098-1102
    26, 30, 31,              // F3: dt, Va, gam             :1112-1115
    39, 43, 44, 45, 47, 50,  // F4: dt, Va, gam, chi, CL, T :1125-1130
    52, 56, 57, 58, 59, 60,  // F5: dt, Va, gam, chi, phi, CL :1140-1145
    65, 69, 70, 71, 72, 73,  // F6                          :1155-1160
    78,                      // F7: dt                      :1172
    91};                     // F8: dt                      :1184

template <int WIND>
__device__ __forceinline__ void window_eval(const FgConst &c, const double *__restrict__ s0,
                                            const double *__restrict__ s1, double dt, bool needG,
                                            WindowOut &o) {
    constexpr bool W = (WIND == 1);
    const double z = s0[2], Va = s0[3], gam = s0[4], chi = s0[5], phi = s0[6], CL = s0[7];
    const double dphi = s0[8], dCL = s0[9], T = s0[10];
    double sc, cc, sg, cg, sp, cp;
    sincos(chi, &sc, &cc);
    sincos(gam, &sg, &cg);
    sincos(phi, &sp, &cp);

    // wind, NED <- ENU (src/problem.cpp:522-524, 970-981): Wx = v = -Vref*zs/href with zs = -z,
    // dWx_dz = -dv_dz; every other component is exactly zero under models 0 and 1
    const double Wxz = c.wind_Wxz;
    double Wx = 0.0;
    if (W) {
        const double zs = -z;
        Wx = -2.4 * zs / 10.0;
    }

    // (Wx + Va*cos(chi)*cos(gam)), (Wy + Va*cos(gam)*sin(chi)), (Wz - Va*sin(gam))
    const double Vacc = Va * cc, Vacg = Va * cg, Vasg = Va * sg;
    const double vx = W ? Wx + Vacc * cg : Vacc * cg;
    const double vy = Vacg * sc;
    const double vz = -Vasg;
    // the z-column instances of the reference's repeated wind-gradient brackets
    double az = 0, bz = 0, cz = 0, dz = 0, ez = 0, fz = 0;
    if (W) {
        const double Wxzcc = Wxz * cc, Wxzsc = Wxz * sc;
        az = Wxzcc * cg;          // (dWx_dz*cc*cg - dWz_dz*sg + dWy_dz*cg*sc)
        bz = Wxzcc * sg;          // (dWz_dz*cg + dWx_dz*cc*sg + dWy_dz*sc*sg)
        cz = -Wxzsc;              // (dWy_dz*cc - dWx_dz*sc)
        dz = Wxzcc;               // (dWx_dz*cc + dWy_dz*sc)
        ez = -((Wxz * cg) * sc);  // (dWy_dz*cc*cg - dWx_dz*cg*sc)
        fz = -(Wxzsc * sg);       // (dWy_dz*cc*sg - dWx_dz*sc*sg)
    }

    const double CdT = c.Cd0 + (CL * CL) / c.ARpiee;  // (Cd0 + CL*CL/(AR*pi*ee))
    const double rSV = c.rhoSS * Va;                  // rho*SS*Va
    const double CLrS = CL * c.rho * c.SS;            // CL*rho*SS
    const double CLrSV = CLrS * Va;
    const double Va2 = Va * Va;
    const double Tmm = T / c.mm;
    const double gsg = c.g * sg, gcg = c.g * cg;

    // ---- rhs, src/problem.cpp:1003-1008 ----
    const double drag3 = (rSV * Va * CdT) / c.twomm;
    const double dx3 = W ? Tmm - vz * az - gsg - drag3 : Tmm - gsg - drag3;
    const double n4 = W ? vz * bz - gcg : -gcg;  // (vx*bx + vy*by + vz*bz - g*cos(gam))
    const double dx4 = (n4 + (CLrSV * Va * cp) / c.twomm) / Va;
    const double lift5 = (CLrSV * Va * sp) / c.twomm;
    const double dx5 = W ? -(vz * cz - lift5) / Vacg : -(-lift5) / Vacg;

    // ---- defects, src/problem.cpp:1012-1019 ----
    o.f[0] = s1[0] - vx * dt - s0[0];
    o.f[1] = s1[1] - vy * dt - s0[1];
    o.f[2] = s1[2] - vz * dt - s0[2];
    o.f[3] = s1[3] - dx3 * dt - s0[3];
    o.f[4] = s1[4] - dx4 * dt - s0[4];
    o.f[5] = s1[5] - dx5 * dt - s0[5];
    o.f[6] = s1[6] - dphi * dt - s0[6];
    o.f[7] = s1[7] - dCL * dt - s0[7];
    if (!needG) return;

    // ---- Jacobian rows, src/problem.cpp:1074-1192 ----
    const double Vadt = Va * dt, mdt = -dt;
    double *v = o.v;
    // F1 :1084-1088
    v[0] = -vx;
    v[1] = mdt * cc * cg;
    v[2] = Vadt * cc * sg;
    v[3] = Vadt * cg * sc;
    // F2 :1098-1102
    v[4] = -vy;
    v[5] = mdt * cg * sc;
    v[6] = Vadt * sc * sg;
    v[7] = -(Vadt * cc * cg);
    // F3 :1112-1115
    v[8] = Vasg;
    v[9] = dt * sg;
    v[10] = Vadt * cg;
    // F4 :1125-1130
    {
        const double dragv = (rSV * CdT) / c.mm;
        const double drag11 = (c.rhoSS * Va2 * CdT) / c.twomm;
        v[11] = W ? vz * az - Tmm + gsg + drag11 : -Tmm + gsg + drag11;
        v[12] = W ? dt * (-(sg * az) + dragv) - 1.0 : dt * dragv - 1.0;
        v[13] = W ? mdt * (n4 + Vacg * az) : mdt * n4;
        v[14] = W ? dt * (ez * vz) : 0.0;
        v[15] = (CLrS * Va2 * dt) / c.ARpieemm;
        v[16] = mdt / c.mm;
    }
    // F5 :1140-1145
    {
        const double S5 = n4 + (CLrS * Va2 * cp) / c.twomm;
        const double liftv = (CLrSV * cp) / c.mm;
        v[17] = -S5 / Va;
        v[18] = W ? (dt * S5) / Va2 - (dt * (-(sg * bz) + liftv)) / Va
                  : (dt * S5) / Va2 - (dt * liftv) / Va;
        v[19] = W ? -(dt * (vz * az + gsg - Vacg * bz)) / Va - 1.0 : -(dt * gsg) / Va - 1.0;
        v[20] = W ? -(dt * (fz * vz)) / Va : 0.0;
        v[21] = (CLrSV * dt * sp) / c.twomm;
        v[22] = -(rSV * dt * cp) / c.twomm;
    }
    // F6 :1155-1160
    {
        const double lift6 = (CLrS * Va2 * sp) / c.twomm;
        const double Q = W ? vz * cz - lift6 : -lift6;
        const double sidev = (CLrSV * sp) / c.mm;
        v[23] = Q / Vacg;
        v[24] = W ? -(dt * (sg * cz + sidev)) / Vacg - (dt * Q) / (Va2 * cg)
                  : -(dt * sidev) / Vacg - (dt * Q) / (Va2 * cg);
        v[25] = W ? (dt * sg * Q) / (Va * (cg * cg)) - (dt * (Vacg * cz)) / Vacg
                  : (dt * sg * Q) / (Va * (cg * cg));
        v[26] = W ? -(dt * (vz * dz)) / Vacg - 1.0 : -1.0;
        v[27] = -(CLrSV * dt * cp) / (c.twomm * cg);
        v[28] = -(rSV * dt * sp) / (c.twomm * cg);
    }
    v[29] = -dphi;  // F7 :1172
    v[30] = -dCL;   // F8 :1184
}

// structural constants of a record (tabG zero-initialisation and the +-1 / -dt entries,
// src/problem.cpp:1038, 1084, 1098, 1112, 1170-1171, 1182-1183, 1204), x-dependent entries zeroed
__device__ __forceinline__ void record_init(double *rec, double dt) {
    double2 *r2 = reinterpret_cast<double2 *>(rec);
#pragma unroll
    for (int j = 0; j < REC / 2; j++) r2[j] = make_double2(0.0, 0.0);
    rec[1] = -1.0;   // F1 d/dx
    rec[15] = -1.0;  // F2 d/dy
    rec[29] = -1.0;  // F3 d/dz
    rec[85] = -1.0;  // F7 d/dphi
    rec[87] = -dt;   // F7 d/ddphi
    rec[99] = -1.0;  // F8 d/dCL
    rec[101] = -dt;  // F8 d/ddCL
#pragma unroll
    for (int s = 0; s < PF; s++) rec[13 * s + 12] = 1.0;  // d/d(state s at node k+1)
}

__device__ __forceinline__ void record_store(double *rec, const double *v) {
#pragma unroll
    for (int i = 0; i < NVAR; i++) {
        // adjacent (even, odd) positions go out as one 16-byte store
        if (i + 1 < NVAR && (kVarIdx[i] % 2 == 0) && kVarIdx[i + 1] == kVarIdx[i] + 1) {
            *reinterpret_cast<double2 *>(rec + kVarIdx[i]) = make_double2(v[i], v[i + 1]);
            i++;
        } else {
            rec[kVarIdx[i]] = v[i];
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// coalesced copy of `count` doubles from shared to global by `nthr` threads; 16-byte stores when both
// sides allow it
__device__ __forceinline__ void copy_out(double *__restrict__ dst, const double *__restrict__ src,
                                         int count, int tid, int nthr) {
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
    if (vec) {
        const int pairs = count >> 1;
        for (int i = tid; i < pairs; i += nthr)
            reinterpret_cast<double2 *>(dst)[i] = reinterpret_cast<const double2 *>(src)[i];
        if ((count & 1) && tid == 0) dst[count - 1] = src[count - 1];
    } else {
        for (int i = tid; i < count; i += nthr) dst[i] = src[i];
    }
}

// ---- the kernel -----------------------------------------------------------------------------------
//
// grid.x = B trajectories, blockDim.x = 32*ceil(ts/32) (<= MAXT <= 1024).  Dynamic shared memory:
//   sbuf [max(n, neF + R0) rounded to even]   x slice, later the F / objective-row staging area
//   tile [warps][NPP][REC_LD]                 Jacobian records of NPP consecutive windows per warp
//   red  [2][32]                              cross-warp cost sums
template <int FORM, int WIND, int NPP, int MAXT>
__global__ void __launch_bounds__(MAXT)
fg_batch_kernel(int slot, const double *__restrict__ x, long ldx, double *__restrict__ F, long ldF,
                double *__restrict__ G, long ldG, int needF, int needG, int sbuf_len) {
    extern __shared__ __align__(16) double smem[];
    const FgConst &c = c_fg[slot];
    const int ts = c.ts, n = c.n;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = nthr >> 5;
    double *sbuf = smem;
    double *tile = smem + sbuf_len + (size_t)warp * (NPP * REC_LD);
    double *red = smem + sbuf_len + (size_t)nwarps * (NPP * REC_LD);

    const size_t b = blockIdx.x;
    const double *xb = x + b * ldx;
    double *Fb = F + b * ldF;
    double *Gb = G + b * ldG;

    // ---- stage x: coalesced, 16 bytes per lane when the trajectory is 16-byte aligned ----
    if ((reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
        const double2 *x2 = reinterpret_cast<const double2 *>(xb);
        double2 *s2 = reinterpret_cast<double2 *>(sbuf);
        for (int i = tid; i < (n >> 1); i += nthr) s2[i] = __ldg(x2 + i);
        if ((n & 1) && tid == 0) sbuf[n - 1] = __ldg(xb + n - 1);
    } else {
        for (int i = tid; i < n; i += nthr) sbuf[i] = __ldg(xb + i);
    }
    const double dt_pre = __ldg(xb);  // every lane needs dt before the barrier for record_init
    if (needG && lane < NPP) record_init(tile + lane * REC_LD, dt_pre);
    __syncthreads();

    const double dt = sbuf[0];
    const int k = tid;
    const bool active = k < ts;
    WindowOut o;
    double sumT = 0.0, sump = 0.0;     // cost partial sums
    double r0x = 0.0, r0y = 0.0, r0T = 0.0;  // this node's objective-row entries
    double rex = 0.0, rey = 0.0, reT = 0.0;  // node ts's, held by the thread of window ts-1
    if (active) {
        const double *s0 = sbuf + 1 + PX * k;
        window_eval<WIND>(c, s0, s0 + PX, dt, needG != 0, o);
        // ---- objective terms: src/problemS10.cpp:246-258, 340-372; src/problemG7.cpp:240-241, 370
        const double T = s0[10];
        sumT = T * T;
        r0T = c.kT * T;
        if (FORM == TOLCUDA_FORM_S10) {
            const double ddx = s0[0] - c.xg, ddy = s0[1] - c.yg;
            const double r = sqrt(ddx * ddx + ddy * ddy);
            const double rmR = r - c.rg;
            sump = rmR * rmR;
            r0x = c.kp * rmR * ddx / r;
            r0y = c.kp * rmR * ddy / r;
        }
        if (k == ts - 1) {
            const double *se = s0 + PX;
            const double Te = se[10];
            sumT += Te * Te;
            reT = c.kT * Te;
            if (FORM == TOLCUDA_FORM_S10) {
                const double ddx = se[0] - c.xg, ddy = se[1] - c.yg;
                const double r = sqrt(ddx * ddx + ddy * ddy);
                const double rmR = r - c.rg;
                sump += rmR * rmR;
                rex = c.kp * rmR * ddx / r;
                rey = c.kp * rmR * ddy / r;
            }
        }
    }

    // ---- Jacobian records: NPP windows per pass through the warp's tile, then one coalesced copy ----
    if (needG) {
        const int kwarp = warp * 32;
        double *Grec = Gb + c.R0 + (size_t)REC * kwarp;
        const bool vec = (reinterpret_cast<uintptr_t>(Grec) & 15) == 0;
#pragma unroll 1
        for (int pass = 0; pass < 32 / NPP; pass++) {
            const int kbase = kwarp + pass * NPP;
            if (kbase >= ts) break;
            const int valid = min(NPP, ts - kbase);
            if (active && (lane / NPP) == pass) record_store(tile + (lane % NPP) * REC_LD, o.v);
            __syncwarp();
            double *dst = Grec + (size_t)REC * (pass * NPP);
            if (vec) {
                for (int i = lane; i < valid * (REC / 2); i += 32) {
                    const int node = i / (REC / 2), j = i - node * (REC / 2);
                    reinterpret_cast<double2 *>(dst)[i] =
                        *reinterpret_cast<const double2 *>(tile + node * REC_LD + 2 * j);
                }
            } else {
                for (int i = lane; i < valid * REC; i += 32) {
                    const int node = i / REC, j = i - node * REC;
                    dst[i] = tile[node * REC_LD + j];
                }
            }
            __syncwarp();
        }
    }

    // ---- cost sums: warp shuffle, then across warps ----
    sumT = warp_sum(sumT);
    if (FORM == TOLCUDA_FORM_S10) sump = warp_sum(sump);
    if (lane == 0) {
        red[warp] = sumT;
        red[32 + warp] = sump;
    }
    // endpoints used by the G7 objective / boundary rows (read before sbuf is recycled)
    const double x0 = sbuf[1], y0 = sbuf[2];
    const double xf = sbuf[1 + PX * ts], yf = sbuf[2 + PX * ts];
    double bnd[PX];  // thread 0: node ts minus node 0, per state
    if (tid == 0) {
#pragma unroll
        for (int cidx = 0; cidx < PX; cidx++) bnd[cidx] = sbuf[1 + PX * ts + cidx] - sbuf[1 + cidx];
    }
    __syncthreads();  // every thread is done reading x: sbuf becomes the output staging area

    double *sF = sbuf;            // [neF]
    double *sRow0 = sbuf + c.neF; // [R0] (S10 only; G7's objective row is written directly)
    if (active && needF) {
#pragma unroll
        for (int s = 0; s < PF; s++) sF[1 + PF * k + s] = o.f[s];
    }
    if (FORM == TOLCUDA_FORM_S10) {
        if (active && needG) {
            sRow0[1 + 3 * k] = r0x;
            sRow0[2 + 3 * k] = r0y;
            sRow0[3 + 3 * k] = r0T;
            if (k == ts - 1) {
                sRow0[1 + 3 * ts] = rex;
                sRow0[2 + 3 * ts] = rey;
                sRow0[3 + 3 * ts] = reT;
            }
        }
    } else if (active && needG) {
        // G7 objective row [dt, x_0, y_0, T_0 .. T_{ts-1}, x_ts, y_ts, T_ts], src/problemG7.cpp:343-380
        Gb[3 + k] = r0T;
        if (k == ts - 1) Gb[ts + 5] = reT;
    }
    if (tid == 0) {
        double tT = 0.0, tp = 0.0;
        for (int w = 0; w < nwarps; w++) {
            tT += red[w];
            tp += red[32 + w];
        }
        double *Gbnd = Gb + c.R0 + (size_t)REC * ts;
        if (FORM == TOLCUDA_FORM_S10) {
            if (needF) {
                sF[0] = c.half_kT * tT + c.half_kp * tp + c.kdt * dt;  // src/problemS10.cpp:264
                double *Fbnd = sF + (c.neF - c.nb);                    // src/problemS10.cpp:292-303
#pragma unroll
                for (int cidx = 0; cidx < PX; cidx++) Fbnd[cidx] = bnd[cidx];
                Fbnd[5] = bnd[5] - 2.0 * M_PI;
            }
            if (needG) {
                sRow0[0] = c.kdt;  // src/problemS10.cpp:378-381
                // boundary rows [dt, (0,c), (ts,c)]: the dt entry is uninitialised in the reference
                // (src/problemS10.cpp:397,414) and DEFINED as 0.0 here
#pragma unroll
                for (int cidx = 0; cidx < PX; cidx++) {
                    Gbnd[3 * cidx] = 0.0;
                    Gbnd[3 * cidx + 1] = -1.0;
                    Gbnd[3 * cidx + 2] = 1.0;
                }
            }
        } else {
            const double ddx = xf - x0, ddy = yf - y0;
            const double dist = sqrt(ddx * ddx + ddy * ddy);
            if (needF) {
                sF[0] = c.half_kT * tT + c.kv_ts * dt / dist;  // src/problemG7.cpp:249
                double *Fbnd = sF + (c.neF - c.nb);            // src/problemG7.cpp:276-294
                const double gx = c.xg - x0, gy = c.yg - y0;
                const double dmax = sqrt(gx * gx + gy * gy);
                Fbnd[0] = ddx - dist * c.cos_chid;
                Fbnd[1] = ddy - dist * c.sin_chid;
#pragma unroll
                for (int cidx = 2; cidx < PX; cidx++) Fbnd[cidx] = bnd[cidx];
                Fbnd[11] = dist - dmax;
            }
            if (needG) {
                // objective row ends, src/problemG7.cpp:343-380 (sic: kp, where cost() uses kv)
                const double d3 = dist * dist * dist;
                const double gx0 = c.kp_ts * dt * ddx / d3, gy0 = c.kp_ts * dt * ddy / d3;
                Gb[0] = c.kp_ts / dist;
                Gb[1] = gx0;
                Gb[2] = gy0;
                Gb[ts + 3] = -gx0;
                Gb[ts + 4] = -gy0;
                // boundary rows, src/problemG7.cpp:404-511
                const double ex = ddx / dist, ey = ddy / dist;
                double *p = Gbnd;
                p[0] = 0.0, p[1] = -1.0 + ex * c.cos_chid, p[2] = ey * c.cos_chid;
                p[3] = 1.0 - ex * c.cos_chid, p[4] = -(ey * c.cos_chid);
                p += 5;
                p[0] = 0.0, p[1] = ex * c.sin_chid, p[2] = -1.0 + ey * c.sin_chid;
                p[3] = -(ex * c.sin_chid), p[4] = 1.0 - ey * c.sin_chid;
                p += 5;
#pragma unroll
                for (int cidx = 2; cidx < PX; cidx++) {
                    p[0] = 0.0, p[1] = -1.0, p[2] = 1.0;
                    p += 3;
                }
                p[0] = 0.0, p[1] = -ex, p[2] = -ey, p[3] = ex, p[4] = ey;
            }
        }
    }
    __syncthreads();
    if (needF) copy_out(Fb, sF, c.neF, tid, nthr);
    if (FORM == TOLCUDA_FORM_S10 && needG) copy_out(Gb, sRow0, c.R0, tid, nthr);
}

template <int FORM, int WIND, int NPP, int MAXT>
cudaError_t launch_one(const FgLaunch &L) {
    auto kern = fg_batch_kernel<FORM, WIND, NPP, MAXT>;
    const int nthr = 32 * ((L.ts + 31) / 32);
    const int nwarps = nthr / 32;
    int sbuf_len = L.n > L.neF + L.R0 ? L.n : L.neF + L.R0;
    sbuf_len = (sbuf_len + 1) & ~1;
    const size_t smem = sizeof(double) * ((size_t)sbuf_len + (size_t)nwarps * NPP * REC_LD + 64);
    static size_t configured = 0;  // per instantiation
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<L.B, nthr, smem, L.stream>>>(L.slot, L.x, L.ldx, L.F, L.ldF, L.G, L.ldG, L.needF, L.needG,
                                        sbuf_len);
    return cudaGetLastError();
}

template <int FORM, int WIND>
cudaError_t launch_npp(const FgLaunch &L) {
    // 256-thread bound: up to 255 registers per thread; the 1024-thread variant (ts > 256) is
    // register-capped at 64 and exists for completeness, not speed
    const bool small = L.ts <= 256;
    switch (L.npp) {
    case 8: return small ? launch_one<FORM, WIND, 8, 256>(L) : launch_one<FORM, WIND, 8, 1024>(L);
    case 16: return small ? launch_one<FORM, WIND, 16, 256>(L) : launch_one<FORM, WIND, 16, 1024>(L);
    case 32: return small ? launch_one<FORM, WIND, 32, 256>(L) : launch_one<FORM, WIND, 32, 1024>(L);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

cudaError_t fg_upload_const(int slot, const FgConst &c, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_fg, &c, sizeof(FgConst), sizeof(FgConst) * (size_t)slot,
                                            cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

cudaError_t fg_launch(const FgLaunch &L) {
    if (L.B <= 0) return cudaSuccess;
    if (L.ts < 1 || L.ts > 1024) return cudaErrorInvalidValue;
    if (L.form == TOLCUDA_FORM_S10) {
        if (L.wind == 1) return launch_npp<TOLCUDA_FORM_S10, 1>(L);
        if (L.wind == 0) return launch_npp<TOLCUDA_FORM_S10, 0>(L);
    } else if (L.form == TOLCUDA_FORM_G7) {
        if (L.wind == 1) return launch_npp<TOLCUDA_FORM_G7, 1>(L);
        if (L.wind == 0) return launch_npp<TOLCUDA_FORM_G7, 0>(L);
    }
    return cudaErrorInvalidValue;
}
