// Problem set-up the reference performs before the first callback, restated for the batch driver and
// for a maintainer who wants SNOPT's arrays without building a reference `problem` object:
//   problemG7::InitialCond   src/problemG7.cpp:19-217   (+ RotateYaw :520-542)
//   problemS10::InitialCond  src/problemS10.cpp:19-219
//   problem::setLimits       src/problem.cpp:198-365    (node-0 constants from the constructor :80-134)
// Host code, glibc libm, same expression order as the reference: the results are bit-identical to the
// reference's arrays (tests/test_host_cpu.py compares against fixtures written by the compiled reference).
#include <cmath>

#include "tolcuda_internal.h"

namespace tolcuda {

namespace {
const double kG = 9.81;      // include/problem.h:72
const double kRho = 1.2682;  // include/problem.h:73
}  // namespace

void initial_guess(const tolcuda_config &cfg, double *x) {
    const bool g7 = cfg.formulation == TOLCUDA_G7;
    const int ts = cfg.ts, px = TOLCUDA_PX;
    const double mm = cfg.aircraft[0], SS = cfg.aircraft[2], ee = cfg.aircraft[3], AR = cfg.aircraft[4],
                 Cd0 = cfg.aircraft[5];
    const double xg = cfg.goal[0], yg = cfg.goal[1];
    const double xi = 0, yi = 0, zi = 0;  // src/problem.cpp:83-85, 111-113
    const double g = kG, rho = kRho;
    // "added by will": src/problemG7.cpp:39-43, src/problemS10.cpp:38-42
    const double tfinal = g7 ? 10 : 20;
    const double dt = tfinal / ts;
    const double xAmp = g7 ? 40 : 100, yAmp = g7 ? 0 : 100, zAmp = 0;
    double chi_d = 0.0;
    double t = 0.0;
    const double w_s = 2.0 * M_PI / (tfinal);
    double pre_chi = 0.0, pre_phi = 0.0, pre_CL = 0.0, phidoti = 0.0, CLdoti = 0.0;
    for (int ii = 0; ii <= ts; ii++) {
        double xs, ys, zs, xdot, ydot, zdot, xddot, yddot, zddot;
        if (g7) {  // src/problemG7.cpp:73-81
            xs = xAmp / tfinal * t + xi;
            ys = -yAmp * cos(w_s * t) + yAmp + yi;
            zs = zAmp * cos(w_s * t) - zAmp + zi;
            xdot = xAmp / tfinal;
            ydot = yAmp * w_s * sin(w_s * t);
            zdot = -zAmp * w_s * sin(w_s * t);
            xddot = 0.0;
            yddot = yAmp * w_s * w_s * cos(w_s * t);
            zddot = -zAmp * w_s * w_s * cos(w_s * t);
            // RotateYaw, src/problemG7.cpp:520-542
            chi_d = atan2(yg - yi, xg - xi);
            const double M11 = cos(chi_d), M12 = -sin(chi_d), M13 = 0.0;
            const double M21 = sin(chi_d), M22 = cos(chi_d), M23 = 0.0;
            const double M31 = 0.0, M32 = 0.0, M33 = 1.0;
            const double v0 = xs, v1 = ys, v2 = zs;
            xs = M11 * v0 + M12 * v1 + M13 * v2;
            ys = M21 * v0 + M22 * v1 + M23 * v2;
            zs = M31 * v0 + M32 * v1 + M33 * v2;
        } else {  // src/problemS10.cpp:82-90
            xs = xAmp * sin(w_s * t) - xAmp + xi;
            ys = (-yAmp * cos(w_s * t) + yi);
            zs = zAmp * cos(w_s * t) - zAmp + zi;
            xdot = w_s * xAmp * cos(w_s * t);
            ydot = (w_s * yAmp * sin(w_s * t));
            zdot = -w_s * zAmp * sin(w_s * t);
            xddot = -w_s * w_s * xAmp * sin(w_s * t);
            yddot = (w_s * w_s * yAmp * cos(w_s * t));
            zddot = -w_s * w_s * zAmp * cos(w_s * t);
        }
        const double W1 = 0, W2 = 0, W3 = 0;  // the initial trajectory never leaves z = 0
        const double a1 = xdot - W1, a2 = ydot - W2, a3 = zdot - W3;
        const double mag_a = sqrt(a1 * a1 + a2 * a2 + a3 * a3);
        const double Va = mag_a;
        double chi = g7 ? atan2(a2, a1) + chi_d : atan2(a2, a1);
        const double gam = atan2(-a3, sqrt(a1 * a1 + a2 * a2));
        if (ii > 0) {  // unwrap against the previous node, src/problemG7.cpp:112-129
            double diff_chi = chi - pre_chi;
            while (((diff_chi) < -M_PI) || ((diff_chi) > M_PI)) {
                if (diff_chi < -M_PI) {
                    const double m = ceil((-M_PI - diff_chi) / (2.0 * M_PI));
                    chi = chi + 2.0 * M_PI * m;
                }
                if (diff_chi > M_PI) {
                    const double m = floor((M_PI - diff_chi) / (2.0 * M_PI));
                    chi = chi + 2.0 * M_PI * m;
                }
                diff_chi = chi - pre_chi;
            }
        }
        const double r1_1 = a1 / mag_a, r1_2 = a2 / mag_a, r1_3 = a3 / mag_a;
        const double an1 = -xddot * (r1_1 * r1_1 - 1.0) - r1_1 * r1_2 * yddot - r1_1 * r1_3 * (zddot - g);
        const double an2 = -yddot * (r1_2 * r1_2 - 1.0) - r1_1 * r1_2 * xddot - r1_2 * r1_3 * (zddot - g);
        const double an3 = -(zddot - g) * (r1_3 * r1_3 - 1.0) - r1_1 * r1_3 * xddot - r1_2 * r1_3 * yddot;
        const double mag_an = sqrt(an1 * an1 + an2 * an2 + an3 * an3);
        const double r3_1 = -an1 / mag_an, r3_2 = -an2 / mag_an, r3_3 = -an3 / mag_an;
        const double r2_3 = r3_1 * r1_2 - r3_2 * r1_1;
        const double phi = atan2(r2_3, r3_3);
        const double L = mm * mag_an;
        const double CL = 2.0 * L / (rho * Va * Va * SS);
        const double D = 0.5 * rho * Va * Va * SS * (Cd0 + CL * CL / (M_PI * AR * ee));
        const double T = mm * (r1_1 * xddot + r1_2 * yddot + r1_3 * (zddot - g)) + D;
        if (ii == 0) {
            phidoti = 0.0;
            CLdoti = 0.0;
        } else {
            phidoti = (phi - pre_phi) / dt;
            CLdoti = (CL - pre_CL) / dt;
        }
        pre_phi = phi;
        pre_CL = CL;
        double *s = x + ii * px;
        s[1] = xs, s[2] = ys, s[3] = zs, s[4] = Va, s[5] = gam, s[6] = chi, s[7] = phi, s[8] = CL;
        s[9] = phidoti, s[10] = CLdoti, s[11] = T;
        x[0] = dt;
        pre_chi = chi;
        t = t + dt;
    }
    if (!g7) {  // "Populate phidot,CLdot at t=0", src/problemS10.cpp:208-211
        x[9] = x[ts * px + 9];
        x[10] = x[ts * px + 10];
    }
}

void bounds(const tolcuda_config &cfg, double *xlow, double *xupp, double *Flow, double *Fupp) {
    const bool g7 = cfg.formulation == TOLCUDA_G7;
    const int ts = cfg.ts, px = TOLCUDA_PX, pF = TOLCUDA_PF;
    const int nb = g7 ? 12 : 11;
    const int neF = pF * ts + 1 + nb;
    const double *ac = cfg.aircraft;
    const double CLmin = ac[6], CLmax = ac[7], phimax = ac[8], Vamin = ac[9], Vamax = ac[10],
                 gammamax = ac[11], phidotmax = ac[12], Tmin = ac[13], Tmax = ac[14];
    // limits.param file order: dtmin dtmax xmin xmax ymin ymax zmin zmax (src/parameters.cpp:108-115)
    const double *lm = cfg.limits;
    const double dtmin = lm[0], dtmax = lm[1], xmin = lm[2], xmax = lm[3], ymin = lm[4], ymax = lm[5],
                 zmin = lm[6], zmax = lm[7];
    // node-0 constants hard-coded in the constructor, src/problem.cpp:80-134
    const double xi = 0, yi = 0, zi = 0;
    const double Va1 = 4, Va2 = 50;
    const double gamma1 = g7 ? 0.0 * M_PI / 180.0 : 0, gamma2 = g7 ? 0.0 * M_PI / 180.0 : 0;
    const double chi1 = g7 ? -1e20 * M_PI / 180.0 : -1.7453292519943296e+18;
    const double chi2 = g7 ? 1e20 * M_PI / 180.0 : 1.7453292519943296e+18;
    const double phi1 = g7 ? -90.0 * M_PI / 180.0 : -1.5707963267948966;
    const double phi2 = g7 ? 90.0 * M_PI / 180.0 : 1.5707963267948966;
    const double CL1 = -0.5, CL2 = g7 ? 3.0 : 3;
    const double phidot1 = -3.4906585039886591, phidot2 = 3.4906585039886591;
    const double CLdot1 = -200.0, CLdot2 = 200.0;
    for (int ii = 0; ii <= ts; ii++) {
        double *lo = xlow + ii * px, *up = xupp + ii * px;
        if (ii == 0) {  // src/problem.cpp:254-268
            lo[1] = xi, up[1] = xi;
            lo[2] = yi, up[2] = yi;
            lo[3] = zi, up[3] = zi;
            lo[4] = Va1, up[4] = Va2;
            lo[5] = gamma1, up[5] = gamma2;
            lo[6] = chi1, up[6] = chi2;
            lo[7] = phi1, up[7] = phi2;
            lo[8] = CL1, up[8] = CL2;
            lo[9] = phidot1, up[9] = phidot2;
            lo[10] = CLdot1, up[10] = CLdot2;
            lo[11] = 0, up[11] = 1e20;
            xlow[0] = dtmin, xupp[0] = dtmax;
        } else {  // :272-285 (sic: the CLdot bound is phidotmax)
            lo[1] = xmin, up[1] = xmax;
            lo[2] = ymin, up[2] = ymax;
            lo[3] = zmin, up[3] = zmax;
            lo[4] = Vamin, up[4] = Vamax;
            lo[5] = -gammamax, up[5] = gammamax;
            lo[6] = -1e20, up[6] = 1e20;
            lo[7] = -phimax, up[7] = phimax;
            lo[8] = CLmin, up[8] = CLmax;
            lo[9] = -phidotmax, up[9] = phidotmax;
            lo[10] = -phidotmax, up[10] = phidotmax;
            lo[11] = Tmin, up[11] = Tmax;
        }
    }
    Flow[0] = -1e20, Fupp[0] = 1e20;  // :297-301
    for (int i = 1; i <= pF * ts; i++) Flow[i] = 0.0, Fupp[i] = 0.0;  // :305-315
    for (int b = 0; b < nb; b++) {    // :326-358
        Flow[neF - nb + b] = 0.0;
        Fupp[neF - nb + b] = 0.0;
    }
    if (g7) Flow[neF - 1] = -1e20;  // dist <= dmax
}

}  // namespace tolcuda
