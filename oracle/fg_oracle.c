/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of tol's SNOPT user-function path.
 * See fg_oracle.h for the role of this file and its parity status (PINNED against the compiled,
 * unmodified reference).  It is the checker for libtolcuda and is never linked into it.
 *
 * Restatement rules.  The reference evaluates every Jacobian entry by re-running the whole
 * expression table of its row; here each parenthesised sub-expression of the reference is
 * evaluated ONCE per node and reused.  Because IEEE-754 arithmetic is deterministic, reusing the
 * value of an identical parenthesised sub-expression is bit-identical to recomputing it, so --
 * built with -ffp-contract=off, like the reference's own -O2 x86-64 build -- this file reproduces
 * the reference bit for bit (tests/test_oracle.py).  The association order of every product
 * and sum below is the C left-to-right order of the cited reference line.
 *
 * All file:line citations are into /root/reference/. */
#include "fg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TOLO_G 9.81     /* include/problem.h:72 */
#define TOLO_RHO 1.2682 /* include/problem.h:73 */

/* ---------------------------------------------------------------- dispatch tables (countG) --- */

typedef struct {
    int Fnum, xnum, tf, tx;
} tolo_disp;

/* Would the gradient routine selected for Fnum raise the reference's `Gnonzero` flag for
 * (Fnum, xnum, tf, tx)?  One branch per routine, citing the lines that set the flag. */
static int raises_gnonzero(const tolo_problem *p, int Fnum, int xnum, int tf, int tx) {
    const int ts = p->ts;
    if (Fnum == 0) {
        if (p->formulation == TOLO_S10) /* src/problemS10.cpp:349-382 */
            return xnum == 0 || xnum == 1 || xnum == 10 || xnum == 11;
        /* src/problemG7.cpp:343-380 */
        if ((xnum == 0 || xnum == 1) && (tx == 0 || tx == ts)) return 1;
        return xnum == 10 || xnum == 11;
    }
    if (Fnum <= p->numstates) { /* src/problem.cpp:1074,1197-1206 */
        if (tx == tf) return 1;
        return tx == tf + 1 && xnum == Fnum - 1;
    }
    if (p->formulation == TOLO_S10) { /* src/problemS10.cpp:401-413 */
        if (xnum >= 0 && xnum <= 10 && Fnum == xnum + 9) return tx == 0 || tx == ts;
        return 0;
    }
    /* src/problemG7.cpp:406-511 */
    if (xnum >= 2 && xnum <= 10 && Fnum == xnum + 9) return tx == 0 || tx == ts;
    if ((xnum == 0 || xnum == 1) && (Fnum == 9 || Fnum == 10 || Fnum == 20))
        return tx == 0 || tx == ts;
    return 0;
}

/* src/problem.cpp:813-919, including the "Account for dt" re-pointing of each defect row's dt
 * entry (:883-910).  disp may be NULL. */
static int walk_pattern(const tolo_problem *p, int *iGfun, int *jGvar, tolo_disp *disp) {
    const int pF = p->numstates, px = p->numinp;
    const int n = px * (p->ts + 1) + 1;              /* src/problem.cpp:151 */
    const int neF = pF * p->ts + 1 + p->numbounds;   /* src/problem.cpp:152 */
    int neG = 0, reserve_dt = -1;
    for (int ii = 0; ii < neF; ii++) {
        for (int jj = 0; jj < n; jj++) {
            int Fnum = ii % pF, xnum, tf, tx; /* :827-836 */
            if (Fnum == 0 && ii != 0) Fnum = pF;
            tf = (ii - 1) / pF;
            if (ii >= neF - p->numbounds) Fnum = p->numstates + p->numbounds - (neF - 1 - ii);
            if (jj == 0) { /* :840-849 */
                xnum = px;
                reserve_dt = neG;
            } else {
                xnum = (jj - 1) % px;
            }
            tx = (jj - 1) / px;
            if (raises_gnonzero(p, Fnum, xnum, tf, tx) || xnum == px) { /* :868-879 */
                if (iGfun) iGfun[neG] = ii;
                if (jGvar) jGvar[neG] = jj;
                if (disp) {
                    disp[neG].Fnum = Fnum;
                    disp[neG].xnum = xnum;
                    disp[neG].tf = tf;
                    disp[neG].tx = tx;
                }
                neG++;
            }
            if (xnum == Fnum - 1 && tf == tx && Fnum > 0 && Fnum <= pF) { /* :883-910 */
                if (disp && reserve_dt >= 0) {
                    disp[reserve_dt].Fnum = Fnum;
                    disp[reserve_dt].xnum = px;
                    disp[reserve_dt].tf = tf;
                    disp[reserve_dt].tx = tx;
                }
                reserve_dt = -1;
            }
        }
    }
    return neG;
}

int tolo_pattern(const tolo_problem *p, int *iGfun, int *jGvar) {
    return walk_pattern(p, iGfun, jGvar, NULL);
}

void tolo_dims(const tolo_problem *p, int *n, int *neF, int *neG) {
    if (n) *n = p->numinp * (p->ts + 1) + 1;
    if (neF) *neF = p->numstates * p->ts + 1 + p->numbounds;
    if (neG) *neG = walk_pattern(p, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------------------ modelWind --- */

typedef struct {
    double *u, *v, *w, *du_dx, *du_dy, *du_dz, *dv_dx, *dv_dy, *dv_dz, *dw_dx, *dw_dy, *dw_dz;
    double *store;
} tolo_wind;

static void wind_alloc(tolo_wind *W, int nodes) {
    W->store = (double *)calloc((size_t)12 * nodes, sizeof(double));
    double **f = &W->u;
    for (int i = 0; i < 12; i++) f[i] = W->store + (size_t)i * nodes;
}

/* src/problem.cpp:475-531 (cases 0 and 1; case 3 needs the MongoDB wind cube and is unreachable
 * in the reference as built, src/problem.cpp:63-78) */
static void model_wind(const tolo_problem *p, const double *x, tolo_wind *W) {
    const int nodes = p->ts + 1;
    memset(W->store, 0, sizeof(double) * 12 * nodes);
    if (p->wind_model == 1) {
        const double Vref = 2.4, href = 10;
        for (int i = 0; i < nodes; i++) {
            double zs = -x[i * p->numinp + 3]; /* :522 */
            W->v[i] = -Vref * zs / href;       /* :523 */
            W->dv_dz[i] = -Vref / href;        /* :524 */
        }
    }
    if (p->wind_model == 3) { /* src/problem.cpp:544-695: trilinear interpolation of v and its gradient */
        const int nn = p->grid_nn, nu = p->grid_nu;
        const double dx = p->spacing[0], dy = p->spacing[1], dz = p->spacing[2];
#define GV(i, j, k) p->grid_v[((size_t)(i) * nn + (j)) * nu + (k)]
        for (int ii = 0; ii < nodes; ii++) {
            const double xs = x[ii * p->numinp + 2] + p->datum[0]; /* :551-553, ENU <- NED */
            const double ys = x[ii * p->numinp + 1] + p->datum[1];
            const double zs = -x[ii * p->numinp + 3] + p->datum[2];
            int xi, yi, zi;
            /* :556-572.  (sic) the x search runs to cache_north and the y search to cache_east */
            for (xi = 0; xi < p->grid_nn; xi++)
                if ((xs - p->grid_x[xi]) < dx) break;
            for (yi = 0; yi < p->grid_ne; yi++)
                if ((ys - p->grid_y[yi]) < dy) break;
            for (zi = 0; zi < p->grid_nu; zi++)
                if ((zs - p->grid_z[zi]) < dz) break;
            double vc[8];
            vc[0] = GV(xi, yi, zi), vc[1] = GV(xi + 1, yi, zi), vc[2] = GV(xi, yi + 1, zi);
            vc[3] = GV(xi + 1, yi + 1, zi), vc[4] = GV(xi, yi, zi + 1), vc[5] = GV(xi + 1, yi, zi + 1);
            vc[6] = GV(xi, yi + 1, zi + 1), vc[7] = GV(xi + 1, yi + 1, zi + 1);
            const double xrel = (xs - p->grid_x[xi]), yrel = (ys - p->grid_y[yi]), zrel = (zs - p->grid_z[zi]);
            const double zeta = xrel / dx, eta = yrel / dy, mu = zrel / dz;
            double N[8], NX[8], NY[8], NZ[8];
            N[0] = (1 - zeta) * (1 - eta) * (1 - mu); /* :618-625 */
            N[1] = zeta * (1 - eta) * (1 - mu);
            N[2] = (1 - zeta) * eta * (1 - mu);
            N[3] = zeta * eta * (1 - mu);
            N[4] = (1 - zeta) * (1 - eta) * mu;
            N[5] = zeta * (1 - eta) * mu;
            N[6] = (1 - zeta) * eta * mu;
            N[7] = zeta * eta * mu;
            NX[0] = -((yrel / dy - 1.0) * (zrel / dz - 1.0)) / dx; /* :643-650 */
            NX[1] = ((yrel / dy - 1.0) * (zrel / dz - 1.0)) / dx;
            NX[2] = (yrel * (zrel / dz - 1.0)) / (dx * dy);
            NX[3] = -(yrel * (zrel / dz - 1.0)) / (dx * dy);
            NX[4] = (zrel * (yrel / dy - 1.0)) / (dx * dz);
            NX[5] = -(zrel * (yrel / dy - 1.0)) / (dx * dz);
            NX[6] = -(yrel * zrel) / (dx * dy * dz);
            NX[7] = (yrel * zrel) / (dx * dy * dz);
            NY[0] = -((xrel / dx - 1.0) * (zrel / dz - 1.0)) / dy; /* :653-660 */
            NY[1] = (xrel * (zrel / dz - 1.0)) / (dx * dy);
            NY[2] = ((xrel / dx - 1.0) * (zrel / dz - 1.0)) / dy;
            NY[3] = -(xrel * (zrel / dz - 1.0)) / (dx * dy);
            NY[4] = (zrel * (xrel / dx - 1.0)) / (dy * dz);
            NY[5] = -(xrel * zrel) / (dx * dy * dz);
            NY[6] = -(zrel * (xrel / dx - 1.0)) / (dy * dz);
            NY[7] = (xrel * zrel) / (dx * dy * dz);
            NZ[0] = -((xrel / dx - 1.0) * (yrel / dy - 1.0)) / dz; /* :663-670 */
            NZ[1] = (xrel * (yrel / dy - 1.0)) / (dx * dz);
            NZ[2] = (yrel * (xrel / dx - 1.0)) / (dy * dz);
            NZ[3] = -(xrel * yrel) / (dx * dy * dz);
            NZ[4] = ((xrel / dx - 1.0) * (yrel / dy - 1.0)) / dz;
            NZ[5] = -(xrel * (yrel / dy - 1.0)) / (dx * dz);
            NZ[6] = -(yrel * (xrel / dx - 1.0)) / (dy * dz);
            NZ[7] = (xrel * yrel) / (dx * dy * dz);
            for (int i = 0; i < 8; i++) { /* :631-635, :682-692: only v is interpolated */
                W->v[ii] += N[i] * vc[i];
                W->dv_dx[ii] += NX[i] * vc[i];
                W->dv_dy[ii] += NY[i] * vc[i];
                W->dv_dz[ii] += NZ[i] * vc[i];
            }
        }
#undef GV
    }
}

/* ------------------------------------------------- per-node shared sub-expressions (F and G) --- */

typedef struct {
    double xs, ys, zs, Va, gam, chi, phi, CL, dphi, dCL, T, dt;
    double Wx, Wy, Wz, Wxx, Wxy, Wxz, Wyx, Wyy, Wyz, Wzx, Wzy, Wzz;
    double cc, sc, cg, sg, cp, sp; /* cos/sin of chi, gam, phi */
    double vx, vy, vz;             /* (Wx + Va*cos(chi)*cos(gam)) etc.               */
    double ax, ay, az;             /* (dWx_d? *cc*cg - dWz_d? *sg + dWy_d? *cg*sc)    */
    double bx, by, bz;             /* (dWz_d? *cg + dWx_d? *cc*sg + dWy_d? *sc*sg)    */
    double cx, cy, cz;             /* (dWy_d? *cc - dWx_d? *sc)                       */
} tolo_node;

static void node_load(const tolo_problem *p, const double *x, const tolo_wind *W, int k,
                      tolo_node *q) {
    const double *s = x + k * p->numinp; /* src/problem.cpp:1046-1057 */
    q->xs = s[1], q->ys = s[2], q->zs = s[3], q->Va = s[4], q->gam = s[5], q->chi = s[6];
    q->phi = s[7], q->CL = s[8], q->dphi = s[9], q->dCL = s[10], q->T = s[11], q->dt = x[0];
    /* NED <- ENU, src/problem.cpp:970-981 and :1061-1072 */
    q->Wx = W->v[k], q->Wy = W->u[k], q->Wz = -W->w[k];
    q->Wxx = W->dv_dy[k], q->Wxy = W->dv_dx[k], q->Wxz = -W->dv_dz[k];
    q->Wyx = W->du_dy[k], q->Wyy = W->du_dx[k], q->Wyz = -W->du_dz[k];
    q->Wzx = -W->dw_dy[k], q->Wzy = -W->dw_dx[k], q->Wzz = W->dw_dz[k];
    q->cc = cos(q->chi), q->sc = sin(q->chi);
    q->cg = cos(q->gam), q->sg = sin(q->gam);
    q->cp = cos(q->phi), q->sp = sin(q->phi);
    const double Va = q->Va, cc = q->cc, sc = q->sc, cg = q->cg, sg = q->sg;
    q->vx = q->Wx + Va * cc * cg;
    q->vy = q->Wy + Va * cg * sc;
    q->vz = q->Wz - Va * sg;
    q->ax = q->Wxx * cc * cg - q->Wzx * sg + q->Wyx * cg * sc;
    q->ay = q->Wxy * cc * cg - q->Wzy * sg + q->Wyy * cg * sc;
    q->az = q->Wxz * cc * cg - q->Wzz * sg + q->Wyz * cg * sc;
    q->bx = q->Wzx * cg + q->Wxx * cc * sg + q->Wyx * sc * sg;
    q->by = q->Wzy * cg + q->Wxy * cc * sg + q->Wyy * sc * sg;
    q->bz = q->Wzz * cg + q->Wxz * cc * sg + q->Wyz * sc * sg;
    q->cx = q->Wyx * cc - q->Wxx * sc;
    q->cy = q->Wyy * cc - q->Wxy * sc;
    q->cz = q->Wyz * cc - q->Wxz * sc;
}

/* ---------------------------------------------------------------------------- computeF ------ */

/* src/problemS10.cpp:227-265 */
static void cost_s10(const tolo_problem *p, const double *x, double *F) {
    double sump = 0.0, sumT = 0.0, dt = x[0];
    const double R = p->rg;
    for (int ii = 0; ii <= p->ts; ii++) {
        double xs = x[ii * p->numinp + 1], ys = x[ii * p->numinp + 2], T = x[ii * p->numinp + 11];
        double r = sqrt((xs - p->xg) * (xs - p->xg) + (ys - p->yg) * (ys - p->yg));
        double dR = (r - R) * (r - R);
        sump = sump + dR;
        sumT = sumT + T * T;
    }
    F[0] = 0.5 * p->kT * sumT + 0.5 * p->kp * sump + p->kdt * dt;
}

/* src/problemG7.cpp:225-250 */
static void cost_g7(const tolo_problem *p, const double *x, double *F) {
    double costsum = 0.0, xs = 0.0, ys = 0.0, dt = x[0];
    for (int ii = 0; ii <= p->ts; ii++) {
        xs = x[ii * p->numinp + 1];
        ys = x[ii * p->numinp + 2];
        double T = x[ii * p->numinp + 11];
        costsum = costsum + T * T;
    }
    double delx = xs - x[1], dely = ys - x[2];
    double dist = sqrt(delx * delx + dely * dely);
    F[0] = p->kT * 0.5 * costsum + p->kv * p->ts * dt / dist;
}

/* src/problem.cpp:929-1021 */
static void dynamic_constraints(const tolo_problem *p, const double *x, const tolo_wind *W,
                                double *F) {
    const double g = TOLO_G, rho = TOLO_RHO;
    const double mm = p->mm, SS = p->SS, Cd0 = p->Cd0, AR = p->AR, ee = p->ee;
    for (int k = 0; k < p->ts; k++) {
        tolo_node q;
        node_load(p, x, W, k, &q);
        const double Va = q.Va, CL = q.CL, T = q.T;
        double dx[6];
        dx[0] = q.vx; /* :1003 */
        dx[1] = q.vy; /* :1004 */
        dx[2] = q.vz; /* :1005 */
        dx[3] = T / mm - q.vy * q.ay - q.vz * q.az - q.vx * q.ax - g * q.sg -
                (rho * SS * Va * Va * (Cd0 + CL * CL / (AR * M_PI * ee))) / (2.0 * mm); /* :1006 */
        dx[4] = (q.vx * q.bx + q.vy * q.by + q.vz * q.bz - g * q.cg +
                 (CL * rho * SS * Va * Va * q.cp) / (2 * mm)) /
                Va; /* :1007 */
        dx[5] = -(q.vz * q.cz + q.cx * q.vx + q.vy * q.cy -
                  (CL * rho * SS * Va * Va * q.sp) / (2.0 * mm)) /
                (Va * q.cg); /* :1008 */
        const double *s0 = x + k * p->numinp, *s1 = s0 + p->numinp;
        double *Fk = F + k * p->numstates;
        for (int s = 0; s < 6; s++) Fk[1 + s] = s1[1 + s] - dx[s] * x[0] - s0[1 + s]; /* :1012-1017 */
        Fk[7] = s1[7] - s0[9] * x[0] - s0[7];                                          /* :1018 */
        Fk[8] = s1[8] - s0[10] * x[0] - s0[8];                                         /* :1019 */
    }
}

/* src/problemS10.cpp:273-305 */
static void boundary_s10(const tolo_problem *p, const double *x, double *F, int neF) {
    if (p->numbounds <= 0) return;
    const double chi_m = 2.0 * M_PI, delz = 0.0;
    const double *xe = x + p->ts * p->numinp;
    double *Fb = F + (neF - p->numbounds);
    for (int c = 0; c < 11; c++) Fb[c] = xe[1 + c] - x[1 + c];
    Fb[2] = xe[3] - x[3] - delz;
    Fb[5] = xe[6] - x[6] - chi_m;
}

/* src/problemG7.cpp:258-296 */
static void boundary_g7(const tolo_problem *p, const double *x, double *F, int neF) {
    if (p->numbounds <= 0) return;
    const double *xe = x + p->ts * p->numinp;
    double xf = xe[1], x0 = x[1], yf = xe[2], y0 = x[2];
    double dist = sqrt((xf - x0) * (xf - x0) + (yf - y0) * (yf - y0));
    double dmax = sqrt((p->xg - x0) * (p->xg - x0) + (p->yg - y0) * (p->yg - y0));
    double *Fb = F + (neF - p->numbounds);
    Fb[0] = xf - x0 - dist * cos(p->chi_d);
    Fb[1] = yf - y0 - dist * sin(p->chi_d);
    for (int c = 2; c <= 10; c++) Fb[c] = xe[1 + c] - x[1 + c];
    Fb[11] = dist - dmax;
}

/* ---------------------------------------------------------------------------- computeG ------ */

/* One row of the reference's tabG table: src/problem.cpp:1074-1192.  tab[0..10] = d/d(state c at
 * node k), tab[11] = d/d(dt). */
static void dynamics_row(const tolo_problem *p, const tolo_node *q, int Fnum, double tab[12]) {
    const double g = TOLO_G, rho = TOLO_RHO;
    const double mm = p->mm, SS = p->SS, Cd0 = p->Cd0, AR = p->AR, ee = p->ee;
    const double Va = q->Va, CL = q->CL, T = q->T, dt = q->dt;
    const double cc = q->cc, sc = q->sc, cg = q->cg, sg = q->sg, cp = q->cp, sp = q->sp;
    const double vx = q->vx, vy = q->vy, vz = q->vz;
    const double ax = q->ax, ay = q->ay, az = q->az, bx = q->bx, by = q->by, bz = q->bz;
    const double cx = q->cx, cy = q->cy, cz = q->cz;
    for (int i = 0; i < 12; i++) tab[i] = 0.0;
    switch (Fnum) {
    case 1: /* :1080-1090 */
        tab[0] = -1.0;
        tab[3] = -dt * cc * cg;
        tab[4] = Va * dt * cc * sg;
        tab[5] = Va * dt * cg * sc;
        tab[11] = -q->Wx - Va * cc * cg;
        break;
    case 2: /* :1094-1104 */
        tab[1] = -1.0;
        tab[3] = -dt * cg * sc;
        tab[4] = Va * dt * sc * sg;
        tab[5] = -Va * dt * cc * cg;
        tab[11] = -q->Wy - Va * cg * sc;
        break;
    case 3: /* :1108-1117 */
        tab[2] = -1.0;
        tab[3] = dt * sg;
        tab[4] = Va * dt * cg;
        tab[11] = Va * sg - q->Wz;
        break;
    case 4: { /* :1121-1132 */
        const double ex = q->Wyx * cc * cg - q->Wxx * cg * sc;
        const double ey = q->Wyy * cc * cg - q->Wxy * cg * sc;
        const double ez = q->Wyz * cc * cg - q->Wxz * cg * sc;
        tab[3] = dt * (cc * cg * ax - sg * az + cg * sc * ay +
                       (rho * SS * Va * (Cd0 + (CL * CL) / (AR * M_PI * ee))) / mm) -
                 1.0;
        tab[4] = -dt * (vx * bx + vy * by + vz * bz - g * cg + Va * cg * az + Va * cc * sg * ax +
                        Va * sc * sg * ay);
        tab[5] = dt * (ex * vx + vy * ey + ez * vz + Va * cc * cg * ay - Va * cg * sc * ax);
        tab[7] = (CL * rho * SS * (Va * Va) * dt) / (AR * M_PI * ee * mm);
        tab[10] = -dt / mm;
        tab[11] = vx * ax + vy * ay + vz * az - T / mm + g * sg +
                  (rho * SS * (Va * Va) * (Cd0 + (CL * CL) / (AR * M_PI * ee))) / (2.0 * mm);
        break;
    }
    case 5: { /* :1136-1147 */
        const double fx = q->Wyx * cc * sg - q->Wxx * sc * sg;
        const double fy = q->Wyy * cc * sg - q->Wxy * sc * sg;
        const double fz = q->Wyz * cc * sg - q->Wxz * sc * sg;
        const double S5 = vx * bx + vy * by + vz * bz - g * cg +
                          (CL * rho * SS * (Va * Va) * cp) / (2.0 * mm);
        tab[3] = (dt * S5) / (Va * Va) -
                 (dt * (cc * cg * bx - sg * bz + cg * sc * by + (CL * rho * SS * Va * cp) / mm)) / Va;
        tab[4] = -(dt * (vx * ax + vy * ay + vz * az + g * sg - Va * cg * bz - Va * cc * sg * bx -
                         Va * sc * sg * by)) /
                     Va -
                 1.0;
        tab[5] = -(dt * (fx * vx + vy * fy + fz * vz + Va * cc * cg * by - Va * cg * sc * bx)) / Va;
        tab[6] = (CL * rho * SS * Va * dt * sp) / (2.0 * mm);
        tab[7] = -(rho * SS * Va * dt * cp) / (2.0 * mm);
        tab[11] = -S5 / Va;
        break;
    }
    case 6: { /* :1151-1162 */
        const double dxw = q->Wxx * cc + q->Wyx * sc;
        const double dyw = q->Wxy * cc + q->Wyy * sc;
        const double dzw = q->Wxz * cc + q->Wyz * sc;
        const double Q = vz * cz + cx * vx + vy * cy - (CL * rho * SS * (Va * Va) * sp) / (2.0 * mm);
        tab[3] = -(dt * (sg * cz - cc * cg * cx - cg * sc * cy + (CL * rho * SS * Va * sp) / mm)) /
                     (Va * cg) -
                 (dt * Q) / ((Va * Va) * cg);
        tab[4] = (dt * sg * Q) / (Va * (cg * cg)) -
                 (dt * (Va * cg * cz + Va * cc * sg * cx + Va * sc * sg * cy)) / (Va * cg);
        tab[5] = -(dt * (vz * dzw + dxw * vx + vy * dyw - Va * cc * cg * cy + Va * cg * sc * cx)) /
                     (Va * cg) -
                 1.0;
        tab[6] = -(CL * rho * SS * Va * dt * cp) / (2.0 * mm * cg);
        tab[7] = -(rho * SS * Va * dt * sp) / (2.0 * mm * cg);
        tab[11] = Q / (Va * cg);
        break;
    }
    case 7: /* :1166-1174 */
        tab[6] = -1.0;
        tab[8] = -dt;
        tab[11] = -q->dphi;
        break;
    case 8: /* :1178-1186 */
        tab[7] = -1.0;
        tab[9] = -dt;
        tab[11] = -q->dCL;
        break;
    default:
        break;
    }
}

/* src/problemS10.cpp:314-386 */
static double cost_gradient_s10(const tolo_problem *p, const double *x, int xnum, int tx) {
    const double xs = x[tx * p->numinp + 1], ys = x[tx * p->numinp + 2], T = x[tx * p->numinp + 11];
    const double r = sqrt((xs - p->xg) * (xs - p->xg) + (ys - p->yg) * (ys - p->yg));
    const double R = p->rg;
    double Gs = 0.0;
    if (xnum == 0) Gs = p->kp * (r - R) * (xs - p->xg) / r;
    if (xnum == 1) Gs = p->kp * (r - R) * (ys - p->yg) / r;
    if (xnum == 10) Gs = p->kT * T;
    if (xnum == 11) Gs = p->kdt;
    return Gs;
}

/* src/problemG7.cpp:305-384 */
static double cost_gradient_g7(const tolo_problem *p, const double *x, int xnum, int tx) {
    const int ts = p->ts;
    const double T = x[tx * p->numinp + 11], dt = x[0];
    const double xf = x[ts * p->numinp + 1], x0 = x[1], yf = x[ts * p->numinp + 2], y0 = x[2];
    const double dist = sqrt((xf - x0) * (xf - x0) + (yf - y0) * (yf - y0));
    double Gs = 0.0;
    if (xnum == 0 && tx == 0) Gs = p->kp * ts * dt * (xf - x0) / (dist * dist * dist);
    if (xnum == 0 && tx == ts) Gs = -p->kp * ts * dt * (xf - x0) / (dist * dist * dist);
    if (xnum == 1 && tx == 0) Gs = p->kp * ts * dt * (yf - y0) / (dist * dist * dist);
    if (xnum == 1 && tx == ts) Gs = -p->kp * ts * dt * (yf - y0) / (dist * dist * dist);
    if (xnum == 10) Gs = p->kT * T;
    if (xnum == 11) Gs = p->kp * ts / (dist);
    return Gs;
}

/* src/problemS10.cpp:395-415.  The reference never assigns Gs on the dt column (xnum == 11): that
 * return value is uninitialised memory there; it is DEFINED as 0.0 here. */
static double boundary_gradient_s10(const tolo_problem *p, int Fnum, int xnum, int tx) {
    double Gs = 0.0;
    if (xnum >= 0 && xnum <= 10 && Fnum == xnum + 9) {
        if (tx == 0) Gs = -1.0;
        if (tx == p->ts) Gs = 1.0;
    }
    return Gs;
}

/* src/problemG7.cpp:393-513 */
static double boundary_gradient_g7(const tolo_problem *p, const double *x, int Fnum, int xnum,
                                   int tx) {
    const int ts = p->ts;
    const double xf = x[ts * p->numinp + 1], x0 = x[1], yf = x[ts * p->numinp + 2], y0 = x[2];
    const double dist = sqrt((xf - x0) * (xf - x0) + (yf - y0) * (yf - y0));
    const double chi_d = p->chi_d;
    double Gs = 0;
    if (xnum >= 2 && xnum <= 10 && Fnum == xnum + 9) {
        if (tx == 0) Gs = -1.0;
        if (tx == ts) Gs = 1.0;
    }
    if (xnum == 0 && Fnum == 9) {
        if (tx == 0) Gs = -1.0 + ((xf - x0) / dist) * cos(chi_d);
        if (tx == ts) Gs = 1.0 - ((xf - x0) / dist) * cos(chi_d);
    }
    if (xnum == 1 && Fnum == 9) {
        if (tx == 0) Gs = ((yf - y0) / dist) * cos(chi_d);
        if (tx == ts) Gs = -((yf - y0) / dist) * cos(chi_d);
    }
    if (xnum == 0 && Fnum == 10) {
        if (tx == 0) Gs = ((xf - x0) / dist) * sin(chi_d);
        if (tx == ts) Gs = -((xf - x0) / dist) * sin(chi_d);
    }
    if (xnum == 1 && Fnum == 10) {
        if (tx == 0) Gs = -1.0 + ((yf - y0) / dist) * sin(chi_d);
        if (tx == ts) Gs = 1.0 - ((yf - y0) / dist) * sin(chi_d);
    }
    if (xnum == 0 && Fnum == 20) {
        if (tx == 0) Gs = -(xf - x0) / dist;
        if (tx == ts) Gs = (xf - x0) / dist;
    }
    if (xnum == 1 && Fnum == 20) {
        if (tx == 0) Gs = -(yf - y0) / dist;
        if (tx == ts) Gs = (yf - y0) / dist;
    }
    return Gs;
}

/* src/problem.cpp:782-806 with the per-entry dispatch tables of countG */
static void compute_g(const tolo_problem *p, const double *x, const tolo_wind *W,
                      const tolo_disp *disp, int neG, double *G) {
    tolo_node q;
    double tab[12];
    int node_k = -1, row_F = -1, row_k = -1;
    for (int e = 0; e < neG; e++) {
        const int Fnum = disp[e].Fnum, xnum = disp[e].xnum, tf = disp[e].tf, tx = disp[e].tx;
        if (Fnum == 0) {
            G[e] = p->formulation == TOLO_S10 ? cost_gradient_s10(p, x, xnum, tx)
                                              : cost_gradient_g7(p, x, xnum, tx);
        } else if (Fnum <= p->numstates) {
            double Gs = 0.0; /* src/problem.cpp:1042 */
            if (tx == tf) {
                if (node_k != tx) {
                    node_load(p, x, W, tx, &q);
                    node_k = tx;
                    row_F = -1;
                }
                if (row_F != Fnum || row_k != tx) {
                    dynamics_row(p, &q, Fnum, tab);
                    row_F = Fnum;
                    row_k = tx;
                }
                Gs = tab[xnum]; /* :1197 */
            } else if (tx == tf + 1 && xnum == Fnum - 1) {
                Gs = 1.0; /* :1200-1205 */
            }
            G[e] = Gs;
        } else {
            G[e] = p->formulation == TOLO_S10 ? boundary_gradient_s10(p, Fnum, xnum, tx)
                                              : boundary_gradient_g7(p, x, Fnum, xnum, tx);
        }
    }
}

/* ------------------------------------------------------------------------------ entry points --- */

void tolo_eval_many(const tolo_problem *p, int count, const double *x, long ldx, double *F,
                    long ldF, double *G, long ldG) {
    const int n = p->numinp * (p->ts + 1) + 1;
    const int neF = p->numstates * p->ts + 1 + p->numbounds;
    tolo_disp *disp = NULL;
    int neG = 0;
    if (G) {
        /* every row but row 0 has fewer than 16 entries; row 0 has at most n */
        disp = (tolo_disp *)malloc(sizeof(tolo_disp) * ((size_t)neF * 16 + (size_t)n));
        neG = walk_pattern(p, NULL, NULL, disp);
    }
    tolo_wind W;
    wind_alloc(&W, p->ts + 1);
    for (int b = 0; b < count; b++) {
        const double *xb = x + (size_t)b * ldx;
        model_wind(p, xb, &W); /* src/DefineFG.cpp:24 */
        if (F) {               /* src/problem.cpp:765-774 */
            double *Fb = F + (size_t)b * ldF;
            if (p->formulation == TOLO_S10) cost_s10(p, xb, Fb);
            else cost_g7(p, xb, Fb);
            dynamic_constraints(p, xb, &W, Fb);
            if (p->formulation == TOLO_S10) boundary_s10(p, xb, Fb, neF);
            else boundary_g7(p, xb, Fb, neF);
        }
        if (G) compute_g(p, xb, &W, disp, neG, G + (size_t)b * ldG);
    }
    free(W.store);
    free(disp);
}

void tolo_eval(const tolo_problem *p, const double *x, int needF, double *F, int needG, double *G) {
    tolo_eval_many(p, 1, x, 0, needF > 0 ? F : NULL, 0, needG > 0 ? G : NULL, 0);
}
