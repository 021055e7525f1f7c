/* TEST INFRASTRUCTURE ONLY (oracle/).  C-ABI harness around the UNMODIFIED reference sources.
 *
 * oracle/Makefile compiles /root/reference/src/{DefineFG,arguments,parameters,problem,problemG7,
 * problemS10,snoptProblem,jsoncpp}.cpp where they lie (nothing is copied into this repo) and links
 * them with this file into oracle/_ref/libtolref.so.  The harness
 *   - builds a reference `problemG7` / `problemS10` from the same positional arguments the
 *     reference CLI takes (reference src/arguments.cpp:32-46, src/tol.cpp:5-36),
 *   - exposes the arrays the reference hands to SNOPT (n, neF, neG, iGfun, jGvar, x0, bounds:
 *     reference src/problem.cpp:151-189, 198-365, 813-919),
 *   - evaluates F and G through the reference's own public entry points
 *     modelWind / computeF / computeG (reference include/problem.h:22-24) or through the full
 *     callback DEFINEGusrfg_ (reference src/DefineFG.cpp:9-48).
 *
 * It defines the process-global `prob` that the reference normally gets from src/tol.cpp:3
 * (tol.cpp itself needs CPython 2 and cannot be compiled here).
 *
 * fopen() calls made by the reference objects are routed through __wrap_fopen (ld --wrap=fopen,
 * see Makefile): with null-IO switched on, the four per-call debug dumps (Xoutput/Woutput/
 * Foutput/Goutput.txt) and Ioutput.txt go to /dev/null, which leaves the arithmetic untouched and
 * lets the CPU baseline be timed "arithmetic only"; switched off they are written to the cwd
 * exactly as shipped.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libtolcuda) never does. */
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "problemG7.h"
#include "problemS10.h"

problem *prob = NULL; /* normally defined in reference src/tol.cpp:3 */

namespace {

int g_null_io = 1;

/* reach the protected members without touching the reference sources */
struct ProbeG7 : problemG7 {
    explicit ProbeG7(arguments &a) : problemG7(a) {}
    using problem::F;
    using problem::Flow;
    using problem::Fupp;
    using problem::iGfun;
    using problem::jGvar;
    using problem::n;
    using problem::neF;
    using problem::neG;
    using problem::sn;
    using problem::x;
    using problem::xlow;
    using problem::xupp;
    using problem::ac;
    using problem::gn;
    using problem::lm;
    using problem::xg;
    using problem::yg;
    using problem::zg;
    using problem::rg;
    using problemG7::chi_d;
    using problem::Pwindmodel;
    using problem::cache;
    using problem::cache_east;
    using problem::cache_north;
    using problem::cache_up;
    using problem::xspacing;
    using problem::yspacing;
    using problem::zspacing;
    using problem::EastFromDatum;
    using problem::NorthFromDatum;
    using problem::UpFromDatum;
    using problem::u;
    using problem::v;
    using problem::w;
    using problem::du_dx;
    using problem::du_dy;
    using problem::du_dz;
    using problem::dv_dx;
    using problem::dv_dy;
    using problem::dv_dz;
    using problem::dw_dx;
    using problem::dw_dy;
    using problem::dw_dz;
    typedef problem::winddoc wdoc;

};
struct ProbeS10 : problemS10 {
    explicit ProbeS10(arguments &a) : problemS10(a) {}
    using problem::F;
    using problem::Flow;
    using problem::Fupp;
    using problem::iGfun;
    using problem::jGvar;
    using problem::n;
    using problem::neF;
    using problem::neG;
    using problem::sn;
    using problem::x;
    using problem::xlow;
    using problem::xupp;
    using problem::ac;
    using problem::gn;
    using problem::lm;
    using problem::xg;
    using problem::yg;
    using problem::zg;
    using problem::rg;
    using problem::Pwindmodel;
    using problem::cache;
    using problem::cache_east;
    using problem::cache_north;
    using problem::cache_up;
    using problem::xspacing;
    using problem::yspacing;
    using problem::zspacing;
    using problem::EastFromDatum;
    using problem::NorthFromDatum;
    using problem::UpFromDatum;
    using problem::u;
    using problem::v;
    using problem::w;
    using problem::du_dx;
    using problem::du_dy;
    using problem::du_dz;
    using problem::dv_dx;
    using problem::dv_dy;
    using problem::dv_dz;
    using problem::dw_dx;
    using problem::dw_dy;
    using problem::dw_dz;
    typedef problem::winddoc wdoc;
};

struct Handle {
    ProbeG7 *g7 = NULL;
    ProbeS10 *s10 = NULL;
    problem *base() const { return g7 ? (problem *)g7 : (problem *)s10; }
};

struct CoutSilencer {
    std::streambuf *old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};


template <class P>
static void set_grid(P *p, int ne, int nn, int nu, const double *gx, const double *gy, const double *gz,
                     const double *vals, const double *datum, const double *spacing) {
    p->cache.clear();
    for (int i = 0; i < ne; i++) {
        p->cache.push_back(std::vector<std::vector<typename P::wdoc> >());
        for (int j = 0; j < nn; j++) {
            p->cache[i].push_back(std::vector<typename P::wdoc>());
            for (int k = 0; k < nu; k++) {
                const double *q = vals + 3 * ((size_t)(i * nn + j) * nu + k);
                p->cache[i][j].push_back(typename P::wdoc(gx[i], gy[j], gz[k], q[0], q[1], q[2]));
            }
        }
    }
    p->cache_east = ne, p->cache_north = nn, p->cache_up = nu;
    p->EastFromDatum = datum[0], p->NorthFromDatum = datum[1], p->UpFromDatum = datum[2];
    p->xspacing = spacing[0], p->yspacing = spacing[1], p->zspacing = spacing[2];
    p->Pwindmodel = 3;
}

}  // namespace

extern "C" {

FILE *__real_fopen(const char *path, const char *mode);
FILE *__wrap_fopen(const char *path, const char *mode) {
    if (g_null_io) return __real_fopen("/dev/null", mode);
    return __real_fopen(path, mode);
}

void tolref_set_null_io(int on) { g_null_io = on; }

/* mission: "G7" | "S10"; root_path must end in '/' (paths are concatenated,
 * reference src/parameters.cpp:43,78,103,131).  Returns NULL on failure. */
void *tolref_create(const char *mission, const char *aircraft_name, double east, double north,
                    double up, double east_goal, double north_goal, double up_goal,
                    double radius_goal, const char *root_path) {
    try {
        char b[7][64];
        const double v[7] = {east, north, up, east_goal, north_goal, up_goal, radius_goal};
        for (int i = 0; i < 7; i++) snprintf(b[i], sizeof b[i], "%.17g", v[i]);
        std::string ac(aircraft_name), ms(mission);
        char *argv[11] = {(char *)"tol", b[0], b[1], b[2], b[3], b[4], b[5], b[6],
                          (char *)ac.c_str(), (char *)ms.c_str(), NULL};
        arguments args(argv);
        args.root_path = root_path;
        CoutSilencer quiet;
        Handle *h = new Handle();
        if (ms == "G7")
            h->g7 = new ProbeG7(args);
        else if (ms == "S10")
            h->s10 = new ProbeS10(args);
        else {
            delete h;
            return NULL;
        }
        prob = h->base();
        return h;
    } catch (std::exception &e) {
        fprintf(stderr, "tolref_create: %s\n", e.what());
        return NULL;
    }
}

void tolref_destroy(void *hv) {
    Handle *h = (Handle *)hv;
    if (!h) return;
    if (prob == h->base()) prob = NULL;
    delete h->g7;
    delete h->s10;
    delete h;
}

#define FIELD(h, f) ((h)->g7 ? (h)->g7->f : (h)->s10->f)

void tolref_dims(void *hv, int *n, int *neF, int *neG, int *ts, int *numbounds) {
    Handle *h = (Handle *)hv;
    *n = FIELD(h, n);
    *neF = FIELD(h, neF);
    *neG = FIELD(h, neG);
    *ts = FIELD(h, sn).ts;
    *numbounds = FIELD(h, sn).numbounds;
}

void tolref_pattern(void *hv, int *iGfun, int *jGvar) {
    Handle *h = (Handle *)hv;
    const int neG = FIELD(h, neG);
    memcpy(iGfun, FIELD(h, iGfun), sizeof(int) * neG);
    memcpy(jGvar, FIELD(h, jGvar), sizeof(int) * neG);
}

void tolref_x0(void *hv, double *x) {
    Handle *h = (Handle *)hv;
    memcpy(x, FIELD(h, x), sizeof(double) * FIELD(h, n));
}

void tolref_bounds(void *hv, double *xlow, double *xupp, double *Flow, double *Fupp) {
    Handle *h = (Handle *)hv;
    const int n = FIELD(h, n), neF = FIELD(h, neF);
    memcpy(xlow, FIELD(h, xlow), sizeof(double) * n);
    memcpy(xupp, FIELD(h, xupp), sizeof(double) * n);
    memcpy(Flow, FIELD(h, Flow), sizeof(double) * neF);
    memcpy(Fupp, FIELD(h, Fupp), sizeof(double) * neF);
}

/* the parsed .param values and goal as the reference holds them (reference src/parameters.cpp:42-148,
 * src/problem.cpp:24-27): ac[15] in file order with the three deg->rad conversions applied,
 * gn[5] = kT,kp,kv,ka,kdt, lm[8] = dtmin,dtmax,xmax,ymax,zmax,xmin,ymin,zmin (member order),
 * sn[6] = ts,numinp,numstates,numbounds,opt_tol,feas_tol, goal[4] = xg,yg,zg,rg (NED). */
void tolref_params(void *hv, double *ac, double *gn, double *lm, double *sn, double *goal,
                   int *wind_model) {
    Handle *h = (Handle *)hv;
    const aircraft &a = FIELD(h, ac);
    const double av[15] = {a.mm, a.b, a.SS, a.ee, a.AR, a.Cd0, a.CLmin, a.CLmax, a.phimax, a.Vamin,
                           a.Vamax, a.gammamax, a.phidotmax, a.Tmin, a.Tmax};
    memcpy(ac, av, sizeof av);
    const gain &g = FIELD(h, gn);
    const double gv[5] = {g.kT, g.kp, g.kv, g.ka, g.kdt};
    memcpy(gn, gv, sizeof gv);
    const limit &l = FIELD(h, lm);
    const double lv[8] = {l.dtmin, l.dtmax, l.xmax, l.ymax, l.zmax, l.xmin, l.ymin, l.zmin};
    memcpy(lm, lv, sizeof lv);
    const snopt &s = FIELD(h, sn);
    const double sv[6] = {(double)s.ts, (double)s.numinp, (double)s.numstates, (double)s.numbounds,
                          s.opt_tol, s.feas_tol};
    memcpy(sn, sv, sizeof sv);
    goal[0] = FIELD(h, xg), goal[1] = FIELD(h, yg), goal[2] = FIELD(h, zg), goal[3] = FIELD(h, rg);
    *wind_model = FIELD(h, Pwindmodel);
}

/* Wind model 3 of the reference (src/problem.cpp:544-695) without its MongoDB server: fill the protected
 * wind cache the way cacheWind would (src/problem.cpp:443-459: cache[i][j][k], i < cache_east,
 * j < cache_north, k < cache_up, each entry a winddoc x,y,z,u,v,w) and switch Pwindmodel to 3.  gx[i], gy[j],
 * gz[k] are the grid coordinates (ENU, metres from the datum), vals = 3 values (u,v,w) per grid point in
 * i-major order.  The reference code that consumes the cache runs unmodified. */
void tolref_set_wind_grid(void *hv, int ne, int nn, int nu, const double *gx, const double *gy, const double *gz,
                          const double *vals, const double *datum, const double *spacing) {
    Handle *h = (Handle *)hv;
    if (h->g7) set_grid(h->g7, ne, nn, nu, gx, gy, gz, vals, datum, spacing);
    else set_grid(h->s10, ne, nn, nu, gx, gy, gz, vals, datum, spacing);
}

/* the 12 per-node wind arrays as the last modelWind call left them, out[12][ts+1] in member order
 * u,v,w,du_dx,du_dy,du_dz,dv_dx,dv_dy,dv_dz,dw_dx,dw_dy,dw_dz (include/problem.h:103) */
void tolref_get_wind(void *hv, double *out) {
    Handle *h = (Handle *)hv;
    const int nodes = FIELD(h, sn).ts + 1;
#define CPW(i, f) memcpy(out + (size_t)(i) * nodes, FIELD(h, f).data(), sizeof(double) * nodes)
    CPW(0, u); CPW(1, v); CPW(2, w); CPW(3, du_dx); CPW(4, du_dy); CPW(5, du_dz);
    CPW(6, dv_dx); CPW(7, dv_dy); CPW(8, dv_dz); CPW(9, dw_dx); CPW(10, dw_dy); CPW(11, dw_dz);
#undef CPW
}

double tolref_chi_d(void *hv) {
    Handle *h = (Handle *)hv;
    return h->g7 ? h->g7->chi_d : 0.0;
}

/* arithmetic path only: the three public members DEFINEGusrfg_ calls, in its order */
void tolref_eval(void *hv, const double *x, int needF, double *F, int needG, double *G) {
    Handle *h = (Handle *)hv;
    problem *p = h->base();
    p->modelWind(const_cast<double *>(x));
    if (needF > 0) p->computeF(const_cast<double *>(x), F);
    if (needG > 0) p->computeG(const_cast<double *>(x), G);
}

/* the full snOptA callback exactly as SNOPT would invoke it */
void tolref_usrfun(void *hv, const double *x, int needF, double *F, int needG, double *G) {
    Handle *h = (Handle *)hv;
    prob = h->base();
    int status = 0, n = FIELD(h, n), neF = FIELD(h, neF), neG = FIELD(h, neG), zero = 0;
    DEFINEGusrfg_(&status, &n, const_cast<double *>(x), &needF, &neF, F, &needG, &neG, G, NULL,
                  &zero, NULL, &zero, NULL, &zero);
}

/* evaluate `count` trajectories stored trajectory-major with leading dimensions (timing loops) */
void tolref_eval_many(void *hv, int count, const double *x, long ldx, double *F, long ldF,
                      double *G, long ldG) {
    for (int b = 0; b < count; b++)
        tolref_eval(hv, x + b * ldx, 1, F + b * ldF, 1, G + b * ldG);
}

void tolref_usrfun_many(void *hv, int count, const double *x, long ldx, double *F, long ldF,
                        double *G, long ldG) {
    for (int b = 0; b < count; b++)
        tolref_usrfun(hv, x + b * ldx, 1, F + b * ldF, 1, G + b * ldG);
}

/* the reference's result writers (src/problem.cpp:1247-1365, 1371-1418) on a given state: x and F[0] are
 * copied into the arrays the writers read.  writeJSON honours its file name; writeTXT ignores it and always
 * writes "snopt_output.txt" into the current directory (through fopen: switch null-IO off first). */
void tolref_write_json(void *hv, const double *x, double F0, const char *path) {
    Handle *h = (Handle *)hv;
    memcpy(FIELD(h, x), x, sizeof(double) * FIELD(h, n));
    FIELD(h, F)[0] = F0;
    h->base()->writeJSON(path);
}

void tolref_write_txt(void *hv, const double *x, double F0) {
    Handle *h = (Handle *)hv;
    memcpy(FIELD(h, x), x, sizeof(double) * FIELD(h, n));
    FIELD(h, F)[0] = F0;
    h->base()->writeTXT("snopt_results.txt");
}

} /* extern "C" */
