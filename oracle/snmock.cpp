/* TEST INFRASTRUCTURE ONLY (oracle/).  A stand-in for the SNOPT library (commercial, absent) that lets the
 * UNMODIFIED reference driver -- problem::runSNOPT (src/problem.cpp:1214-1240) -> snoptProblemA::solve
 * (src/snoptProblem.cpp:448-487) -> f_snkera(..., usrfun, ...) -- run end to end, so that the drop-in claim
 * "link libtolcuda's DEFINEGusrfg_ instead of src/DefineFG.cpp and change nothing else" can be executed.
 *
 * It is NOT an optimiser.  f_snkera calls the user function the way SNOPT does (snFunA argument list, 1-based
 * iGfun/jGvar in the caller's arrays, Status = 1 on the first call, 0 in between, 2 on the last, needF/needG
 * varying, cu/iu/ru as handed over), takes a few damped steps along the objective gradient it finds in G,
 * clipped to the bounds, and logs every call (status, needF, needG, x, F, G) to the file named by
 * $SNMOCK_LOG.  Two builds share it: tol_dropin_ref (reference DefineFG.cpp) and tol_dropin_cuda
 * (libtolcuda); their logs must agree call by call.  The other f_sn* entry points do nothing except
 * f_snmema, which reports a workspace size. */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {

/* argument lists as declared in reference include/snopt/snopt.h:60-66 (snFunA) and :93-216 (f_sn*); the
 * header itself is not included so that the entry points tol never reaches can stay one-liners */
typedef void (*snFunA)(int *Status, int *n, double x[], int *needF, int *neF, double F[], int *needG, int *neG,
                       double G[], char cu[], int *lencu, int iu[], int *leniu, double ru[], int *lenru);
typedef void (*isnLog)(void);
typedef isnLog isnLog2, isqLog, isnSTOP;

void f_sninit(const char *, int *, int *, int *, int *, int *, double *, int *) {}
void f_snspec(const char *, int *, int *inform, int *, int *, double *, int *) { *inform = 101; }
void f_sngetc(const char *, int *, char *, int *, int *errors, int *, int *, double *, int *) { *errors = 0; }
void f_sngeti(const char *, int *, int *ivalue, int *errors, int *, int *, double *, int *) { *ivalue = 0, *errors = 0; }
void f_sngetr(const char *, int *, double *rvalue, int *errors, int *, int *, double *, int *) { *rvalue = 0, *errors = 0; }
void f_snset(const char *, int *, int *errors, int *, int *, double *, int *) { *errors = 0; }
void f_snseti(const char *, int *, int *, int *errors, int *, int *, double *, int *) { *errors = 0; }
void f_snsetr(const char *, int *, double *, int *errors, int *, int *, double *, int *) { *errors = 0; }
void f_snsetprint(const char *, int *, int *, int *, int *, double *, int *) {}
void f_snend(int *) {}
void f_snmema(int *inform, int *, int *, int *, int *, int *miniw, int *minrw, int *, int *, double *, int *) {
    *inform = 104, *miniw = 500, *minrw = 500;
}

void f_snkera(int *start, const char *name, int *nf, int *n, double *objadd, int *objrow, snFunA usrfun, isnLog,
              isnLog2, isqLog, isnSTOP, int *iAfun, int *jAvar, int *neA, double *A, int *iGfun, int *jGvar,
              int *neG, double *xlow, double *xupp, double *flow, double *fupp, double *x, int *xstate,
              double *xmul, double *f, int *fstate, double *fmul, int *inform, int *ns, int *ninf, double *sinf,
              int *miniw, int *minrw, int *iu, int *leniu, double *ru, int *lenru, int *, int *, double *, int *) {
    (void)start, (void)name, (void)objadd, (void)iAfun, (void)jAvar, (void)neA, (void)A, (void)flow, (void)fupp;
    (void)xstate, (void)xmul, (void)fstate, (void)fmul;
    const int N = *n, NF = *nf, NG = *neG;
    std::vector<double> G(NG, 0.0);
    FILE *log = NULL;
    if (const char *p = std::getenv("SNMOCK_LOG")) log = std::fopen(p, "wb");
    if (log) {
        const int hdr[4] = {N, NF, NG, *objrow};
        std::fwrite(hdr, sizeof(int), 4, log);
        std::fwrite(iGfun, sizeof(int), NG, log);  // as SNOPT sees them: 1-based
        std::fwrite(jGvar, sizeof(int), NG, log);
    }
    const int steps = std::getenv("SNMOCK_STEPS") ? std::atoi(std::getenv("SNMOCK_STEPS")) : 6;
    char cu[8] = {0};
    int lencu = 0;
    for (int it = 0; it <= steps; it++) {
        int status = it == 0 ? 1 : (it == steps ? 2 : 0);
        int needF = (it % 3 != 2), needG = (it % 3 != 1);  // F+G, F only, G only, F+G, ...
        if (it == 0 || it == steps) needF = needG = 1;
        usrfun(&status, n, x, &needF, nf, f, &needG, neG, G.data(), cu, &lencu, iu, leniu, ru, lenru);
        if (log) {
            const int rec[3] = {status, needF, needG};
            std::fwrite(rec, sizeof(int), 3, log);
            std::fwrite(x, sizeof(double), N, log);
            std::fwrite(f, sizeof(double), NF, log);
            std::fwrite(G.data(), sizeof(double), NG, log);
        }
        if (status < 0) {  // the user function asked to stop
            *inform = 71;
            break;
        }
        if (it < steps) {  // damped step along -d(objective)/dx, clipped to the bounds
            for (int e = 0; e < NG; e++)
                if (iGfun[e] == *objrow) {
                    const int j = jGvar[e] - 1;
                    double v = x[j] - 1e-4 * std::tanh(G[e]);
                    x[j] = std::fmin(std::fmax(v, xlow[j]), xupp[j]);
                }
        }
        *inform = 1;
    }
    if (log) std::fclose(log);
    *ns = 0, *ninf = 0, *sinf = 0.0, *miniw = 500, *minrw = 500;
}

/* never reached by tol (setNeG is called, src/problem.cpp:1224), present so that snoptProblem.cpp links */
void f_snjac(int *info, int *, int *, snFunA, double *, double *, double *, int *, int *, int *, int *, double *,
             int *, int *, int *, int *, int *miniw, int *minrw, int *, int *, double *, int *, int *, int *,
             double *, int *) {
    *info = 102, *miniw = 500, *minrw = 500;
}
void f_snopta(void) {}
void f_snoptb(void) {}
void f_snkerb(void) {}
void f_snoptc(void) {}
void f_snkerc(void) {}
void f_snmem(void) {}

} /* extern "C" */
