/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of tol's SNOPT user-function path.
 *
 * This is the CHECKER for libtolcuda, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Parity status: PINNED -- the
 * restatement is compared bit-for-bit with the unmodified reference compiled by oracle/Makefile
 * (oracle/_ref/libtolref.so) in tests/test_oracle.py and against the committed fixtures in
 * tests/golden/ (generated from that same compiled reference by oracle/gen_golden.py).  The
 * reference ships no tests or golden vectors of its own.
 *
 * All file:line citations are into /root/reference/. */
#ifndef TOL_FG_ORACLE_H_
#define TOL_FG_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { TOLO_G7 = 7, TOLO_S10 = 10 };

typedef struct tolo_problem {
    int formulation; /* TOLO_G7 | TOLO_S10                      (src/tol.cpp:5-36)            */
    int ts;          /* time segments                            (src/parameters.cpp:139)      */
    int numinp;      /* px = 11                                  (src/parameters.cpp:140)      */
    int numstates;   /* pF = 8                                   (src/parameters.cpp:141)      */
    int numbounds;   /* 12 (G7) | 11 (S10)                       (src/parameters.cpp:142)      */
    int wind_model;  /* Pwindmodel: 0 none, 1 linear layer, 3 cube (src/problem.cpp:475-695)   */
    double mm, SS, ee, AR, Cd0; /* aircraft                      (src/parameters.cpp:48-53)    */
    double kT, kp, kv, kdt;     /* gains                         (src/parameters.cpp:83-87)    */
    double xg, yg, rg;          /* goal, NED                     (src/problem.cpp:24-27)       */
    double chi_d;               /* G7 course angle               (src/problemG7.cpp:524)       */
    /* wind model 3 only: the cached wind cube of src/problem.cpp:443-459 (cache[i][j][k], i < ne,
     * j < nn, k < nu), of which modelWind case 3 (:544-695) interpolates the v component alone */
    int grid_ne, grid_nn, grid_nu;
    const double *grid_x, *grid_y, *grid_z; /* cache[i][0][0].x, cache[0][j][0].y, cache[0][0][k].z    */
    const double *grid_v;                   /* cache[i][j][k].v, i-major                               */
    double datum[3];                        /* EastFromDatum, NorthFromDatum, UpFromDatum              */
    double spacing[3];                      /* xspacing, yspacing, zspacing (include/problem.h:87-89)  */
} tolo_problem;

/* n, neF as src/problem.cpp:151-152; neG by counting the pattern walk */
void tolo_dims(const tolo_problem *p, int *n, int *neF, int *neG);

/* literal restatement of countG's dense (row, column) walk (src/problem.cpp:813-919): fills
 * iGfun/jGvar (0-based, row-major) and returns neG.  Pass NULL arrays to count only. */
int tolo_pattern(const tolo_problem *p, int *iGfun, int *jGvar);

/* modelWind + computeF + computeG as DEFINEGusrfg_ sequences them (src/DefineFG.cpp:24-38), without
 * the debug dumps.  G is produced per pattern entry through the same (Fnum, xnum, tf, tx) dispatch
 * computeG uses (src/problem.cpp:782-806).  The 11 S10 entries the reference leaves uninitialised
 * (src/problemS10.cpp:397,414) are DEFINED as 0.0 here. */
void tolo_eval(const tolo_problem *p, const double *x, int needF, double *F, int needG, double *G);

/* count trajectories, trajectory-major with leading dimensions */
void tolo_eval_many(const tolo_problem *p, int count, const double *x, long ldx, double *F,
                    long ldF, double *G, long ldG);

#ifdef __cplusplus
}
#endif
#endif
