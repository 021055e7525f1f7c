/* Per-call latency of the drop-in callback from C (no Python in the loop): DEFINEGusrfg_ with SNOPT's argument
 * list on the reference's own initial guess.
 *   gcc -O2 -I include -o tools/exp/latency tools/exp/latency.c -Ltol_b200 -ltolcuda -Wl,-rpath,'$ORIGIN/../../tol_b200'
 *   tools/exp/latency oracle/_ref/params/ tempest S10 [ts] */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "tolcuda.h"

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int ts = argc > 4 ? atoi(argv[4]) : 0;
    tolcuda_handle h;
    if (tolcuda_create_from_files(argv[1], argv[2], argv[3], 0, 0, 70, 0, -100, 0, 100, ts, 0, &h)) {
        fprintf(stderr, "%s\n", tolcuda_last_error());
        return 1;
    }
    int n, neF, neG;
    tolcuda_dims(h, &n, &neF, &neG);
    tolcuda_config cfg;
    tolcuda_get_config(h, &cfg);
    double *x = malloc(8 * n), *F = malloc(8 * neF), *G = malloc(8 * neG);
    tolcuda_problem_initial_guess(&cfg, x);
    tolcuda_bind_global(h);
    int status = 0, zero = 0;
    const char *lab[3] = {"F+G", "F", "G"};
    for (int m = 0; m < 3; m++) {
        int needF = m != 2, needG = m != 1;
        for (int i = 0; i < 200; i++)
            DEFINEGusrfg_(&status, &n, x, &needF, &neF, F, &needG, &neG, G, NULL, &zero, NULL, &zero, NULL, &zero);
        const int reps = 5000;
        const double t0 = now();
        for (int i = 0; i < reps; i++)
            DEFINEGusrfg_(&status, &n, x, &needF, &neF, F, &needG, &neG, G, NULL, &zero, NULL, &zero, NULL, &zero);
        printf("%s/%s ts=%d %-3s: %.2f us per callback (F[0]=%.17g, status %d)\n", argv[3], argv[2], cfg.ts, lab[m],
               1e6 * (now() - t0) / reps, F[0], status);
    }
    tolcuda_destroy(h);
    return 0;
}
