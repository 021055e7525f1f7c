// Closed form of the sparsity pattern reference problem::countG builds by walking the dense
// neF x n grid and probing the gradient routines' Gnonzero flag (src/problem.cpp:813-919).
//
// Row-major, 0-based, strictly increasing in (row, column):
//   row 0 (objective)      S10: dt, then (x, y, T) of every node 0..ts     src/problemS10.cpp:349-382
//                          G7 : dt, x_0, y_0, T_0..T_{ts-1}, x_ts, y_ts, T_ts   src/problemG7.cpp:343-380
//   row 1+8k+s (defect s of window k): dt, the 11 columns of node k, column s of node k+1
//                                                                            src/problem.cpp:868,1074,1200
//   boundary row b         S10: dt, (node 0, state b), (node ts, state b)    src/problemS10.cpp:401-413
//                          G7 : rows 0,1,11: dt, x_0, y_0, x_ts, y_ts; rows 2..10 as S10
//                                                                            src/problemG7.cpp:406-511
// The dt column is in every row because countG keeps an entry when `xnum == px` (src/problem.cpp:868).
#include "tolcuda_internal.h"

namespace tolcuda {

static inline int col(int k, int c) { return 1 + TOLCUDA_PX * k + c; }

void pattern_dims(int form, int ts, int *n, int *neF, int *neG, int *R0, int *nbG) {
    const int nb = form == TOLCUDA_FORM_G7 ? 12 : 11;
    const int r0 = form == TOLCUDA_FORM_G7 ? ts + 6 : 3 * ts + 4;
    const int bg = form == TOLCUDA_FORM_G7 ? 42 : 33;
    if (n) *n = TOLCUDA_PX * (ts + 1) + 1;       // src/problem.cpp:151
    if (neF) *neF = TOLCUDA_PF * ts + 1 + nb;    // src/problem.cpp:152
    if (R0) *R0 = r0;
    if (nbG) *nbG = bg;
    if (neG) *neG = r0 + TOLCUDA_REC * ts + bg;  // 105*ts+48 (G7), 107*ts+37 (S10)
}

void pattern_build(int form, int ts, std::vector<int> &iG, std::vector<int> &jG) {
    int n, neF, neG, R0, nbG;
    pattern_dims(form, ts, &n, &neF, &neG, &R0, &nbG);
    const int nb = form == TOLCUDA_FORM_G7 ? 12 : 11;
    iG.clear();
    jG.clear();
    iG.reserve(neG);
    jG.reserve(neG);
    auto put = [&](int i, int j) {
        iG.push_back(i);
        jG.push_back(j);
    };
    put(0, 0);
    if (form == TOLCUDA_FORM_S10) {
        for (int k = 0; k <= ts; k++) {
            put(0, col(k, 0));
            put(0, col(k, 1));
            put(0, col(k, 10));
        }
    } else {
        put(0, col(0, 0));
        put(0, col(0, 1));
        for (int k = 0; k < ts; k++) put(0, col(k, 10));
        put(0, col(ts, 0));
        put(0, col(ts, 1));
        put(0, col(ts, 10));
    }
    for (int k = 0; k < ts; k++)
        for (int s = 0; s < TOLCUDA_PF; s++) {
            const int row = 1 + TOLCUDA_PF * k + s;
            put(row, 0);
            for (int c = 0; c < TOLCUDA_PX; c++) put(row, col(k, c));
            put(row, col(k + 1, s));
        }
    const int rb = neF - nb;
    for (int b = 0; b < nb; b++) {
        put(rb + b, 0);
        if (form == TOLCUDA_FORM_G7 && (b == 0 || b == 1 || b == 11)) {
            put(rb + b, col(0, 0));
            put(rb + b, col(0, 1));
            put(rb + b, col(ts, 0));
            put(rb + b, col(ts, 1));
        } else {
            put(rb + b, col(0, b));
            put(rb + b, col(ts, b));
        }
    }
}

// Column-compressed view of the same pattern: colptr[n+1], rowidx[neG] and, for every CSC position p, the
// position perm[p] of that entry in coordinate (row-major) order.  Entries of a column keep their row order
// (a stable counting sort by column), which is what a CSC consumer expects.
void pattern_csc(int form, int ts, std::vector<int> &colptr, std::vector<int> &rowidx, std::vector<int> &perm) {
    int n, neF, neG;
    pattern_dims(form, ts, &n, &neF, &neG, nullptr, nullptr);
    std::vector<int> iG, jG;
    pattern_build(form, ts, iG, jG);
    colptr.assign(n + 1, 0);
    for (int e = 0; e < neG; e++) colptr[jG[e] + 1]++;
    for (int j = 0; j < n; j++) colptr[j + 1] += colptr[j];
    rowidx.assign(neG, 0);
    perm.assign(neG, 0);
    std::vector<int> next(colptr.begin(), colptr.end() - 1);
    for (int e = 0; e < neG; e++) {
        const int p = next[jG[e]]++;
        rowidx[p] = iG[e];
        perm[p] = e;
    }
}

}  // namespace tolcuda
