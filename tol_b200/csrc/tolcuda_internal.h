// Internal declarations shared by the host-side sources of libtolcuda.
#ifndef TOLCUDA_INTERNAL_H_
#define TOLCUDA_INTERNAL_H_

#include <cmath>
#include <functional>
#include <string>
#include <vector>

#include "../../include/tolcuda.h"
#include "fg_const.h"

#ifndef TOLCUDA_FORM_G7
#define TOLCUDA_FORM_G7 7
#define TOLCUDA_FORM_S10 10
#endif

namespace tolcuda {

void set_error(const std::string &msg);
// tolcuda_api.cpp: synchronous cudaMemcpy on `device`; kind 1 = host to device, 2 = device to host
int tolcuda_copy_raw(int device, void *dst, const void *src, size_t bytes, int kind);

void pattern_dims(int form, int ts, int *n, int *neF, int *neG, int *R0, int *nbG);
void pattern_build(int form, int ts, std::vector<int> &iG, std::vector<int> &jG);
void pattern_csc(int form, int ts, std::vector<int> &colptr, std::vector<int> &rowidx, std::vector<int> &perm);

void initial_guess(const tolcuda_config &cfg, double *x);
void bounds(const tolcuda_config &cfg, double *xlow, double *xupp, double *Flow, double *Fupp);

// compact G rows (compact.cpp): length of a row, expansion of one row / of B rows on a host thread pool
long compact_len(int form, int ts);
void expand_row(int form, int ts, const double *src, double *dst);
void expand_row(int form, int ts, const double *src, double *dst, bool wide);  // wide: AVX-512 line stores
void expand_row_cached(int form, int ts, const double *src, double *dst);     // ordinary (cache-allocating) stores

class HostPool {
public:
    explicit HostPool(int threads);
    ~HostPool();
    HostPool(const HostPool &) = delete;
    HostPool &operator=(const HostPool &) = delete;
    int threads() const { return threads_; }
    // fn(i) for i in [0, n) on the pool's threads and the caller's; returns when all are done
    void parallel_for(long n, const std::function<void(long)> &fn);
    // TOLCUDA_HOST_THREADS, else the cores this process may run on divided by LOCAL_WORLD_SIZE
    static int default_threads();

private:
    struct Impl;
    Impl *impl_;
    int threads_;
};

void expand_rows(HostPool &pool, int form, int ts, long B, const double *Gc, long ldGc, double *G, long ldG);

// results.cpp: the reference's two result files
int write_results_json(const tolcuda_config &cfg, const char *aircraft, const char *mission, double east,
                       double north, double up, const double *x, double final_cost, const char *path);
int write_results_txt(const tolcuda_config &cfg, const double *x, double final_cost, const char *path);

// dumps.cpp: the reference callback's per-call dump files
int write_value_dump(const std::string &path, const double *v, long count);
int write_wind_dump(const std::string &path, int wind_model, int ts, const double *x);

int read_params(const std::string &path, std::vector<double> &out);
int read_aircraft(const std::string &root, const std::string &name, double ac[15]);
int read_gains(const std::string &root, const std::string &mission, double gn[5]);
int read_limits(const std::string &root, const std::string &mission, double lm[8]);
int read_snopt(const std::string &root, const std::string &mission, double sn[6]);

}  // namespace tolcuda

#endif
