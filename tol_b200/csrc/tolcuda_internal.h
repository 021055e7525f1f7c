// Internal declarations shared by the host-side sources of libtolcuda.
#ifndef TOLCUDA_INTERNAL_H_
#define TOLCUDA_INTERNAL_H_

#include <cmath>
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

void pattern_dims(int form, int ts, int *n, int *neF, int *neG, int *R0, int *nbG);
void pattern_build(int form, int ts, std::vector<int> &iG, std::vector<int> &jG);

void initial_guess(const tolcuda_config &cfg, double *x);
void bounds(const tolcuda_config &cfg, double *xlow, double *xupp, double *Flow, double *Fupp);

int read_params(const std::string &path, std::vector<double> &out);
int read_aircraft(const std::string &root, const std::string &name, double ac[15]);
int read_gains(const std::string &root, const std::string &mission, double gn[5]);
int read_limits(const std::string &root, const std::string &mission, double lm[8]);
int read_snopt(const std::string &root, const std::string &mission, double sn[6]);

}  // namespace tolcuda

#endif
