// Readers for the reference's `value // comment` parameter files, bit-compatible with
// parameters::readparams (src/parameters.cpp:14-34) and the typed views built on it
// (aircraft :42-69, gain :77-94, limit :102-122, snopt :130-148).
//
// Behaviours kept on purpose:
//   * the reference's delimiter is the multi-character literal '//' narrowed to char, i.e. '/':
//     a line is cut at its FIRST '/', not at "//";
//   * the remaining text goes through std::stod: leading white space skipped, the longest valid
//     floating literal parsed, anything after it ignored (the shipped files carry a literal
//     backslash-n after each number and, for skywalker.param, a trailing CR);
//   * a line that does not begin with a number (the header comment, blank lines) or whose value is
//     out of double range makes stod throw, and the reference skips that line.
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <string>

#include "tolcuda_internal.h"

namespace tolcuda {

int read_params(const std::string &path, std::vector<double> &out) {
    out.clear();
    std::ifstream in(path);
    if (!in.is_open()) return TOLCUDA_EIO;
    std::string line;
    while (std::getline(in, line)) {
        const std::string item = line.substr(0, line.find('/'));
        const char *s = item.c_str();
        char *end = nullptr;
        errno = 0;
        const double v = std::strtod(s, &end);
        if (end == s || errno == ERANGE) continue;
        out.push_back(v);
    }
    return 0;
}

static int read_exact(const std::string &path, size_t want, std::vector<double> &v) {
    int e = read_params(path, v);
    if (e) {
        set_error("cannot open parameter file " + path);
        return e;
    }
    if (v.size() != want) {  // reference: std::length_error("Wrong number of parameters ...")
        set_error("wrong number of parameters in " + path + ": got " + std::to_string(v.size()) +
                  ", need " + std::to_string(want));
        return TOLCUDA_EIO;
    }
    return 0;
}

int read_aircraft(const std::string &root, const std::string &name, double ac[15]) {
    std::vector<double> v;
    int e = read_exact(root + "aircraft/" + name + ".param", 15, v);
    if (e) return e;
    for (int i = 0; i < 15; i++) ac[i] = v[i];
    ac[8] = v[8] * M_PI / 180.0;    // phimax     src/parameters.cpp:56
    ac[11] = v[11] * M_PI / 180.0;  // gammamax   :59
    ac[12] = v[12] * M_PI / 180.0;  // phidotmax  :60
    return 0;
}

int read_gains(const std::string &root, const std::string &mission, double gn[5]) {
    std::vector<double> v;
    int e = read_exact(root + "problems/" + mission + "/gains.param", 5, v);
    if (e) return e;
    for (int i = 0; i < 5; i++) gn[i] = v[i];
    return 0;
}

int read_limits(const std::string &root, const std::string &mission, double lm[8]) {
    std::vector<double> v;
    int e = read_exact(root + "/problems/" + mission + "/limits.param", 8, v);
    if (e) return e;
    for (int i = 0; i < 8; i++) lm[i] = v[i];  // file order: dtmin dtmax xmin xmax ymin ymax zmin zmax
    return 0;
}

int read_snopt(const std::string &root, const std::string &mission, double sn[6]) {
    std::vector<double> v;
    int e = read_exact(root + "/problems/" + mission + "/snopt.param", 6, v);
    if (e) return e;
    for (int i = 0; i < 6; i++) sn[i] = v[i];
    return 0;
}

}  // namespace tolcuda
