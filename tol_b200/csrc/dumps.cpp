// The reference callback's per-call dump files (src/DefineFG.cpp:16-46, src/problem.cpp:740-756), host only.
// The reference rewrites Xoutput.txt / Woutput.txt / Foutput.txt / Goutput.txt in its working directory on every
// call of the user function; matlab/@plotSNOPT/plotSNOPT.m:108-125 polls Xoutput.txt to draw the iterates while
// SNOPT runs.  Here the files are written only on request (tolcuda_set_dump_dir), one formatted buffer and one
// write per file.
#include <cstdio>
#include <string>
#include <vector>

#include "tolcuda_internal.h"

namespace tolcuda {

namespace {

int flush_to(const std::string &path, const std::vector<char> &buf, size_t len) {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) {
        set_error("cannot open " + path + " for writing");
        return TOLCUDA_EIO;
    }
    const bool ok = std::fwrite(buf.data(), 1, len, f) == len;
    if (std::fclose(f) != 0 || !ok) {
        set_error("short write to " + path);
        return TOLCUDA_EIO;
    }
    return 0;
}

}  // namespace

// one value per line, "%.14f\n" (src/DefineFG.cpp:18-20, 32-34, 43-45)
int write_value_dump(const std::string &path, const double *v, long count) {
    // a finite double printed with %.14f takes at most 1 + 309 + 1 + 14 characters
    std::vector<char> buf;
    size_t len = 0;
    for (long i = 0; i < count; i++) {
        if (buf.size() - len < 400) buf.resize(buf.size() + (1 << 16) + 400);
        len += (size_t)std::snprintf(buf.data() + len, 400, "%.14f\n", v[i]);
    }
    return flush_to(path, buf, len);
}

// one line per node, the 12 wind arrays u v w du_dx du_dy du_dz dv_dx dv_dy dv_dz dw_dx dw_dy dw_dz as "%.6f "
// (src/problem.cpp:740-756); values of modelWind cases 0 and 1 (src/problem.cpp:480-531)
int write_wind_dump(const std::string &path, int wind_model, int ts, const double *x) {
    if (wind_model != TOLCUDA_WIND_NONE && wind_model != TOLCUDA_WIND_LINEAR_LAYER) {
        set_error("Woutput.txt is written for wind models 0 and 1 only");
        return TOLCUDA_EUNSUPPORTED;
    }
    std::vector<char> buf;
    size_t len = 0;
    for (int i = 0; i <= ts; i++) {
        double w[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (wind_model == TOLCUDA_WIND_LINEAR_LAYER) {
            const double Vref = 2.4, href = 10;          // :504-505
            const double zs = -x[i * TOLCUDA_PX + 3];    // ENU <- NED, :521
            w[1] = -Vref * zs / href;                    // v, :522
            w[8] = -Vref / href;                         // dv_dz, :523
        }
        if (buf.size() - len < 12 * 400) buf.resize(buf.size() + (1 << 16) + 12 * 400);
        for (int j = 0; j < 12; j++)
            len += (size_t)std::snprintf(buf.data() + len, 400, j < 11 ? "%.6f " : "%.6f\n", w[j]);
    }
    return flush_to(path, buf, len);
}

}  // namespace tolcuda
