// Result writers: the two files reference `tol` leaves behind after a solve, byte for byte.
//
//   snopt_results.json   problem::writeJSON  src/problem.cpp:1247-1365 (written by src/tol.cpp:30 and read
//                        back by the mission layer, msl/mission.py:204-240)
//   snopt_output.txt     problem::writeTXT   src/problem.cpp:1371-1418 (call commented out in src/tol.cpp:29)
//
// The reference builds a jsoncpp document and prints it with Json::StyledWriter (vendored jsoncpp,
// src/jsoncpp.cpp:4297-4469).  What that writer does to this document is restated here directly:
// object members in byte order of their names, three spaces per level, `"name" : value`, reals as %.17g,
// an array on one line (`[ a, b ]`) only if it has fewer than 25 elements and the line stays under 74
// characters, otherwise one element per line.
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "tolcuda_internal.h"

namespace tolcuda {

namespace {

struct Node {
    enum Kind { REAL, INT, STRING, ARRAY, OBJECT } kind = REAL;
    double real = 0.0;
    long integer = 0;
    std::string str;
    std::vector<double> arr;
    std::map<std::string, Node> obj;  // std::string order == jsoncpp's CZString order (memcmp, then length)

    Node() {}
    static Node of(double v) { Node n; n.kind = REAL, n.real = v; return n; }
    static Node of_int(long v) { Node n; n.kind = INT, n.integer = v; return n; }
    static Node of(const std::string &v) { Node n; n.kind = STRING, n.str = v; return n; }
    static Node of(const std::vector<double> &v) { Node n; n.kind = ARRAY, n.arr = v; return n; }
    Node &operator[](const std::string &k) {
        kind = OBJECT;
        return obj[k];
    }
};

// src/jsoncpp.cpp:4036-4071 (valueToString(double)): %.17g, non-finite values as jsoncpp spells them
std::string real_text(double v) {
    char buf[40];
    if (std::isfinite(v)) std::snprintf(buf, sizeof buf, "%.17g", v);
    else if (v != v) std::snprintf(buf, sizeof buf, "null");
    else std::snprintf(buf, sizeof buf, v < 0 ? "-1e+9999" : "1e+9999");
    for (char *p = buf; *p; p++)
        if (*p == ',') *p = '.';  // fixNumericLocale
    return buf;
}

// src/jsoncpp.cpp:4075-4140 (valueToQuotedString) for the characters that can occur in names here
std::string quoted(const std::string &s) {
    std::string out = "\"";
    for (unsigned char ch : s) {
        switch (ch) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (ch < 0x20) {
                    char u[8];
                    std::snprintf(u, sizeof u, "\\u%04X", ch);
                    out += u;
                } else {
                    out += (char)ch;
                }
        }
    }
    return out + "\"";
}

class Styled {
public:
    std::string doc;
    void value(const Node &n) {
        switch (n.kind) {
            case Node::REAL: doc += real_text(n.real); break;
            case Node::INT: doc += std::to_string(n.integer); break;
            case Node::STRING: doc += quoted(n.str); break;
            case Node::ARRAY: array(n.arr); break;
            case Node::OBJECT: object(n); break;
        }
    }

private:
    std::string indent_;
    static constexpr int kMargin = 74, kIndent = 3;  // StyledWriter::StyledWriter, :4297-4298

    void write_indent() {  // :4448-4457
        if (!doc.empty()) {
            const char last = doc.back();
            if (last == ' ') return;
            if (last != '\n') doc += '\n';
        }
        doc += indent_;
    }
    void object(const Node &n) {  // :4341-4366
        if (n.obj.empty()) {
            doc += "{}";
            return;
        }
        write_indent();
        doc += '{';
        indent_ += std::string(kIndent, ' ');
        size_t i = 0;
        for (const auto &kv : n.obj) {
            write_indent();
            doc += quoted(kv.first);
            doc += " : ";
            value(kv.second);
            if (++i < n.obj.size()) doc += ',';
        }
        indent_.resize(indent_.size() - kIndent);
        write_indent();
        doc += '}';
    }
    void array(const std::vector<double> &a) {  // :4370-4439
        if (a.empty()) {
            doc += "[]";
            return;
        }
        std::vector<std::string> txt;
        bool multi = (int)a.size() * 3 >= kMargin;
        if (!multi) {
            int len = 4 + ((int)a.size() - 1) * 2;
            for (double v : a) {
                txt.push_back(real_text(v));
                len += (int)txt.back().size();
            }
            multi = len >= kMargin;
        }
        if (multi) {
            write_indent();
            doc += '[';
            indent_ += std::string(kIndent, ' ');
            for (size_t i = 0; i < a.size(); i++) {
                write_indent();
                doc += txt.empty() ? real_text(a[i]) : txt[i];
                if (i + 1 < a.size()) doc += ',';
            }
            indent_.resize(indent_.size() - kIndent);
            write_indent();
            doc += ']';
        } else {
            doc += "[ ";
            for (size_t i = 0; i < txt.size(); i++) {
                if (i) doc += ", ";
                doc += txt[i];
            }
            doc += " ]";
        }
    }
};

int put_file(const char *path, const std::string &text) {
    FILE *f = std::fopen(path, "w");
    if (!f) {
        set_error(std::string("cannot write ") + path);
        return TOLCUDA_EIO;
    }
    const bool ok = std::fwrite(text.data(), 1, text.size(), f) == text.size();
    if (std::fclose(f) != 0 || !ok) {
        set_error(std::string("short write to ") + path);
        return TOLCUDA_EIO;
    }
    return 0;
}

}  // namespace

int write_results_json(const tolcuda_config &cfg, const char *aircraft, const char *mission, double east,
                       double north, double up, const double *x, double final_cost, const char *path) {
    const int ts = cfg.ts, px = TOLCUDA_PX;
    static const char *names[TOLCUDA_PX] = {"x", "y", "z", "Va", "gam", "chi", "phi", "CL", "dphi", "dCL", "T"};
    std::vector<double> col[TOLCUDA_PX], time_arr;
    double times = 0;  // src/problem.cpp:1276-1291
    for (int ii = 0; ii <= ts; ii++) {
        time_arr.push_back(times);
        for (int c = 0; c < px; c++) col[c].push_back(x[1 + c + ii * px]);
        times = times + x[0];
    }
    Node r;
    r["args"]["east"] = Node::of(east);  // :1293-1301
    r["args"]["north"] = Node::of(north);
    r["args"]["up"] = Node::of(up);
    r["args"]["xg"] = Node::of(cfg.goal[0]);
    r["args"]["yg"] = Node::of(cfg.goal[1]);
    r["args"]["zg"] = Node::of(cfg.goal[2]);
    r["args"]["rd"] = Node::of(cfg.goal[3]);
    r["args"]["aircraft"] = Node::of(std::string(aircraft));
    r["args"]["problem"] = Node::of(std::string(mission));
    r["problem"] = Node::of(std::string(mission));
    r["FinalCost"] = Node::of(final_cost);  // :1305-1306
    r["dt"] = Node::of(x[0]);
    r["trajectory"]["time"] = Node::of(time_arr);  // :1307-1318
    for (int c = 0; c < px; c++) r["trajectory"][names[c]] = Node::of(col[c]);
    static const char *acn[15] = {"mass", "b", "S", "e", "AR", "Cd0", "CLmin", "CLmax", "phimax", "Vamin",
                                  "Vamax", "gammamax", "dphimax", "Tmin", "Tmax"};  // :1320-1335
    r["aircraft"]["name"] = Node::of(std::string(aircraft));
    for (int i = 0; i < 15; i++) r["aircraft"][acn[i]] = Node::of(cfg.aircraft[i]);
    static const char *gnn[5] = {"kT", "kp", "kv", "ka", "kdt"};  // :1337-1341
    for (int i = 0; i < 5; i++) r["gains"][gnn[i]] = Node::of(cfg.gains[i]);
    static const char *lmn[8] = {"dtmin", "dtmax", "xmin", "xmax", "ymin", "ymax", "zmin", "zmax"};  // :1343-1350
    for (int i = 0; i < 8; i++) r["limits"][lmn[i]] = Node::of(cfg.limits[i]);
    r["snopt"]["ts"] = Node::of_int(ts);  // :1352-1357
    r["snopt"]["numinp"] = Node::of_int(TOLCUDA_PX);
    r["snopt"]["numstates"] = Node::of_int(TOLCUDA_PF);
    r["snopt"]["numbounds"] = Node::of_int(cfg.formulation == TOLCUDA_FORM_G7 ? 12 : 11);
    r["snopt"]["opt_tol"] = Node::of(cfg.solver_tol[0]);
    r["snopt"]["feas_tol"] = Node::of(cfg.solver_tol[1]);
    Styled w;
    w.value(r);
    w.doc += "\n";  // StyledWriter::write, :4300-4309
    return put_file(path, w.doc);
}

int write_results_txt(const tolcuda_config &cfg, const double *x, double final_cost, const char *path) {
    const int ts = cfg.ts, px = TOLCUDA_PX;
    std::string out;
    char buf[64];
    const double tfinal = 10, dt = tfinal / ts;  // src/problem.cpp:1381-1382
    out += "% SNOPT Output: Thesis Optimization \n";
    std::snprintf(buf, sizeof buf, "%4.2f", tfinal);
    out += std::string("% Simulation: tf_i = ") + buf + " s, dt_i = ";
    std::snprintf(buf, sizeof buf, "%4.2f", dt);
    out += std::string(buf) + " s \n";
    out += "% ";
    for (const char *h : {"time", "x", "y", "z", "Va", "gamma", "chi", "phi", "CL", "dphi", "dCL", "T", "dt"})
        out += std::string(h) + " \t \t";
    out += "Final Cost \n";
    out += "ProblemS10 \n";  // sic, for every mission (:1397)
    double time_s = 0.0;
    auto put = [&](double v, const char *tail) {
        std::snprintf(buf, sizeof buf, "%-4.7e ", v);
        out += buf;
        out += tail;
    };
    for (int ii = 0; ii <= ts; ii++) {  // :1399-1416
        put(time_s, "\t");
        for (int c = 0; c < px; c++) put(x[1 + c + ii * px], "\t");
        put(x[0], "\t");
        put(final_cost, "\n");
        time_s = time_s + x[0];
    }
    return put_file(path, out);
}

}  // namespace tolcuda
