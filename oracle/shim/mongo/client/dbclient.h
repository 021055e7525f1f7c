/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the legacy mongo-cxx-driver header that
 * /root/reference/include/problem.h:6 includes.  The driver is not in the reference tree nor in
 * this image.  connect() throws, which drives the UNMODIFIED reference into its own documented
 * fallback `Pwindmodel = 1` (reference src/problem.cpp:63-78) -- the only wind model the
 * reference can run without its MongoDB wind server.  Nothing here is product code. */
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

namespace mongo {

struct BSONElement {
    double numberDouble() const { return 0.0; }
};

struct BSONObj {
    int nFields() const { return 0; }
    BSONElement getField(const char *) const { return BSONElement(); }
};

struct Query {};

/* sink for the `MONGO_QUERY("x" << GTE << lo << LTE << hi)` builder expressions
 * (reference src/problem.cpp:371-460, never executed because connect() throws) */
struct QueryBuilderSink {
    template <class T> QueryBuilderSink &operator<<(const T &) { return *this; }
    operator Query() const { return Query(); }
};

static const int GTE = 0;
static const int LTE = 1;

struct DBClientCursor {
    BSONObj next() { return BSONObj(); }
};

struct DBClientConnection {
    void connect(const std::string &) { throw std::runtime_error("oracle shim: no wind database"); }
    unsigned long long count(const std::string &) { return 0; }
    BSONObj distinct(const std::string &, const std::string &, Query) { return BSONObj(); }
    std::auto_ptr<DBClientCursor> query(const std::string &, Query, int) {
        return std::auto_ptr<DBClientCursor>(new DBClientCursor());
    }
};

namespace client {
inline void initialize() {}
}  // namespace client

}  // namespace mongo

#define MONGO_QUERY(x) ((mongo::QueryBuilderSink() << x))
