// Fused evaluate + gather (include/tolcuda.h, tolcuda_gather_*): the shards of a batch, evaluated on several GPUs,
// land as F and G rows in ONE GPU's memory while they are being computed.  SURVEY.md 8f-4; nothing of the kind in
// the reference (SNOPT is host-only).
//
// Built from the library's own public primitives and nothing else -- tolcuda_eval_batch with device pointers into
// the owner's buffer (the kernels' F stores and TMA bulk copies go straight over NVLink), TOLCUDA_COMPACT_G for the
// peers' G (NVLink is the slower side: a third of the bytes), tolcuda_stream_signal / _wait for the per-chunk
// hand-over and tolcuda_expand_compact_g_device on the owner -- so the protocol is one implementation for a C++
// driver with several devices (tolbatch --gather-gpu) and for one process per GPU (tol_b200/dist.py, which only
// carries the 64-byte handle between the processes).
//
// Buffer on the owner:  F rows [B][ldF] | G rows [B][ldG] | peers' compact rows [B][ldC] | flags uint32 [world][16]
#include <algorithm>
#include <cstdint>
#include <new>
#include <string>
#include <vector>

#include "tolcuda_internal.h"

using namespace tolcuda;

namespace {
constexpr int MAX_CHUNKS = 16;
constexpr size_t FLAG_BYTES = 4096;  // world <= 64
}  // namespace

struct tolcuda_gather {
    tolcuda_handle h = nullptr;
    int device = 0, rank = 0, world = 1, dst = 0;
    long B = 0, ldF = 0, ldG = 0, ldC = 0;
    int n = 0, neF = 0, neG = 0, form = 0, ts = 0;
    char *base = nullptr;  // the owner's buffer as addressed from this participant's device
    size_t bytes = 0;
    bool owner = false, ipc = false;
    unsigned epoch = 0;
    double *Fp() const { return reinterpret_cast<double *>(base); }
    double *Gp() const { return Fp() + (size_t)B * ldF; }
    double *Cp() const { return Gp() + (size_t)B * ldG; }
    char *flags() const { return base + bytes - FLAG_BYTES; }
};

namespace {

void shard(long B, int rank, int world, long *b0, long *b1) {  // tol_b200.synth.shard_range
    const long per = (B + world - 1) / world;
    *b0 = std::min(B, rank * per);
    *b1 = std::min(B, *b0 + per);
}

int init_common(tolcuda_gather *g, tolcuda_handle h, long B, int rank, int world, int dst) {
    if (!h || B < 1 || world < 1 || world > 64 || rank < 0 || rank >= world || dst < 0 || dst >= world) {
        set_error("tolcuda_gather: need a context, B >= 1, 1 <= world <= 64, ranks inside the world");
        return TOLCUDA_EINVAL;
    }
    tolcuda_config cfg;
    int e = tolcuda_get_config(h, &cfg);
    if (e) return e;
    g->h = h, g->device = cfg.device, g->rank = rank, g->world = world, g->dst = dst, g->B = B;
    g->form = cfg.formulation, g->ts = cfg.ts;
    tolcuda_dims(h, &g->n, &g->neF, &g->neG);
    g->ldF = tolcuda_padded_ld(g->neF), g->ldG = tolcuda_padded_ld(g->neG);
    g->ldC = tolcuda_padded_ld(tolcuda_compact_len(g->form, g->ts));
    g->bytes = sizeof(double) * (size_t)B * (g->ldF + g->ldG + (world > 1 ? g->ldC : 0)) + FLAG_BYTES;
    return 0;
}

}  // namespace

extern "C" {

int tolcuda_gather_create(tolcuda_handle h, long B, int world, int dst, tolcuda_gather_handle *out,
                          unsigned char *ipc_handle) {
    if (!out) return TOLCUDA_EINVAL;
    *out = nullptr;
    tolcuda_gather *g = new (std::nothrow) tolcuda_gather();
    if (!g) return TOLCUDA_ENOMEM;
    int e = init_common(g, h, B, dst, world, dst);
    void *p = nullptr;
    if (!e) e = tolcuda_device_alloc(g->device, g->bytes, &p);
    if (!e) {
        g->base = static_cast<char *>(p), g->owner = true;
        const std::vector<char> zeros(FLAG_BYTES, 0);  // the chunk flags start at zero
        e = tolcuda_copy_to_device(g->device, g->flags(), zeros.data(), FLAG_BYTES);
    }
    if (!e && ipc_handle) e = tolcuda_ipc_export(g->device, g->base, ipc_handle);
    if (e) {
        if (p) tolcuda_device_free(g->device, p);
        delete g;
        return e;
    }
    *out = g;
    return 0;
}

int tolcuda_gather_attach(tolcuda_handle h, long B, int world, int rank, int dst, tolcuda_gather_handle owner,
                          const unsigned char *ipc_handle, tolcuda_gather_handle *out) {
    if (!out || (!owner && !ipc_handle)) return TOLCUDA_EINVAL;
    *out = nullptr;
    tolcuda_gather *g = new (std::nothrow) tolcuda_gather();
    if (!g) return TOLCUDA_ENOMEM;
    int e = init_common(g, h, B, rank, world, dst);
    if (!e && rank == dst) {
        set_error("tolcuda_gather_attach: the gathering rank uses tolcuda_gather_create");
        e = TOLCUDA_EINVAL;
    }
    if (!e && owner) {  // same process: peer access to the owner's allocation
        if (owner->B != B || owner->world != world || owner->bytes != g->bytes || owner->form != g->form || owner->ts != g->ts) {
            set_error("tolcuda_gather_attach: the owner's gather was created for another batch or problem");
            e = TOLCUDA_EINVAL;
        }
        if (!e) e = tolcuda_enable_peer(g->device, owner->device);
        if (!e) g->base = owner->base;
    } else if (!e) {  // another process: map the exported allocation on this device
        void *p = nullptr;
        e = tolcuda_ipc_open(g->device, ipc_handle, &p);
        if (!e) g->base = static_cast<char *>(p), g->ipc = true;
    }
    if (e) {
        delete g;
        return e;
    }
    *out = g;
    return 0;
}

int tolcuda_gather_close(tolcuda_gather_handle g) {
    if (!g) return 0;
    int e = 0;
    if (g->owner) e = tolcuda_device_free(g->device, g->base);
    else if (g->ipc) e = tolcuda_ipc_close(g->device, g->base);
    delete g;
    return e;
}

static void pieces(long q0, long q1, int chunks, std::vector<std::pair<long, long>> &out) {
    out.clear();
    const long per = std::max<long>(1, (q1 - q0 + chunks - 1) / chunks);
    for (long a = q0; a < q1; a += per) out.emplace_back(a, std::min(q1, a + per));
}

int tolcuda_gather_send(tolcuda_gather_handle g, const double *x, long ldx, int chunks) {
    if (!g || g->owner || !x || ldx < g->n) return TOLCUDA_EINVAL;
    chunks = std::max(1, std::min(chunks, MAX_CHUNKS));
    g->epoch++;
    long b0, b1;
    shard(g->B, g->rank, g->world, &b0, &b1);
    std::vector<std::pair<long, long>> pc;
    pieces(b0, b1, chunks, pc);
    const int fl = TOLCUDA_NEED_F | TOLCUDA_NEED_G | TOLCUDA_DEVICE_PTRS | TOLCUDA_NO_SYNC | TOLCUDA_COMPACT_G;
    for (size_t ci = 0; ci < pc.size(); ci++) {
        const long a = pc[ci].first, e = pc[ci].second;
        int rc = tolcuda_eval_batch(g->h, (int)(e - a), x + (size_t)(a - b0) * ldx, ldx, g->Fp() + (size_t)a * g->ldF, g->ldF,
                                    g->Cp() + (size_t)a * g->ldC, g->ldC, fl);
        if (!rc) rc = tolcuda_stream_signal(g->h, g->flags() + 4 * (g->rank * MAX_CHUNKS + (int)ci), g->epoch);
        if (rc) return rc;
    }
    return 0;
}

int tolcuda_gather_collect(tolcuda_gather_handle g, const double *x, long ldx, int chunks, double **F, long *ldF,
                           double **G, long *ldG) {
    if (!g || !g->owner) return TOLCUDA_EINVAL;
    chunks = std::max(1, std::min(chunks, MAX_CHUNKS));
    g->epoch++;
    long b0, b1;
    shard(g->B, g->rank, g->world, &b0, &b1);
    if (b1 > b0) {
        if (!x || ldx < g->n) return TOLCUDA_EINVAL;
        int rc = tolcuda_eval_batch(g->h, (int)(b1 - b0), x, ldx, g->Fp() + (size_t)b0 * g->ldF, g->ldF,
                                    g->Gp() + (size_t)b0 * g->ldG, g->ldG,
                                    TOLCUDA_NEED_F | TOLCUDA_NEED_G | TOLCUDA_DEVICE_PTRS | TOLCUDA_NO_SYNC);
        if (rc) return rc;
    }
    // chunk-major: the peers send concurrently, so their chunks c arrive at about the same time
    std::vector<std::pair<long, long>> pc;
    for (int ci = 0; ci < chunks; ci++)
        for (int q = 0; q < g->world; q++) {
            if (q == g->rank) continue;
            long q0, q1;
            shard(g->B, q, g->world, &q0, &q1);
            pieces(q0, q1, chunks, pc);
            if (ci >= (int)pc.size()) continue;
            const long a = pc[ci].first, e = pc[ci].second;
            int rc = tolcuda_stream_wait(g->h, g->flags() + 4 * (q * MAX_CHUNKS + ci), g->epoch);
            if (!rc)
                rc = tolcuda_expand_compact_g_device(g->h, e - a, g->Cp() + (size_t)a * g->ldC, g->ldC,
                                                     g->Gp() + (size_t)a * g->ldG, g->ldG, TOLCUDA_NO_SYNC);
            if (rc) return rc;
        }
    int rc = tolcuda_synchronize(g->h);
    if (rc) return rc;
    if (F) *F = g->Fp();
    if (ldF) *ldF = g->ldF;
    if (G) *G = g->Gp();
    if (ldG) *ldG = g->ldG;
    return 0;
}

int tolcuda_gather_buffer(tolcuda_gather_handle g, void **base, size_t *bytes) {
    if (!g) return TOLCUDA_EINVAL;
    if (base) *base = g->base;
    if (bytes) *bytes = g->bytes;
    return 0;
}

int tolcuda_copy_to_device(int device, void *dst, const void *src, size_t bytes) {
    return tolcuda_copy_raw(device, dst, src, bytes, 1);
}

int tolcuda_copy_to_host(int device, void *dst, const void *src, size_t bytes) {
    return tolcuda_copy_raw(device, dst, src, bytes, 2);
}

}  // extern "C"
