"""One process per GPU, trajectories sharded by index, no collective on the data path (SURVEY.md
section 8e).  torch.distributed is used for the three things around the path: agreeing on the shard
ranges, the max-over-ranks of the timings, and the optional host-side gather of per-trajectory rows to
rank 0 (the `north_star`'s "host-side gather of per-trajectory F/G").  Two device-side gathers exist beside it
(SURVEY.md section 8f-4): compact rows through one NCCL gather (eval_and_gather_device) and the fused form, in
which every rank's F/G kernel writes its rows straight into the gathering GPU's memory over NVLink
(eval_and_gather_peer)."""
import numpy as np
import torch
import torch.distributed as dist

from .synth import shard_range


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def my_shard(B):
    """[b0, b1) of the global batch owned by this rank"""
    r, w = world()
    return shard_range(B, r, w)


def max_over_ranks(value, device="cpu"):
    r, w = world()
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_rows(local, B):
    """local: this rank's rows [b1-b0, m] (numpy, float64).  Returns the full [B, m] array on rank 0
    (None elsewhere).  Rows are placed by trajectory index, so the result does not depend on the
    number of ranks."""
    r, w = world()
    if w == 1:
        return local
    m = local.shape[1]
    per = (B + w - 1) // w
    pad = np.zeros((per, m))
    pad[:local.shape[0]] = local
    mine = torch.from_numpy(pad)
    parts = [torch.empty_like(mine) for _ in range(w)] if r == 0 else None
    dist.gather(mine, parts, dst=0)
    if r != 0:
        return None
    out = np.empty((B, m))
    for q in range(w):
        b0, b1 = shard_range(B, q, w)
        out[b0:b1] = parts[q].numpy()[:b1 - b0]
    return out


def eval_and_gather_device(ev, X, B, dst=0):
    """Device-side gather: every rank evaluates its shard X (torch CUDA tensor
    [b1-b0, >= n]) into F rows and COMPACT G rows, the shards are gathered on GPU `dst` with one NCCL gather
    per array -- G crosses NVLink as compact rows, a third of the bytes -- and expanded there into rows in
    coordinate order by the device expansion kernel.  Returns (F [B, neF], G [B, neG]) as CUDA tensors on rank
    `dst`, (None, None) elsewhere; rows are placed by trajectory index."""
    r, w = world()
    nb = X.shape[0]
    per = (B + w - 1) // w
    Lc = ev.compact_len
    # the kernel runs on the stream torch has queued these allocations (and the tail fill) on: Evaluator follows
    # torch's current stream unless the caller pinned another one, in which case the caller orders the two
    Fl = torch.empty(per, ev.neF, dtype=torch.float64, device=X.device)
    Gl = torch.empty(per, Lc, dtype=torch.float64, device=X.device)
    Fl[nb:].zero_()  # rows of the padded shard no trajectory owns
    Gl[nb:].zero_()
    if nb:
        ev.eval_batch_device(X, Fl[:nb], Gl[:nb], compact_rows=True)
    if w == 1:
        G = torch.empty(B, ev.neG, dtype=torch.float64, device=X.device)
        ev.expand_compact_device(Gl[:B], G)
        return Fl[:B], G
    Fp = [torch.empty_like(Fl) for _ in range(w)] if r == dst else None
    Gp = [torch.empty_like(Gl) for _ in range(w)] if r == dst else None
    dist.gather(Fl, Fp, dst=dst)
    dist.gather(Gl, Gp, dst=dst)
    if r != dst:
        return None, None
    torch.cuda.current_stream().synchronize()  # the gathers were issued from torch's stream; the expansions below are ordered after them
    F = torch.empty(B, ev.neF, dtype=torch.float64, device=X.device)
    G = torch.empty(B, ev.neG, dtype=torch.float64, device=X.device)
    for q in range(w):
        b0, b1 = shard_range(B, q, w)
        if b1 > b0:
            F[b0:b1] = Fp[q][:b1 - b0]
            ev.expand_compact_device(Gp[q][:b1 - b0], G[b0:b1], sync=False)
    ev.synchronize()
    return F, G


class Gather:
    """tolcuda_gather_* (include/tolcuda.h, gather.cpp) for one process per GPU: rank `dst` creates the gather and its
    64-byte IPC handle travels through torch.distributed; every other rank attaches on its own device.  Collective."""

    def __init__(self, ev, B, dst=0):
        import ctypes as C
        from . import lib as _l
        r, w = world()
        self.ev, self.B, self.dst, self.rank, self.world, self.L = ev, B, dst, r, w, ev.L
        self.g = C.c_void_p()
        handle = C.create_string_buffer(64)
        if r == dst:
            _l.check(self.L.tolcuda_gather_create(ev.h, B, w, dst, C.byref(self.g), handle))
        box = [handle.raw if r == dst else None]
        if w > 1:
            dist.broadcast_object_list(box, src=dst)
        if r != dst:
            _l.check(self.L.tolcuda_gather_attach(ev.h, B, w, r, dst, None, box[0], C.byref(self.g)))
        base, nbytes = C.c_void_p(), C.c_size_t()
        _l.check(self.L.tolcuda_gather_buffer(self.g, C.byref(base), C.byref(nbytes)))
        self.ptr, self.nbytes = base.value, nbytes.value

    def run(self, X, chunks=4):
        """every rank: evaluate the local shard X into the gathering GPU's rows; returns (F [B, ldF], G [B, ldG]) as
        CUDA tensors on rank dst (views of the gather's buffer, valid until close()), (None, None) elsewhere.  Ends
        with a barrier: the staging region is free again when it returns."""
        import ctypes as C
        from . import lib as _l
        from .evaluator import PeerBuffer
        F = G = None
        self.ev._follow_torch(X)  # what torch has queued on X (and, on the owner, on the buffer) comes first
        xp = C.c_void_p(X.data_ptr()) if X.shape[0] else None
        if self.rank != self.dst:
            if X.shape[0]:
                _l.check(self.L.tolcuda_gather_send(self.g, xp, X.stride(0), int(chunks)))
                self.ev.synchronize()
        else:
            Fp, Gp, ldF, ldG = C.c_void_p(), C.c_void_p(), C.c_long(), C.c_long()
            _l.check(self.L.tolcuda_gather_collect(self.g, xp, X.stride(0) if X.shape[0] else self.ev.n, int(chunks),
                                                   C.byref(Fp), C.byref(ldF), C.byref(Gp), C.byref(ldG)))
            view = PeerBuffer(self.ev.device, self.ptr, self.nbytes, None)
            F = view.tensor((Fp.value - self.ptr) // 8, self.B, ldF.value)
            G = view.tensor((Gp.value - self.ptr) // 8, self.B, ldG.value)
        if self.world > 1:
            dist.barrier()
        return F, G

    def close(self):
        if getattr(self, "g", None):
            self.L.tolcuda_gather_close(self.g)
            self.g = None

    __del__ = close


def eval_and_gather_peer(ev, X, B, dst=0, out=None, compact=True, chunks=4):
    """Fused evaluate + gather: rank `dst` owns one buffer holding all B rows of F and
    G; every other rank maps it over CUDA IPC (NVLink peer access) and its F/G kernel stores its shard's rows
    directly into it -- F with coalesced stores, G with the kernel's TMA bulk copies -- so the transfer happens
    record by record while the shard is being computed: no send buffer, no copy or collective afterwards.
    torch.distributed only carries the 64-byte handle and the closing barrier.
      compact=True   the library's own protocol, tolcuda_gather_* (class Gather): the peers' G crosses NVLink as
                     COMPACT rows (a third of the bytes: NVLink, at 0.9 TB/s per direction, is the slower side)
                     into a staging region of the same buffer in `chunks` launches, each followed by a
                     stream-ordered flag in the owner's memory; the owner's stream waits on the flag and expands
                     that chunk (expand_kernel.cu) while the next ones are still arriving -- no host in between
      compact=False  the peers write full rows at their final place; nothing runs on `dst` afterwards
    Returns (F [B, ldF], G [B, ldG], buffer) as CUDA tensors on rank `dst` (rows by trajectory index, padded
    leading dimensions), (None, None, None) elsewhere.  `out`: what open_peer_buffer returned on this rank, to
    reuse the allocation and its mappings over many calls (mapping a handle costs far more than a launch)."""
    from .evaluator import padded_ld
    r, w = world()
    b0, b1 = shard_range(B, r, w)
    assert X.shape[0] == b1 - b0
    if compact and w > 1:
        g = out if out is not None else Gather(ev, B, dst)
        assert isinstance(g, Gather), "the buffer was opened for another layout"
        F, G = g.run(X, chunks)
        if r != dst:
            if out is None:
                g.close()
            return None, None, None
        return F, G, g
    ldF, ldG = padded_ld(ev.neF), padded_ld(ev.neG)
    nbytes = peer_buffer_bytes(ev, B, False)
    buf = out if out is not None else open_peer_buffer(ev, B, dst, False)
    assert not isinstance(buf, Gather) and buf.nbytes >= nbytes, "the buffer was opened for another layout"
    if b1 > b0:
        ev.eval_batch_ptrs(b1 - b0, X.data_ptr(), X.stride(0), buf.ptr + 8 * b0 * ldF, ldF,
                           buf.ptr + 8 * (B * ldF + b0 * ldG), ldG, sync=True)
    F = G = None
    if r == dst:
        F, G = buf.tensor(0, B, ldF), buf.tensor(B * ldF, B, ldG)
    if w > 1:
        dist.barrier()  # every shard has landed in the owner's memory
        if r != dst:
            if out is None:
                buf.close()
            return None, None, None
    return F, G, buf



def open_peer_buffer(ev, B, dst=0, staged=True):
    """collective: what eval_and_gather_peer reuses over many calls -- a Gather (the library's protocol: compact rows,
    staging region, chunk flags) when `staged` and there are peers, else a plain PeerBuffer for full rows in place
    (rank `dst` allocates and exports it, every other rank maps it on its own device).  close() it when done."""
    from .evaluator import PeerBuffer
    r, w = world()
    if staged and w > 1:
        return Gather(ev, B, dst)
    nbytes = peer_buffer_bytes(ev, B, False)
    buf = PeerBuffer.alloc(ev.device, nbytes) if r == dst else None
    if w > 1:
        box = [buf.handle if r == dst else None]
        dist.broadcast_object_list(box, src=dst)
        if r != dst:
            buf = PeerBuffer.open(ev.device, box[0], nbytes)
    return buf


def peer_buffer_bytes(ev, B, staged=True):
    """bytes of a full-rows PeerBuffer: F rows | G rows (the staged layout is the library's, tolcuda_gather_buffer)"""
    from .evaluator import padded_ld
    assert not staged or world()[1] == 1
    return 8 * B * (padded_ld(ev.neF) + padded_ld(ev.neG))
