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
    Fl = torch.zeros(per, ev.neF, dtype=torch.float64, device=X.device)
    Gl = torch.zeros(per, Lc, dtype=torch.float64, device=X.device)
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
    torch.cuda.current_stream().synchronize()  # the gathers ran on torch's stream, the expansion runs on the context's
    F = torch.empty(B, ev.neF, dtype=torch.float64, device=X.device)
    G = torch.empty(B, ev.neG, dtype=torch.float64, device=X.device)
    for q in range(w):
        b0, b1 = shard_range(B, q, w)
        if b1 > b0:
            F[b0:b1] = Fp[q][:b1 - b0]
            ev.expand_compact_device(Gp[q][:b1 - b0], G[b0:b1], sync=False)
    ev.synchronize()
    return F, G


def eval_and_gather_peer(ev, X, B, dst=0, out=None, compact=True, chunks=4):
    """Fused evaluate + gather: rank `dst` owns one buffer holding all B rows of F and
    G; every other rank maps it over CUDA IPC (NVLink peer access) and its F/G kernel stores its shard's rows
    directly into it -- F with coalesced stores, G with the kernel's TMA bulk copies -- so the transfer happens
    record by record while the shard is being computed: no send buffer, no copy or collective afterwards.
    torch.distributed only carries the 64-byte handle and the closing barrier.
      compact=True   the peers' G crosses NVLink as COMPACT rows (a third of the bytes: NVLink, at 0.9 TB/s per
                     direction, is the slower side) into a staging region of the same buffer, and `dst` expands
                     them into rows in coordinate order at HBM speed (expand_kernel.cu).  A peer's shard goes in
                     `chunks` launches, each followed by a stream-ordered flag written into the owner's memory
                     (tolcuda_stream_signal); the owner's stream waits on the flag (tolcuda_stream_wait) and
                     expands that chunk while the next ones are still arriving -- no host in between
      compact=False  the peers write full rows at their final place; nothing runs on `dst` afterwards
    Returns (F [B, ldF], G [B, ldG], buffer) as CUDA tensors on rank `dst` (rows by trajectory index, padded
    leading dimensions), (None, None, None) elsewhere.  `out`: what open_peer_buffer returned on this rank, to
    reuse the allocation and its mappings over many calls (mapping a handle costs far more than a launch)."""
    from .evaluator import padded_ld
    r, w = world()
    b0, b1 = shard_range(B, r, w)
    ldF, ldG, ldC = padded_ld(ev.neF), padded_ld(ev.neG), padded_ld(ev.compact_len)
    staged = compact and w > 1
    chunks = max(1, min(int(chunks), MAX_CHUNKS))
    nbytes = peer_buffer_bytes(ev, B, staged)
    buf = out if out is not None else open_peer_buffer(ev, B, dst, staged)
    assert buf.nbytes >= nbytes
    buf.epoch = getattr(buf, "epoch", 0) + 1  # every rank calls in step, so the counters agree
    assert not staged or buf.nbytes == nbytes, "the buffer was opened for another layout"
    flags = buf.ptr + nbytes - FLAG_BYTES  # uint32 [w][MAX_CHUNKS], zeroed by open_peer_buffer

    def pieces(q0, q1):
        per = max(1, (q1 - q0 + chunks - 1) // chunks)
        return [(a, min(q1, a + per)) for a in range(q0, q1, per)]

    if b1 > b0:
        assert X.shape[0] == b1 - b0
        if staged and r != dst:
            for ci, (a, e) in enumerate(pieces(b0, b1)):
                ev.eval_batch_ptrs(e - a, X[a - b0:].data_ptr(), X.stride(0), buf.ptr + 8 * a * ldF, ldF,
                                   buf.ptr + 8 * (B * (ldF + ldG) + a * ldC), ldC, compact_rows=True, sync=False)
                ev.stream_signal(flags + 4 * (r * MAX_CHUNKS + ci), buf.epoch)
            ev.synchronize()
        else:
            ev.eval_batch_ptrs(b1 - b0, X.data_ptr(), X.stride(0), buf.ptr + 8 * b0 * ldF, ldF,
                               buf.ptr + 8 * (B * ldF + b0 * ldG), ldG, sync=staged is False or w == 1)
    F = G = None
    if r == dst:
        F, G = buf.tensor(0, B, ldF), buf.tensor(B * ldF, B, ldG)
        if staged:
            Gc = buf.tensor(B * (ldF + ldG), B, ldC)
            # chunk-major: the peers send concurrently, so their chunks c arrive at about the same time
            todo = [(ci, q, a, e) for q in range(w) if q != dst
                    for ci, (a, e) in enumerate(pieces(*shard_range(B, q, w)))]
            for ci, q, a, e in sorted(todo):
                ev.stream_wait(flags + 4 * (q * MAX_CHUNKS + ci), buf.epoch)
                ev.expand_compact_device(Gc[a:e], G[a:e], sync=False)
            ev.synchronize()
    if w > 1:
        dist.barrier()  # every shard has landed (and been expanded) in the owner's memory; the staging region is free again
        if r != dst:
            if out is None:
                buf.close()
            return None, None, None
    return F, G, buf


MAX_CHUNKS = 16


def open_peer_buffer(ev, B, dst=0, staged=True):
    """collective: rank `dst` allocates the gather buffer (peer_buffer_bytes) and exports it, every other rank
    maps it on its own device; returns this rank's PeerBuffer (close() it when done)"""
    from .evaluator import PeerBuffer
    r, w = world()
    nbytes = peer_buffer_bytes(ev, B, staged and w > 1)
    buf = PeerBuffer.alloc(ev.device, nbytes) if r == dst else None
    if r == dst:  # the chunk flags start at zero
        buf.tensor((nbytes - FLAG_BYTES) // 8, 1, FLAG_BYTES // 8).zero_()
        torch.cuda.synchronize(ev.device)
    if w > 1:
        box = [buf.handle if r == dst else None]
        dist.broadcast_object_list(box, src=dst)
        if r != dst:
            buf = PeerBuffer.open(ev.device, box[0], nbytes)
    return buf


def peer_buffer_bytes(ev, B, staged=True):
    """bytes of the gathering rank's buffer: F rows | G rows | (staged) compact rows of the peers"""
    from .evaluator import padded_ld
    ldF, ldG, ldC = padded_ld(ev.neF), padded_ld(ev.neG), padded_ld(ev.compact_len)
    return 8 * B * (ldF + ldG + (ldC if staged else 0)) + FLAG_BYTES


FLAG_BYTES = 4096  # uint32 [world][MAX_CHUNKS] chunk flags behind the staging region (world <= 64)
