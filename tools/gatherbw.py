"""Device-side gather of the shards of N GPUs on one of them, three ways, timed as the max
over ranks of the wall time between two barriers (kernels, transfers and the owner's expansion included):

  nccl     tol_b200.dist.eval_and_gather_device: compact rows, one NCCL gather per array, expansion on the owner
  peer     tol_b200.dist.eval_and_gather_peer(compact=True): the peers' kernels store F and compact G rows straight
           into the owner's memory over NVLink in 4 chunks (peer1: 1, peer8: 8), each followed by a stream-ordered
           flag the owner's stream waits on before it expands that chunk
  peerfull eval_and_gather_peer(compact=False): the peers' kernels store full rows at their final place

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/gatherbw.py [B_total] [fixture]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import tol_b200 as T  # noqa: E402
import tol_b200.dist as D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
name = sys.argv[2] if len(sys.argv) > 2 else "S10_tempest_ts200"
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dist.get_world_size()
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
ev = T.Evaluator.from_golden(g, device=local)
b0, b1 = D.my_shard(B)
U = 64  # distinct trajectories, repeated (inputs do not change the timing)
Xu = torch.from_numpy(T.synth.batch(g["x"][0], 7, 0, U)).cuda()
X = Xu[(torch.arange(b0, b1, device="cuda")) % U].contiguous()


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


res = {}
res["nccl"] = timed(lambda: D.eval_and_gather_device(ev, X, B, dst=0))
ref = D.eval_and_gather_device(ev, X, B, dst=0)
for label, compact, chunks in (("peer1", True, 1), ("peer", True, 4), ("peer8", True, 8), ("peerfull", False, 1)):
    buf = D.open_peer_buffer(ev, B, 0, compact)
    res[label] = timed(lambda: D.eval_and_gather_peer(ev, X, B, dst=0, out=buf, compact=compact, chunks=chunks))
    F, G, _ = D.eval_and_gather_peer(ev, X, B, dst=0, out=buf, compact=compact, chunks=chunks)
    if rank == 0:
        same = torch.equal(F[:, :ev.neF], ref[0]) and torch.equal(G[:, :ev.neG], ref[1])
        res[label + "_bitexact_vs_nccl"] = bool(same)
    del F, G
    dist.barrier()
    buf.close()
if rank == 0:
    rows = 8.0 * B * (ev.neF + ev.neG)
    print("%s  B=%d on %d GPUs -> GPU 0   (rows gathered: %.2f GB)" % (name, B, w, rows / 1e9))
    for k in ("nccl", "peer1", "peer", "peer8", "peerfull"):
        print("  %-9s %8.3f ms   %6.1f GB/s of gathered rows   %.3e node-evals/s" % (
            k, res[k] * 1e3, rows / res[k] / 1e9, B * int(g["ts"]) / res[k]))
    print("  bit-exact vs nccl path:", [res.get(k + "_bitexact_vs_nccl") for k in ("peer1", "peer", "peer8", "peerfull")])
ev.close()
dist.destroy_process_group()
