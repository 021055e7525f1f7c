#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric: FP64 F+G node-evals/s = trajectories x collocation windows evaluated per second.
Workload (N=1 and per GPU at N>1, weak scaling): S10 / tempest, B = 65,536 trajectories x ts = 200
windows (BASELINE.json configs[3], the configuration the metric's 60 %-of-HBM target is quoted on;
13.2 GB, fits one B200), synthetic inputs per SURVEY.md section 8d.  One "step" = one F+G pass over
the whole batch = one kernel launch.

  value     device-resident: x, F, G live in HBM; K launches timed with CUDA events on the launching
            stream; 13.2 GB touched per step >> 126 MB L2, so no L2 flush is needed.
  e2e       the same pass through tolcuda_eval_batch with HOST (pinned) x, F, G: chunked
            H2D -> kernel -> D2H inside the timed region; full F and G rows are in host memory at the end
            of every step.  Default path: G crosses PCIe as compact rows and the library's host threads
            write the caller's rows; e2e.full_g_copy is the same call with every G value crossing PCIe.
  roofline  algorithmic bytes 8*(n+neF+neG) per trajectory / average launch duration / measured HBM peak.
  cpu_baseline (N=1)  the reference's own CPU path (oracle/_ref, unmodified sources, -O2) on all host
            cores as independent processes, on a bounded sample of the same batch.

--impl reference runs only that CPU arm and prints the same JSON shape with "impl": "reference"."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "fp64_FG_node_evals_per_s"
UNIT = "node-evals/s"
WORKLOADS = {
    # name: (fixture with x0 + parameters, seed0, default batch)
    "S10_tempest_ts200_B65536": ("S10_tempest_ts200", 20270000, 65536),
    "G7_skywalker_ts100_B4096": ("G7_skywalker_ts100", 20260000, 4096),
}


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


# ------------------------------------------------------------------------------ CPU reference arm

_W = {}


def _worker_init(fixture, use_ref):
    """one reference (or port) problem object per process: the reference is not thread-safe
    (process-global `prob`, member scratch)"""
    g = golden(fixture)
    _W["g"] = g
    if use_ref:
        import refclient as R
        goal = g["goal_enu"]
        ts = int(g["ts"])
        _W["p"] = R.RefProblem(str(g["mission"]), str(g["aircraft"]), tuple(g["enu"]), tuple(goal),
                               ts=None if ts == 100 else ts, null_io=True)
    else:
        import portclient as P
        _W["p"] = P.PortProblem(str(g["mission"]), int(g["ts"]), g["ac"], g["gn"], g["goal_ned"],
                                int(g["wind_model"]))


def _worker_run(args):
    seed0, b0, count = args
    import tol_b200.synth as synth
    p, g = _W["p"], _W["g"]
    key = (seed0, b0, count)
    if _W.get("key") != key:  # inputs and outputs are prepared once, outside the timed step
        _W["key"] = key
        _W["X"] = synth.batch(g["x"][0], seed0, b0, b0 + count)
        _W["F"] = np.empty((count, p.neF))
        _W["G"] = np.empty((count, p.neG))
    X, F, G = _W["X"], _W["F"], _W["G"]
    t0 = time.perf_counter()
    p.eval_many(X, F, G)
    dt = time.perf_counter() - t0
    return dt, float(F[:, 0].sum())


class CpuArm:
    """the reference's CPU implementation of the path on all usable host cores"""

    def __init__(self, fixture, seed0, per_worker):
        import multiprocessing as mp
        import portclient as P
        import refclient as R
        self.fixture, self.seed0, self.per_worker = fixture, seed0, per_worker
        self.use_ref = R.available()
        if not self.use_ref and not P.available():
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        self.cores = max(1, min(len(os.sched_getaffinity(0)), 128))
        self.kind = "reference" if self.use_ref else "port"
        self.pool = mp.get_context("fork").Pool(self.cores, _worker_init, (fixture, self.use_ref))
        self.ts = int(golden(fixture)["ts"])

    def step(self):
        """every worker evaluates `per_worker` trajectories; returns node-evals/s of the step"""
        jobs = [(self.seed0, w * self.per_worker, self.per_worker) for w in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_worker_run, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return self.cores * self.per_worker * self.ts / wall, wall, max(r[0] for r in res)

    def sample(self):
        what = ("unmodified reference modelWind+computeF+computeG (oracle/_ref, g++ -O2, debug dumps to /dev/null)"
                if self.use_ref else "oracle/fg_oracle.c port (non-redundant restatement, gcc -O2)")
        return "%d processes x %d trajectories x %d windows per step; %s" % (
            self.cores, self.per_worker, self.ts, what)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args, wl_name):
    fixture, seed0, _ = WORKLOADS[wl_name]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(fixture, seed0, args.cpu_sample)
    for _ in range(max(1, min(args.warmup, 2))):
        arm.step()
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        arm.step()
        total += arm.cores * arm.per_worker * arm.ts
    wall = time.perf_counter() - t0
    value = total / wall
    g = golden(fixture)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl_name, "mission": str(g["mission"]), "ts": int(g["ts"]),
                   "step": "bounded sample: " + arm.sample()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.sample()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    arm.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm

class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], 0, set()
        for t, ln in self.rows:
            if not (t0 <= t <= t1 + 0.2):
                continue
            f = [s.strip() for s in ln.split(",")]
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(wl_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, recorded from an
    `ncu --set full` capture under profiles/ (never measured during a bench run)"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(wl_name)
        except ValueError:
            return None
    return None


def scaled_traffic(wl_name, alg_bytes):
    t = ncu_traffic(wl_name)
    return None if not t else t["ratio"] * alg_bytes


def run_ours(args, wl_name):
    import torch
    import torch.distributed as dist
    import tol_b200 as T
    from tol_b200.evaluator import padded_ld

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libtolcuda has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fixture, seed0, Bdef = WORKLOADS[wl_name]
    B = args.batch or Bdef
    g = golden(fixture)
    ev = T.Evaluator.from_golden(g, device=local)
    ts, n, neF, neG = int(g["ts"]), ev.n, ev.neF, ev.neG
    ldx, ldF, ldG = padded_ld(n), padded_ld(neF), padded_ld(neG)

    # weak scaling: every rank owns B trajectories, global indices [rank*B, (rank+1)*B)
    Xh = torch.zeros(B, ldx, dtype=torch.float64, pin_memory=True)
    T.synth.batch(g["x"][0], seed0, rank * B, (rank + 1) * B, out=Xh.numpy())
    Xd = Xh.cuda()
    Fd = torch.empty(B, ldF, dtype=torch.float64, device="cuda")
    Gd = torch.empty(B, ldG, dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream()  # kernels and the timing events share this stream
    torch.cuda.set_stream(stream)
    ev.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: K launches, CUDA events on the launching stream
    for _ in range(args.warmup):
        ev.eval_batch_device(Xd, Fd, Gd, sync=False)
    barrier()
    sampler = ClockSampler(local)
    time.sleep(0.3)
    l0 = ev.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        ev.eval_batch_device(Xd, Fd, Gd, sync=False)
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    ms_dev = e0.elapsed_time(e1)
    launches = ev.launches - l0
    clocks = sampler.stop(tw0, tw1)

    # parity spot check of what the timed launches left in HBM (outside the timed region)
    checked = None
    if rank == 0:
        try:
            import portclient as P
            if not P.available():
                subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
            port = P.PortProblem(str(g["mission"]), ts, g["ac"], g["gn"], g["goal_ned"], int(g["wind_model"]))
            rows = np.linspace(0, B - 1, 8).astype(int)
            Xs = np.ascontiguousarray(Xh.numpy()[rows, :n])
            Fr, Gr = np.empty((rows.size, neF)), np.empty((rows.size, neG))
            port.eval_many(Xs, Fr, Gr)
            idx = torch.from_numpy(rows).cuda()
            Fg, Gg = Fd[idx, :neF].cpu().numpy(), Gd[idx, :neG].cpu().numpy()
            checked = bool((np.abs(Fg - Fr) <= 1e-14 + 1e-12 * np.abs(Fr)).all()
                           and (np.abs(Gg - Gr) <= 1e-14 + 1e-12 * np.abs(Gr)).all())
        except Exception as exc:  # the check is a courtesy; never hide the numbers behind it
            checked = "unavailable: %s" % exc

    # ---- end to end: host (pinned) buffers through tolcuda_eval_batch, copies inside the timed region
    del Fd, Gd
    torch.cuda.empty_cache()
    ev.use_own_stream()
    host_threads = max(1, min(8, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))
    ev.set_host_threads(host_threads)
    Fh = torch.empty(B, ldF, dtype=torch.float64, pin_memory=True)
    Gh = torch.empty(B, ldG, dtype=torch.float64, pin_memory=True)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_run(steps, full_copy):
        for _ in range(min(args.warmup, 2)):
            ev.eval_batch_host(Xh.numpy(), Fh.numpy(), Gh.numpy(), full_copy=full_copy)
        barrier()
        l0_ = ev.launches
        t0_ = time.perf_counter()
        for _ in range(steps):
            ev.eval_batch_host(Xh.numpy(), Fh.numpy(), Gh.numpy(), full_copy=full_copy)
        torch.cuda.synchronize()
        t_ = time.perf_counter() - t0_
        barrier()
        return t_, ev.launches - l0_

    # default path: compact G rows across PCIe + expansion on host threads; then, for comparison, the same
    # call with TOLCUDA_FULL_G_COPY (every G value crosses PCIe)
    t_e2e, e2e_launches = e2e_run(e2e_steps, False)
    full_steps = max(1, min(e2e_steps, 2))
    t_full, _ = e2e_run(full_steps, True)
    if rank == 0 and checked is True:  # the rows the e2e path left in host memory, against the oracle rows
        checked = bool((np.abs(Fh.numpy()[rows, :neF] - Fr) <= 1e-14 + 1e-12 * np.abs(Fr)).all()
                       and (np.abs(Gh.numpy()[rows, :neG] - Gr) <= 1e-14 + 1e-12 * np.abs(Gr)).all())

    # max over ranks
    tt = torch.tensor([ms_dev, t_e2e, t_full], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_dev, t_e2e, t_full = float(tt[0]), float(tt[1]), float(tt[2])

    units_step = world * B * ts
    value = units_step * args.steps / (ms_dev * 1e-3)
    e2e_value = units_step * e2e_steps / t_e2e
    alg_bytes = 8.0 * (n + neF + neG) * B  # per launch (one rank)
    launch_ms = ms_dev / args.steps
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "mission": str(g["mission"]), "aircraft": str(g["aircraft"]),
                   "ts": ts, "trajectories_per_gpu": B, "n": n, "neF": neF, "neG": neG,
                   "parallelism": "shard by trajectory index, no collective",
                   "l2": "%.1f GB touched per step, far larger than the 126 MB L2; no flush needed" % (alg_bytes / 1e9),
                   "inputs": "x0*(1+0.05u)+0.01u', PCG64(seed0+b), seed0=%d (SURVEY.md 8d)" % seed0,
                   "parity_spot_check": checked},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n * B * world,
                "d2h_bytes_per_step": 8 * (neF + padded_ld(ev.compact_len)) * B * world,
                "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps, "gpu_launches": e2e_launches,
                "api": "tolcuda_eval_batch(TOLCUDA_HOST_PTRS), pinned host x/F/G; full F and G rows in host memory "
                       "at the end of every step; G crosses PCIe as compact rows (x-dependent values only) and "
                       "host threads write the rows in coordinate order, structural constants as literals",
                "host_threads": host_threads,
                "full_g_copy": {"value": units_step * full_steps / t_full, "unit": UNIT, "steps": full_steps,
                                "ms_per_step": 1e3 * t_full / full_steps,
                                "d2h_bytes_per_step": 8 * (neF + neG) * B * world,
                                "api": "same call with TOLCUDA_FULL_G_COPY: every G value crosses PCIe"}},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": scaled_traffic(wl_name, alg_bytes), "traffic_source": "profiles/roofline_traffic.json (ncu --set full capture of the same kernel, ratio to algorithmic bytes applied)", "peak_source": peak_src,
                     "kernel": "fg_cta_kernel<%s, wind %d, PLAIN, runs of 2 trajectories per CTA>" % (str(g["mission"]), int(g["wind_model"])), "algorithmic_bytes_per_launch": alg_bytes,
                     "launch_ms": launch_ms},
        "clocks": clocks,
    }
    ev.close()
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            del Xd, Fh, Gh
            arm = CpuArm(fixture, seed0, args.cpu_sample)
            arm.step()
            v, wall, _ = arm.step()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                    "sample": arm.sample(), "seconds": wall}
            arm.close()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S10_tempest_ts200_B65536", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override trajectories per GPU (experiments)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1000, help="trajectories per CPU process per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
