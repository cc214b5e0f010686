#!/usr/bin/env python
"""bench.py — BoomerAMG-PCG solve throughput (MDOF/s) on B200, BASELINE.json's metric.

    python bench.py --gpus 1 --steps K --warmup W            # hb200 arm (this repository)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path

A "step" is one complete AMG-PCG solve (x0 = 0, b = ones, tol 1e-8, two-norm stopping, V(1,1)
l1-Jacobi, HMIS + ext+i hierarchy from the reference's own BoomerAMGSetup) of
`ij -27pt -n 256 256 256 -solver 1 -rlx 18` (BASELINE.json configs[1]); with N GPUs each rank
owns one 256^3 brick (weak scaling, `-P` process grid as ij lays it out).
value = global rows / solve time / 1e6.  The hierarchy setup (reference, CPU) and its upload
are timed separately and reported in `config`, never inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hb200", choices=["hb200", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=256, help="brick edge per GPU")
    ap.add_argument("--problem", default="27pt", choices=["27pt", "laplacian", "vardifconv"])
    ap.add_argument("--solver", default="pcg", choices=sorted(SOLVERS))
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl", "peer"],
                    help="N>1 halo transport: NVLink peer puts, NCCL send/recv, or peer when every rank can map every peer")
    ap.add_argument("--cpu-iters", type=int, default=3, help="iterations of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-timeout", type=float, default=420.0,
                    help="seconds a stage of the hb200 arm may take before the rank gives up")
    ap.add_argument("--spmv-only", action="store_true",
                    help="BASELINE.json configs[4]: the ParCSR SpMV loop of `ij -solver -1 -nmv 100` (no AMG setup), "
                         "GB/s against the HBM peak, stored format and general CSR kernel")
    ap.add_argument("--nmv", type=int, default=100, help="matvecs per step of --spmv-only (ij -nmv)")
    ap.add_argument("--format", default="auto", choices=["auto", "csr"],
                    help="csr: every block stays in plain CSR (HB200_NO_PAT / NO_SELL / NO_CSR16): the general-matrix "
                         "solve next to the structured one")
    ap.add_argument("--global-size", action="store_true",
                    help="--n is the edge of the GLOBAL grid (strong scaling; BASELINE.json configs[3]: "
                         "-vardifconv -n 256^3 -P 2 2 2)")
    ap.add_argument("--no-e2e-ij", action="store_true",
                    help="skip the `oracle/_ref/ij_b200` leg (the solve time the unmodified ij driver prints)")
    ap.add_argument("--coarsen-type", type=int, default=-1,
                    help="BoomerAMG coarsening of the reference setup (-1: ij's CPU default, HMIS = 10; 8: PMIS, "
                         "the only one hypre's own device setup has — used to compare with baseline/_ref/ij_hypre_cuda)")
    ap.add_argument("--mpi-worker", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


# --solver: ij's solver id, metric name, label, reference-side runner, device-side class
SOLVERS = {
    "pcg": (1, "amg_pcg_solve_mdof_per_s", "BoomerAMG-PCG"),
    "gmres": (3, "amg_gmres_solve_mdof_per_s", "BoomerAMG-GMRES(5)"),
    "bicgstab": (9, "amg_bicgstab_solve_mdof_per_s", "BoomerAMG-BiCGSTAB"),
    "cogmres": (16, "amg_cogmres_solve_mdof_per_s", "BoomerAMG-COGMRES(5)"),
    "flexgmres": (61, "amg_flexgmres_solve_mdof_per_s", "BoomerAMG-FlexGMRES(5)"),
    "lgmres": (51, "amg_lgmres_solve_mdof_per_s", "BoomerAMG-LGMRES(5, 2 augmentation vectors)"),
}


def reference_solve(pb, args, max_iter):
    """the reference's own solve of the configured Krylov driver (oracle/ref_bridge.c: ij's settings)"""
    if args.solver == "pcg":
        return pb.pcg(precond="amg", tol=args.tol, max_iter=max_iter, two_norm=1)
    if args.solver == "gmres":
        return pb.gmres(precond="amg", tol=args.tol, max_iter=max_iter, k_dim=5)
    return pb.krylov_ext(args.solver, precond="amg", tol=args.tol, max_iter=max_iter, k_dim=5)


def device_solver(hb, args):
    if args.solver == "pcg":
        return hb.ParCSRPCG(tol=args.tol, max_iter=100, two_norm=1, logging=1)
    if args.solver == "gmres":
        return hb.ParCSRGMRES(tol=args.tol, max_iter=100, k_dim=5, logging=1)
    if args.solver == "bicgstab":
        return hb.ParCSRBiCGSTAB(tol=args.tol, max_iter=100, logging=1)
    if args.solver == "cogmres":
        return hb.ParCSRCOGMRES(tol=args.tol, max_iter=100, k_dim=5, logging=1)
    if args.solver == "lgmres":
        return hb.ParCSRLGMRES(tol=args.tol, max_iter=100, k_dim=5, aug_dim=2, logging=1)
    return hb.ParCSRFlexGMRES(tol=args.tol, max_iter=100, k_dim=5, logging=1)


def grid_of(args, world):
    """global grid and process grid of a run: weak scaling (one n^3 brick per GPU) unless --global-size"""
    P = PGRID[world]
    n = args.n
    if args.global_size:
        return (n, n, n), P
    return (n * P[0], n * P[1], n * P[2]), P


def workload_string(args, gn, P):
    """the ij command line this run stands for — the same text on both arms (hb200 / reference)"""
    solver = -1 if args.spmv_only else SOLVERS[args.solver][0]
    tail = f" -nmv {args.nmv} -x0rand" if args.spmv_only else " -rlx 18"
    if args.coarsen_type == 8 and not args.spmv_only:
        tail += " -pmis"
    return f"ij -{args.problem} -n {gn[0]} {gn[1]} {gn[2]} -P {P[0]} {P[1]} {P[2]} -solver {solver}{tail}"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(args, rank, world):
    from oracle import refbridge as rb
    mpi = world > 1
    rb.load(mpi=mpi)
    if world > 1:
        rb.set_num_threads(max(1, (os.cpu_count() or world) // world))
    gn, P = grid_of(args, world)
    t0 = time.time()
    pb = rb.Problem(args.problem, gn, P=P, mpi=mpi, x0rand=args.spmv_only)
    gen_s = time.time() - t0
    setup_s = 0.0 if args.spmv_only else pb.setup_amg(relax_type=18, coarsen_type=args.coarsen_type)
    return rb, pb, gn, gen_s, setup_s


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified hypre
    3.1.0 CPU build, OpenMP; for N > 1 its MPI build on oracle/minimpi with N ranks x cores/N
    threads, i.e. the "CPU MPI+OpenMP build" of BASELINE.json), same config / metric / unit.
    One complete solve first (real iteration count, full-size parity anchor); each timed step is
    a bounded sample: `cpu_iters` PCG iterations of the same solve, scaled to the full iteration
    count by the per-iteration cost."""
    rank, world, local = dist_env()
    if world > 1 and rank != 0 and not args.mpi_worker:
        return
    N = args.gpus
    if N > 1 and not args.mpi_worker:
        # rank 0 of the torchrun job launches the N-rank CPU reference on the host cores
        mpirun = os.path.join(ROOT, "oracle", "_ref", "mpirun")
        cmd = [mpirun, "-np", str(N), sys.executable, os.path.abspath(__file__), "--impl", "reference",
               "--mpi-worker", "--gpus", str(N), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--n", str(args.n), "--problem", args.problem, "--tol", str(args.tol), "--solver", args.solver,
               "--cpu-iters", str(args.cpu_iters), "--nmv", str(args.nmv), "--coarsen-type", str(args.coarsen_type)]
        cmd += ["--spmv-only"] if args.spmv_only else []
        cmd += ["--global-size"] if args.global_size else []
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            print(json.dumps({"impl": "reference", "unavailable":
                              f"N-rank CPU reference failed (rc {r.returncode}): {r.stderr[-300:]}"}), flush=True)
            return
        print(lines[-1], flush=True)
        return
    from oracle import refbridge as rb
    mpi = N > 1
    rb.load(mpi=mpi)
    ncores = os.cpu_count() or 1
    if mpi:
        rb.set_num_threads(max(1, ncores // N))
    myrank = rb.load().rb_comm_rank() if mpi else 0
    gn, P = grid_of(args, N)
    pb = rb.Problem(args.problem, gn, P=P, mpi=mpi, x0rand=args.spmv_only)
    threads = rb.num_threads()
    if args.spmv_only:
        # ij -solver -1: the loop of HYPRE_ParCSRMatrixMatvec(1, A, x, 0, b) (ij.c:4774-4777) on the host cores
        nmv = max(1, min(args.nmv, 10))
        times = [pb.matvec_time(nmv) for _ in range(args.warmup + args.steps)][args.warmup:]
        t = float(np.mean(times))
        v0 = pb.level_view(0, 0)
        nnz_g = None
        rows = pb.global_rows
        by = 12.0 * (v0.diag_nnz + v0.offd_nnz) + 20.0 * v0.num_rows          # this rank's share
        if myrank != 0:
            return
        line = {"impl": "reference", "metric": "parcsr_spmv_gb_per_s", "value": by * N / t / 1e9, "unit": "GB/s",
                "n_gpus": N, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * args.nmv,
                "higher_is_better": True, "scaling": "strong" if args.global_size else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(args, gn, P), "rows": rows, "ms_per_spmv": t * 1e3,
                           "reference_build": f"CPU, {N} rank(s) x {threads} OpenMP threads"},
                "cpu_baseline": {"value": by * N / t / 1e9, "unit": "GB/s", "cores": N * threads, "kind": "reference",
                                 "sample": f"{nmv} matvecs per step"},
                "e2e": {"value": by * N / t / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    setup_s = pb.setup_amg(relax_type=18, coarsen_type=args.coarsen_type)
    solve = lambda mi: reference_solve(pb, args, mi)
    full = solve(100)
    its_full = full["iterations"]
    k = args.cpu_iters if args.cpu_iters > 0 else 3
    times = []
    for s in range(args.warmup + args.steps):
        r = solve(k)
        if s >= args.warmup:
            times.append(r["seconds"])
    t_k = float(np.mean(times)) if times else full["seconds"] * (k + 1) / (its_full + 1)
    t_full = t_k * (its_full + 1) / (k + 1)
    rows = pb.global_rows
    val = rows / t_full / 1e6
    if myrank != 0:
        return
    line = {
        "impl": "reference", "metric": SOLVERS[args.solver][1],
        "value": val, "unit": "MDOF/s",
        "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_full * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.global_size else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args, gn, P),
                   "reference_build": f"CPU, {N} rank(s) x {threads} OpenMP threads",
                   "rows": rows, "setup_s": setup_s, "iterations": its_full,
                   "ms_per_iteration": t_full * 1e3 / max(its_full, 1),
                   "final_rel_res": full["final_rel_res"], "full_solve_s": full["seconds"]},
        "cpu_baseline": {"value": val, "unit": "MDOF/s", "cores": N * threads, "kind": "reference",
                         "sample": f"{k} PCG iterations of the same solve per step ({t_k:.3f} s), scaled by "
                                   f"({its_full}+1)/({k}+1) to the full {its_full}-iteration solve "
                                   f"(one complete solve measured: {full['seconds']:.2f} s)"},
        "e2e": {"value": val, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ij_dropin(args, gn, P):
    """The path a hypre user times: the UNMODIFIED reference driver (src/test/ij.c) linked in front of
    libHYPRE_b200.so, pageable host vectors, upload inside its Setup phase; the solve time is the one ij
    prints itself ("PCG Solve: wall clock time", ij.c:6151-6160).  One rank."""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "ij_b200")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ij_b200 not built (needs /root/reference at build time)"}
    solver = SOLVERS[args.solver][0]
    cmd = [exe, f"-{args.problem}", "-n", str(gn[0]), str(gn[1]), str(gn[2]), "-solver", str(solver), "-rlx", "18",
           "-tol", str(args.tol)]
    env = dict(os.environ, HYPRE_B200_VERBOSE="1")
    t0 = time.time()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=os.path.dirname(exe))
    except subprocess.TimeoutExpired:
        return {"unavailable": "ij_b200 timed out"}
    wall = time.time() - t0
    out = r.stdout
    sec = {}
    for name in ("Setup", "Solve"):
        m = re.search(r"(?:PCG|GMRES) %s:\s*\n\s*wall clock time = ([0-9.eE+-]+) seconds" % name, out)
        if m:
            sec[name] = float(m.group(1))
    its = re.findall(r"Iterations = (\d+)", out)
    res = re.findall(r"Final (?:GMRES )?Relative Residual Norm = ([0-9.eE+-]+)", out)
    dev = re.findall(r"on device: (\d+) its, ([0-9.]+) ms", r.stderr)
    if r.returncode != 0 or "Solve" not in sec or not dev:
        return {"unavailable": f"ij_b200 rc {r.returncode}: {(r.stderr or out)[-200:]}"}
    rows = gn[0] * gn[1] * gn[2]
    return {"command": "ij_b200 " + " ".join(cmd[1:]), "solve_s_printed_by_ij": sec["Solve"],
            "setup_s_printed_by_ij": sec.get("Setup"), "value": rows / sec["Solve"] / 1e6, "unit": "MDOF/s",
            "iterations": int(its[-1]) if its else None, "final_rel_res": float(res[-1]) if res else None,
            "device_solve_ms": float(dev[-1][1]), "process_wall_s": wall,
            "note": "host vectors are pageable (hypre's own allocations); the hierarchy upload happens inside "
                    "HYPRE_PCGSetup (hypre_shim.c), not inside the solve ij times"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.format == "csr":
        # the general-matrix path: no structured re-encoding of any block (read at upload time)
        os.environ["HB200_NO_PAT"] = "1"
        os.environ["HB200_NO_SELL"] = "1"
        os.environ["HB200_NO_CSR16"] = "1"
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hb200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    import hypre_b200 as hb
    from hypre_b200._lib import lib, check
    hb.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = [hb.comm_get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        hb.comm_init(rank, world, uid[0])
        # auto: NVLink peer puts (the whole V-cycle stays one CUDA graph) when every rank can map every
        # peer, else NCCL send/recv; --halo nccl / peer force one
        mode = {"nccl": 0, "peer": 1, "auto": 2}[args.halo]
        check(lib.hb200_set_halo_mode(mode))

    # a stage that stops making progress (a rank lost in a collective) must not hold the box:
    # every stage re-arms the watchdog; on expiry the rank says where it was and leaves
    import threading
    wd = {"stage": "init", "deadline": time.time() + args.stage_timeout}

    def stage(name, factor=1.0):
        wd["stage"] = name
        wd["deadline"] = time.time() + factor * args.stage_timeout
        if os.environ.get("HB200_TRACE"):
            print(f"[bench rank {rank}] stage: {name}", file=sys.stderr, flush=True)

    def watchdog():
        while True:
            time.sleep(5.0)
            if time.time() > wd["deadline"]:
                print(f"bench.py: rank {rank} made no progress in stage '{wd['stage']}' for "
                      f"{args.stage_timeout:.0f} s, giving up", file=sys.stderr, flush=True)
                os._exit(3)

    threading.Thread(target=watchdog, daemon=True).start()

    def barrier():
        torch.cuda.synchronize()
        hb.sync()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = peaks()
    NOMINAL = 8000.0
    stream_holder = {}

    def stream():
        if "s" not in stream_holder:
            stream_holder["s"] = torch.cuda.ExternalStream(lib.hb200_compute_stream())
        return stream_holder["s"]

    def time_spmv(M, reps=20, xs=None, ys=None):
        if xs is None:
            xs = torch.randn(max(M.num_cols, 1), dtype=torch.float64, device="cuda")
        if ys is None:
            ys = torch.empty(max(M.num_rows, 1), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        for _ in range(3):
            check(lib.hb200_parcsr_matvec(M.handle, 1.0, xs.data_ptr(), 0.0, ys.data_ptr(), ys.data_ptr()))
        hb.sync()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream())
        for _ in range(reps):
            check(lib.hb200_parcsr_matvec(M.handle, 1.0, xs.data_ptr(), 0.0, ys.data_ptr(), ys.data_ptr()))
        s1.record(stream())
        hb.sync()
        return s0.elapsed_time(s1) / reps

    def csr_model(M):
        # 12 B per nonzero + row pointer + x + y, diag block (the offd block adds
        # 12 B per nonzero and the halo 16 B per sent entry: < 1 % at these sizes)
        return 12.0 * M.diag_nnz + 4.0 * (M.num_rows + 1) + 8.0 * M.num_cols + 8.0 * M.num_rows

    def stored_model(M):
        """algorithmic bytes of one y = A x on the diag block in the format it is stored in"""
        fi = M.format_info()
        n, nnz = M.num_rows, M.diag_nnz
        if fi["kernel"] in (7, 9):
            by = 1.0 * n + 8.0 * M.num_cols + 8.0 * n + 12.0 * fi["pattern_entries"] \
                + 12.0 * fi["pattern_irregular_nnz"] + 8.0 * fi["pattern_irregular_rows"]
            if M.num_rows != M.num_cols:
                by += 4.0 * n                       # first column of every row (rectangular blocks: P, P^T)
            irr = fi["pattern_irregular_rows"]
            geo = fi["kernel"] == 9 and fi.get("box_geo") and os.environ.get("HB200_BOX_NO_GEO", "0") in ("", "0")
            if geo:
                by -= 1.0 * n                       # the grid-box kernel reads no row codes: x + y only
            name = ("spmv_box<EPI_AXPBY,GEO> (row-pattern format, compact-stencil kernel on a grid box, x + y only" if geo
                    else "spmv_box<EPI_AXPBY> (row-pattern format, compact-stencil kernel, 1 B/row + x + y" if fi["kernel"] == 9
                    else "spmv_pat<EPI_AXPBY> (row-pattern format, 1 B/row + x + y") \
                + (f"; {irr} irregular rows in CSR)" if irr else ")")
        elif fi["kernel"] == 6:
            by = float(fi["sell_entries"]) * fi["sell_bytes_per_entry"] + 8.0 * (n / 32.0) + 4.0 * n \
                + 8.0 * M.num_cols + 8.0 * n
            name = f"spmv_sell<EPI_AXPBY> (packed SELL-32, {fi['sell_bytes_per_entry']} B/nonzero)"
        elif fi["kernel"] == 8:
            by = csr_model(M) - 2.0 * nnz
            name = "spmv_vector<EPI_AXPBY,K,I16> (CSR, 16-bit column offsets, 10 B/nonzero)"
        else:
            by = csr_model(M)
            name = "spmv_vector<EPI_AXPBY,K> (CSR, 12 B/nonzero)"
        return by, name

    def kernel_entry(level, M, passes):
        """one y = A x over the launch time measured here; on N > 1 the time includes the halo and the offd pass"""
        n, nnz = M.num_rows, M.diag_nnz
        by, name = stored_model(M)
        ms_k = time_spmv(M)
        gbs = by / (ms_k * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": name.replace(" (", f" on A_{level} (", 1), "level": level, "rows": n, "nnz": nnz,
                "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
                "ms_per_launch": ms_k, "bytes_per_launch": by, "bytes_per_nnz": by / max(nnz, 1),
                "csr_equivalent_gbs": csr_model(M) / (ms_k * 1e-3) / 1e9,
                "launches_per_iteration": passes, "ms_per_iteration": passes * ms_k, "traffic": None}

    gn, P = grid_of(args, world)

    # =====================================================================================
    # --spmv-only: BASELINE.json configs[4], the loop of HYPRE_ParCSRMatrixMatvec(1, A, x, 0, b)
    # =====================================================================================
    if args.spmv_only:
        stage("matrix generation (CPU)", 2.0)
        rb, pb, gn, gen_s, _ = build_problem(args, rank, world)
        stage("upload")
        t0 = time.time()
        A = hb.ParCSRMatrix.from_view(pb.level_view(0, 0))
        hb.sync()
        upload_s = time.time() - t0
        nloc, rows = A.num_rows, pb.global_rows
        xh = torch.from_numpy(np.array(pb.x0)).pin_memory()      # seeded random x (ij -x0rand, par_vector.c:441-455)
        x = xh.cuda()
        y = torch.empty(max(nloc, 1), dtype=torch.float64, device="cuda")
        yh = torch.empty(max(nloc, 1), dtype=torch.float64).pin_memory()
        stage("warm-up")
        for _ in range(max(args.warmup, 3)):
            check(lib.hb200_parcsr_matvec(A.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
        # parity at full size: the reference's own matvec on the same x (element-wise, 1e-12)
        yref = pb.matvec(1.0, np.array(pb.x0), 0.0)
        hb.sync()
        den = max_over_ranks(float(np.max(np.abs(yref))) if yref.size else 0.0)
        err = max_over_ranks(float(np.max(np.abs(y[:nloc].cpu().numpy() - yref))) / den if yref.size else 0.0)
        stage("timed loop")
        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = hb.launch_count()
        e0.record(stream())
        for _ in range(args.steps):
            for _ in range(args.nmv):
                check(lib.hb200_parcsr_matvec(A.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
        e1.record(stream())
        barrier()
        launches = hb.launch_count() - l0
        ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        ms_mv = ms_step / args.nmv
        # e2e: x from pinned host memory in, y back out, every step (nmv matvecs per step as ij runs them)
        stage("e2e loop")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            x.copy_(xh, non_blocking=True)
            torch.cuda.synchronize()
            for _ in range(args.nmv):
                check(lib.hb200_parcsr_matvec(A.handle, 1.0, x.data_ptr(), 0.0, y.data_ptr(), y.data_ptr()))
            hb.sync()
            yh.copy_(y, non_blocking=False)
        barrier()
        e2e_ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        clocks = sampler.stop() if sampler else None
        by_csr_g = sum_over_ranks(csr_model(A) + 12.0 * A.offd_nnz)
        by_st, st_name = stored_model(A)
        by_st_g = sum_over_ranks(by_st + 12.0 * A.offd_nnz)
        value = by_csr_g / (ms_mv * 1e-3) / 1e9
        stage("kernel kinds")
        ms_stored = time_spmv(A, xs=x, ys=y)
        kinds = {"stored": {"kernel": st_name, "ms": ms_stored, "bytes": by_st,
                            "achieved_gbs": by_st / (ms_stored * 1e-3) / 1e9,
                            "csr_equivalent_gbs": csr_model(A) / (ms_stored * 1e-3) / 1e9}}
        if A.format_info()["kernel"] != 1:
            A.set_spmv_kernel(1, 0)
            ms_csr = time_spmv(A, xs=x, ys=y)
            A.set_spmv_kernel(0, 0)
        else:
            ms_csr = ms_stored
        kinds["csr"] = {"kernel": "spmv_vector<EPI_AXPBY,K> (general CSR, 12 B/nonzero)", "ms": ms_csr,
                        "bytes": csr_model(A), "achieved_gbs": csr_model(A) / (ms_csr * 1e-3) / 1e9}
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            stage("cpu baseline")
            t_ref = pb.matvec_time(5)
            cpu = {"value": csr_model(A) / t_ref / 1e9, "unit": "GB/s", "cores": rb.num_threads(), "kind": "reference",
                   "sample": f"5 HYPRE_ParCSRMatrixMatvec calls of the reference on the host cores ({t_ref * 1e3:.1f} ms each)"}
        if rank == 0:
            csr_gbs = kinds["csr"]["achieved_gbs"]
            line = {
                "metric": "parcsr_spmv_gb_per_s", "value": value, "unit": "GB/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if args.global_size else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_string(args, gn, P), "rows": rows, "rows_per_gpu": nloc,
                           "nnz_A0_per_gpu": A.num_nonzeros, "ms_per_spmv": ms_mv, "spmv_per_step": args.nmv,
                           "value_is": "CSR-model bytes of SURVEY section 8(d) (12.75 B/nnz, all ranks) / time of one "
                                       "ParCSR matvec in the format the upload chose; `roofline` is the general CSR kernel",
                           "stored_format_gbs": by_st_g / (ms_mv * 1e-3) / 1e9,
                           "parity_vs_reference_max_rel_err": err, "format": args.format,
                           "halo": (["nccl", "peer"][lib.hb200_halo_mode()] if world > 1 else None),
                           "cache": "inputs larger than L2 (A alone is %.1f GB)" % (12.0 * A.num_nonzeros / 1e9),
                           "generate_s": gen_s, "upload_s": upload_s, "kernel_kinds": kinds,
                           "timing": "CUDA events on the hb200 compute stream, max over ranks"},
                "e2e": {"value": by_csr_g / (e2e_ms_step / args.nmv * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_ms_step,
                        "h2d_bytes_per_step": 8 * A.num_cols, "d2h_bytes_per_step": 8 * nloc,
                        "api": "hb200_parcsr_matvec (the call behind HYPRE_ParCSRMatrixMatvec), x in / y out per step"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": kinds["csr"]["kernel"] + " on A_0", "achieved": csr_gbs,
                             "peak": peak, "unit": "GB/s", "frac": csr_gbs / peak, "peak_source": peak_src,
                             "frac_of_nominal_8TBs": csr_gbs / NOMINAL, "ms_per_launch": ms_csr,
                             "bytes_per_launch": csr_model(A), "bytes_per_nnz": csr_model(A) / max(A.diag_nnz, 1),
                             "traffic": None, "stored_format": kinds["stored"]},
                "cpu_baseline": cpu, "clocks": clocks,
            }
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- hierarchy from the reference's own setup (CPU), uploaded once: timed separately
    stage("reference setup (CPU)", 2.0)      # N ranks share the host cores: the one stage that scales with N
    rb, pb, gn, gen_s, setup_s = build_problem(args, rank, world)
    stage("upload")
    t0 = time.time()
    hier = pb.hierarchy()
    if rank == 0 and os.environ.get("HB200_BENCH_LEVELS"):
        for l, lv in enumerate(hier["levels"]):
            for nm in ("A", "P"):
                M = lv[nm]
                if M is None:
                    continue
                oi = M.arrays()["offd_i"]
                nbr = int((np.diff(oi) > 0).sum()) if oi is not None else 0
                print(f"[levels] L{l} {nm}: rows {M.num_rows} diag_nnz {M.diag_nnz} offd_nnz {M.offd_nnz} "
                      f"offd_rows {nbr} cols_offd {M.num_cols_offd} sends {M.num_sends} recvs {M.num_recvs}",
                      file=sys.stderr)
    mats, amg = hb.amg_from_hierarchy(hier, use_graph=not args.no_graph)
    hb.sync()
    upload_s = time.time() - t0
    A = mats[0][0]
    nloc = A.num_rows
    rows = pb.global_rows
    nnz0 = A.num_nonzeros
    b_host = torch.from_numpy(np.array(pb.b)).pin_memory()
    x_host = torch.zeros(nloc, dtype=torch.float64).pin_memory()
    b = b_host.cuda()
    x = torch.zeros(nloc, dtype=torch.float64, device="cuda")

    solver = device_solver(hb, args)
    solver.set_precond(amg)

    def step_dev():
        check(lib.hb200_vec_set(x.data_ptr(), 0.0, nloc))
        return solver.solve(A, b, x)

    def step_host():
        x_host.zero_()
        return solver.solve(A, b_host.numpy(), x_host.numpy())

    stage("warm-up solves")
    for _ in range(args.warmup):
        res = step_dev()
    its = res.num_iterations if args.warmup else None
    stage("timed solves")

    # ---- timed region: K solves, device-resident inputs
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream())
    for _ in range(args.steps):
        res = step_dev()
        launches += int(res.kernel_launches)
    e1.record(stream())
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    its = res.num_iterations
    relres = res.rel_residual_norm
    value = rows / (ms * 1e-3) / 1e6

    # ---- e2e: same solve through the host-buffer entry point (H2D b, x0; D2H x inside)
    stage("e2e solves")
    for _ in range(min(2, args.warmup)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- roofline: every level's SpMV kernel is timed alone with CUDA events on the library's
    # compute stream; `roofline` is the one with the largest share of an iteration.  A_0 of a
    # constant-coefficient stencil runs in the row-pattern format (1 B per row), other structured
    # operators in packed SELL (2 or 9 B per nonzero), everything else (every coarse level, every
    # unstructured matrix) in CSR through spmv_vector; the general CSR kernel is also timed on A_0
    # (`roofline.csr`).
    stage("per-level kernel timing")
    # one PCG iteration launches the A_0 kernel 3 times (Krylov matvec, residual, post-smoothing; the
    # pre-smoothing sweep starts from a zero guess and reads no matrix) and the A_l kernel, l >= 1, twice
    # (the matvec is collective on N > 1: the levels timed are fixed, not chosen from rank-local sizes)
    per_level = []
    for l, (Al, _) in enumerate(mats[:min(4, len(mats))]):
        per_level.append(kernel_entry(l, Al, 3 if l == 0 else 2))
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    # of the same workload (profiles/); a number measured under the profiler is not taken live here
    traffic_tab = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(traffic_tab) and args.problem == "27pt" and args.n == 256 and world == 1 and per_level:
        with open(traffic_tab) as f:
            tab = json.load(f)
        for key, ent in tab.items():
            if per_level[0]["kernel"].startswith(key) and not A.format_info()["pattern_irregular_rows"]:
                per_level[0]["traffic"] = ent["bytes"]
                per_level[0]["traffic_source"] = ent["source"]
        if per_level[0]["traffic"] is None and per_level[0]["kernel"].startswith("spmv_box<EPI_AXPBY,GEO>"):
            # no capture of this variant yet (written after the round's last GPU session): not a measured number
            ent = tab.get("spmv_box<EPI_AXPBY> on A_0")
            if ent:
                per_level[0]["traffic_note"] = (f"no ncu capture of the GEO variant yet; the predicated variant it replaces moved "
                                                f"{ent['bytes'] / 1e6:.0f} MB per launch ({ent['source']}), "
                                                f"{per_level[0]['rows'] / 1e6:.1f} MB of it row codes, which this one does not read")
    roofline = max(per_level, key=lambda e: e["ms_per_iteration"])
    roofline = dict(roofline, note="the level kernel with the largest share of the iteration; every level in `levels`")
    roofline["frac_of_nominal_8TBs"] = roofline["achieved"] / NOMINAL
    roofline["levels"] = per_level
    if A.format_info()["kernel"] != 1:
        A.set_spmv_kernel(1, 0)
        ms_csr = time_spmv(A)
        A.set_spmv_kernel(0, 0)
    else:
        ms_csr = per_level[0]["ms_per_launch"]
    cb = csr_model(A)
    roofline["csr"] = {"kernel": "spmv_vector<EPI_AXPBY,K> on A_0 (the general CSR kernel, 12 B/nonzero)",
                       "achieved": cb / (ms_csr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": cb / (ms_csr * 1e-3) / 1e9 / peak,
                       "frac_of_nominal_8TBs": cb / (ms_csr * 1e-3) / 1e9 / NOMINAL, "ms_per_launch": ms_csr,
                       "bytes_per_launch": cb, "bytes_per_nnz": cb / max(nnz0, 1)}
    # whole-solve traffic: bytes of every matrix pass of one iteration in its STORED format and in the CSR
    # model (3 x A_0, 2 x A_l, 2 x P_l per V(1,1)-PCG iteration) + ~90 B/row of vector traffic on the fine
    # level, over the measured time of an iteration
    stored_b, csr_b = 0.0, 0.0
    for l, (Al, Pl) in enumerate(mats):
        passes = 3 if l == 0 else 2
        stored_b += passes * stored_model(Al)[0]
        csr_b += passes * csr_model(Al)
        if Pl is not None:
            pb_st = stored_model(Pl)[0]
            stored_b += 2 * pb_st
            csr_b += 2 * csr_model(Pl)
    vec_b = 90.0 * nloc
    ms_it = ms / max(its, 1)
    roofline["whole_iteration"] = {
        "ms": ms_it, "bytes_stored_formats": stored_b + vec_b, "bytes_csr_model": csr_b + vec_b,
        "gbs_stored_formats": (stored_b + vec_b) / (ms_it * 1e-3) / 1e9,
        "gbs_csr_model": (csr_b + vec_b) / (ms_it * 1e-3) / 1e9,
        "frac_stored_formats": (stored_b + vec_b) / (ms_it * 1e-3) / 1e9 / peak,
        "note": "per rank; matrix passes of one PCG iteration (3 x A_0, 2 x A_l, 2 x P_l) + 90 B/row of vectors"}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own solve, bounded sample
    stage("cpu baseline")
    cpu = None
    ref_parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_solve = lambda mi: reference_solve(pb, args, mi)
        if args.cpu_iters <= 0 or args.n <= 256:
            # the whole reference solve on the host cores (~10-20 s at 256^3): also the full-size
            # parity evidence (iteration count, final residual)
            r = ref_solve(100)
            t_full = r["seconds"]
            sample = (f"the complete reference solve ({r['iterations']} iterations, {t_full:.2f} s) "
                      "on the host cores, OpenMP")
            ref_parity = {"reference_iterations": r["iterations"], "reference_final_rel_res": r["final_rel_res"],
                          "hb200_iterations": its, "hb200_final_rel_res": relres}
        else:
            k = args.cpu_iters
            r = ref_solve(k)
            t_full = r["seconds"] * (its + 1) / (k + 1)
            sample = (f"{k} iterations of the same solve on the host cores ({r['seconds']:.3f} s), "
                      f"scaled by ({its}+1)/({k}+1) to the full {its}-iteration solve")
        cpu = {"value": rows / t_full / 1e6, "unit": "MDOF/s", "cores": rb.num_threads(),
               "kind": "reference", "sample": sample, "seconds": t_full}

    # ---- the unmodified ij driver in front of the shim (rank 0, N = 1): what a hypre user measures
    e2e_ij = None
    if rank == 0 and world == 1 and not args.no_e2e_ij and args.format == "auto":
        stage("ij drop-in")
        e2e_ij = run_ij_dropin(args, gn, P)

    if rank == 0:
        line = {
            "metric": SOLVERS[args.solver][1],
            "value": value, "unit": "MDOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong" if args.global_size else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_string(args, gn, P),
                "solver": SOLVERS[args.solver][2]
                          + ", HMIS + ext+i, l1-Jacobi V(1,1); hierarchy from the reference's BoomerAMGSetup, "
                            "uploaded once (not timed)",
                "rows": rows, "rows_per_gpu": nloc, "nnz_A0_per_gpu": nnz0, "levels": pb.num_levels,
                "iterations": its, "ms_per_iteration": ms / max(its, 1),
                "iterations_note": "weak scaling: the reference's own hierarchy needs more iterations on the larger "
                                   "global problem (25 at 256^3, 32 at 512^3 for 27pt), which caps the efficiency of "
                                   "this metric at their ratio; ms_per_iteration separates the two effects",
                "final_rel_res": relres, "tol": args.tol, "parity_vs_reference": ref_parity, "format": args.format,
                "cache": "inputs larger than L2 (A_0 alone is %.1f GB)" % (12.0 * nnz0 / 1e9),
                "setup_s_reference_cpu": setup_s, "generate_s": gen_s, "upload_s": upload_s,
                "cuda_graph_vcycle": (not args.no_graph) and (world == 1 or lib.hb200_halo_mode() == 1
                                                              or bool(os.environ.get("HB200_GRAPH_NCCL"))),
                "halo": (["nccl", "peer"][lib.hb200_halo_mode()] if world > 1 else None),
                "timing": "CUDA events on the hb200 compute stream, max over ranks",
            },
            "e2e": {"value": rows / (e2e_ms * 1e-3) / 1e6, "unit": "MDOF/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 16 * nloc, "d2h_bytes_per_step": 8 * nloc,
                    "api": "hb200_pcg_solve_host (host b, x; the call behind HYPRE_PCGSolve)",
                    "ij_dropin": e2e_ij},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
