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
    ap.add_argument("--solver", default="pcg", choices=["pcg", "gmres"])
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl", "peer"],
                    help="N>1 halo transport: NVLink peer puts, NCCL send/recv, or peer when every rank can map every peer")
    ap.add_argument("--cpu-iters", type=int, default=3, help="iterations of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-timeout", type=float, default=420.0,
                    help="seconds a stage of the hb200 arm may take before the rank gives up")
    ap.add_argument("--spmv-only", action="store_true", help="configs[4]: SpMV bandwidth line")
    ap.add_argument("--mpi-worker", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


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
    P = PGRID[world]
    n = args.n
    gn = (n * P[0], n * P[1], n * P[2])
    t0 = time.time()
    pb = rb.Problem(args.problem, gn, P=P, mpi=mpi)
    gen_s = time.time() - t0
    setup_s = pb.setup_amg(relax_type=18)
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
               "--n", str(args.n), "--problem", args.problem, "--tol", str(args.tol),
               "--cpu-iters", str(args.cpu_iters)]
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
    n = args.n
    P = PGRID[N]
    gn = (n * P[0], n * P[1], n * P[2])
    pb = rb.Problem(args.problem, gn, P=P, mpi=mpi)
    setup_s = pb.setup_amg(relax_type=18)
    threads = rb.num_threads()
    full = pb.pcg(precond="amg", tol=args.tol, max_iter=100, two_norm=1)
    its_full = full["iterations"]
    k = args.cpu_iters if args.cpu_iters > 0 else 3
    times = []
    for s in range(args.warmup + args.steps):
        r = pb.pcg(precond="amg", tol=args.tol, max_iter=k, two_norm=1)
        if s >= args.warmup:
            times.append(r["seconds"])
    t_k = float(np.mean(times)) if times else full["seconds"] * (k + 1) / (its_full + 1)
    t_full = t_k * (its_full + 1) / (k + 1)
    rows = pb.global_rows
    val = rows / t_full / 1e6
    if myrank != 0:
        return
    line = {
        "impl": "reference", "metric": "amg_pcg_solve_mdof_per_s", "value": val, "unit": "MDOF/s",
        "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"ij -{args.problem} -n {gn[0]} {gn[1]} {gn[2]} -P {P[0]} {P[1]} {P[2]} -solver 1 "
                               "-rlx 18 (BoomerAMG-PCG, HMIS + ext+i, l1-Jacobi V(1,1)), reference CPU build "
                               f"({N} rank(s) x {threads} OpenMP threads)",
                   "rows": rows, "setup_s": setup_s, "iterations": its_full,
                   "final_rel_res": full["final_rel_res"], "full_solve_s": full["seconds"]},
        "cpu_baseline": {"value": val, "unit": "MDOF/s", "cores": N * threads, "kind": "reference",
                         "sample": f"{k} PCG iterations of the same solve per step ({t_k:.3f} s), scaled by "
                                   f"({its_full}+1)/({k}+1) to the full {its_full}-iteration solve "
                                   f"(one complete solve measured: {full['seconds']:.2f} s)"},
        "e2e": {"value": val, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
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
        # auto: the NVLink peer-put halo (whole V-cycle in one CUDA graph) where this round measured
        # it on hardware (N = 2: 105.5 ms against 111.9 ms over NCCL on the same library state,
        # profiles/r1_multi_gpu.md); the NCCL send/recv halo, the configuration measured at N = 8,
        # for every other N.  --halo peer forces the former.
        mode = {"nccl": 0, "peer": 1, "auto": 2 if world == 2 else 0}[args.halo]
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
    stream = torch.cuda.ExternalStream(lib.hb200_compute_stream())

    if args.solver == "pcg":
        solver = hb.ParCSRPCG(tol=args.tol, max_iter=100, two_norm=1, logging=1)
    else:
        solver = hb.ParCSRGMRES(tol=args.tol, max_iter=100, k_dim=5, logging=1)
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
    e0.record(stream)
    for _ in range(args.steps):
        res = step_dev()
        launches += int(res.kernel_launches)
    e1.record(stream)
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
    # as `roofline_csr`.
    stage("per-level kernel timing")
    peak, peak_src = peaks()

    def time_spmv(M, reps=20):
        xs = torch.randn(max(M.num_cols, 1), dtype=torch.float64, device="cuda")
        ys = torch.empty(max(M.num_rows, 1), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        for _ in range(3):
            check(lib.hb200_parcsr_matvec(M.handle, 1.0, xs.data_ptr(), 0.0, ys.data_ptr(), ys.data_ptr()))
        hb.sync()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(reps):
            check(lib.hb200_parcsr_matvec(M.handle, 1.0, xs.data_ptr(), 0.0, ys.data_ptr(), ys.data_ptr()))
        s1.record(stream)
        hb.sync()
        return s0.elapsed_time(s1) / reps

    def csr_model(M):
        return 12.0 * M.diag_nnz + 4.0 * (M.num_rows + 1) + 8.0 * M.num_cols + 8.0 * M.num_rows

    def kernel_entry(level, M, passes):
        """algorithmic bytes of one y = A x on the diag block in its stored format
        over the launch time measured here; on N > 1 the time includes the halo and the offd pass"""
        fi = M.format_info()
        n, nnz = M.num_rows, M.diag_nnz
        if fi["kernel"] == 7:
            by = 1.0 * n + 8.0 * M.num_cols + 8.0 * n + 12.0 * fi["pattern_entries"] \
                + 12.0 * fi["pattern_irregular_nnz"] + 8.0 * fi["pattern_irregular_rows"]
            irr = fi["pattern_irregular_rows"]
            name = f"spmv_pat<EPI_AXPBY> on A_{level} (row-pattern format, 1 B/row + x + y" \
                + (f"; {irr} irregular rows in CSR)" if irr else ")")
        elif fi["kernel"] == 6:
            by = float(fi["sell_entries"]) * fi["sell_bytes_per_entry"] + 8.0 * (n / 32.0) + 4.0 * n \
                + 8.0 * M.num_cols + 8.0 * n
            name = f"spmv_sell<EPI_AXPBY> on A_{level} (packed SELL-32, {fi['sell_bytes_per_entry']} B/nonzero)"
        elif fi["kernel"] == 8:
            by = csr_model(M) - 2.0 * nnz
            name = f"spmv_vector<EPI_AXPBY,K,I16> on A_{level} (CSR, 16-bit column offsets, 10 B/nonzero)"
        else:
            by = csr_model(M)
            name = f"spmv_vector<EPI_AXPBY,K> on A_{level} (CSR, 12 B/nonzero)"
        ms_k = time_spmv(M)
        gbs = by / (ms_k * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": name, "level": level, "rows": n, "nnz": nnz,
                "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
                "ms_per_launch": ms_k, "bytes_per_launch": by, "bytes_per_nnz": by / max(nnz, 1),
                "csr_equivalent_gbs": csr_model(M) / (ms_k * 1e-3) / 1e9,
                "launches_per_iteration": passes, "ms_per_iteration": passes * ms_k, "traffic": None}

    # one PCG iteration launches the A_0 kernel 3 times (Krylov matvec, residual, post-smoothing; the
    # pre-smoothing sweep starts from a zero guess and reads no matrix) and the A_l kernel, l >= 1, twice
    # (the matvec is collective on N > 1: the levels timed are fixed, not chosen from rank-local sizes)
    per_level = []
    for l, (Al, _) in enumerate(mats[:min(4, len(mats))]):
        per_level.append(kernel_entry(l, Al, 3 if l == 0 else 2))
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    # of the same workload (profiles/r1_ncu_spmv_kernels.md); a number measured under the profiler is
    # not taken live here
    if args.problem == "27pt" and args.n == 256 and world == 1:
        ncu_traffic = {7: 285242880.0 + 97056768.0, 6: 1140867000.0 + 131572224.0, 1: 5595142000.0 + 138026752.0}
        fi0 = A.format_info()
        if per_level and fi0["kernel"] in ncu_traffic and not fi0["pattern_irregular_rows"]:
            per_level[0]["traffic"] = ncu_traffic[fi0["kernel"]]
            per_level[0]["traffic_source"] = "ncu --set full, profiles/r1_ncu_spmv_kernels.md (capture 2)"
    roofline = max(per_level, key=lambda e: e["ms_per_iteration"])
    roofline = dict(roofline, note="the level kernel with the largest share of the iteration; all levels in roofline_levels")
    roofline_csr = None
    if A.format_info()["kernel"] != 1:
        A.set_spmv_kernel(1, 0)
        ms_csr = time_spmv(A)
        A.set_spmv_kernel(0, 0)
        cb = csr_model(A)
        roofline_csr = {"bound": "hbm", "kernel": "spmv_vector<EPI_AXPBY,K> on A_0 (general CSR path)",
                        "achieved": cb / (ms_csr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": cb / (ms_csr * 1e-3) / 1e9 / peak, "ms_per_launch": ms_csr,
                        "bytes_per_launch": cb, "bytes_per_nnz": cb / max(nnz0, 1), "traffic": None}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own solve, bounded sample
    stage("cpu baseline")
    cpu = None
    ref_parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.cpu_iters <= 0 or args.n <= 256:
            # the whole reference solve on the host cores (~10-20 s at 256^3): also the full-size
            # parity evidence (iteration count, final residual)
            r = pb.pcg(precond="amg", tol=args.tol, max_iter=100, two_norm=1) if args.solver == "pcg" \
                else pb.gmres(precond="amg", tol=args.tol, max_iter=100, k_dim=5)
            t_full = r["seconds"]
            sample = (f"the complete reference solve ({r['iterations']} iterations, {t_full:.2f} s) "
                      "on the host cores, OpenMP")
            ref_parity = {"reference_iterations": r["iterations"], "reference_final_rel_res": r["final_rel_res"],
                          "hb200_iterations": its, "hb200_final_rel_res": relres}
        else:
            k = args.cpu_iters
            r = pb.pcg(precond="amg", tol=args.tol, max_iter=k, two_norm=1)
            t_full = r["seconds"] * (its + 1) / (k + 1)
            sample = (f"{k} PCG iterations of the same solve on the host cores ({r['seconds']:.3f} s), "
                      f"scaled by ({its}+1)/({k}+1) to the full {its}-iteration solve")
        cpu = {"value": rows / t_full / 1e6, "unit": "MDOF/s", "cores": rb.num_threads(),
               "kind": "reference", "sample": sample, "seconds": t_full}

    if rank == 0:
        n = args.n
        line = {
            "metric": "amg_pcg_solve_mdof_per_s", "value": value, "unit": "MDOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"ij -{args.problem} -n {gn[0]} {gn[1]} {gn[2]} -P {PGRID[world][0]} "
                            f"{PGRID[world][1]} {PGRID[world][2]} -solver {1 if args.solver == 'pcg' else 3} "
                            "-rlx 18 (BoomerAMG-PCG, HMIS + ext+i, l1-Jacobi V(1,1)); hierarchy from the "
                            "reference's BoomerAMGSetup, uploaded once (not timed)",
                "rows": rows, "rows_per_gpu": nloc, "nnz_A0_per_gpu": nnz0, "levels": pb.num_levels,
                "iterations": its, "final_rel_res": relres, "tol": args.tol, "parity_vs_reference": ref_parity,
                "cache": "inputs larger than L2 (A_0 alone is %.1f GB)" % (12.0 * nnz0 / 1e9),
                "setup_s_reference_cpu": setup_s, "generate_s": gen_s, "upload_s": upload_s,
                "cuda_graph_vcycle": (not args.no_graph) and (world == 1 or lib.hb200_halo_mode() == 1),
                "halo": (["nccl", "peer"][lib.hb200_halo_mode()] if world > 1 else None),
                "timing": "CUDA events on the hb200 compute stream, max over ranks",
            },
            "e2e": {"value": rows / (e2e_ms * 1e-3) / 1e6, "unit": "MDOF/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 16 * nloc, "d2h_bytes_per_step": 8 * nloc,
                    "api": "hb200_pcg_solve_host (host b, x; the call behind HYPRE_PCGSolve)"},
            "gpu_launches": launches,
            "roofline": roofline,
            "roofline_csr": roofline_csr,
            "roofline_levels": per_level,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
