"""Test infrastructure: the reference's own regression jobs for this path, run through the drop-in.

Each case is one line of src/test/TEST_ij/smoother.jobs or solvers.jobs (the `ij` command line and
the rank count) with the iteration count and final relative residual the reference keeps for it in
smoother.saved / solvers.saved (compared by the reference's own runtest.sh at the printed precision).
The UNMODIFIED ij driver linked in front of hypre_shim.c runs the job; the shim hands the solve to
libhb200 (on a GPU: oracle/_ref/ij_b200_mpi; in the CPU suite: the host emulation of the kernels,
oracle/_ref/ij_b200_emu_mpi).  `device` says whether the job's solver / smoother is on the B200 path
(the shim must then report "on device") or outside it (the shim must pass it to the reference).

    python tests/ref_golden_jobs.py [emu|gpu|ref] [case ids ...]     # prints one line per case
"""
import os
import re
import signal
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BIN = {"emu": "ij_b200_emu_mpi", "gpu": "ij_b200_mpi", "ref": "ij_refmpi"}

# id, ranks, ij arguments, iterations, final relative residual, on the B200 path
# smoother.jobs / smoother.saved: stand-alone BoomerAMG jobs get `-pout 1` (ij's default print level 3
# asks for per-cycle output, which the shim leaves to the reference; the numbers do not depend on it)
CASES = [
    ("smoother.4", 4, "-rhsrand -solver 1 -rlx 6 -n 20 20 10 -P 2 2 1 -w -10", 6, 5.846604e-09, True),
    ("smoother.8", 3, "-rhsrand -n 15 30 10 -rlx 0 -CF 1 -pout 1", 11, 7.457693e-09, True),
    ("smoother.9", 3, "-rhsrand -n 15 30 10 -rlx 18 -CF 1 -pout 1", 17, 4.979125e-09, True),
    ("smoother.10", 3, "-rhsrand -n 15 30 10 -rlx 18 -pout 1", 23, 8.254191e-09, True),
    ("smoother.11", 4, "-rhsrand -solver 1 -rlx 8 -n 20 20 10 -P 2 2 1", 6, 2.509163e-09, True),
    ("smoother.11.1", 4, "-rhsrand -solver 1 -rlx 88 -n 20 20 10 -P 2 2 1", 6, 2.881634e-09, True),
    ("smoother.11.2", 4, "-rhsrand -solver 1 -rlx 89 -n 20 20 10 -P 2 2 1", 6, 3.877871e-10, True),
    ("smoother.12", 4, "-rhsrand -solver 1 -rlx 16 -n 20 20 10 -P 2 2 1", 6, 2.510130e-09, True),
    ("smoother.13", 4, "-rhsrand -solver 1 -rlx 16 -cheby_order 3 -n 20 20 10 -P 2 2 1", 5, 6.702200e-09, True),
    ("smoother.14", 4, "-rhsrand -solver 1 -rlx 17 -n 20 20 10 -P 2 2 1", 6, 5.044385e-10, False),
    # rlx 15 (CG smoother): the hierarchy is outside the path, so the cycle stays in the reference and is
    # called back as the preconditioner of the device PCG; the smoother's inner PCG solves run on device too
    ("smoother.15", 4, "-rhsrand -solver 1 -rlx 15 -n 20 20 10 -P 2 2 1", 15, 5.807749e-09, True),
    ("smoother.16", 4, "-rhsrand -solver 1 -rlx 16 -cheby_scale 0 -n 20 20 20 -P 2 2 1 -27pt", 6, 1.555966e-09, True),
    ("smoother.17", 4, "-rhsrand -solver 1 -rlx 16 -cheby_variant 1 -n 20 20 20 -P 2 2 1", 7, 2.088732e-09, True),
    ("smoother.18", 4, "-solver 3 -rlx 16 -cheby_eig_est 0 -n 40 40 20 -P 2 2 1 -difconv -a 10 10 10", 11, 8.192864e-09, True),
    ("smoother.20", 4, "-solver 1 -rlx 16 -cheby_eig_est 5 -n 40 40 20 -P 2 2 1 -vardifconv -eps 0.1", 11, 3.089502e-09, True),
    ("smoother.21", 4, "-solver 1 -rlx 16 -cheby_eig_est 10 -cheby_scale 0 -n 40 40 20 -P 2 2 1", 8, 8.065309e-10, True),
    ("smoother.22", 4, "-solver 1 -rlx 16 -cheby_eig_est 10 -cheby_scale 1 -n 40 40 20 -P 2 2 1", 7, 7.310897e-09, True),
    ("smoother.23", 4, "-solver 1 -rlx 16 -cheby_eig_est 0 -cheby_scale 1 -n 40 40 20 -P 2 2 1", 8, 2.608713e-09, True),
    ("smoother.24", 4, "-solver 1 -rlx 16 -cheby_eig_est 0 -cheby_scale 0 -n 40 40 20 -P 2 2 1", 9, 3.848198e-09, True),
    # solvers.jobs / solvers.saved
    ("solvers.0", 2, "-solver 1 -rhsrand", 7, 3.095059e-09, True),
    ("solvers.2", 2, "-solver 3 -rhsrand", 7, 4.842561e-09, True),
    ("solvers.19", 3, "-n 23 29 31 -solver 1 -rhsrand -precon_cycles 2 -rlx 18", 8, 2.463625e-09, True),
    ("solvers.20", 4, "-n 23 29 31 -solver 3 -rhsrand -precon_cycles 3 -rlx 18", 7, 5.912905e-10, True),
]


def run_case(case, how="emu", timeout=900):
    """returns dict(its, res, device, seconds, ok, why)"""
    cid, nranks, args, its_ref, res_ref, device = case
    exe = os.path.join(REF, BIN[how])
    cmd = [os.path.join(REF, "mpirun"), "-np", str(nranks), exe, *args.split()]
    env = dict(os.environ, OMP_NUM_THREADS="1", HYPRE_B200_VERBOSE="1")
    t0 = time.time()
    # own process group, so that a timeout takes the ranks down with their launcher
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=REF, env=env, start_new_session=True)
    out = {"id": cid, "seconds": 0.0, "ok": False, "why": "", "its": None, "res": None, "device": None}
    try:
        so, se = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        p.communicate()
        out["why"] = f"no result after {timeout} s"
        return out
    r = subprocess.CompletedProcess(cmd, p.returncode, so, se)
    out["seconds"] = time.time() - t0
    if r.returncode != 0:
        out["why"] = f"exit code {r.returncode}: {r.stderr[-800:]}"
        return out
    its = re.findall(r"Iterations = (\d+)", r.stdout)
    res = re.findall(r"Final (?:GMRES )?Relative Residual Norm = ([0-9.eE+-]+)", r.stdout)
    if not its or not res:
        out["why"] = "no iteration count / residual in the output"
        return out
    out["its"], out["res"] = int(its[-1]), float(res[-1])
    out["device"] = "on device" in r.stderr
    # the reference's own check compares the printed text; the device path adds its products in another
    # order than the 1-thread CPU loops where a warp shares a row, so the residual gets a small band
    if out["its"] != its_ref:
        out["why"] = f"iterations {out['its']} != {its_ref}"
    elif abs(out["res"] - res_ref) > 1e-5 * res_ref:
        out["why"] = f"residual {out['res']:.6e} != {res_ref:.6e}"
    elif how != "ref" and out["device"] != device:
        out["why"] = f"device path {'expected' if device else 'not expected'}: {r.stderr[-600:]}"
    else:
        out["ok"] = True
    return out


def main():
    how = sys.argv[1] if len(sys.argv) > 1 else "emu"
    ids = set(sys.argv[2:])
    bad = 0
    for case in CASES:
        if ids and case[0] not in ids:
            continue
        o = run_case(case, how)
        exact = o["res"] is not None and f"{o['res']:.6e}" == f"{case[4]:.6e}"
        print(f"{case[0]:14s} np={case[1]} its {o['its']} (saved {case[3]}) residual {o['res']} (saved {case[4]:.6e}"
              f"{', same digits' if exact else ''}) device={o['device']} {o['seconds']:.1f}s "
              f"{'ok' if o['ok'] else 'FAIL: ' + o['why']}", flush=True)
        bad += 0 if o["ok"] else 1
    print(f"{bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
