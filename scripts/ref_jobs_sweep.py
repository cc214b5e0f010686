"""No GPU needed: every `ij` line of the reference's TEST_ij/*.jobs files that touches this path, run twice —
through the reference itself (oracle/_ref/ij_refmpi) and through the drop-in (the same unmodified driver in
front of hypre_shim.c; `emu`: host emulation of the kernels, `gpu`: the real library) — and compared:
every "Iterations = N" line must be equal and every final relative residual must agree to 1e-5 relative.
Whether the shim ran the solve on the device or handed it back to the reference is recorded per job.

usage: python scripts/ref_jobs_sweep.py [emu|gpu] [file.jobs ...] [--max-seconds S] [--jobs J] [--max-rows R] [--min-rows R]
                                        [--omp T]
--omp T: both sides run with OMP_NUM_THREADS = T and the drop-in with HYPRE_B200_GS_CHUNKS=host: the hybrid Gauss-Seidel
smoothers then follow the reference's T-thread semantics on the device (one launch per sweep).
"""
import os
import re
import signal
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
TEST_IJ = "/root/reference/src/test/TEST_ij"
DEFAULT_FILES = ["default", "solvers", "smoother", "coarsening", "interp", "agg_interp", "posneg", "nonmixedint", "cheby", "lazy"]
SKIP_FLAGS = ("-fromfile", "-fromparcsrfile", "-fromonecsrfile", "-rhsfromfile", "-rhsparcsrfile", "-print", "-printbin",
              "-frombinfile", "-rhsfromonefile", "-x0fromfile", "-auxfromfile", "-exec_device", "-mm_vendor", "-indexList")


def read_jobs(path):
    text = open(path).read()
    lines = [l for l in text.splitlines() if not l.lstrip().startswith("#")]
    joined = " ".join(l.rstrip("\\").strip() for l in lines)
    out = []
    for m in re.finditer(r"mpirun\s+-np\s+(\d+)\s+\./ij\s+(.*?)\s*>+\s*(\S+)", joined):
        out.append((m.group(3), int(m.group(1)), m.group(2).split()))
    return out


OMP = 1


def run(binary, nranks, args, timeout):
    cmd = [os.path.join(REF, "mpirun"), "-np", str(nranks), os.path.join(REF, binary), *args]
    env = dict(os.environ, OMP_NUM_THREADS=str(OMP), HYPRE_B200_VERBOSE="1")
    if OMP > 1:
        env["HYPRE_B200_GS_CHUNKS"] = "host"
    t0 = time.time()
    # own process group, so that a timeout takes the ranks down with their launcher
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=REF, env=env, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        # SIGTERM first: the emulation unlinks its shared-memory arena on it (a SIGKILLed rank leaves 256 MB of tmpfs behind)
        os.killpg(p.pid, signal.SIGTERM)
        try:
            p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            os.killpg(p.pid, signal.SIGKILL)
            p.communicate()
        return None, None, "", time.time() - t0, "timeout"
    r = subprocess.CompletedProcess(cmd, p.returncode, out, err)
    its = [int(x) for x in re.findall(r"Iterations = (\d+)", r.stdout)]
    res = [float(x) for x in re.findall(r"Final (?:[A-Za-z]+ )?Relative Residual Norm = ([0-9.eE+-]+)", r.stdout)]
    return its, res, r.stderr, time.time() - t0, ("" if r.returncode == 0 else f"exit code {r.returncode}")


def one(job, how, timeout):
    name, nranks, args = job
    its_r, res_r, _, t_r, err_r = run("ij_refmpi", nranks, args, timeout)
    if err_r or not its_r:
        return name, "skipped", f"reference: {err_r or 'no solve in the output'}"
    its_d, res_d, stderr, t_d, err_d = run("ij_b200_emu_mpi" if how == "emu" else "ij_b200_mpi", nranks, args, timeout)
    where = "device" if "on device" in stderr else "reference"
    why = re.findall(r"\[hypre_b200\] (.*?): not on the B200 path", stderr)
    note = f"np={nranks} its {its_d} res {res_d} ran on {where}" + (f" ({why[0]})" if why and where == "reference" else "") + f" {t_d:.0f}s"
    if err_d:
        return name, "FAIL", f"{err_d}: {stderr[-400:]} | {' '.join(args)}"
    if its_d != its_r:
        return name, "FAIL", f"iterations {its_d} != reference {its_r} | {note} | {' '.join(args)}"
    for a, b in zip(res_d, res_r):
        if abs(a - b) > 1e-5 * abs(b) + 1e-300:
            return name, "FAIL", f"residual {a} != reference {b} | {note} | {' '.join(args)}"
    return name, "ok", note + " | " + " ".join(args)


def main():
    argv = sys.argv[1:]
    how = "emu"
    max_seconds, njobs, timeout, max_rows, min_rows = 1e9, 2, 400, 40000, 0
    files = []
    while argv:
        a = argv.pop(0)
        if a in ("emu", "gpu"):
            how = a
        elif a == "--max-seconds":
            max_seconds = float(argv.pop(0))
        elif a == "--jobs":
            njobs = int(argv.pop(0))
        elif a == "--max-rows":
            max_rows = int(argv.pop(0))
        elif a == "--min-rows":
            min_rows = int(argv.pop(0))
        elif a == "--timeout":
            timeout = float(argv.pop(0))
        elif a == "--omp":
            global OMP
            OMP = int(argv.pop(0))
        else:
            files.append(a.replace(".jobs", ""))
    jobs = []
    for f in files or DEFAULT_FILES:
        for j in read_jobs(os.path.join(TEST_IJ, f + ".jobs")):
            if any(s in j[2] for s in SKIP_FLAGS):
                continue
            args = list(j[2])
            # (stand-alone BoomerAMG, ij's default solver 0, prints per cycle at ij's default print level 3: the shim
            # prints those tables itself since the end of round 2, the jobs run as they are written)
            rows = 1000
            if "-n" in args:
                k = args.index("-n")
                rows = int(args[k + 1]) * int(args[k + 2]) * int(args[k + 3])
            if (how == "emu" and rows > max_rows) or rows < min_rows:
                continue                      # minutes per job on the host emulation
            jobs.append((j[0], j[1], args))
    t0 = time.time()
    counts = {"ok": 0, "FAIL": 0, "skipped": 0, "not run": 0}

    def guarded(j):
        if time.time() - t0 > max_seconds:
            return j[0], "not run", "time budget"
        return one(j, how, timeout)

    with ThreadPoolExecutor(njobs) as ex:
        for name, status, note in ex.map(guarded, jobs):
            counts[status] += 1
            print(f"{name:22s} {status:8s} {note}", flush=True)
    print(f"# {len(jobs)} jobs: {counts}")
    return 1 if counts["FAIL"] else 0


if __name__ == "__main__":
    sys.exit(main())
