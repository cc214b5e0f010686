"""Device-code diff against a commit: compiles every .cu of hypre_b200/csrc at <commit> and in the
working tree to sm_100a cubins (no GPU needed) and compares the SASS of every kernel, instruction by
instruction.  Used to show that a host-side refactor left the GPU-verified kernels untouched.

usage: python scripts/sass_diff.py <commit> [file.cu ...]
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-DHB200_WITH_NCCL"]


def kernels(cubin):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", line)
        if m and cur is not None:
            res[cur].append(m.group(1))
    return res


def main():
    commit = sys.argv[1]
    csrc = os.path.join(ROOT, "hypre_b200", "csrc")
    files = sys.argv[2:] or sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
    tmp = tempfile.mkdtemp(prefix="sassdiff_")
    try:
        for side in ("old", "new"):
            d = os.path.join(tmp, side, "hypre_b200", "csrc")
            os.makedirs(d)
            os.makedirs(os.path.join(tmp, side, "include"))
            names = subprocess.run(["git", "-C", ROOT, "ls-tree", "--name-only", commit, "hypre_b200/csrc/"],
                                   capture_output=True, text=True, check=True).stdout.split()
            for path in names + ["include/hb200.h"]:
                dst = os.path.join(tmp, side, path)
                if side == "old":
                    blob = subprocess.run(["git", "-C", ROOT, "show", f"{commit}:{path}"], capture_output=True)
                    if blob.returncode == 0:
                        open(dst, "wb").write(blob.stdout)
                elif os.path.exists(os.path.join(ROOT, path)):
                    shutil.copy(os.path.join(ROOT, path), dst)
            for f in os.listdir(csrc):          # files new since the commit
                if side == "new" and os.path.isfile(os.path.join(csrc, f)) and not os.path.exists(os.path.join(d, f)):
                    shutil.copy(os.path.join(csrc, f), os.path.join(d, f))
        worst = 0
        for f in files:
            sass = {}
            for side in ("old", "new"):
                src = os.path.join(tmp, side, "hypre_b200", "csrc", f)
                if not os.path.exists(src):
                    sass[side] = None
                    continue
                cubin = os.path.join(tmp, f"{side}_{f}.cubin")
                r = subprocess.run(["nvcc", *FLAGS, "-cubin", "-o", cubin, src], capture_output=True, text=True,
                                   cwd=os.path.dirname(src))
                if r.returncode != 0:
                    print(f"{f}: nvcc failed on the {side} side\n{r.stderr[-1500:]}")
                    sass[side] = None
                    continue
                sass[side] = kernels(cubin)
            if sass["old"] is None or sass["new"] is None:
                print(f"{f}: only on one side, skipped")
                continue
            o, n = sass["old"], sass["new"]
            changed = [k for k in o if k in n and o[k] != n[k]]
            gone = [k for k in o if k not in n]
            added = [k for k in n if k not in o]
            # a kernel that only changed its mangled name (e.g. a new defaulted template parameter):
            # matched by body
            bodies = {}
            for k in added:
                bodies.setdefault(tuple(n[k]), []).append(k)
            renamed = [k for k in gone if bodies.get(tuple(o[k]))]
            gone = [k for k in gone if k not in renamed]
            print(f"{f}: {len(o)} kernels at {commit}, {len(n)} now; identical {len(o) - len(changed) - len(gone) - len(renamed)}, "
                  f"renamed with identical SASS {len(renamed)}, changed {len(changed)}, removed {len(gone)}, "
                  f"added {len(added) - len(renamed)}")
            for k in changed[:10]:
                print("   changed:", k[:110])
            for k in gone[:10]:
                print("   removed:", k[:110])
            worst = max(worst, len(changed) + len(gone))
        return 1 if worst else 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
