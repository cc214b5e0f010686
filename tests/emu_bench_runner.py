"""Test infrastructure: runs the repo's bench.py unchanged against the host emulation of the kernels
(child of tests/test_emu_kernels.py; needs HB200_EMU_TEST=1).  It checks the CODE PATH of the bench —
argument handling, stages and watchdog, halo-mode choice, per-level timing, the JSON contract — on a
machine without a GPU; the numbers it prints are meaningless and are never kept."""
import os
import runpy
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emu_env  # noqa: E402

emu_env.activate()
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)

    def synchronize(self):
        pass


torch.cuda.Event = _Event
torch.cuda.ExternalStream = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
_orig_init = dist.init_process_group


def _init(backend=None, **kw):
    kw.pop("device_id", None)
    return _orig_init("gloo", **kw)


dist.init_process_group = _init
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
