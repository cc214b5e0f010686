"""Worker of test_emu_kernels.py::test_peer_halo_watchdog (2 host processes, emulated kernels):
rank 0 runs one exchange that rank 1 never joins; its polling kernel has to give up after
HB200_HALO_TIMEOUT_S and the library has to report the failure instead of hanging."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import emu_env
    emu_env.activate()
    dist.init_process_group("gloo")
    import hypre_b200 as hb
    from hypre_b200._lib import lib, check, HB200Error
    from oracle import refbridge as rb
    hb.init(0)
    uid = [hb.comm_get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    hb.comm_init(rank, world, uid[0])
    check(lib.hb200_set_halo_mode(1))
    rb.load(mpi=True)
    rb.set_num_threads(1)
    pb = rb.Problem("laplacian", (16, 8, 8), P=(2, 1, 1), mpi=True)
    pb.setup_amg(relax_type=18)
    mats, amg = hb.amg_from_hierarchy(pb.hierarchy())
    A = mats[0][0]
    x = torch.ones(A.num_cols, dtype=torch.float64).cuda()
    y = torch.zeros(A.num_rows, dtype=torch.float64).cuda()
    A.matvec(1.0, x, 0.0, y)          # both ranks: builds the plan, one complete exchange
    hb.sync()
    dist.barrier()
    if rank == 0:
        t0 = time.time()
        try:
            A.matvec(1.0, x, 0.0, y)  # rank 1 never sends: the wait kernel must time out
            hb.sync()
            print("NO ERROR RAISED", flush=True)
        except HB200Error as e:
            dt = time.time() - t0
            assert "timed out" in str(e), str(e)
            assert dt < 30.0, dt
            print(f"WATCHDOG OK after {dt:.1f} s: {e}", flush=True)
    else:
        time.sleep(4.0)
    dist.barrier()
    dist.destroy_process_group()
    # the protocol state is broken on purpose: no orderly teardown — but the emulated IPC arena of this process (a POSIX
    # shared-memory object the runtime unlinks at a normal exit) must not stay behind
    import glob
    for f in glob.glob("/dev/shm/hb_emu_%d_*" % os.getpid()):
        try:
            os.unlink(f)
        except OSError:
            pass
    os._exit(0)


if __name__ == "__main__":
    main()
