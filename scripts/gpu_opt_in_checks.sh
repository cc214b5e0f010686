#!/bin/bash
# First GPU call of the next round: the code paths that were finished after this round's last GPU
# session and have only run in the host emulation so far.
#   1. wide row-pattern kernel (HB200_PAT_WIDE=1) and fused Krylov dots (HB200_FUSED_DOTS=1) on hardware;
#   2. bench with and without fused dots (N = 1);
#   3. with >= 2 GPUs: N = 2 bench with the wide format on (default) and off;
#   (4. and 5. at the end of the file)
TAG=${1:-optin}
NG=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
HB200_PAT_WIDE=1 timeout 600 python -m pytest -q -x -s tests/emu_wide_case.py 2>&1 | tail -3
for v in 1 ""; do
  echo "== fused dots: '${v}'"
  HB200_FUSED_DOTS=$v timeout 600 python -m pytest -q -x -s tests/emu_fused_dots_case.py 2>&1 | grep "REPORT\|passed\|failed"
done
HB200_FUSED_DOTS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pcg or amg" 2>&1 | tail -2
for v in "" 1; do
  if [ -n "$v" ]; then export HB200_FUSED_DOTS=1; else unset HB200_FUSED_DOTS; fi
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_fused_${v:-0}.log 2>&1
  grep '^{' $OUT/bench_fused_${v:-0}.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused=${v:-0}', d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['iterations'], d['config']['final_rel_res'])"
done
unset HB200_FUSED_DOTS
if [ "$NG" -ge 2 ]; then
  for fw in "" 1; do
    if [ -n "$fw" ]; then export HB200_FUSE_WAIT=1; else unset HB200_FUSE_WAIT; fi
    timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "2ranks_peer" 2>&1 | tail -1
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench2_fusewait_${fw:-0}.log 2>&1
    grep '^{' $OUT/bench2_fusewait_${fw:-0}.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fuse_wait=${fw:-0}', d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['iterations'], d['config']['final_rel_res'])"
  done
  unset HB200_FUSE_WAIT
  for w in on off; do
    if [ "$w" = off ]; then export HB200_NO_PAT_WIDE=1; else unset HB200_NO_PAT_WIDE; fi
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench2_wide_$w.log 2>&1
    grep '^{' $OUT/bench2_wide_$w.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('wide=$w', d['value'], d['ms_per_step'], d['config']['iterations'], d['config']['final_rel_res'], [ (e['kernel'][:28], round(e['ms_per_launch'],3)) for e in d['roofline_levels']])"
  done
fi
#   4. upload path after the deferred formats: per-block times of the 256^3 hierarchy (HB200_TRACE lines);
#   5. with >= 4 GPUs: the reference's regression jobs through ij_b200_mpi on hardware (tests/ref_golden_jobs.py).
HB200_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_upload_trace.log 2>&1
grep "upload\|stored transpose" $OUT/bench_upload_trace.log | head -40
if [ "$NG" -ge 4 ]; then
  timeout 900 python tests/ref_golden_jobs.py gpu 2>&1 | cut -c1-260 | tee $OUT/reference_regression_jobs_gpu.txt | tail -26
elif [ "$NG" -ge 2 ]; then
  timeout 600 python tests/ref_golden_jobs.py gpu solvers.0 solvers.2 2>&1 | cut -c1-260
fi
