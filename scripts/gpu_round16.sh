#!/bin/bash
# N-GPU bench line only
TAG=${1:-r16}
NG=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 3 --warmup 3 > $OUT/bench.log 2>&1
grep '^{' $OUT/bench.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['config']['iterations'], d['config']['final_rel_res'], d['config']['halo'], d['config']['cuda_graph_vcycle'], d['config']['upload_s'], d['config']['setup_s_reference_cpu'])
for e in d['roofline_levels']: print(e['kernel'][:120], round(e['ms_per_launch'],4), round(e['frac'],3))
"
tail -3 $OUT/bench.log | cut -c1-300
