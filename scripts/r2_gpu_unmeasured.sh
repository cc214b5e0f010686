#!/bin/bash
# First GPU session after round 2: hardware numbers for the three pieces written after the round's GPU budget was
# spent: spmv_box<...,GEO>, the chunked hybrid Gauss-Seidel, the f4 Krylov drivers.
#   gpurun --timeout 1500 -- 'bash scripts/r2_gpu_unmeasured.sh r3a'
TAG=${1:-r3a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD
{
echo "== parity of the new code on hardware"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "krylov_ext or mass_inner or hybrid_gs or reference_thread_count or matvec or ij_dropin" tests/test_ij_formats.py 2>&1 | tail -5

echo "== spmv_box: GEO against the predicated kernel of the round-2 profiles (27-pt, one SpMV launch)"
for n in 128 192 256 384; do
   for geo in 0 1; do
      HB200_BOX_NO_GEO=$((1 - geo)) timeout 300 python bench.py --spmv-only --n $n --steps 20 --warmup 5 --no-cpu-baseline > $OUT/spmv_${n}_geo${geo}.json 2> $OUT/spmv_${n}_geo${geo}.err
      python - <<EOF
import json
d = json.loads([l for l in open("$OUT/spmv_${n}_geo${geo}.json") if l.startswith("{")][-1])
print("n=$n geo=$geo", d["roofline"]["kernel"][:40], "ms", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"])
EOF
   done
done

echo "== default solve: GEO on / off, the Jacobi sweep through the box kernel on / off"
for cfg in "HB200_BOX_NO_GEO=1" "HB200_BOX_NO_GEO=0" "HB200_BOX_JACOBI=1"; do
   env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-ij > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err
   python - <<EOF
import json
d = json.loads([l for l in open("$OUT/bench_$cfg.json") if l.startswith("{")][-1])
print("$cfg", "ms/solve", d["ms_per_step"], "its", d["config"]["iterations"], "res", d["config"]["final_rel_res"], "A0 ms", d["roofline"]["ms_per_launch"])
EOF
done

echo "== hypre's default smoother (hybrid l1-GS 13/14) through the drop-in at 128^3: wavefronts against chunks"
for mode in "" host auto 4736 18944; do
   /usr/bin/time -f "%e s wall" env OMP_NUM_THREADS=16 HYPRE_B200_GS_CHUNKS=$mode HYPRE_B200_VERBOSE=1 timeout 900 \
      oracle/_ref/ij_b200 -27pt -n 128 128 128 -solver 1 2>&1 | grep -i "on device\|Iterations\|Final Rel\|wall" | sed "s/^/chunks='$mode': /"
done

echo "== f4 drivers on config 4 (vardifconv 256^3): GMRES against COGMRES / FlexGMRES / LGMRES / BiCGSTAB"
for sv in gmres cogmres flexgmres lgmres bicgstab; do
   timeout 900 python bench.py --problem vardifconv --solver $sv --steps 5 --warmup 2 --no-e2e-ij > $OUT/vdc_$sv.json 2> $OUT/vdc_$sv.err
   python - <<EOF
import json
d = json.loads([l for l in open("$OUT/vdc_$sv.json") if l.startswith("{")][-1])
print("$sv", "ms/solve", d["ms_per_step"], "its", d["config"]["iterations"], "parity", d["config"].get("parity_vs_reference"))
EOF
done

echo "== upload of the 27-pt 256^3 hierarchy: single-core host passes (the round-2 trace) against the threaded ones"
for cfg in "HB200_UPLOAD_THREADS=1 HB200_ANALYSIS_THREADS=1" "HB200_UPLOAD_THREADS=4 HB200_ANALYSIS_THREADS=8"; do
   env $cfg HB200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-ij > $OUT/upload.json 2> $OUT/upload.err
   echo "$cfg: upload_s $(python -c "import json;print(json.loads([l for l in open('$OUT/upload.json') if l.startswith('{')][-1])['config']['upload_s'])")"
   grep "CSR copy\|row patterns" $OUT/upload.err | head -6
done

echo "== ncu: the GEO kernel (one capture), launch list of the default solve"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_box -c 1 -o $OUT/ncu_spmv_box_geo \
   python bench.py --spmv-only --n 256 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_geo.log 2>&1
ncu -i $OUT/ncu_spmv_box_geo.ncu-rep --page raw --csv > $OUT/ncu_spmv_box_geo_raw.csv 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-ij > $OUT/launches.log 2>&1
} 2>&1 | tee $OUT/session.log
