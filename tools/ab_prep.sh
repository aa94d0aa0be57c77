#!/bin/bash
# Developer tool (run under gpurun): A/B of the k_cert_prep variants (M3D_PREP_ILP builds under tools/_alt/,
# CTAs per SM), then parity + stress + default bench + captures with the fastest one installed.
# The variant libraries are built beforehand in the build container (git-ignored, they travel with the snapshot):
#   cd macaque_3d_pose_estimation_b200/csrc && for v in 0 1 2; do
#     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DM3D_PREP_ILP=$v \
#          -c -o /tmp/ransac_v$v.o m3d_ransac.cu
#     nvcc -shared -o ../../tools/_alt/libm3d_ilp$v.so /tmp/ransac_v$v.o $(ls _obj/*.o | grep -v m3d_ransac.o); done
# (libm3d.so itself is the M3D_PREP_ILP=3 build.)  Result of the round-2 run: profiles/r02g_ab_prep.txt.
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $1"; }
L=macaque_3d_pose_estimation_b200/csrc/libm3d.so
cp $L tools/_alt/libm3d_ilp3.so
: > gpurun_out/ab_prep.txt
one() {  # lib-variant ctas
  cp tools/_alt/libm3d_ilp$1.so $L
  M3D_CERT_SETUP_CTAS=$2 python bench.py --only --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$1" "$2" <<'P' | tee -a gpurun_out/ab_prep.txt
import json, sys
d = json.loads([l for l in open('gpurun_out/sw.json') if l.startswith('{')][0])
k = d['roofline']['kernels']
print(sys.argv[1], sys.argv[2], round(d['ms_per_step'], 3), round(k['k_cert_prep']['ms_per_step'], 3), '%.4g' % d['value'])
P
}
one 3 5; one 3 4; one 1 5; one 1 4; one 2 5; one 0 5
el sweep
best=$(sort -k3 -g gpurun_out/ab_prep.txt | head -1)
echo "best: $best"
v=$(echo $best | cut -d' ' -f1); c=$(echo $best | cut -d' ' -f2)
cp tools/_alt/libm3d_ilp$v.so $L
export M3D_CERT_SETUP_CTAS=$c
echo "$v $c" > gpurun_out/ab_prep_choice.txt
timeout 200 python -m pytest tests/test_ransac_cert.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q -k "ransac or cert or stage or step4 or smoke" > gpurun_out/t_r02g.log 2>&1; tail -2 gpurun_out/t_r02g.log; el tests
timeout 150 python tools/stress_cert.py > gpurun_out/stress_r02g.log 2>&1; tail -1 gpurun_out/stress_r02g.log; cp gpurun_out/stress_cert.json gpurun_out/stress_cert_r02g.json 2>/dev/null; el stress
timeout 200 python bench.py > gpurun_out/bench_default7.json 2> gpurun_out/bench_default7.err; tail -c 200 gpurun_out/bench_default7.json; el bench
timeout 120 bash tools/profile_round.sh r02g; el profile
