#!/bin/bash
# one validation call: targeted parity, sweep, stress, default bench, profile
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $1"; }
timeout 200 python -m pytest tests/test_ransac_cert.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q -k "ransac or cert or stage or step4 or smoke" > gpurun_out/t_r02f.log 2>&1; tail -2 gpurun_out/t_r02f.log; el tests
bash tools/sweep_cert.sh "X=0" "M3D_CERT_SETUP_CTAS=6" "M3D_CERT_SETUP_CTAS=4" 2>&1 | tail -6; el sweep
timeout 150 python tools/stress_cert.py > gpurun_out/stress_r02f.log 2>&1; tail -1 gpurun_out/stress_r02f.log; cp gpurun_out/stress_cert.json gpurun_out/stress_cert_r02f.json 2>/dev/null; el stress
timeout 200 python bench.py > gpurun_out/bench_default6.json 2> gpurun_out/bench_default6.err; tail -c 300 gpurun_out/bench_default6.json; el bench
timeout 120 bash tools/profile_round.sh r02f; el profile
