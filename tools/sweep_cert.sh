run() { env "$@" python bench.py --only --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/sw.json 2> gpurun_out/sw.err; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/sw.json') if l.startswith('{')][0]); print(sys.argv[1:], '%.4g'%d['value'], round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})" "$@"; }
run M3D_CERT_LANE_LIMIT=16
run M3D_CERT_LANE_LIMIT=32
run M3D_CERT_LANE_LIMIT=16 M3D_CERT_SETUP_CTAS=4
run M3D_CERT_LANE_LIMIT=16 M3D_CERT_SETUP_CTAS=2
run M3D_CERT_LANE_LIMIT=16 M3D_CERT_CTAS=4
run M3D_CERT_LANE_LIMIT=16 M3D_CERT_CTAS=2
