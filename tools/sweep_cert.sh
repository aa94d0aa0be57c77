# usage: bash tools/sweep_cert.sh "ENV=.. [--bench-flag ..]" ...   (one quoted spec per run)
run() { env $(echo "$1" | tr ' ' '\n' | grep = | grep -v '^--' | tr '\n' ' ') python bench.py --only --steps 5 --warmup 3 --no-cpu --no-e2e $(echo "$1" | tr ' ' '\n' | grep '^--' | tr '\n' ' ' | sed 's/=/ /g') > gpurun_out/sw.json 2> gpurun_out/sw.err; tail -2 gpurun_out/sw.err; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/sw.json') if l.startswith('{')][0]); print(sys.argv[1], '%.4g'%d['value'], round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})" "$1"; }
for v in "$@"; do run "$v"; done
