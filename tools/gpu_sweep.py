#!/usr/bin/env python
"""Developer tool: run bench.py under several env-selected kernel variants on the GPU box
and print one summary line each (also appended to gpurun_out/sweep.log)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env, args):
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=e,
                       capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    if not line:
        return None, p.stderr[-400:]
    return json.loads(line[-1]), ""


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "sweep.log"), "a")
    specs = sys.argv[1:]
    # spec: workload:frames:ENV=VAL[,ENV=VAL]
    for spec in specs:
        wl, frames, envs = (spec.split(":") + [""])[:3]
        env = dict(kv.split("=") for kv in envs.split(",") if kv)
        d, err = run(env, ["--steps", "5", "--warmup", "3", "--frames", frames, "--workload", wl,
                           "--no-cpu", "--no-e2e"])
        if d is None:
            msg = "%s FAILED %s" % (spec, err)
        else:
            msg = "%s value=%.4g ms=%.3f frac=%.4f subsets=%s clk=%s" % (
                spec, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("mean_subsets_per_point"),
                d["clocks"]["sm_mhz"] if d.get("clocks") else None)
        print(msg, flush=True)
        log.write(msg + "\n")


if __name__ == "__main__":
    main()
