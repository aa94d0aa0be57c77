#!/usr/bin/env python
"""Developer tool: attribute the per-instruction counts of an ncu report to CUDA source lines,
joining the report's SASS page with `nvdisasm -g` of the in-tree library.
Usage: python tools/ncu_lines.py report.ncu-rep <mangled-kernel-substring> [n_points]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "macaque_3d_pose_estimation_b200", "csrc", "libm3d.so")


def disasm_lines(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, active, line = None, False, None
        for ln in txt.splitlines():
            m = re.match(r"^\.text\.(\S+):", ln)
            if m:
                active = kernel_sub in m.group(1)
                line = None
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
                continue
            if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
                out.append(line)
        if out:
            return out
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    npts = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 1.0
    lines = disasm_lines(ksub)
    flt = []
    for a in sys.argv:
        if a.startswith("--kernel="):          # demangled-name regex when the report holds several kernels
            flt = ["-k", "regex:" + a[len("--kernel="):]]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + flt,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    inst = []
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > iI and r[iI].isdigit():
            inst.append((int(r[iI]), int(r[iT]), int(r[iS])))
            stalls.append({n: int(r[i]) for i, n in stall_cols if r[i].isdigit() and int(r[i])})
    print("sass instructions: report %d, disassembly %d" % (len(inst), len(lines)))
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    allst = collections.Counter()
    for (a, b, c), ln, st in zip(inst, lines, stalls):
        key = (ln[0], ln[1]) if ln else ("?", 0)
        agg[key][0] += a
        agg[key][1] += b
        agg[key][2] += c
        agg[key][3].update(st)
        allst.update(st)
    print("stall samples:", ", ".join("%s=%d" % kv for kv in allst.most_common(8)))
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[2] for v in agg.values())
    print("%-22s %6s %10s %8s %8s" % ("file:line", "%inst", "warp/pt", "thr/pt", "%samples"))
    order = -2 if "--by-samples" in sys.argv else -1
    srt = sorted(agg.items(), key=(lambda kv: -kv[1][2]) if "--by-samples" in sys.argv else (lambda kv: -kv[1][0]))
    if "--all" in sys.argv:
        srt = sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1]))
    for key, v in (srt if "--all" in sys.argv else srt[:60]):
        top = ",".join("%s:%d" % kv for kv in v[3].most_common(3))
        print("%-22s %6.2f %10.1f %8.0f %8.2f  %s" % ("%s:%d" % key, 100.0 * v[0] / tot, v[0] / npts, v[1] / npts,
                                                      100.0 * v[2] / max(tots, 1), top))


if __name__ == "__main__":
    main()
