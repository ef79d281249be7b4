"""Attribute executed warp-instructions and stall samples of one kernel to CUDA source lines.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep iter_bwd_kernel [top_n]

Joins ncu's per-SASS-instruction source page with nvdisasm's line table of the SAME build of libtef_b200.so
(ncu's own CUDA view needs the GUI)."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    mangled = kern + "ILb0E" if not kern.endswith("E") else kern          # templated kernels: the non-deterministic instance
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    so = os.path.join(ROOT, "taming_event_flow_b200", "libtef_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    lines = None
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
        m = re.search(r"\.text\.(\S*%s\S*):\n(.*?)(?=\n//-{10,}|\Z)" % mangled, out, re.S)
        if m:
            lines = m.group(2).splitlines()
            break
    assert lines is not None, "kernel not found in the library"
    table, cur = [], ("?", 0)
    for ln in lines:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            table.append(cur)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr = rows[1]
    data = rows[2:]
    iE, iW = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    if len(data) > len(table):
        data = data[:len(table)]
    if len(data) != len(table):
        print("WARNING: %d SASS instructions in the report vs %d in the library (different build?)" % (len(data), len(table)))
    n = min(len(data), len(table))
    ex, st = collections.Counter(), collections.Counter()
    for i in range(n):
        ex[table[i]] += int(data[i][iE])
        st[table[i]] += int(data[i][iW])
    te, ts = sum(ex.values()), sum(st.values())
    src = {}
    print("%-22s %6s %7s %7s  %s" % ("file", "line", "exec%", "stall%", "source"))
    for key, v in ex.most_common(top):
        f, l = key
        if f not in src:
            cand = glob.glob(os.path.join(ROOT, "taming_event_flow_b200", "csrc", f))
            src[f] = open(cand[0]).read().splitlines() if cand else []
        text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
        print("%-22s %6d %6.2f%% %6.2f%%  %s" % (f, l, 100.0 * v / te, 100.0 * st[key] / max(ts, 1), text))


if __name__ == "__main__":
    main()
