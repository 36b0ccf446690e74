#!/usr/bin/env python
"""hotlines.py — where a kernel's instructions go, by source line (development tool).

Joins the per-instruction counters of an ncu report (`ncu -i X.ncu-rep --page source --csv --print-source sass`) with the
line table of the same kernel in the built library (`nvdisasm -g`), and prints executed warp instructions, thread
instructions and stall samples per source line and per file — the CLI source page only lists the kernel's main file,
and this kernel lives in headers.

    python tools/hotlines.py gpurun_out/prof_r02_final.ncu-rep quoridor_lockstep_kernelILb0ELi5 [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel):
    """[(file, line)] per instruction of `kernel`, in address order, from nvdisasm -g over the library's cubins"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "montecarlotreesearch_b200", "libmctsb.so")], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in out:
            continue
        table, cur, inside = [], ("?", 0), False
        for ln in out.splitlines():
            if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
                inside = kernel in ln
            if not inside:
                continue
            mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                table.append(cur)
        if table:
            return table
    raise SystemExit("kernel %s not found in the library" % kernel)


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = {n: i for i, n in enumerate(rows[hdr])}
    body = [r for r in rows[hdr + 1:] if len(r) > cols["Thread Instructions Executed"]]
    table = line_table(kernel)
    if len(table) != len(body):
        print("warning: %d instructions in the report, %d in the library (different builds?)" % (len(body), len(table)), file=sys.stderr)
    by_line, by_file = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r, key in zip(body, table):
        vals = [int(float(r[cols["Instructions Executed"]] or 0)), int(float(r[cols["Thread Instructions Executed"]] or 0)), int(float(r[cols["# Samples"]] or 0))]
        for acc in (by_line[key], by_file[key[0]], tot):
            for j in range(3):
                acc[j] += vals[j]
    print("total: %.3f G warp instr, %.3f G thread instr (%.1f lanes), %d samples" % (tot[0] / 1e9, tot[1] / 1e9, tot[1] / max(1, tot[0]), tot[2]))
    for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-34s %6.2f %% of warp instr, %6.2f %% of samples, %.1f lanes" % (f, 100 * v[0] / tot[0], 100 * v[2] / max(1, tot[2]), v[1] / max(1, v[0])))
    print("top lines:")
    for (f, l), v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-30s:%-4d %6.2f %% instr  %6.2f %% samples  %.1f lanes" % (f, l, 100 * v[0] / tot[0], 100 * v[2] / max(1, tot[2]), v[1] / max(1, v[0])))


if __name__ == "__main__":
    main()
