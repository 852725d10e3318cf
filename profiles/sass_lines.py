"""Attribute an ncu `--page source` (SASS) capture to CUDA source lines, with no GPU.

ncu's CSV export of the source page is SASS-only; `nvdisasm -g` of the cubin embedded in libdcr.so (compiled with
-lineinfo) gives the line of every instruction offset.  Joining the two by offset yields, per source line, the share
of executed warp-instructions and of stall samples — what guides the kernel tuning recorded in profiles/.

    python profiles/sass_lines.py <report.ncu-rep> <libdcr.so> [kernel-name-substring] [top N]

The .so must be the build that was profiled.
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(so_path):
    """{mangled function name: [(offset, file, line, inlined_chain)]} for every kernel in the .so."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, capture_output=True)
    table = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        fn, cur = None, None
        for ln in out.splitlines():
            m = re.match(r"^\s*\.text\.(\S+):", ln)
            if m:
                fn = m.group(1)
                table.setdefault(fn, [])
                cur = None
                continue
            m = re.match(r'^\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and fn is not None:
                table[fn].append((int(m.group(1), 16), cur, m.group(2).strip()))
    return table


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        return name


def ncu_source(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            elif r:
                cur["rows"].append(r)
    return blocks


def main():
    report, so = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    table = {demangle(k): v for k, v in line_table(so).items()}
    def norm(s):
        s = s.replace("(bool)1", "true").replace("(bool)0", "false")
        s = re.sub(r"\((?:int|unsigned int|long|unsigned long)\)", "", s)
        s = re.sub(r"^void\s+", "", s.strip())
        return re.sub(r"\s|dcr::", "", s)

    src_cache = {}
    seen = set()
    for b in ncu_source(report):
        if (b["name"], len(b["rows"])) in seen:      # ncu lists every launch; identical code -> print once
            continue
        seen.add((b["name"], len(b["rows"])))
        if want and want not in b["name"]:
            continue
        key = [k for k in table if norm(k).startswith(norm(b["name"]).split("(")[0]) and
               norm(k).split("(")[0] == norm(b["name"]).split("(")[0]]
        if not key:
            print("no line table for", b["name"])
            continue
        lines = table[key[0]]
        hdr = b["hdr"]
        ai, si, ii, ti = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), \
            hdr.index("Thread Instructions Executed")
        base = int(b["rows"][0][ai], 16)
        off2line = {o: l for o, l, _ in lines}
        agg = defaultdict(lambda: [0, 0, 0])
        for r in b["rows"]:
            off = int(r[ai], 16) - base
            l = off2line.get(off)
            a = agg[l]
            a[0] += int(r[si] or 0)
            a[1] += int(r[ii] or 0)
            a[2] += int(r[ti] or 0)
        ts = sum(a[0] for a in agg.values()) or 1
        tinst = sum(a[1] for a in agg.values()) or 1
        print(f"== {b['name']}: {tinst:,} warp-instructions, {ts:,} samples")
        print("   samples%  inst%  thr/inst  file:line  source")
        for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            text = ""
            if l is not None:
                path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", l[0])
                if path not in src_cache:
                    src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
                if 0 < l[1] <= len(src_cache[path]):
                    text = src_cache[path][l[1] - 1].strip()[:90]
            where = f"{l[0]}:{l[1]}" if l else "?"
            print(f"   {100 * a[0] / ts:7.2f} {100 * a[1] / tinst:6.2f} {a[2] / max(a[1], 1):8.1f}  {where:24s} {text}")


if __name__ == "__main__":
    main()
