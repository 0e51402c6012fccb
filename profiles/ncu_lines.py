"""Stall samples of an ncu capture (--set full --import-source on) summed per CUDA source line.

    ncu -i gpurun_out/<rep>.ncu-rep --page source --csv > /tmp/src.csv
    python profiles/ncu_lines.py /tmp/src.csv [kernel-name-substring] [top N]

The source page of the CSV export is per SASS instruction; the line table comes from the cubins of the built library
(cuobjdump -xelf + nvdisasm -g), matched by the offset of the instruction inside its function.
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "revs-admm_b200", "librevs_admm.so")


def line_tables():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    maps = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        if "-" in os.path.basename(cubin).split(".")[0]:
            continue                       # the linked image repeats the per-file cubins
        out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur, line, fname = None, None, None
        for l in out.split("\n"):
            m = re.match(r"\.text\.(\S+):", l)
            if m:
                cur, line = m.group(1), None
                maps[cur] = {}
                continue
            m = re.search(r'//## File "(.*?)", line (\d+)', l)
            if m:
                fname, line = m.group(1), int(m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m and cur:
                maps[cur][int(m.group(1), 16)] = (fname, line)
    return maps


def demangled_key(maps, name):
    out = subprocess.run(["cu++filt"] + list(maps), capture_output=True, text=True).stdout.split("\n")
    for k, d in zip(maps, out):
        if d.replace(" ", "") == name.replace(" ", "").replace("(int)", "").replace("(bool)", ""):
            return k
    flat = lambda s: re.sub(r"\(int\)|\(bool\)|\s|void", "", s)
    for k, d in zip(maps, out):
        if flat(d) == flat(name):
            return k
    return None


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    maps = line_tables()
    secs = []
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            secs.append([r[1], None, []])
        elif r and r[0] == "Address" and secs:
            secs[-1][1] = r
        elif r and secs and secs[-1][1] is not None:
            secs[-1][2].append(r)
    seen = set()
    srcs = {}
    for name, h, body in secs:
        if want not in name or not body:
            continue
        key = demangled_key(maps, name)
        sig = (name, len(body), body[0][0])
        if key is None or sig in seen:
            continue
        seen.add(sig)
        mp = maps[key]
        base = int(body[0][0], 16)
        si = h.index("# Samples")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        per, per_st, tot = collections.Counter(), collections.defaultdict(collections.Counter), 0
        for r in body:
            loc = mp.get(int(r[0], 16) - base, (None, None))
            n = int(r[si] or 0)
            per[loc] += n
            tot += n
            for i in stall:
                v = int(r[i] or 0)
                if v:
                    per_st[loc][h[i][6:]] += v
        print(f"===== {name}: {tot} samples")
        for loc, n in per.most_common(top):
            f, ln = loc
            text = ""
            if f and os.path.exists(f):
                if f not in srcs:
                    srcs[f] = open(f).read().split("\n")
                text = srcs[f][ln - 1].strip()[:100] if ln and ln <= len(srcs[f]) else ""
            st = ", ".join(f"{k}:{v}" for k, v in per_st[loc].most_common(3))
            print(f"{n:6d} {100 * n / max(tot, 1):5.1f}%  {os.path.basename(f) if f else '?'}:{ln}: {text}   [{st}]")


if __name__ == "__main__":
    main()
