#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <object.o> <kernel substring> [source.cu] [top]
Joins `ncu --page source --csv` (dynamic "Instructions Executed" / stall samples per SASS instruction,
in program order) with `nvdisasm -g` of the same build (SASS instruction -> source line, same order)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, kern = sys.argv[1:4]
srcfile = sys.argv[4] if len(sys.argv) > 4 else None
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
iE, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
dyn = []
for r in rows[start + 1:]:
    if len(r) <= iE or r[0] == "Address" or r[0] == "Kernel Name":
        break
    dyn.append((r[iSrc].strip(), int(r[iE] or 0), int(r[iS] or 0)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
best = None
for f in re.split(r"\n\s*\.text\.", sass):
    name = f.split("\n", 1)[0]
    cur, lines = None, []
    for line in f.split("\n"):
        m = re.search(r'//## File "(.*)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            lines.append(cur)
    if len(lines) == len(dyn) and kern.split("<")[0].replace("regex:", "") in name or (best is None and len(lines) == len(dyn)):
        best = (name, lines)
if best is None:
    sys.exit("no function with %d SASS instructions found (different build?)" % len(dyn))
name, lines = best
tot = sum(d[1] for d in dyn)
tots = sum(d[2] for d in dyn)
by = collections.Counter(); bys = collections.Counter()
for (src, e, s), ln in zip(dyn, lines):
    by[ln] += e; bys[ln] += s
text = open(srcfile).read().split("\n") if srcfile else None
print("%s\n total warp-instructions %d, stall samples %d" % (name[:80], tot, tots))
for ln, c in by.most_common(top):
    t = ""
    if text and ln and ln[0] == os.path.basename(srcfile) and ln[1] <= len(text):
        t = text[ln[1] - 1].strip()[:90]
    print("  %-28s %6.2f%% instr %6.2f%% samples  %s" % ("%s:%s" % ln if ln else "?", 100.0 * c / tot, 100.0 * bys[ln] / max(tots, 1), t))
