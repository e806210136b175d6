"""Attribute an ncu source page (cuda,sass csv) to files / functions of this repo by matching line text."""
import csv, sys, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
files = {}
for rel in ("a.i.gar_b200/csrc/agar_simple.cuh", "a.i.gar_b200/csrc/agar_dev.cuh", "a.i.gar_b200/csrc/agar_bots.cuh",
            "a.i.gar_b200/csrc/agar_b200.cu", "include/agar_math.h"):
    files[os.path.basename(rel)] = open(os.path.join(ROOT, rel)).read().split("\n")
def func_of(fname, line):
    src = files[fname]
    for i in range(min(line, len(src)) - 1, -1, -1):
        m = re.match(r"^(?:template.*\n)?(?:DEV|DEVN|AGAR_HD|__global__|static|__device__)[^(]*?([A-Za-z_0-9]+)\s*\(", src[i])
        if m and not src[i].startswith(" "):
            return m.group(1)
    return "?"
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; agg = {}
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or r[0] in ("", "..."): continue
    try: line = int(r[0])
    except ValueError: continue
    text = r[1].strip()
    d = dict(zip(hdr[4:], r[4:]))
    def num(k):
        try: return float(d.get(k, "0") or 0)
        except ValueError: return 0.0
    owner = None
    for fname, src in files.items():
        if line - 1 < len(src) and src[line - 1].strip() == text:
            owner = fname; break
    key = (owner or "other", func_of(owner, line) if owner else "?")
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += num("Instructions Executed"); a[1] += num("Thread Instructions Executed"); a[2] += num("# Samples")
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%-16s %-24s inst %5.1f%%  samples %5.1f%%  avg lanes %4.1f" % (k[0], k[1], 100 * a[0] / tot, 100 * a[2] / max(tots, 1), a[1] / max(a[0], 1)))
