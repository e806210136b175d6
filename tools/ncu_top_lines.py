"""Summarise an ncu source page (--print-source cuda,sass --csv): top source lines by instructions and samples."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
cur = None; out = []; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or r[0] in ("", "...") : continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    def num(k):
        try: return float(d.get(k, "0") or 0)
        except ValueError: return 0.0
    out.append((cur, line, r[1].strip()[:90], num("Instructions Executed"), num("# Samples"), num("Avg. Threads Executed"), num("L1 Conflicts Shared N-Way")))
tot_i = sum(o[3] for o in out); tot_s = sum(o[4] for o in out)
print("total instructions %.3e samples %d" % (tot_i, tot_s))
for key, name in ((3, "instructions"), (4, "samples")):
    print("== top by", name)
    for o in sorted(out, key=lambda o: -o[key])[:topn]:
        print("%-14s %4d  inst %5.1f%%  smp %5.1f%%  thr %4.1f  | %s" % (o[0], o[1], 100*o[3]/tot_i, 100*o[4]/max(tot_s,1), o[5], o[2]))
