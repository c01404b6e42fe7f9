"""Per-source-line aggregation of an ncu report's source page (stall samples, executed instructions).
Usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, io, linecache, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    d = dict(zip(hdr, r))
    try: s = int(d.get("# Samples", "0") or 0)
    except ValueError: continue
    key = (cur, d["Line No"])
    a = agg.setdefault(key, [0, 0])
    a[0] += s
    try: a[1] += int(d.get("Instructions Executed", "0") or 0)
    except ValueError: pass
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print("total samples", tot, "instructions", toti)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    try: src = linecache.getline(f, int(ln)).strip()[:100]
    except Exception: src = ""
    print(f"{a[0]:6d} {100*a[0]/max(tot,1):5.1f}%  inst {a[1]:9d} {100*a[1]/max(toti,1):5.1f}%  {f.split('/')[-1]}:{ln}  {src}")
