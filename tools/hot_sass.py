"""Per-opcode executed-instruction mix of the hot loop from `ncu --page source --csv --print-source sass`."""
import csv, sys
from collections import Counter
path, want, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 300000
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for s in secs:
    if want not in s["name"]:
        continue
    hdr = s["rows"][0]; data = [r for r in s["rows"][1:] if len(r) > 10]
    ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in data)
    print(s["name"][:80], "total/unit %.1f" % (tot / per))
    hot = [r for r in data if int(r[ie]) >= thr]
    print(len(hot), "hot instrs; sum/unit %.1f" % (sum(int(r[ie]) for r in hot) / per))
    c = Counter()
    for r in hot:
        t = r[src].split()
        op = t[1] if t[0].startswith('@') else t[0]
        c[op.split('.')[0]] += int(r[ie])
    print(", ".join("%s %.1f" % (k, v / per) for k, v in c.most_common(40)))
    open('/tmp/hot.txt', 'w').write("\n".join("%8d %5s  %s" % (int(r[ie]), r[smp], r[src]) for r in hot))
    break
