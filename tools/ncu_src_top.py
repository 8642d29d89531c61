"""Top stall-sample SASS lines of a kernel from an ncu report: python tools/ncu_src_top.py rep.ncu-rep regex [n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# possibly several kernels: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks[:1]:
    hdr = b["rows"][0]
    si, src = hdr.index("# Samples"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in b["rows"][1:]:
        try:
            v = int(r[si])
        except Exception:
            continue
        top = sorted(((int(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:2]
        data.append((v, r[src].strip()[:90], top))
    tot = sum(d[0] for d in data)
    print(b["name"][:100], "total samples", tot)
    agg = {}
    for r in b["rows"][1:]:
        for i in stall:
            try:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
            except Exception:
                pass
    print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    for j, (v, s, top) in enumerate(data):
        pass
    order = sorted(range(len(data)), key=lambda j: -data[j][0])[:n]
    for j in sorted(order):
        v, s, top = data[j]
        print("%5d %5.1f%%  #%-4d %-90s %s" % (v, 100.0 * v / max(tot, 1), j, s, top))
