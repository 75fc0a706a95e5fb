"""Summarise an `ncu --page source --csv` dump: executed warp-instructions per opcode, stall reasons, loop body size."""
import csv, collections, sys
path, denom = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, stalls, tot, hist = collections.Counter(), collections.Counter(), 0, collections.Counter()
for r in data:
    if r[0] == "Kernel Name":
        break
    if len(r) <= ie:
        continue
    toks = [x for x in r[ia].split() if not x.startswith("@")]
    op = toks[0].split(".")[0]
    e = int(r[ie]); ops[op] += e; tot += e; hist[e] += 1
    for i in st:
        try: stalls[hdr[i]] += int(r[i])
        except ValueError: pass
print(f"total executed warp-instructions {tot}  ({tot / denom:.1f} per unit)")
for o, e in ops.most_common(24):
    print(f"  {o:12s} {e:10d} {e / denom:7.1f}")
s = sum(stalls.values())
print("stalls:", ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in stalls.most_common(9)))
print("exec-count histogram (count: #sass lines):", sorted(hist.items(), key=lambda kv: -kv[1])[:6])
