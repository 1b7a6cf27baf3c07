"""Condense `ncu --page source --csv` (SASS view) of ONE kernel: totals of the stall
reasons, instruction mix, and the hottest instructions.
    ncu -i x.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
    python tools/ncu_source_summary.py src.csv [launch_index]
"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split into per-launch sections (each starts with a "Kernel Name" row)
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
s = secs[which]
hdr, body = s["rows"][0], [r for r in s["rows"][1:] if len(r) == len(s["rows"][0])]
col = {h: i for i, h in enumerate(hdr)}
def num(r, h):
    try:
        return float(r[col[h]])
    except (ValueError, KeyError):
        return 0.0
print(s["name"], "launch", which, "SASS instructions", len(body))
tot_inst = sum(num(r, "Instructions Executed") for r in body)
tot_samp = sum(num(r, "# Samples") for r in body)
print(f"warp instructions executed {tot_inst:.4g}, stall samples {tot_samp:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((sum(num(r, h) for r in body), h) for h in stalls), reverse=True)
print("stall reasons (all samples):", ", ".join(f"{h[6:]} {v / max(tot_samp, 1) * 100:.1f}%" for v, h in agg[:8]))
mix = collections.Counter()
for r in body:
    op = r[col["Source"]].split()
    op = [t for t in op if not t.startswith("@")]
    mix[op[0].split(".")[0] if op else "?"] += num(r, "Instructions Executed")
print("instruction mix:", ", ".join(f"{k} {v / tot_inst * 100:.1f}%" for k, v in mix.most_common(14)))
sh = sum(num(r, "L1 Wavefronts Shared") for r in body)
shi = sum(num(r, "L1 Wavefronts Shared Ideal") for r in body)
print(f"shared wavefronts {sh:.4g} (ideal {shi:.4g}), global tag requests {sum(num(r, 'L1 Tag Requests Global') for r in body):.4g}")
print("hottest instructions by stall samples:")
for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:18]:
    top = max(stalls, key=lambda h: num(r, h))
    print(f"  {num(r, '# Samples'):6.0f}  {r[col['Source']].strip()[:70]:70s} {top[6:]}")
