"""Per-CUDA-source-line totals from `ncu --page source --csv --print-source cuda,sass`
(first launch in the file): warp instructions executed and stall samples per line.
    python tools/ncu_lines.py src_cuda.csv [top_n]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, started = [], False
hdr = None
for r in rows:
    if r and r[0] == "File Path":
        if started and r[1].endswith(".cu") and out:
            break
        continue
    if r and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        started = True
        continue
    if hdr and r and r[0].strip().isdigit() and len(r) == len(hdr):
        try:
            float(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        out.append(r)
col = {h: i for i, h in enumerate(hdr)}
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
ti = sum(float(r[ci]) for r in out)
ts = sum(float(r[cs]) for r in out)
print(f"lines {len(out)}, warp instructions {ti:.4g}, samples {ts:.0f}")
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(out, key=lambda r: -float(r[cs]))[:top]:
    best = max(stall, key=lambda h: float(r[col[h]] or 0))
    print(f"{int(r[0]):5d} inst {float(r[ci]) / ti * 100:5.1f}%  samples {float(r[cs]) / ts * 100:5.1f}%  {best[6:]:12s} {r[1].strip()[:90]}")
