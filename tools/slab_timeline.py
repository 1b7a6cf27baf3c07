"""Per-rank launch timelines of the slab step (UBGL_TIMELINE=<prefix>, common.cuh) -> where each rank's time goes:
kernel time by kind, idle gaps between launches, the halo pushes that waited longest and what ran before them.
    python tools/slab_timeline.py gpurun_out/s2_tl 8
"""
import csv, sys
KINDS = ["other", "fill", "accum", "diffuse", "vbc", "advect", "divergence", "sinks", "rbgs", "zgbc", "residual", "norm",
         "restrict", "prolong", "coarsen", "pbc", "gradient", "prestep", "advdiv", "mg_pre", "mg_post", "mg_coarse",
         "finish", "halo_push", "halo_wait", "colocate", "tracers", "items", "terrain"]
prefix, n = sys.argv[1], int(sys.argv[2])
for r in range(n):
    rows = [(KINDS[int(k)], int(l), float(s), float(d)) for k, l, s, d in list(csv.reader(open(f"{prefix}.{r}.csv")))[1:]]
    # one step = the launches between two prestep groups; split at "prestep" following a non-prestep
    starts = [i for i, x in enumerate(rows) if x[0] == "prestep" and (i == 0 or rows[i - 1][0] != "prestep")]
    starts.append(len(rows))
    i0, i1 = starts[-3], starts[-2]  # the second-last full step
    step = rows[i0:i1 + 1]
    span = step[-1][2] - step[0][2]
    busy, by = 0.0, {}
    gaps = []
    for j, (k, l, s, d) in enumerate(step[:-1]):
        busy += d
        by[k] = by.get(k, 0.0) + d
        g = step[j + 1][2] - (s + d)
        gaps.append((g, k, l, step[j + 1][0], step[j + 1][1]))
    gaps.sort(reverse=True)
    halo = sorted(((d, l, step[j - 1][0], step[j - 1][1]) for j, (k, l, s, d) in enumerate(step[:-1]) if k == "halo_push"), reverse=True)
    print(f"rank {r}: step span {span:.3f} ms, kernels {busy:.3f} ms, idle {span - busy:.3f} ms, launches {len(step) - 1}")
    print("   by kind:", {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
    print("   largest gaps (ms, after kind/level -> before kind/level):", [(round(g, 4), a, b, c, d) for g, a, b, c, d in gaps[:6]])
    print("   sum of gaps > 5 us:", round(sum(g for g, *_ in gaps if g > 0.005), 3), " count", sum(1 for g, *_ in gaps if g > 0.005))
    print("   slowest halo pushes (ms, level, preceded by):", [(round(d, 4), l, a, b) for d, l, a, b in halo[:8]])
