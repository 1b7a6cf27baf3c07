"""profiles/traffic_<workload>.json from an `ncu --set full ... --page raw --csv` dump of ONE step:
DRAM read+write bytes per launch of every (kernel kind, MG level) bench.py's roofline names.
    python tools/ncu_traffic.py raw.csv channel8192 profiles/r01g_ncu_full_summary.txt
"""
import csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1])))
workload, src = sys.argv[2], sys.argv[3]
hdr, units = rows[0], rows[1]
c = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[c[name]].replace(",", ""))
    u = units[c[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
kinds = [("k_prestep", "prestep_fused"), ("k_advect", "advect"), ("k_divergence", "divergence"),
         ("k_gradient_save", "finish_fused"), ("k_mg_tail", "mg_coarse_fused")]
acc, level = {}, {"pre": 0, "post": None}
n_pre = 0
for r in rows[2:]:
    name = r[c["Kernel Name"]]
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    key = None
    if "k_mg_run<0" in name or "k_mg_run<(int)0" in name or "k_mg_tile<3, 0" in name:
        if acc.get("_cycle_done"):
            continue
        key = f"mg_pre_fused:{n_pre}"
        n_pre += 1
    elif "k_mg_run<1" in name or "k_mg_run<(int)1" in name or "k_mg_tile<3, 1" in name:
        if acc.get("_cycle_done"):
            continue
        n_pre -= 1
        key = f"mg_post_fused:{n_pre}"
        if n_pre == 0:
            acc["_cycle_done"] = True  # first V-cycle only
    else:
        for pat, kind in kinds:
            if pat in name and not (kind == "mg_coarse_fused" and acc.get("_cycle_done")):
                key = f"{kind}:{n_pre if kind == 'mg_coarse_fused' else 0}"
    if key:
        acc.setdefault(key, []).append(b)
out = {"workload": workload, "source": f"{src} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
       "traffic_bytes_per_launch": {k: sum(v) / len(v) for k, v in acc.items() if not k.startswith("_")}}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"traffic_{workload}.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
