"""profiles/traffic.json from the ncu --set full captures of k_hour and k_commit: DRAM bytes per launch (read + write)."""
import csv, io, json, subprocess, sys
out = {}
for name, rep in (("k_hour", sys.argv[1]), ("k_commit", sys.argv[2])):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    vals = []
    for r in rows[2:]:
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            k = h.index(key)
            v = float(r[k].replace(",", ""))
            u = units[k].lower()
            tot += v * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1.0)
        vals.append(tot)
    out[name] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals), "kernel": rows[2][h.index("Kernel Name")]}
out["pass_dram_bytes_per_launch"] = out["k_hour"]["dram_bytes_per_launch"] + out["k_commit"]["dram_bytes_per_launch"]
out["workload"] = "10m"
out["source"] = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none, k_hour / k_commit at a work hour (h = 11, 12)"
print(json.dumps(out, indent=1))
