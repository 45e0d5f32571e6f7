"""profiles/r02_dram_traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch,
keyed workload -> kernel family (the names bench.py's family_of() produces).

    python scripts/ncu_traffic_json.py ml1m:sbr_mlp2_bwd=gpurun_out/r02_mlp2_bwd.ncu-rep ml1m:sbr_mlp2_fwd=... 
merges into the existing file."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = json.load(open(path)) if os.path.exists(path) else {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=")
    workload, family = key.split(":")
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                                          text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    tot, n, dur, names = 0.0, 0, 0.0, set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        b = sum(float(d[m]) * UNIT[u[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tot += b
        n += 1
        dur += float(d["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u["gpu__time_duration.sum"], 1.0)
        names.add(d["Kernel Name"].split("(")[0])
    out.setdefault(workload, {})[family] = dict(dram_bytes_per_launch=tot / n, launches_captured=n,
                                                avg_us_under_ncu=dur / n, kernel=sorted(names)[0],
                                                source="profiles/" + os.path.basename(rep).replace(".ncu-rep", "_ncu_raw.txt"))
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
