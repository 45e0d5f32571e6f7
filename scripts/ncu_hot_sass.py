"""Top SASS instructions by warp-stall samples of one kernel in an .ncu-rep (`ncu -i X --page source --csv`)."""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for n, r in enumerate(rows[2:]):
    try:
        s = int(r[ix["# Samples"]] or 0)
    except Exception:
        continue
    recs.append((s, n, r))
tot = sum(s for s, _, _ in recs)
print("total samples", tot, "instructions", len(recs), "executed warp-insts", sum(int(r[ix["Instructions Executed"]] or 0) for _, _, r in recs))
for s, n, r in sorted(recs, key=lambda t: -t[0])[:top]:
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{s:6d} {100.0 * s / tot:5.1f}%  #{n:5d} exec={r[ix['Instructions Executed']]:>8s}  {r[ix['Source']][:90]:90s} {why}")
