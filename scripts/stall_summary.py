"""Per-instruction stall summary of an ncu report captured with --import-source on (experiment tooling).
    python scripts/stall_summary.py gpurun_out/prof.ncu-rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stall}; n = 0
for r in data:
    for s in stall:
        try: tot[s] += int(r[ix[s]])
        except ValueError: pass
    n += int(r[ix["# Samples"]] or 0)
print("total samples", n, " instructions/warp-row (executed sum / 131072):", sum(int(r[ix["Instructions Executed"]] or 0) for r in data) / 131072)
for s, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v * 100 > n: print(f"  {s:26s}{100 * v / n:5.1f}%")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    st = sorted(((int(r[ix[s]] or 0), s) for s in stall), reverse=True)[:2]
    print(f"  {100 * int(r[ix['# Samples']]) / n:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {st[0][1]}:{st[0][0]}")
