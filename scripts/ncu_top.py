"""Top stall-sampled SASS instructions of an ncu report: python scripts/ncu_top.py rep.ncu-rep [N] (runs `ncu -i ... --page source --csv`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h = {k: i for i, k in enumerate(b["hdr"])}
    sc, ex = h["# Samples"], h["Instructions Executed"]
    stall_cols = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    data = []
    for i, r in enumerate(b["rows"]):
        try:
            data.append((int(r[sc]), i, r))
        except Exception:
            pass
    tot = sum(v for v, _, _ in data)
    print("==", b["name"][:100], "samples", tot, "instrs", len(data))
    agg = {k: 0 for k in stall_cols}
    for v, i, r in data:
        for k in stall_cols:
            try:
                agg[k] += int(r[h[k]])
            except Exception:
                pass
    print("  stalls:", ", ".join(f"{k[6:]}={100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for v, i, r in sorted(data, key=lambda x: -x[0])[:topn]:
        st = sorted(((int(r[h[k]]) if r[h[k]].isdigit() else 0, k[6:]) for k in stall_cols), reverse=True)[:2]
        print(f"{v:7d} {100 * v / max(tot, 1):5.1f}% #{i:5d} exec={r[ex]:>8s} {r[h['Source']].strip()[:70]:70s} {st}")
