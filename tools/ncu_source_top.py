"""Top stall lines of an .ncu-rep source page: python tools/ncu_source_top.py rep [N] [kernel-substr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = blk.splitlines()
    name = lines[0]
    if len(sys.argv) > 3 and sys.argv[3] not in name:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    si = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
    src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = 0
    for r in rows[1:]:
        if len(r) <= si: continue
        try: s = int(r[si])
        except ValueError: continue
        tot += s
        data.append((s, r))
    print("==", name[:100], "total samples", tot)
    agg = {}
    for s, r in data:
        for i in stall_cols:
            try: agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
            except ValueError: pass
    print("   stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for s, r in sorted(data, key=lambda x: -x[0])[:N]:
        top = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stall_cols), reverse=True)[:2]
        print("   %6d %5.1f%%  ex=%-8s %-70s %s" % (s, 100.0 * s / max(tot, 1), r[ex], r[src].strip()[:70], top))
