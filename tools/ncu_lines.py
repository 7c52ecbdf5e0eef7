"""Instructions executed / stall samples per CUDA source line of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on):  python tools/ncu_lines.py rep kernel-substring [N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, tot, tots = {}, 0, 0
fpath = fn = None
hdr = None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ex = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr is None or fn is None or pat not in fn:
        continue
    if r[0] != "":
        cur = (fpath, r[0], r[1].strip()[:95]); continue
    try:
        e = int(r[ex]); s = int(r[sm]) if r[sm].isdigit() else 0
    except (ValueError, IndexError):
        continue
    tot += e; tots += s
    a = agg.setdefault(cur, [0, 0]); a[0] += e; a[1] += s
print("kernel ~", pat, "instructions", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print("%9d %5.1f%%  smp %5.1f%%  %s:%s  %s" % (v[0], 100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(tots, 1), k[0], k[1], k[2]))
