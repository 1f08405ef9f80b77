"""Per-CUDA-source-line stall samples / instructions from an ncu report (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fn = None
fpath = None
hdr = None
agg = {}
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        fn = row[1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr and row[0] not in ("", "Line No") and fn and (pat in fn):
        ix = {h: i for i, h in enumerate(hdr)}
        try:
            st = int(row[ix["Warp Stall Sampling (All Samples)"]])
            ex = int(row[ix["Instructions Executed"]])
        except Exception:
            continue
        key = (fn[:60], fpath, int(row[0]), row[1].strip()[:90])
        a = agg.setdefault(key, [0, 0])
        a[0] += st
        a[1] += ex
byfn = {}
for (f, fp, ln, src), (st, ex) in agg.items():
    byfn.setdefault(f, []).append((st, ex, fp, ln, src))
for f, rows in byfn.items():
    tst = sum(r[0] for r in rows) or 1
    tex = sum(r[1] for r in rows) or 1
    print("=" * 110)
    print(f, " stall samples", tst, " warp-instructions", tex)
    for st, ex, fp, ln, src in sorted(rows, key=lambda r: -r[0])[:topn]:
        print(f"  {100 * st / tst:5.1f}%  exec {100 * ex / tex:5.1f}%  {fp}:{ln:<4d} {src}")
