"""Summarise an ncu --page source --csv dump: instruction mix and hottest SASS lines per kernel."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None
hdr = None
blocks = {}
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern = row[1]
        hdr = None
        continue
    if row[0] == "Address":
        hdr = row
        blocks[kern] = (hdr, [])
        continue
    if kern and hdr:
        blocks[kern][1].append(row)
for kern, (hdr, rows) in blocks.items():
    if pat and pat not in kern:
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    ex_i, st_i, src_i = ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"], ix["Source"]
    th_i = ix["Avg. Threads Executed"]
    data = []
    for r in rows:
        try:
            data.append((int(r[ex_i]), int(r[st_i]), r[src_i].strip(), float(r[th_i])))
        except Exception:
            pass
    tot_ex = sum(d[0] for d in data) or 1
    tot_st = sum(d[1] for d in data) or 1
    print("=" * 100)
    print(kern[:110])
    print(f"total warp-instructions {tot_ex}  stall samples {tot_st}  SASS lines {len(data)}")
    agg, ags = collections.Counter(), collections.Counter()
    for ex, st, src, th in data:
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        agg[op.split(".")[0]] += ex
        ags[op.split(".")[0]] += st
    print("-- instruction mix (share of executed / share of stall samples)")
    for k, v in agg.most_common(18):
        print(f"   {k:10s} {100 * v / tot_ex:5.1f}%  {100 * ags[k] / tot_st:5.1f}%")
    print(f"-- top {topn} lines by stall samples")
    for n, (ex, st, src, th) in enumerate(sorted(data, key=lambda d: -d[1])[:topn]):
        print(f"   {100 * st / tot_st:5.1f}%  exec {100 * ex / tot_ex:4.1f}%  thr {th:4.1f}  {src[:80]}")
