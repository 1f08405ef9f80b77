"""Per memory instruction: executed count, L1 tag requests and L2 theoretical sectors (ncu --page source --csv), grouped
by opcode and access size; top instructions by L1 tag requests.  usage: ncu_mem.py report.ncu-rep [kernel-substring]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, hdr, blocks = None, None, {}
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern, hdr = row[1], None
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
    print("=" * 100)
    print(kern[:120])
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    items = []
    for r in rows:
        try:
            ex = int(r[ix["Instructions Executed"]])
            tag = int(r[ix["L1 Tag Requests Global"]] or 0)
            l2 = int(r[ix["L2 Theoretical Sectors Global"]] or 0)
            l2i = int(r[ix["L2 Theoretical Sectors Global Ideal"]] or 0)
        except Exception:
            continue
        if tag == 0 and l2 == 0:
            continue
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        a = agg[op]
        a[0] += ex
        a[1] += tag
        a[2] += l2
        a[3] += l2i
        items.append((tag, ex, l2, l2i, src))
    print("opcode                       executed     L1 tag req   tags/instr   L2 sectors   ideal")
    for op, (ex, tag, l2, l2i) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {op:24s} {ex:12d} {tag:12d} {tag / max(ex, 1):8.2f} {l2:12d} {l2i:12d}")
    print("top instructions by L1 tag requests")
    for tag, ex, l2, l2i, src in sorted(items, reverse=True)[:24]:
        print(f"  tags {tag:11d} exec {ex:10d} ({tag / max(ex, 1):5.2f}/instr) L2 {l2:11d}/{l2i:11d}  {src[:70]}")
