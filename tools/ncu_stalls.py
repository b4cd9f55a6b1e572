"""Aggregate the per-instruction stall samples of an ncu source-page CSV export.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
  python tools/ncu_stalls.py src.csv [top]
"""
import csv
import collections
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    # the export holds one block per captured launch: "Kernel Name" line, header, instructions
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "ins": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["ins"].append(r)
    b = blocks[0]
    h = {n: i for i, n in enumerate(b["hdr"])}
    stall_cols = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    nsamp = 0
    opmix = collections.Counter()
    for r in b["ins"]:
        for c in stall_cols:
            tot[c] += int(r[h[c]] or 0)
        nsamp += int(r[h["# Samples"]] or 0)
        op = r[h["Source"]].split()[0] if not r[h["Source"]].strip().startswith("@") else r[h["Source"]].split()[1]
        opmix[op.split(".")[0]] += int(r[h["Instructions Executed"]] or 0)
    print(f"kernel: {b['name'][:90]}  launches in file: {len(blocks)}  samples: {nsamp}")
    print("stall reasons (all samples):")
    for c, v in tot.most_common(8):
        print(f"   {c:28s} {v:8d}  {100.0 * v / max(nsamp, 1):5.1f}%")
    allinst = sum(opmix.values())
    print("executed warp-instruction mix:")
    for op, v in opmix.most_common(10):
        print(f"   {op:10s} {v:12d}  {100.0 * v / allinst:5.1f}%")
    print(f"top {top} instructions by samples:")
    order = sorted(b["ins"], key=lambda r: -int(r[h["# Samples"]] or 0))[:top]
    for r in order:
        reasons = sorted(((int(r[h[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
        rs = ", ".join(f"{c[6:]}={v}" for v, c in reasons if v)
        print(f"   {int(r[h['# Samples']]):7d}  {r[h['Source']].strip()[:70]:70s}  {rs}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
