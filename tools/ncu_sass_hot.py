"""Per-instruction view of an `ncu --page source --csv` dump: opcode histogram weighted by
executed count, and the top stall sites.  Usage: python tools/ncu_sass_hot.py src.csv [top]"""
import csv
import collections
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > idx["Instructions Executed"]
            and r[idx["Instructions Executed"]].isdigit()]
    tot = sum(int(r[idx["Instructions Executed"]]) for r in data)
    samples = sum(int(r[idx["# Samples"]]) for r in data)
    ops = collections.Counter()
    for r in data:
        op = r[idx["Source"]].strip().split()
        op = [o for o in op if not o.startswith("@")]
        ops[op[0].split(".")[0] if op else "?"] += int(r[idx["Instructions Executed"]])
    print(f"total warp-instructions {tot:.3e}, stall samples {samples}")
    print("opcode mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(18)))
    print("top stall sites (samples, executed, SASS):")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top]:
        print(f"  {int(r[idx['# Samples']]):7d} {100 * int(r[idx['# Samples']]) / max(samples, 1):5.1f}%  "
              f"{int(r[idx['Instructions Executed']]):12d}  {r[idx['Source']].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
