"""Top stall sites of one kernel in an .ncu-rep: python tools/ncu_hot.py rep.ncu-rep <kernel ID> [topN]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]   # one section per kernel
    hi = heads[2 * int(kid)]
    end = heads[2 * int(kid) + 1] - 1 if 2 * int(kid) + 1 < len(heads) else len(rows)
    h = rows[hi]
    print("section:", rows[hi - 1][:2])
    si, src = h.index("# Samples"), h.index("Source")
    ie = h.index("Instructions Executed")
    body = [r for r in rows[hi + 1:end] if len(r) == len(h)]
    tot = sum(int(r[si]) for r in body)
    print(f"kernel {kid}: {len(body)} SASS instructions, {tot} samples")
    s0, s1 = h.index("stall_barrier"), h.index("stall_wait")
    agg = {}
    for r in body:
        for j in range(s0, s1 + 1):
            if r[j] not in ("", "0"):
                agg[h[j]] = agg.get(h[j], 0) + int(r[j])
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:10])
    top = sorted(enumerate(body), key=lambda t: -int(t[1][si]))[:topn]
    for i, r in sorted(top):
        st = {h[j][6:]: r[j] for j in range(s0, s1 + 1) if r[j] not in ("", "0")}
        print(f"{i:5d} {r[src].strip()[:64]:64s} smp={r[si]:>4s} exe={r[ie]:>7s} {st}")


if __name__ == "__main__":
    main()
