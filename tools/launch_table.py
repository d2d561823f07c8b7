"""Turn an ncu `--metrics gpu__time_duration.sum --csv` launch list of tools/one_solve.py into a
per-launch table of the FIRST estimator evaluation: python tools/launch_table.py launches.csv [out.txt]"""
import csv
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from profile_solve import labels  # noqa: E402


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    names = [r[4] for r in rows]
    ns = [float(r[-1]) for r in rows]
    pro, per = labels()
    # the launch list may include weight-packing kernels first; align on the end
    n_per = len(per)
    # take the first complete estimator evaluation: it starts at the first STATS conv GEMM
    first = next(i for i, nm in enumerate(names) if "gemm_tc_kernel<256, 0" in nm)
    last = list(zip(names, ns))[first:first + n_per]
    out = []
    tot = 0.0
    by = {}
    for lab, (nm, t) in zip(per, last):
        short = nm.split("(")[0].replace("void mtts::", "")
        out.append(f"{lab:16s} {t/1e3:8.2f} us  {short}")
        tot += t
        k = lab.split(".")[1]
        by[k] = by.get(k, 0.0) + t
    out.append(f"sum of kernel durations (ncu, serialised, cold L2): {tot/1e3:.1f} us over {n_per} launches")
    out.append("by layer type: " + ", ".join(f"{k}={v/1e3:.1f}" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))
    txt = "\n".join(out)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")


if __name__ == "__main__":
    main()
