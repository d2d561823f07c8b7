"""Summarise an .ncu-rep (ncu --set full) into one line per kernel: python tools/ncu_summary.py rep.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "dur_us", 1e-3), ("sm__cycles_active.avg", "sm_active_cyc", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct_active", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr_pct", 1),
        ("dram__bytes_read.sum", "dram_rd_MB", 1e-6), ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
        ("lts__t_bytes.sum", "l2_MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
        ("sm__inst_executed.avg.per_cycle_active", "ipc", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1)]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("mtts::", "")
        parts = [f"{r[hdr.index('ID')]:>3s} {name:28s}"]
        for key, lab, sc in WANT:
            if key in hdr:
                i = hdr.index(key)
                try:
                    v = float(r[i].replace(",", ""))
                    u = units[i]
                    if lab.endswith("_MB"):
                        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    elif lab == "dur_us":
                        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
                    parts.append(f"{lab}={v:.2f}")
                except ValueError:
                    pass
        out.append(" ".join(parts))
    txt = "\n".join(out)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")


if __name__ == "__main__":
    main()
