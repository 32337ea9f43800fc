"""Text summary of `ncu --set full` captures: one CSV row per profiled launch with the counters the roofline discussion
uses (BASELINE.json north_star: dram__bytes.sum.per_second and sm__inst_executed_pipe_tensor against the chip's peaks).

    python tools/ncu_summary.py profiles/r02_kernels.csv  a.ncu-rep [b.ncu-rep ...]

The .ncu-rep binaries themselves are NOT tracked (tens of MB each); this table is what gets committed."""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    ("ms", "gpu__time_duration.sum"),
    ("sm_mhz", "gpc__cycles_elapsed.avg.per_second"),
    ("dram_read_bytes", "dram__bytes_read.sum"),
    ("dram_write_bytes", "dram__bytes_write.sum"),
    ("dram_bytes_per_second", "dram__bytes.sum.per_second"),
    ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_inst", "sm__inst_executed_pipe_tensor.sum"),
    ("issue_slots_busy_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("l1_global_load_sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("registers", "launch__registers_per_thread"),
    ("smem_dyn_bytes", "launch__shared_mem_per_block_dynamic"),
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6,
         "ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6, "Ghz": 1e3, "Mhz": 1, "cycle/nsecond": 1e3, "cycle/usecond": 1}


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    peaks = {}
    ppath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    hbm = float(peaks.get("hbm_gbs", 6583.2)) * 1e9
    rows_out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            rec = {"capture": os.path.basename(rep), "kernel": vals[hdr.index("Kernel Name")][:60],
                   "grid": vals[hdr.index("Grid Size")], "block": vals[hdr.index("Block Size")]}
            for name, metric in WANT:
                v = None
                for i, h in enumerate(hdr):
                    if h == metric or h.endswith("." + metric):
                        try:
                            v = float(vals[i].replace(",", "")) * SCALE.get(units[i], 1)
                        except ValueError:
                            v = None
                        break
                rec[name] = v
            if rec.get("dram_read_bytes") is not None and rec.get("dram_write_bytes") is not None and rec.get("ms"):
                bps = (rec["dram_read_bytes"] + rec["dram_write_bytes"]) / (rec["ms"] * 1e-3)
                rec["dram_bytes_per_second"] = rec["dram_bytes_per_second"] or bps
                rec["dram_pct_of_measured_hbm_peak"] = 100.0 * bps / hbm
            rows_out.append(rec)
    cols = ["capture", "kernel", "grid", "block"] + [n for n, _ in WANT] + ["dram_pct_of_measured_hbm_peak"]
    with open(out_path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for r in rows_out:
            w.writerow({k: ("" if r.get(k) is None else (("%.6g" % r[k]) if isinstance(r[k], float) else r[k])) for k in cols})
    print("wrote %s (%d rows)" % (out_path, len(rows_out)))


if __name__ == "__main__":
    main()
