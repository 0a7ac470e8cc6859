#!/usr/bin/env python
"""Reads ncu reports (ncu -i ... --page raw --csv) and writes the metrics the profiles/ summaries quote.

  python scripts/ncu_extract.py out.csv report1.ncu-rep [report2.ncu-rep ...]

One row per captured launch: kernel, duration, DRAM bytes, L2 hit rate, issue-slot / tensor-pipe / LSU utilisation,
registers, and the top warp-stall reasons (ratio per issue-active cycle)."""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefront_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    for r in body:
        d = {"report": path, "kernel": r[hdr.index("Kernel Name")][:80]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in WANT:
                v = r[i].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    continue
                if WANT[h].startswith("dram_r") or WANT[h].startswith("dram_w"):
                    v *= UNIT_SCALE.get(units[i], 1.0)
                if WANT[h] == "duration_ms":
                    v *= {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6,
                          "second": 1e3}.get(units[i], 1.0)
                d[WANT[h]] = v
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
        d["top_stalls"] = " ".join(f"{k}={v:.2f}" for k, v in top)
        yield d


def main():
    out, reports = sys.argv[1], sys.argv[2:]
    rows = [d for p in reports for d in rows_of(p)]
    cols = ["report", "kernel"] + list(dict.fromkeys(WANT.values())) + ["top_stalls"]
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for d in rows:
            w.writerow({c: d.get(c, "") for c in cols})
    for d in rows:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k != "report"})


if __name__ == "__main__":
    main()
