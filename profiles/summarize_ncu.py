#!/usr/bin/env python
"""Turn `ncu --set full` reports of tests/perf_ops.py (one report per shape) into the two committed summaries:

    python profiles/summarize_ncu.py r02 s1=gpurun_out/prof_s1.ncu-rep s2=... s3=...

  profiles/<tag>_ncu_full_summary.json  per (shape, kernel): mean duration, DRAM bytes, issue/warp activity, stalls
  profiles/<tag>_ncu_traffic.json       dram__bytes_read.sum + dram__bytes_write.sum of the MAIN kernel of each op,
                                        one launch, keyed "op,C,r" (what bench.py reports as roofline.traffic)
The reports were captured with  ncu --set full --import-source on --clock-control none -k regex:<kernels>  on a
command that had already exited 0 without ncu (B200_PROFILING.md)."""
import csv
import io
import json
import os
import subprocess
import sys

SHAPES = {"s1": (128, 32), "s2": (256, 16), "s3": (256, 8)}
MAIN = {"vox_fwd_cell_kernel": "vox_fwd", "devox_fwd_kernel": "devox_fwd", "devox_bwd_cell_kernel": "devox_bwd",
        "row_gather_kernel": "vox_bwd", "tri_fwd_kernel": "devox_fwd", "tri_bwd_kernel": "devox_bwd"}   # r02: tri.cu walks
COLS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__inst_executed.sum": "inst_executed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_wavefront_pct",
    "smsp__inst_executed.sum": "warp_inst_executed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
}


def main():
    tag = sys.argv[1]
    out, traffic = {}, {}
    for arg in sys.argv[2:]:
        shape, rep = arg.split("=")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        acc = {}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("pcfm::", "")
            key = f"{name} grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}"
            d = acc.setdefault(key, {"launches": 0})
            d["launches"] += 1
            for col, short in COLS.items():
                if col in hdr and r[hdr.index(col)] not in ("", "n/a"):
                    v = float(r[hdr.index(col)].replace(",", ""))
                    u = units[hdr.index(col)]
                    if short.endswith("_MB"):
                        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if short == "time_us":
                        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                    d[short] = d.get(short, 0.0) + v
        for key, d in acc.items():
            n = d.pop("launches")
            out[f"{shape}: {key}"] = {"launches": n, **{k: round(v / n, 3) for k, v in d.items()}}
            base = key.split("<")[0].split(" ")[0]
            if base in MAIN and shape in SHAPES:
                c, r = SHAPES[shape]
                traffic[f"{MAIN[base]},{c},{r}"] = int((d["dram_read_MB"] + d["dram_write_MB"]) / n * 1e6)
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump(out, open(os.path.join(here, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    json.dump({"source": "ncu --set full --clock-control none on tests/perf_ops.py --only <shape> --no-variants --iters 1 "
                         "(B=8, N=20000); bytes of the MAIN kernel of each op, mean per launch; see "
                         f"profiles/{tag}_ncu_full_summary.json and profiles/summarize_ncu.py",
               "dram_bytes_per_launch": traffic}, open(os.path.join(here, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
