"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.txt ["note"] [traffic-key samples-per-launch]
With the last two arguments the DRAM traffic per sample of the first kernel in the report is merged into profiles/ncu_traffic.json
under `traffic-key` (bench.py reads it for roofline.traffic)."""
import json
import os
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ", "lts__throughput", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "smsp__average_warps_issue_stalled", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum ", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum ",
        "smsp__sass_thread_inst_executed_op_dfma", "smsp__sass_thread_inst_executed_op_dmul", "smsp__sass_thread_inst_executed_op_dadd", "sm__cycles_elapsed.avg ",
        "sm__sass_inst_executed_op_shared", "derived__smsp__sass_thread_inst_executed_op_d"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {rep}", f"# {note}", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"kernel: {name}")
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k.strip()) if k.endswith(" ") else (k in h) for k in KEYS):
                lines.append(f"  {h:90s} {v:>18s} {u}")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print(out, len(lines), "lines")
    if len(sys.argv) > 5:
        key, nsamp = sys.argv[4], float(sys.argv[5])
        r = rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        path = os.path.join(os.path.dirname(out) or ".", "ncu_traffic.json")
        table = json.load(open(path)) if os.path.exists(path) else {}
        table[key] = {"dram_bytes_per_sample": tot / nsamp, "samples_per_launch": nsamp, "report": os.path.basename(rep),
                      "summary": os.path.basename(out), "kernel": r[hdr.index("Kernel Name")]}
        json.dump(table, open(path, "w"), indent=1, sort_keys=True)
        print(path, key, tot / nsamp, "B/sample")


if __name__ == "__main__":
    main()
