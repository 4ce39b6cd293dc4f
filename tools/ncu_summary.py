"""Turn an ncu report into the tracked evidence under profiles/: the details page of one kernel as a markdown table and
its DRAM bytes per launch in profiles/ncu_traffic.json (read by bench.py into roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof_search_X.ncu-rep search_fs256_kernel r01fs c2 "<command that was profiled>"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, kernel, tag, workload = sys.argv[1:5]
    cmd = sys.argv[5] if len(sys.argv) > 5 else ""
    det = page(rep, "details")
    hdr = det[0]
    col = {h: i for i, h in enumerate(hdr)}
    rows = [r for r in det[1:] if kernel in r[col["Kernel Name"]]]
    first_id = rows[0][col["ID"]]
    rows = [r for r in rows if r[col["ID"]] == first_id]
    kname = rows[0][col["Kernel Name"]]
    lines = [f"# ncu --set full, {kname}, capture {tag}", "",
             f"Command: `{cmd}`" if cmd else "",
             f"(B200, workload {workload}; first captured launch; the `.ncu-rep` stays in gpurun_out/, this is its details page).", "",
             "| section | metric | value | unit |", "|---|---|---|---|"]
    for r in rows:
        if not r[col["Metric Name"]]:
            continue
        lines.append(f"| {r[col['Section Name']]} | {r[col['Metric Name']]} | {r[col['Metric Value']]} | {r[col['Metric Unit']]} |")
    raw = page(rep, "raw")
    rcol = {h: i for i, h in enumerate(raw[0])}
    units = raw[1]
    rr = [r for r in raw[2:] if kernel in r[rcol["Kernel Name"]]][0]

    def val(metric):
        v, u = float(rr[rcol[metric]].replace(",", "")), units[rcol[metric]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    extra = ["", "Raw metrics used elsewhere:", ""]
    for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
              "launch__grid_size", "launch__waves_per_multiprocessor",
              "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"):
        if m in rcol:
            extra.append(f"* `{m}` = {rr[rcol[m]]} {units[rcol[m]]}")
    short = kernel.split("<")[0]
    with open(os.path.join(ROOT, "profiles", f"{tag}_{short}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines + extra) + "\n")
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    t = json.load(open(tpath))
    def num(metric):
        return float(rr[rcol[metric]].replace(",", "")) if metric in rcol else None
    dur = num("gpu__time_duration.sum")
    if dur is not None:
        dur *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[rcol["gpu__time_duration.sum"]], 1.0)
    t[short] = {"dram_bytes_per_launch": int(dram), "capture": tag, "workload": workload, "n_gpus": 1,
                # what bench.py turns into roofline.executed_fma_pipe_frac (FMA-pipe busy cycles of the capture, rescaled to the
                # live kernel time) -- measured, not hand-counted
                "duration_us": dur,
                "pipe_fma_cycles_active_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                "inst_executed_pipe_fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "lsu_wavefronts_pct_of_peak": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                "registers_per_thread": num("launch__registers_per_thread")}
    json.dump(t, open(tpath, "w"), indent=2)
    print(short, "dram bytes per launch", int(dram))


if __name__ == "__main__":
    main()
