"""Pipe / issue / stall view of an `ncu --set full` report: python profiles/ncu_pipes.py report.ncu-rep
(per captured launch: issue-slot use, per-pipe utilisation and the warp-stall breakdown)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ("issue_active", "inst_executed_pipe_", "pipe_fma", "pipe_alu", "pipe_fmaheavy", "pipe_xu", "pipe_lsu", "pipe_uniform",
        "issue_stalled", "inst_executed.sum", "inst_executed.avg.per_cycle", "warp_issue_stalled", "cycles_active.avg",
        "thread_inst_executed_per_inst_executed", "gpu__time_duration")
ki = hdr.index("Kernel Name")
for r in data:
    print(r[ki].split("(")[0].replace("void ", "").strip())
    for i, h in enumerate(hdr):
        if any(k in h for k in keys):
            try:
                if float(r[i].replace(",", "")) == 0.0:
                    continue
            except ValueError:
                pass
            print(f"    {h:110s} {r[i]:>16s} {units[i]}")
    print()
