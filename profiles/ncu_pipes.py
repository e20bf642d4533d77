"""Pipe / issue / stall view of an `ncu --set full` report: python profiles/ncu_pipes.py report.ncu-rep
(per captured launch: issue-slot use, per-pipe utilisation and the warp-stall breakdown)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "_per_issue_active.ratio", "gpu__time_duration.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
ki = hdr.index("Kernel Name")
for r in data:
    print(r[ki].split("(")[0].replace("void ", "").strip())
    for i, h in enumerate(hdr):
        if any(h == k or (k.startswith('_') and h.endswith(k)) for k in keys):
            try:
                if float(r[i].replace(",", "")) == 0.0:
                    continue
            except ValueError:
                pass
            print(f"    {h:110s} {r[i]:>16s} {units[i]}")
    print()
