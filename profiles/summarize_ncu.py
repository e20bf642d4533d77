"""Summarise an `ncu --set full` report: python profiles/summarize_ncu.py report.ncu-rep > out.txt
(runs `ncu -i ... --page raw --csv`; one block per captured launch)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    ("time", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram throughput % of peak", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active % (realtime)", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tensor-memory (operand fetch) cycles active %", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tmem pipe inst % ", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"),
    ("issue slots used %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fma pipe inst %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu pipe inst %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("local-memory (spill) load requests", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("sm throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsu shared wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("shared bank conflicts (ld)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("shared bank conflicts (st)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("registers / thread", "launch__registers_per_thread"),
    ("blocks / SM limit (smem)", "launch__occupancy_limit_shared_mem"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
]
ki = hdr.index("Kernel Name")
for r in data:
    print(r[ki].split("(")[0].replace("void ", "").strip())
    for label, name in want:
        if name in hdr:
            i = hdr.index(name)
            print(f"    {label:46s} {r[i]:>16s} {units[i]}")
    try:
        t = float(r[hdr.index("gpu__time_duration.sum")]); tu = units[hdr.index("gpu__time_duration.sum")]
        rd = float(r[hdr.index("dram__bytes_read.sum")]); ru = units[hdr.index("dram__bytes_read.sum")]
        wr = float(r[hdr.index("dram__bytes_write.sum")]); wu = units[hdr.index("dram__bytes_write.sum")]
        sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ts = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        print(f"    {'dram GB/s (read + write) / time':46s} {(rd * sc[ru] + wr * sc[wu]) / (t * ts[tu]) / 1e9:16.1f} GB/s")
    except Exception:
        pass
    print()
