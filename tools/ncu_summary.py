#!/usr/bin/env python
"""Turn an ncu report (ncu --set full --import-source on ...) into the text summary committed under profiles/:
the headline metrics of the details page, the stall-reason totals and the top stalled instructions of the source page.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt      (runs here: ncu -i needs no GPU)
"""
import csv
import io
import subprocess
import sys

KEEP = ("Duration", "SM Frequency", "Elapsed Cycles", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Achieved Active Warps",
        "Issue Slots Busy", "Executed Ipc", "No Eligible", "Eligible Warps", "Active Warps Per Scheduler", "Warp Cycles Per Issued", "L1/TEX Hit",
        "L2 Hit Rate", "Mem Busy", "DRAM Throughput", "Memory Throughput", "Grid Size", "Block Size", "Shared Memory", "Waves Per SM",
        "Executed Instructions", "Issued Instructions", "fused and", "highest-utilized pipeline", "FP32 peak")


def main():
    rep = sys.argv[1]
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    print("# " + rep)
    for line in det.splitlines():
        if any(k in line for k in KEEP) or line.strip().startswith(("void ", "lfb::")) or "Section:" in line and "Speed Of Light" in line:
            print(line.rstrip())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) >= 3:
        hdr, vals = rows[0], rows[2]
        print("\n# raw counters (per launch)")
        for h, v in zip(hdr, vals):
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
                     "sm__inst_executed_pipe_xu.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
                     "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "launch__registers_per_thread", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
                     "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"):
                print(f"{h} = {v}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hi is None:
        return
    hdr, data = rows[hi], rows[hi + 1:]
    idx = {h: i for i, h in enumerate(hdr)}
    num = lambda r, c: int(r[idx[c]] or 0)  # noqa: E731
    total = sum(num(r, "# Samples") for r in data)
    print(f"\n# source page: {len(data)} SASS instructions, {sum(num(r, 'Instructions Executed') for r in data)} warp-instructions executed, {total} stall samples")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: sum(num(r, h) for r in data) for h in stalls}
    print("stall samples by reason: " + ", ".join(f"{h[6:]} {v} ({100.0 * v / max(total, 1):.1f} %)" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    print("\n# top 25 instructions by stall samples (index, SASS, samples, dominant reasons, executions)")
    for i in sorted(sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:25]):
        r = data[i]
        why = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{i:5d}  {r[idx['Source']].strip()[:64]:64s} {num(r, '# Samples'):5d}  " + " ".join(f"{n}:{v}" for v, n in why if v) + f"  exec {num(r, 'Instructions Executed')}")


if __name__ == "__main__":
    main()
