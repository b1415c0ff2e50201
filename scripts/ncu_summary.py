"""Compact text summary of one `ncu --set full` report (for profiles/): duration, tensor-pipe activity, issue-slot
use, DRAM traffic, shared-memory conflicts, top stall reasons, and the source lines that issue the most instructions.
usage: python scripts/ncu_summary.py report.ncu-rep [n_lines]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum",
    # the L1 data pipe is shared by the LSU (global / shared accesses of the threads) and the tensor core's operand reads
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
print("== %s" % rep.split("/")[-1])
for k in want:
    if k in m:
        print("%-82s %s %s" % (k, m[k][0], m[k][1]))
print("-- stall reasons (warps stalled per issue-active cycle)")
st = [(float(v[0]), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for v, k in sorted(st, reverse=True)[:8]:
    print("  %-40s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = list(csv.reader(io.StringIO(ncu("--page", "source", "--print-source", "cuda,sass", "--csv"))))
cur, agg = None, []
for r in src:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("", "Line No", "Function Name"):
        continue
    try:
        agg.append((int(r[7]), int(r[4]), cur, int(r[0]), r[1].strip()[:88]))
    except ValueError:
        pass
tot = sum(a[0] for a in agg) or 1
print("-- source lines by warp instructions executed (total %d)" % tot)
for a in sorted(agg, reverse=True)[:nlines]:
    print("  %5.1f%%  samples=%5d  %s:%d  %s" % (100.0 * a[0] / tot, a[1], a[2], a[3], a[4]))
