"""Print the metrics DESIGN.md / profiles/ quote from an .ncu-rep (run where ncu is installed, no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex-for-source-page]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for r in rows[2:]:
    print("----", r[idx["Kernel Name"]][:100])
    for w in WANT:
        if w in idx:
            print(f"  {w} = {r[idx[w]]} {units[idx[w]]}")
    for h in hdr:
        if "smsp__average_warp" in h and "issue_stalled" in h and "not_issued" not in h:
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.5:
                print("   stall", h.split("issue_stalled_")[1].split("_per")[0], round(v, 2))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [hdr.index(c) for c in ("stall_long_sb", "stall_lg", "stall_mio", "stall_short_sb", "stall_wait")]
    body = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            break
        body.append(r)
    tot = sum(int(r[i_s] or 0) for r in body)
    print("source page: %d instructions, %d samples" % (len(body), tot))
    for k, r in enumerate(body):
        s = int(r[i_s] or 0)
        if s > tot * 0.01:
            print(f"  {k:4d} {r[i_src][:70]:70s} samples {s:6d} exec {r[i_ex]:>8s} long_sb/lg/mio/short/wait",
                  [r[c] for c in cols])
