"""Turn an ncu report (.ncu-rep brought back from the GPU box under gpurun_out/) into the small text summaries
that are committed under profiles/:   python profiles/summarize_ncu.py gpurun_out/X.ncu-rep profiles/NAME

  NAME.metrics.csv   the raw-page metrics the roofline discussion uses (time, dram bytes, throughputs, issue, ...)
  NAME.stalls.txt    warp-stall breakdown, instruction totals and the hottest SASS instructions (source page)
"""
import collections
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared_atom.sum",
        "sm__cycles_elapsed.avg.per_second", "smsp__thread_inst_executed_per_inst_executed.ratio")


def page(rep, name):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout.splitlines()))


def main(rep, out):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    with open(out + ".metrics.csv", "w") as f:
        for row in raw[2:]:
            name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            f.write(f"# kernel,{name}\nmetric,unit,value\n")
            for i, h in enumerate(hdr):
                if h in KEEP:
                    f.write(f"{h},{units[i]},{row[i]}\n")
    src = page(rep, "source")
    h2, data = src[1], src[2:]
    ia, isrc, ismp = h2.index("Instructions Executed"), h2.index("Source"), h2.index("# Samples")
    cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for i in cols:
            tot[h2[i]] += int(r[i] or 0)
    s = sum(tot.values()) or 1
    with open(out + ".stalls.txt", "w") as f:
        f.write(src[0][1] + "\n")
        f.write("warp-stall samples (all): " + ", ".join(f"{k[6:]} {v / s * 100:.1f}%" for k, v in tot.most_common(10)) + "\n")
        f.write(f"SASS instructions: {len(data)}, warp-instructions executed: {sum(int(r[ia]) for r in data)}\n")
        f.write("hottest instructions (samples, executed, SASS, top stall reasons):\n")
        for k in sorted(range(len(data)), key=lambda k: -int(data[k][ismp]))[:16]:
            r = data[k]
            st = {h2[i][6:]: int(r[i] or 0) for i in cols if int(r[i] or 0) > 0}
            f.write(f"  {r[ismp]:>7} {r[ia]:>10}  {r[isrc].strip()[:80]:<80} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
