"""Compact per-kernel summary of an .ncu-rep (ncu --set full capture): duration, DRAM bytes / throughput, tensor-pipe and
issue activity, occupancy, registers, top warp-stall reasons, instruction count.   python tools/ncu_brief.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"{d['Kernel Name']}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}")
        seen = set()
        for k, label in KEYS:
            if k in d and d[k] not in ("", None) and label not in seen:
                seen.add(label)
                print(f"    {label:34s} {d[k]} {u.get(k, '')}")
        st = []
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(v.replace(",", "")), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("    warp stalls (cycles per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
        print()


if __name__ == "__main__":
    main()
