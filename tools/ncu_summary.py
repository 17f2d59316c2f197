"""Turn an .ncu-rep (ncu --set full ...) into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r1c_scan_full.ncu-rep [kernel-regex] > profiles/<name>.txt

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and prints, per
matching kernel launch, the metrics the roofline argument rests on: duration, DRAM bytes
read / written (the `traffic` of bench.py's roofline object), DRAM and tensor-pipe
utilisation, L2 hit rate, occupancy-limiting resources and the top stall reasons."""
import csv
import io
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if raw.returncode != 0:
        sys.exit(raw.stderr)
    rows = list(csv.reader(io.StringIO(raw.stdout)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --page raw summary (tools/ncu_summary.py)")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print(f"\n## {name}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
        for m in WANT:
            if m in col and r[col[m]] != "":
                print(f"{m}\t{r[col[m]]}\t{units[col[m]]}")
        stalls = []
        for h, i in col.items():
            mm = STALL.fullmatch(h)
            if mm and r[i] not in ("", "n/a"):
                try:
                    stalls.append((float(r[i]), mm.group(1)))
                except ValueError:
                    pass
        for v, n in sorted(stalls, reverse=True)[:5]:
            print(f"stall_{n}_per_issue\t{v:.3f}\twarps")
        try:
            rd = float(r[col["dram__bytes_read.sum"]])
            wr = float(r[col["dram__bytes_write.sum"]])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            tot = rd * scale[units[col["dram__bytes_read.sum"]]] + wr * scale[units[col["dram__bytes_write.sum"]]]
            print(f"traffic_bytes(read+write)\t{tot:.0f}\tbyte")
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
