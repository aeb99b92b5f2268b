"""Filters the raw page of an ncu report down to the metrics quoted in DESIGN.md / read by bench.py.

    python tools/ncu_metrics.py gpurun_out/x.ncu-rep "header comment" > profiles/rNN_....metrics.csv
One block per profiled launch: name, block/grid size, then `metric,unit,value` rows."""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    rep = sys.argv[1]
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("Kernel Name,,%s" % r[col["Kernel Name"]].replace(",", ";"))
        print("Block Size,,%s" % r[col["Block Size"]].replace(",", ";"))
        print("Grid Size,,%s" % r[col["Grid Size"]].replace(",", ";"))
        for h in hdr:
            if h in KEEP or any(h.endswith(k) for k in KEEP):
                print("%s,%s,%s" % (h, units[col[h]], r[col[h]]))


if __name__ == "__main__":
    main()
