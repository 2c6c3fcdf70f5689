"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
    python profiles/summarize.py raw gpurun_out/prof.ncu-rep > profiles/rNN_kernel_ncu.txt

The tensor-pipe metrics are not part of `--set full` on sm_100: capture with
    ncu --set full --metrics $(python profiles/summarize.py tensor-metrics) ...
"""
TENSOR_METRICS = ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,"
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,"
                  "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,"
                  "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum.pct_of_peak_sustained_elapsed,"
                  "sm__inst_executed_pipe_tc.sum")
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg", "sm__inst_executed_pipe_tc.sum",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time over {sum(v[0] for v in agg.values())} launches (ncu gpu__time_duration.sum, "
          "cold-cache + serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} n={v[0]:5d} total_ms={v[1]:10.3f} share={100 * v[1] / tot:5.1f}% avg_ms={v[1] / v[0]:8.4f}")
    print(f"total_ms={tot:.3f}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[name_col][:100])
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"{h} [{units[i]}]: " + " | ".join(r[i] for r in rows[2:]))


if __name__ == "__main__":
    if sys.argv[1] == "tensor-metrics":
        print(TENSOR_METRICS)
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
