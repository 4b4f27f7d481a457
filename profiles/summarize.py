"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_rNN.csv  > profiles/launches_rNN.txt
  python profiles/summarize.py raw      gpurun_out/kernels_rNN.ncu-rep > profiles/kernels_rNN.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        val = float(r[vi].replace(",", ""))
        val = val / 1e3 if r[ui] == "ns" else val * 1e3 if r[ui] == "ms" else val
        a = agg.setdefault(re.sub(r"\(.*", "", r[ki])[:90], [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# %d launches, %.1f us total" % (sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s n=%4d total=%10.1f us avg=%9.1f us share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


def raw(path):
    """``path``: an .ncu-rep, or the CSV that ``ncu -i X.ncu-rep --page raw --csv`` printed (large reports are converted on
    the GPU box and only the CSV is brought back)."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    seen = collections.Counter()
    every = "--all" in sys.argv
    print("# ncu --set full --clock-control none; %s" % ("every captured launch" if every else "first instance of each kernel"))
    for r in data:
        name = re.sub(r"\(CUtensorMap.*|\(const.*|\(float.*", "", r[ki])[:120].replace("(int)", "").replace("(bool)", "")
        seen[name] += 1
        if seen[name] > 1 and not every:
            continue
        print("==== " + name)
        for k in KEYS:
            if k in hdr:
                print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
