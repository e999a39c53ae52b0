"""Key metrics of an .ncu-rep (ncu -i ... --page details --csv) as a small text table.  python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEEP = ("Duration", "SM Frequency", "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput", "Compute (SM) Throughput",
        "Registers Per Thread", "Dynamic Shared Memory Per Block", "Theoretical Occupancy", "Achieved Occupancy", "L2 Hit Rate",
        "Executed Ipc Active", "Cluster Size", "Grid Size", "Block Size", "Mem Busy", "Max Bandwidth")
RAW = ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "sm__cycles_elapsed.avg")
path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.DictReader(out.splitlines()))
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
cur = None
k = -1
for r in rows:
    if r["ID"] != cur:
        cur = r["ID"]
        k += 1
        print(f"== launch {cur}: {r['Kernel Name'][:90]}")
        if k < len(data):
            for m in RAW:
                if m in hdr:
                    i = hdr.index(m)
                    print(f"  {m:<66} {units[i]:<12} {data[k][i]}")
    if r["Metric Name"] in KEEP:
        print(f"  {r['Metric Name']:<66} {r['Metric Unit']:<12} {r['Metric Value']}")
