"""profiles/r2_gemm_pair_traffic.json from an `ncu --set full` capture of gemm_pair_kernel (tools/gpu_final.sh):
mean DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and tensor-pipe activity per launch of the plain-epilogue
LoRA kernel, stamped with the sha1 of the kernel source so that bench.py can refuse a capture of another source.
  python tools/make_traffic_json.py gpurun_out/<capture>.ncu-rep"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, data = rr[0], rr[2:]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
units = rr[1]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


rows = []
for r in data:
    name = r[col["Kernel Name"]]
    if "gemm_pair_kernel" not in name:
        continue
    b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    rows.append({"kernel": name.split("(")[0].replace("void vpt::", ""), "dram_bytes": b,
                 "duration": r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]],
                 "tensor_pipe_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])})
plain = [x for x in rows if x["kernel"].rstrip(">").endswith("0")]
src = os.path.join(ROOT, "vision_pt_b200", "csrc", "gemm_pair.cuh")
out = {"kernel": "gemm_pair_kernel<192, true, 0>", "launches_captured": len(plain),
       "dram_bytes_per_launch": sum(x["dram_bytes"] for x in plain) / max(1, len(plain)),
       "tensor_pipe_pct_mean": sum(x["tensor_pipe_pct"] for x in plain) / max(1, len(plain)),
       "kernel_source_sha1": hashlib.sha1(open(src, "rb").read()).hexdigest()[:12],
       "capture": os.path.basename(rep), "how": "ncu --set full --clock-control none, one eager JiT-B/16 batch-64 step (tools/profile_step.py); "
       "ncu serialises kernels, so inputs a producer just wrote are partly served by the 126 MB L2",
       "all_launches": rows}
path = os.path.join(ROOT, "profiles", "r2_gemm_pair_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(path, out["dram_bytes_per_launch"] / 1e6, "MB/launch over", len(plain), "launches; tensor pipe", out["tensor_pipe_pct_mean"])
