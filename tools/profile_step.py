"""One eager training step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import train as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="JiT-B/16")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()

net = T.build_jit_qlora(args.model, device="cuda", seed=42)
step = T.JiTQLoRATrainStep(net, args.batch, args.res, args.res, use_graph=False)
host = T.synthetic_batch(args.batch, args.res, args.res)
step.image.copy_(host[0]); step.class_ids.copy_(host[1]); step.attention_mask.copy_(host[2])
for _ in range(args.warmup):
    step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(step.loss))
