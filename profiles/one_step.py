#!/usr/bin/env python
"""One C2 update step between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --set full --clock-control none --import-source on \\
      -o gpurun_out/r2_step python profiles/one_step.py [workload]
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
      --log-file gpurun_out/r2_launches.csv python profiles/one_step.py [workload]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.Workload(name, 1, 0, dev, 1, None)
for _ in range(4):
    wl.step(False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
wl.step(False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step of", name)
