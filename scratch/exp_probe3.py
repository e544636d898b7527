import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
r = bench.exp_config(dev, 4096, 10, 3, cpu_sample=False, pipelined=True)
print(json.dumps({k: r[k] for k in ("ms_per_step", "graphs_per_s", "design_overlapped_and_step_captured", "spectral_design")}, indent=1)[:1500])
