"""GPU probe: BED decode kernel GB/s (bench.py's block alone)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hibag_b200 import api
api.set_device(0); torch.cuda.set_device(0)
print(json.dumps(bench.bench_bed_decode(api, torch, torch.device("cuda", 0)), indent=1))
