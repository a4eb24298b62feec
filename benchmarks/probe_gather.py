"""the all-distinct resample probe of bench.py on its own (for ncu: `-k regex:k_gather`): near-uniform weights, reference
layout (1540 B per particle), L2 flushed between iterations -- algorithmic bytes == DRAM traffic."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print(json.dumps(bench.resample_probe(1000, 1024, torch.device("cuda", 0), iters=3)))
