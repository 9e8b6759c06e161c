"""Config 2 (full InstructBLIP-FlanT5-XL, Wanda 2:4) once, for ncu launch lists: python scripts/full_model_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
import bench
from vlmc import native

native.load()
dev = torch.device("cuda", 0)
r = bench.full_model_wanda_nm(torch, native, dev, reps=1)
print({k: v for k, v in r.items() if k != "workload"})
