"""One K23 launch (T = 4096, q_proj shape) for ncu.  python scripts/lora_linear_once.py [T R C]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
T, R, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 4096, 4096)
dt = torch.bfloat16
W = (torch.randn(R, C, device="cuda") * 0.02).to(dt)
A = torch.randn(8, C, device="cuda") * 0.1
B = torch.randn(R, 8, device="cuda") * 0.1
mask = torch.rand(R, C, device="cuda") < 0.5
x = (torch.randn(T, C, device="cuda") * 0.5).to(dt)
for _ in range(3):
    y = native.sparselora_linear_forward(x, W, A, B, 2.0, mask, True)
torch.cuda.synchronize()
print(float(y.float().abs().max()))
