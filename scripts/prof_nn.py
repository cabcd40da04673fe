"""One nn_argmax call per impl (for ncu)."""
import sys
import torch
from starst3r_b200 import match
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
A = torch.nn.functional.normalize(torch.randn(M, 24, generator=g), dim=-1).to(dev)
B = torch.nn.functional.normalize(torch.randn(262144, 24, generator=g), dim=-1).to(dev)
for _ in range(2):
    match.nn_argmax(A, B, impl=sys.argv[1])
torch.cuda.synchronize()
