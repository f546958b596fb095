"""One SigLipLoss fwd + bwd at the headline shape (for ncu launch lists): python tools/siglip_case.py [n] [dim]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    print(bench.siglip_times(torch.device("cuda:0"), n=n, dim=dim, reps=2))
