"""Seeded synthetic inputs shared by the fixtures, the tests and bench.py
(SURVEY.md section 8d).  TEST INFRASTRUCTURE ONLY."""
import hashlib

import numpy as np
import torch


def cloud(kind, B, N, seed):
    """'U' = uniform cube [-1,1)^3, 'S' = unit sphere surface, 'C' = a few tight clusters with
    duplicated points (adversarial for spatial pruning); fp32, CPU generator."""
    g = torch.Generator().manual_seed(seed)
    if kind == "U":
        return torch.rand(B, N, 3, generator=g) * 2 - 1
    if kind == "C":
        centres = torch.rand(B, 5, 3, generator=g) * 20 - 10
        which = torch.randint(0, 5, (B, N), generator=g)
        p = torch.gather(centres, 1, which.unsqueeze(-1).expand(-1, -1, 3)) + 0.01 * torch.randn(B, N, 3, generator=g)
        p[:, N // 2:] = p[:, : N - N // 2]  # every point of the first half appears twice
        return p.contiguous()
    p = torch.randn(B, N, 3, generator=g)
    return p / p.norm(dim=-1, keepdim=True)


def digest(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
