"""Shared builders for the parity tests: the same seeded model on the CUDA product path and in the
CPU oracle."""
import numpy as np
import torch

from miso_b200 import synth
from oracle import oracle as O

SMALL_BOUND = [[-2.0, 2.0], [-1.0, 1.0], [-2.0, 2.0]]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b||; absolute when the reference is ~0."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    nb = b.norm().item()
    d = (a - b).norm().item()
    return d / nb if nb > 1e-30 else d


def make_pair(bound=SMALL_BOUND, n_levels=2, fdim=4, base_cell=0.5, scale=5, std=0.1, seed=0, device="cuda",
              num_poses=4, dtype=torch.float32, fix=True):
    """(GridNet on `device`, OracleGridNet first-order, OracleGridNet second-order) with identical weights."""
    from miso_b200.models import GridNet
    cfg = synth.model_cfg(bound, n_levels=n_levels, feature_dim=fdim, base_cell_size=base_cell,
                          per_level_scale=scale, num_poses=num_poses, fix=fix)
    net = GridNet(cfg, device=device)
    g = torch.Generator().manual_seed(seed)
    feats = []
    for lvl in net.features:
        f = torch.randn(lvl.feature.shape, generator=g) * std
        feats.append(f)
        with torch.no_grad():
            lvl.feature.copy_(f.to(device))
    sd = synth.decoder_weights(n_levels * fdim, seed=seed)
    net.decoder.load_state_dict(sd)
    dec = O.make_decoder(n_levels * fdim)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in sd.items()})
    if not fix:
        for p in dec.parameters():
            p.requires_grad_(True)
    o1 = O.OracleGridNet(bound, feats, dec, second_order=False)
    o2 = O.OracleGridNet(bound, feats, dec, second_order=True)
    return net, o1, o2


def points_in(bound, n, seed=0, scale=1.1):
    """Uniform points in `scale` x the bound (so some fall outside and exercise zeros padding)."""
    g = torch.Generator().manual_seed(seed)
    b = torch.tensor(bound)
    c = (b[:, 0] + b[:, 1]) / 2
    h = (b[:, 1] - b[:, 0]) / 2
    return (torch.rand(n, 3, generator=g) * 2 - 1) * h * scale + c
