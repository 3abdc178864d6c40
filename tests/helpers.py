"""Shared test helpers: build OUR tiny model with the fixture's deterministic weights."""
import numpy as np
import torch

import cwfa_b200
from oracle.weights import deterministic_fill, seeded_randn


def build_tiny_model(fx, device="cpu"):
    """The tiny config of BASELINE.json configs[0] (D=16,S=64,3 steps) with the golden fixture's
    weights: deterministic_fill seeds from fx['config'], permutations + PermuteDim axes from fx."""
    cfg = fx["config"]
    D, S, MAX = cfg["D"], cfg["S"], cfg["MAX"]
    np.random.seed(0)
    torch.manual_seed(0)
    m = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=MAX)
    seeds = cfg["seeds"]
    for n in range(MAX - 1):
        sd = deterministic_fill(m.conv_inn[n].state_dict(), seeds["inn"] + n)
        sd.update({k: v.clone() for k, v in fx["perms"][n].items()})
        m.conv_inn[n].load_state_dict(sd)
        for node in fx["specs"][n]["nodes"]:
            if node["type"] == "perm_dim":
                m.conv_inn[n].module_list[node["idx"]].dims_to_permute = [1, node["axis"]]
        m.cond_nets[n].load_state_dict(deterministic_fill(m.cond_nets[n].state_dict(), seeds["cond"] + n))
    m.cond_nets[-1].load_state_dict(deterministic_fill(m.cond_nets[-1].state_dict(), seeds["lrnn"]))
    return m.to(device)


def tiny_inputs(fx):
    cfg = fx["config"]
    D, S, MAX = cfg["D"], cfg["S"], cfg["MAX"]
    views = seeded_randn((1, 29, S, S), cfg["seeds"]["views"])
    mean_vols = [seeded_randn((1, D // 2 ** (n + 1), S, S), cfg["seeds"]["mean"] + n, 0.1) for n in range(MAX - 1)]
    mean_vols.append(seeded_randn((1, D // 2 ** (MAX - 1), S, S), cfg["seeds"]["mean"] + MAX - 1, 0.1))
    return views, mean_vols
