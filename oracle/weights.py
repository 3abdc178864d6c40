"""Deterministic, machine-independent weight fill for parity tests (TEST INFRASTRUCTURE).

The reference's random init cannot travel to the GPU box (the reference itself is absent
there) and full state_dicts are too large to commit (the LRNN alone is ~100 MB), so both
the golden generator (tests/golden/make_golden.py, which runs the UNMODIFIED reference)
and the parity tests overwrite every floating-point entry of a ``state_dict`` with values
drawn from a CPU ``torch.Generator`` seeded by (seed, crc32(key)).  Integer entries
(permutations, num_batches_tracked) are left alone; permutations are stored in the
fixtures.  Scales follow the layer type so activations stay O(1) through the trunk.
"""
import math
import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def deterministic_fill(sd: dict, seed: int = 0) -> dict:
    """Returns a new dict with the same keys/shapes/dtypes, floats refilled."""
    out = {}
    for key in sd:
        v = sd[key]
        if not torch.is_floating_point(v):
            out[key] = v.clone()
            continue
        g = _gen(seed, key)
        leaf = key.split(".")[-1]
        is_norm_like = v.dim() == 1 or (v.dim() == 3 and leaf in ("weight", "bias") and ".m.1." in key)
        if leaf == "running_var":
            t = _uniform(v.shape, 0.5, 1.5, g)
        elif leaf == "running_mean":
            t = _uniform(v.shape, -0.2, 0.2, g)
        elif v.numel() == 1 and leaf == "weight":            # PReLU slope
            t = _uniform(v.shape, 0.1, 0.4, g)
        elif leaf == "bias":
            t = _uniform(v.shape, -0.1, 0.1, g)
        elif is_norm_like and leaf == "weight":               # BN / LayerNorm gains
            t = _uniform(v.shape, 0.5, 1.5, g)
        elif leaf == "weight":                                # conv / convT / conv1d / conv3d
            fan_in = v[0].numel() if v.dim() > 1 else v.numel()
            if "up.weight" in key:                            # ConvTranspose2d: (in, out, kh, kw)
                fan_in = v.shape[0] * v.shape[2] * v.shape[3] / 4.0
            b = math.sqrt(3.0 / max(fan_in, 1))
            t = _uniform(v.shape, -b, b, g)
        else:
            t = _uniform(v.shape, -0.1, 0.1, g)
        out[key] = t.to(v.dtype)
    return out


def seeded_randn(shape, seed: int, scale: float = 1.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float32) * scale
