"""Yardstick for the bf16 / fp16 tolerances at the FULL config: the CPU oracle re-run with every convolution operand rounded to
the half type (tests/helpers.py:half_operand_emulation -- what a tensor-core path with exact fp32 accumulation computes),
compared with the probe of the unmodified reference (tests/golden/r2_full.pt).  Its per-level error is the error the ARITHMETIC
TYPE causes on this network; the GPU tests allow the kernels max(SURVEY tolerance, 1.5 x this).

    python tests/golden/make_emulation_yardstick.py        ->  tests/golden/r2_full_emu.pt   (needs no reference, ~5 min of CPU)

The conditioning net's Conv3d pair is outside the emulation (the engine evaluates it as two banded 2-D convolutions with a
half-precision hidden tensor), so the yardstick is, if anything, slightly optimistic."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import build_full_model, full_inputs, half_operand_emulation, probe_errors   # noqa: E402
from oracle import cwfa_oracle as O                                                       # noqa: E402
from oracle.weights import seeded_randn                                                   # noqa: E402


def main():
    fx = torch.load(os.path.join(HERE, "r2_full.pt"), weights_only=False)
    om = build_full_model(fx).export_for_oracle()
    views, mvs = full_inputs(fx)
    L = len(om["levels"])
    seeds = fx["config"]["seeds"]
    out = {}
    for kind in ("bf16", "fp16"):
        with torch.no_grad(), half_operand_emulation(kind):
            outs, jacs = O.reconstruct(om, views, mvs, bn_mode="batch", return_all=True)
            x, vB = seeded_randn((1, 96, 512, 512), seeds["fwd_x"]), seeded_randn((1, 29, 512, 512), seeds["fwd_views"])
            res = O.forward_nll(om, x, vB, mvs)
        for n in range(L + 1):
            key = "inv/lrnn" if n == L else f"inv/vol{n}"
            e = probe_errors(outs[n], fx[key])
            out[f"{kind}/{key}"] = e[1]
            if n < L:
                r = float(fx[f"inv/jac{n}"][0])
                out[f"{kind}/inv/jac{n}"] = abs(float(jacs[n][0]) - r) / abs(r)
        for n, r in enumerate(res):
            out[f"{kind}/fwd/z{n}"] = probe_errors(r["z"], fx[f"fwd/0/z{n}"])[1]
            rj = float(fx[f"fwd/0/jac{n}"][0])
            out[f"{kind}/fwd/jac{n}"] = abs(float(r["logdet"][0]) - rj) / abs(rj)
        print(kind, {k: f"{v:.2e}" for k, v in out.items() if k.startswith(kind)}, flush=True)
    torch.save(out, os.path.join(HERE, "r2_full_emu.pt"))


if __name__ == "__main__":
    main()
