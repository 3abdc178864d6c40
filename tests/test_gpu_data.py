"""GPU: the rows either side of the path: lenslet-view extraction, the reference's GT-pyramid helper,
checkpoint round trip in the reference's file format."""
import os

import pytest
import torch

from conftest import max_abs, rel_l2
from helpers import build_tiny_model, tiny_inputs
from oracle import cwfa_oracle as O
from oracle.weights import seeded_randn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_extract_views_vs_reference_golden(golden_modules):
    from cwfa_b200.data import extract_views
    img = seeded_randn((2, 1, 90, 100), 30)
    coords = golden_modules["extract_views/coords"].tolist()
    out = extract_views(img.to(DEV), coords, [64, 64])
    assert torch.equal(out.cpu(), golden_modules["extract_views/out"])            # bit-exact gather
    out16 = extract_views(img.half().to(DEV), coords, [64, 64])
    assert torch.equal(out16.cpu(), O.extract_views(img.half(), coords, [64, 64]).float())
    mean, std = 0.25, 1.75
    outn = extract_views(img.to(DEV), coords, [64, 64], mean, std)
    assert torch.equal(outn.cpu(), (golden_modules["extract_views/out"] - mean) / std)


def test_extract_views_full_size_property():
    """2160^2 sensor image -> 29 x 512 x 512 views: every view equals the plain slice of the image."""
    from cwfa_b200.data import extract_views
    g = torch.Generator().manual_seed(0)
    img = torch.rand((1, 1, 2160, 2160), generator=g)
    coords = [[300 + 370 * (i // 6), 280 + 330 * (i % 6)] for i in range(29)]
    out = extract_views(img.to(DEV), coords, [512, 512]).cpu()
    assert torch.equal(out, O.extract_views(img, coords, [512, 512]))


@pytest.mark.parametrize("side", [128, 256])
def test_extract_views_wide_views_clipped_windows(side):
    """Wide views (the 16-byte store path): windows clipped at every border, an image width that is not a multiple of 4 (the
    source misalignment then changes from row to row), batch 2, fp32 / fp16 images, with and without normalisation -- bit-exact
    against the oracle's restatement of XLFMDataset.extract_views."""
    from cwfa_b200.data import extract_views
    Hi, Wi = 700, 1002
    img = seeded_randn((2, 1, Hi, Wi), 31)
    h = side // 2
    coords = [[10, 17], [Hi - 5, Wi - 9], [h, h], [Hi - h, Wi - h], [h + 1, Wi - 3], [Hi - 2, h + 3], [350, 501], [351, 502], [352, 503]]
    ref = O.extract_views(img, coords, [side, side])
    assert torch.equal(extract_views(img.to(DEV), coords, [side, side]).cpu(), ref)
    mean, std = 0.25, 1.75
    assert torch.equal(extract_views(img.to(DEV), coords, [side, side], mean, std).cpu(), (ref - mean) / std)
    assert torch.equal(extract_views(img.half().to(DEV), coords, [side, side]).cpu(), O.extract_views(img.half(), coords, [side, side]).float())


def test_evaluate_inn_forward_vs_reference_golden(golden_tiny):
    model = build_tiny_model(golden_tiny, DEV)
    cfg = golden_tiny["config"]
    gt = seeded_randn((2, cfg["D"], cfg["S"], cfg["S"]), 4).to(DEV)
    losses, cache, prior, ljs = model.evaluate_INN_forward(gt, fix_empty_depths=False)
    assert rel_l2(torch.stack(losses), golden_tiny["evalfwd/losses"]) < 1e-4
    assert rel_l2(torch.stack(prior), golden_tiny["evalfwd/prior"]) < 1e-4
    assert rel_l2(torch.stack(ljs), golden_tiny["evalfwd/logjac"]) < 1e-4
    assert rel_l2(cache[cfg["MAX"] - 1], golden_tiny["evalfwd/gt_last"]) < 1e-6


def test_checkpoint_round_trip_reference_format(golden_tiny, tmp_path):
    from cwfa_b200 import CWFAModel
    from cwfa_b200.data import load_checkpoints, load_INN_steps, save_checkpoints
    model = build_tiny_model(golden_tiny, DEV)
    stats = (0.1, 1.2, 0.1, 1.2, 0.3, 2.0)
    save_checkpoints(model, str(tmp_path), epoch=7, training_statistics=stats)
    save_checkpoints(model, str(tmp_path), epoch=3, training_statistics=stats)
    found = load_INN_steps(str(tmp_path))
    assert sorted(found) == [1, 2, 3] and all(v[0] == 7 for v in found.values())       # highest epoch wins
    data = torch.load(found[1][1], weights_only=False)
    assert {"epoch", "args", "INN_state_dict", "condition_state_dict", "optimizer_state_dict", "training_statistics"} <= set(data)
    cfg = golden_tiny["config"]
    fresh = CWFAModel(n_depths=cfg["D"], volume_side_size=cfg["S"], INN_max_down_steps=cfg["MAX"], seed=123).to(DEV)
    assert load_checkpoints(fresh, str(tmp_path)) == stats
    views, mean_vols = tiny_inputs(golden_tiny)
    mv = [t.to(DEV) for t in mean_vols]
    assert torch.equal(fresh.reconstruct(views.to(DEV), mv), model.reconstruct(views.to(DEV), mv))
