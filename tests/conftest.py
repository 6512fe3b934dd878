import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def golden_index():
    with open(os.path.join(GOLDEN_DIR, "index.json")) as f:
        return json.load(f)["cases"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def build_case(name):
    """(cfg, state_dict, x, golden arrays, meta) for a committed fixture."""
    from layoutdit_b200.config import DiTConfig
    from layoutdit_b200.synth import make_state_dict, synthetic_pages
    meta = golden_index()[name]
    cfg = DiTConfig(**meta["config"])
    sd = make_state_dict(cfg, meta["weight_seed"], meta["stress"])
    x = synthetic_pages(meta["batch"], meta["height"], meta["width"], meta["input_seed"])
    return cfg, sd, x, load_golden(name), meta


def compare_to_golden(feats, gold, meta, rel_fro, max_abs_rel, keys=("p2", "p3", "p4", "p5")):
    """Compare an OrderedDict of taps with a fixture.  Returns {tap: (rel_fro, max_abs/absmax)}."""
    import torch
    out = {}
    for k in keys:
        v = feats[k].detach().float().cpu().contiguous().numpy()
        assert tuple(v.shape) == tuple(gold[k + "_shape"]), (k, v.shape, gold[k + "_shape"])
        if k in gold:
            ref, got = gold[k], v
        else:
            ref, got = gold[k + "_samples"], v.reshape(-1)[:: meta["sample_stride"]]
        err = np.linalg.norm((got - ref).astype(np.float64)) / max(np.linalg.norm(ref.astype(np.float64)), 1e-30)
        mx = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)
        out[k] = (float(err), float(mx))
        assert err <= rel_fro, f"{k}: rel-Frobenius {err:.3e} > {rel_fro:.1e}"
        assert mx <= max_abs_rel, f"{k}: max-abs/absmax {mx:.3e} > {max_abs_rel:.1e}"
    return out


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def fpn_golden_index():
    with open(os.path.join(GOLDEN_DIR, "index_fpn.json")) as f:
        return json.load(f)["cases"]


def build_fpn_case(name):
    """(cfg, backbone state dict, fpn state dict, x, golden arrays, meta) for a committed DiTWithFPN fixture."""
    from layoutdit_b200.config import DiTConfig
    from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, synthetic_pages
    meta = fpn_golden_index()[name]
    cfg = DiTConfig(**meta["config"])
    sd = make_state_dict(cfg, meta["weight_seed"], meta["stress"])
    fsd = make_fpn_state_dict(cfg.hidden_size, 256, meta["fpn_seed"], meta["stress"])
    x = synthetic_pages(meta["batch"], meta["height"], meta["width"], meta["input_seed"])
    return cfg, sd, fsd, x, load_golden(name), meta


def transform_golden_index():
    with open(os.path.join(GOLDEN_DIR, "index_transform.json")) as f:
        return json.load(f)["cases"]


def build_transform_case(name):
    """(raw pages, golden samples, shape, meta) for a committed input-transform fixture."""
    from layoutdit_b200.synth import raw_pages
    meta = transform_golden_index()[name]
    gold = load_golden(name)
    return raw_pages([tuple(s) for s in meta["sizes"]], meta["seed"]), gold["samples"], tuple(gold["shape"]), meta
