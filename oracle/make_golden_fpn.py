"""Generate tests/golden/fpn_*.npz by running THE REFERENCE'S OWN ``DiTWithFPN`` (build container only).

Run from the repo root:  ``python oracle/make_golden_fpn.py``

Imports ``layoutdit.modeling.dit_backbone.DiTWithFPN`` from /root/reference/src unmodified (the class at
R:src/layoutdit/modeling/dit_backbone.py:65-95, which instantiates torchvision's FeaturePyramidNetwork);
only the hub fetch at :26-31 is replaced by a local ``BeitConfig``, exactly as ``oracle/make_golden.py``
does.  Weights and inputs come from ``layoutdit_b200.synth`` (bit-stable numpy streams), so a fixture stores
the case description and the reference's outputs only.  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from layoutdit_b200.config import DiTConfig, dit_base  # noqa: E402
from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, synthetic_pages  # noqa: E402
from oracle import hf_reference  # noqa: E402

TINY = dict(hidden_size=128, num_hidden_layers=6, num_attention_heads=2, intermediate_size=256, image_size=64)

# name -> (config, backbone weight seed, fpn weight seed, stress, batch, H, W, input seed, store-full?)
CASES = {
    "fpn_tiny_native": (DiTConfig(**TINY), 11, 21, True, 2, 64, 64, 201, True),
    "fpn_tiny_odd":    (DiTConfig(**TINY), 12, 22, True, 1, 80, 112, 202, True),
    "fpn_tiny_w0":     (DiTConfig(**TINY), 15, 23, False, 1, 96, 64, 203, True),
    "fpn_base_224_w1": (dit_base(), 1, 24, True, 1, 224, 224, 1234, False),
}
SAMPLE_STRIDE = 97


def reference_model(cfg: DiTConfig):
    import transformers
    from layoutdit.modeling import dit_backbone as ref

    hf_cfg = hf_reference.hf_config(cfg.to_dict())
    orig = transformers.AutoConfig.from_pretrained
    ref.AutoConfig.from_pretrained = staticmethod(lambda *a, **k: hf_cfg)
    try:
        m = ref.DiTWithFPN(pretrained=False)
    finally:
        ref.AutoConfig.from_pretrained = orig
    return m.eval()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    index_path = os.path.join(out_dir, "index_fpn.json")
    index = {}
    for name, (cfg, wseed, fseed, stress, B, H, W, xseed, full) in CASES.items():
        sd = make_state_dict(cfg, wseed, stress)
        fsd = make_fpn_state_dict(cfg.hidden_size, 256, fseed, stress)
        x = synthetic_pages(B, H, W, xseed)
        ref = reference_model(cfg)
        ref.backbone.dit.load_state_dict(sd, strict=True)
        ref.fpn.load_state_dict(fsd, strict=True)
        with torch.no_grad():
            feats = ref(x)
        assert list(feats.keys()) == ["p2", "p3", "p4", "p5", "pool"]
        arrays = {}
        for k, v in feats.items():
            a = v.contiguous().numpy().astype(np.float32)
            arrays[k + "_shape"] = np.asarray(a.shape, dtype=np.int64)
            if full or k in ("p5", "pool"):
                arrays[k] = a
            else:
                arrays[k + "_samples"] = a.reshape(-1)[::SAMPLE_STRIDE].copy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        index[name] = dict(config=cfg.to_dict(), weight_seed=wseed, fpn_seed=fseed, stress=stress, batch=B, height=H,
                           width=W, input_seed=xseed, full=full, sample_stride=SAMPLE_STRIDE)
        print(name, {k: tuple(v.shape) for k, v in feats.items()})
    with open(index_path, "w") as f:
        json.dump(dict(generator="oracle/make_golden_fpn.py",
                       reference="/root/reference/src/layoutdit/modeling/dit_backbone.py (DiTWithFPN)",
                       torchvision=__import__("torchvision").__version__, torch=torch.__version__, cases=index), f, indent=1)


if __name__ == "__main__":
    main()
