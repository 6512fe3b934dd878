"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF (build container only).

Run from the repo root:  ``python oracle/make_golden.py``

Imports ``layoutdit.modeling.dit_backbone.DiTBackbone`` from /root/reference/src
unmodified.  The only thing replaced is the network fetch at
R:src/layoutdit/modeling/dit_backbone.py:26-31: ``AutoConfig.from_pretrained`` returns a
local ``BeitConfig`` and ``AutoModel.from_pretrained`` is never reached
(``pretrained=False``).  Weights and inputs come from ``layoutdit_b200.synth`` (numpy
PCG64, bit-stable), so a fixture stores only the case description and the reference's
outputs.  /root/reference does not exist on the GPU box: nothing at test time reads it.
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

from layoutdit_b200.config import DiTConfig, dit_base, dit_large  # noqa: E402
from layoutdit_b200.synth import make_state_dict, synthetic_pages  # noqa: E402
from oracle import hf_reference  # noqa: E402

TINY = dict(hidden_size=128, num_hidden_layers=6, num_attention_heads=2, intermediate_size=256, image_size=64)

# name -> (config, weight seed, stress, batch, H, W, input seed, store-full?)
CASES = {
    "tiny_abs_native":   (DiTConfig(**TINY), 11, True, 2, 64, 64, 101, True),
    "tiny_abs_interp":   (DiTConfig(**TINY), 11, True, 1, 96, 64, 102, True),
    "tiny_abs_odd":      (DiTConfig(**TINY), 12, True, 1, 80, 112, 103, True),
    "tiny_relpos":       (DiTConfig(**TINY, use_absolute_position_embeddings=False,
                                    use_relative_position_bias=True), 13, True, 2, 64, 64, 104, True),
    "tiny_relpos_interp": (DiTConfig(**TINY, use_absolute_position_embeddings=False,
                                     use_relative_position_bias=True), 13, True, 1, 96, 64, 105, True),
    "tiny_shared_relpos": (DiTConfig(**TINY, use_absolute_position_embeddings=False,
                                     use_shared_relative_position_bias=True), 14, True, 1, 64, 96, 106, True),
    "tiny_w0":           (DiTConfig(**TINY), 15, False, 1, 64, 64, 107, True),
    "base_224_w0":       (dit_base(), 0, False, 1, 224, 224, 1234, False),
    "base_224_w1":       (dit_base(), 1, True, 1, 224, 224, 1234, False),
    "base_320x224_w1":   (dit_base(), 1, True, 1, 320, 224, 1235, False),
    "large_224_w1":      (dit_large(), 2, True, 1, 224, 224, 1236, False),
    # relative-position bias at the real size (N = 197, 12 heads, window 27 x 27 + 3): per-layer and shared tables
    "base_224_relpos_w1": (dit_base(use_absolute_position_embeddings=False, use_relative_position_bias=True),
                           3, True, 1, 224, 224, 1237, False),
    "base_224_shared_relpos_w1": (dit_base(use_absolute_position_embeddings=False, use_shared_relative_position_bias=True),
                                  4, True, 1, 224, 224, 1238, False),
    # BASELINE config 3 geometry (N = 1025, position table interpolated 14 x 14 -> 32 x 32)
    "base_512_w1":       (dit_base(), 1, True, 1, 512, 512, 1239, False),
}

SAMPLE_STRIDE = 97  # prime; flattened outputs are sampled every SAMPLE_STRIDE elements


def reference_backbone(cfg: DiTConfig):
    """Instantiate the reference's own class with the hub fetch replaced."""
    import transformers
    from layoutdit.modeling import dit_backbone as ref

    hf_cfg = hf_reference.hf_config(cfg.to_dict())
    orig = transformers.AutoConfig.from_pretrained
    ref.AutoConfig.from_pretrained = staticmethod(lambda *a, **k: hf_cfg)
    try:
        m = ref.DiTBackbone(pretrained=False)
    finally:
        ref.AutoConfig.from_pretrained = orig
    return m.eval()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)
    index = {}
    only = set(sys.argv[1:])          # optional: regenerate just these cases (the others keep their index entries)
    index_path = os.path.join(out_dir, "index.json")
    if only and os.path.exists(index_path):
        index = json.load(open(index_path))["cases"]
    for name, (cfg, wseed, stress, B, H, W, xseed, full) in CASES.items():
        if only and name not in only:
            continue
        sd = make_state_dict(cfg, wseed, stress)
        x = synthetic_pages(B, H, W, xseed)
        ref = reference_backbone(cfg)
        missing = ref.dit.load_state_dict(sd, strict=True)
        with torch.no_grad():
            feats = ref(x)
            twin = hf_reference.build(cfg.to_dict(), sd)(x)
        arrays = {}
        for k, v in feats.items():
            assert torch.equal(v, twin[k]), f"hf_reference twin differs from the reference on {name}/{k}"
            a = v.contiguous().numpy().astype(np.float32)
            arrays[k + "_shape"] = np.asarray(a.shape, dtype=np.int64)
            arrays[k + "_sum"] = np.asarray(a.astype(np.float64).sum())
            arrays[k + "_sumsq"] = np.asarray((a.astype(np.float64) ** 2).sum())
            if full:
                arrays[k] = a
            else:
                arrays[k + "_samples"] = a.reshape(-1)[::SAMPLE_STRIDE].copy()
        if not full:
            arrays["p5"] = feats["p5"].contiguous().numpy().astype(np.float32)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        index[name] = dict(config=cfg.to_dict(), weight_seed=wseed, stress=stress, batch=B,
                           height=H, width=W, input_seed=xseed, full=full,
                           sample_stride=SAMPLE_STRIDE)
        print(name, {k: tuple(v.shape) for k, v in feats.items()}, missing)
    with open(os.path.join(out_dir, "index.json"), "w") as f:
        json.dump(dict(generator="oracle/make_golden.py",
                       reference="/root/reference/src/layoutdit/modeling/dit_backbone.py (DiTBackbone)",
                       transformers=__import__("transformers").__version__,
                       torch=torch.__version__, cases=index), f, indent=1)


if __name__ == "__main__":
    main()
