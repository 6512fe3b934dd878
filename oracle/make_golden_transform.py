"""Generate tests/golden/transform_*.npz from the transform object inside THE REFERENCE'S OWN detector
(build container only): ``LayoutDetectionModel(ModelConfig()).model.transform`` -- the
``GeneralizedRCNNTransform`` the reference configures at R:src/layoutdit/modeling/model.py:44-56 -- run in eval
mode on seeded raw pages.  Only the hub fetch is replaced (as in oracle/make_golden.py).
Run from the repo root:  ``python oracle/make_golden_transform.py``
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

from layoutdit_b200.config import dit_base  # noqa: E402
from layoutdit_b200.synth import raw_pages  # noqa: E402
from oracle import hf_reference  # noqa: E402

# name -> (list of (H, W), seed)
CASES = {
    "transform_mixed": ([(300, 212), (224, 224), (500, 640), (97, 131)], 301),
    "transform_1024": ([(1024, 1024), (1024, 768)], 302),
    "transform_small": ([(16, 16), (1, 1), (33, 500)], 303),
}
STRIDE = 7


def reference_transform():
    import transformers
    from layoutdit.configuration.model_config import ModelConfig
    from layoutdit.modeling import dit_backbone as ref
    from layoutdit.modeling.model import LayoutDetectionModel
    hf_cfg = hf_reference.hf_config(dit_base(num_hidden_layers=1).to_dict())
    orig = transformers.AutoConfig.from_pretrained
    ref.AutoConfig.from_pretrained = staticmethod(lambda *a, **k: hf_cfg)
    orig_model = ref.AutoModel.from_pretrained
    ref.AutoModel.from_pretrained = staticmethod(lambda *a, **k: ref.AutoModel.from_config(hf_cfg))
    try:
        m = LayoutDetectionModel(ModelConfig())
    finally:
        ref.AutoConfig.from_pretrained = orig
        ref.AutoModel.from_pretrained = orig_model
    return m.eval().model.transform


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    tr = reference_transform()
    index = {}
    for name, (sizes, seed) in CASES.items():
        pages = raw_pages(sizes, seed)
        with torch.no_grad():
            images, _ = tr([p.clone() for p in pages])
        t = images.tensors.contiguous().numpy().astype(np.float32)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), shape=np.asarray(t.shape, dtype=np.int64),
                            samples=t.reshape(-1)[::STRIDE].copy())
        index[name] = dict(sizes=sizes, seed=seed, stride=STRIDE, image_sizes=[list(s) for s in images.image_sizes])
        print(name, t.shape, images.image_sizes)
    with open(os.path.join(out_dir, "index_transform.json"), "w") as f:
        json.dump(dict(generator="oracle/make_golden_transform.py",
                       reference="LayoutDetectionModel(ModelConfig()).model.transform (R:src/layoutdit/modeling/model.py:44-56)",
                       torchvision=__import__("torchvision").__version__, cases=index), f, indent=1)


if __name__ == "__main__":
    main()
