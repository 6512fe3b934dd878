"""GPU: the drop-in inside its real consumer, and the on-disk formats next to the path.

* ``layoutdit_b200.DiTWithFPN`` as the ``backbone`` of torchvision's ``FasterRCNN``, built exactly as the reference
  builds its detector (R:src/layoutdit/modeling/model.py:33-56: MultiScaleRoIAlign over p2..p5 + pool, the anchor
  generator of R:configuration/model_config.py, fixed_size (224, 224), mean/std 0.5), eval mode, against the same
  detector around the reference's backbone (HF ``BeitModel`` + torchvision FPN in fp32 on the same GPU);
* fp32 FPN maps (``out_dtype=torch.float32``): the store format of the convolution epilogue, bit-identical after
  rounding to the bf16 maps;
* checkpoint file -> module -> forward against the oracle;
* BASELINE configs 3 and 4 at their 8-GPU per-rank batch (4 pages of 512x512; DiT-large, 8 pages) against the oracle,
  DiT-large batch independence at the full batch of 64.
"""
from collections import OrderedDict

import pytest
import torch
import torch.nn as nn

from layoutdit_b200 import DiTBackbone, DiTWithFPN, checkpoint
from layoutdit_b200.config import DiTConfig, dit_base, dit_large
from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, raw_pages, synthetic_pages
from oracle import dit_oracle, fpn_oracle, hf_reference

pytestmark = pytest.mark.gpu

SMALL224 = dict(hidden_size=128, num_hidden_layers=4, num_attention_heads=2, intermediate_size=256, image_size=224)


def _rel_fro(got, ref):
    return float((got.double().cpu() - ref.double().cpu()).norm() / ref.double().cpu().norm().clamp_min(1e-30))


class _ReferenceDiTWithFPN(nn.Module):
    """R:dit_backbone.py:65-95 with the hub fetch replaced: HF backbone twin + torchvision's own FPN, fp32."""

    def __init__(self, cfg, sd, fsd):
        super().__init__()
        from torchvision.ops import FeaturePyramidNetwork
        from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
        self.backbone = hf_reference.build(cfg.to_dict(), sd)
        self.fpn = FeaturePyramidNetwork([cfg.hidden_size] * 4, 256, extra_blocks=LastLevelMaxPool())
        self.fpn.load_state_dict(fsd, strict=True)
        self.out_channels = 256

    def forward(self, x):
        return self.fpn(self.backbone(x))


def _detector(backbone):
    """R:model.py:33-56."""
    from torchvision.models.detection import FasterRCNN
    from torchvision.models.detection.rpn import AnchorGenerator
    from torchvision.ops import MultiScaleRoIAlign
    roi = MultiScaleRoIAlign(featmap_names=["p2", "p3", "p4", "p5", "pool"], output_size=7, sampling_ratio=2)
    anchors = AnchorGenerator(sizes=((32,), (64,), (128,), (256,), (512,)), aspect_ratios=((0.5, 1.0, 2.0),) * 5)
    return FasterRCNN(backbone, num_classes=5 + 1, rpn_anchor_generator=anchors, box_roi_pool=roi, max_size=224, min_size=224,
                      fixed_size=(224, 224), image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5))


def test_fpn_fp32_output_is_the_unrounded_bf16_output(cuda_device):
    cfg = DiTConfig(**SMALL224)
    sd, fsd = make_state_dict(cfg, 21, True), make_fpn_state_dict(cfg.hidden_size, 256, 22, True)
    x = synthetic_pages(3, 224, 224, 7).cuda()
    a = DiTWithFPN(pretrained=False, config=cfg, state_dict=sd, fpn_state_dict=fsd).cuda().eval()(x)
    b = DiTWithFPN(pretrained=False, config=cfg, state_dict=sd, fpn_state_dict=fsd, out_dtype=torch.float32).cuda().eval()(x)
    assert list(b.keys()) == ["p2", "p3", "p4", "p5", "pool"]
    for k in a:
        assert b[k].dtype == torch.float32 and a[k].dtype == torch.bfloat16 and a[k].shape == b[k].shape
        assert torch.isfinite(b[k]).all()
        assert torch.equal(b[k].to(torch.bfloat16), a[k]), k     # same accumulators, only the store format differs
    ref = fpn_oracle.dit_with_fpn_forward(sd, fsd, cfg.to_dict(), x.cpu())
    for k, r in ref.items():
        assert _rel_fro(b[k], r) < 1e-2


def test_drop_in_backbone_inside_faster_rcnn(cuda_device):
    cfg = DiTConfig(**SMALL224)
    sd, fsd = make_state_dict(cfg, 31, True), make_fpn_state_dict(cfg.hidden_size, 256, 32, True)
    ours = _detector(DiTWithFPN(pretrained=False, config=cfg, state_dict=sd, fpn_state_dict=fsd, out_dtype=torch.float32))
    torch.manual_seed(5)
    theirs = _detector(_ReferenceDiTWithFPN(cfg, sd, fsd))
    # identical detection heads on both sides (random init, R trains them from scratch)
    ours.rpn.load_state_dict(theirs.rpn.state_dict())
    ours.roi_heads.load_state_dict(theirs.roi_heads.state_dict())
    ours, theirs = ours.cuda().eval(), theirs.cuda().eval()
    pages = [p.cuda() for p in raw_pages([(300, 260), (224, 224), (512, 400)], 9)]

    cap = {}
    def hook(tag):
        def fn(_m, _inp, out):
            cap[tag] = out
        return fn
    h1 = ours.rpn.head.register_forward_hook(hook("ours"))
    h2 = theirs.rpn.head.register_forward_hook(hook("theirs"))
    with torch.no_grad():
        det_o = ours(pages)
        det_t = theirs(pages)
    h1.remove(); h2.remove()

    # RPN head outputs (objectness logits and box deltas per level): a deterministic function of the FPN maps
    for part in (0, 1):
        for lo, lt in zip(cap["ours"][part], cap["theirs"][part]):
            assert lo.shape == lt.shape and torch.isfinite(lo).all()
            e = _rel_fro(lo, lt)
            assert e < 2e-2, f"rpn head output {part}: rel-Frobenius {e:.3e}"
    # detections: well-formed, and the confident reference boxes are found again (bf16 features move scores a little,
    # so ordering and the exact NMS survivors may differ)
    from torchvision.ops import box_iou
    assert len(det_o) == len(det_t) == len(pages)
    matched = total = 0
    for do, dt in zip(det_o, det_t):
        assert set(do.keys()) == {"boxes", "labels", "scores"} and do["boxes"].dtype == torch.float32
        assert torch.isfinite(do["boxes"]).all() and torch.isfinite(do["scores"]).all()
        top = dt["scores"].argsort(descending=True)[:10]
        if len(top) == 0 or len(do["boxes"]) == 0:
            continue
        iou = box_iou(dt["boxes"][top], do["boxes"])
        same = dt["labels"][top][:, None] == do["labels"][None, :]
        matched += int(((iou > 0.8) & same).any(dim=1).sum())
        total += len(top)
    print(f"faster-rcnn: {matched}/{total} of the reference's top boxes re-found")
    assert total == 0 or matched >= 0.7 * total


@pytest.mark.parametrize("ext", ["pth", "safetensors"])
def test_checkpoint_file_to_forward(cuda_device, tmp_path, ext):
    """SURVEY 8 f4 on the GPU: a LayoutDiT whole-model checkpoint on disk -> build_from_checkpoint -> forward == oracle."""
    cfg = DiTConfig(hidden_size=128, num_hidden_layers=6, num_attention_heads=2, intermediate_size=256, image_size=64)
    sd, fsd = make_state_dict(cfg, 41, True), make_fpn_state_dict(cfg.hidden_size, 256, 42, True)
    src = DiTWithFPN(pretrained=False, config=cfg, state_dict=sd, fpn_state_dict=fsd)
    path = str(tmp_path / f"layoutdit_epoch3.{ext}")
    checkpoint.save_checkpoint(src, path, layout="layoutdit")
    m = checkpoint.build_from_checkpoint(path).cuda().eval()
    assert isinstance(m, DiTWithFPN) and m.backbone.config == cfg
    x = synthetic_pages(2, 64, 64, 3)
    got = m(x.cuda())
    ref = fpn_oracle.dit_with_fpn_forward(sd, fsd, cfg.to_dict(), x)
    for k, r in ref.items():
        assert _rel_fro(got[k].float(), r) < 1e-2, k
    # the real DiT export layout (BeitForMaskedImageModeling: beit.* without pooler) into the bare backbone
    mim = {"beit." + k: v for k, v in sd.items() if not k.startswith("pooler.")}
    mim["lm_head.weight"] = torch.zeros(16, cfg.hidden_size)
    path2 = str(tmp_path / "pytorch_model.bin")
    torch.save(mim, path2)
    bb = DiTBackbone(pretrained=path2, config=cfg).cuda().eval()
    ref2 = dit_oracle.dit_backbone_forward(sd, cfg.to_dict(), x)
    got2 = bb(x.cuda())
    for k, r in ref2.items():
        assert _rel_fro(got2[k].float(), r) < 1e-2, k


@pytest.mark.parametrize("name,cfg,B,H,W", [("config 3 per rank", dit_base(), 4, 512, 512), ("config 4 per rank", dit_large(), 8, 224, 224)])
def test_strong_scaling_per_rank_batches_match_oracle(cuda_device, name, cfg, B, H, W):
    """BASELINE configs 3 / 4 on 8 GPUs leave 4 pages of 512x512 (M = 4100 rows) or 8 pages (DiT-large, M = 1576 rows) per
    rank: the under-filled persistent schedules must give the same numbers as the oracle."""
    sd = make_state_dict(cfg, 1, True)
    x = synthetic_pages(B, H, W, 55)
    got = DiTBackbone(pretrained=False, config=cfg, state_dict=sd).cuda().eval()(x.cuda())
    sel = [0, B - 1]
    ref = dit_oracle.dit_backbone_forward(sd, cfg.to_dict(), x[sel])
    for k, r in ref.items():
        e = _rel_fro(got[k][sel].float(), r)
        print(name, k, f"{e:.2e}")
        assert e < 1e-2


def test_dit_large_full_batch_is_batch_independent(cuda_device):
    cfg = dit_large()
    sd = make_state_dict(cfg, 2, True)
    x = synthetic_pages(64, 224, 224, 1236)
    m = DiTBackbone(pretrained=False, config=cfg, state_dict=sd).cuda().eval()
    full = m(x.cuda())
    solo = m(x[41:42].cuda())
    for k in full:
        assert torch.isfinite(full[k].float()).all()
        assert torch.equal(full[k][41:42], solo[k]), k
