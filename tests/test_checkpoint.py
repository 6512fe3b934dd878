"""CPU: checkpoint formats either side of the hot path."""
import pytest
import torch

from layoutdit_b200 import DiTBackbone, DiTWithFPN, checkpoint
from layoutdit_b200.config import DiTConfig, dit_base, dit_large
from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict
from oracle import hf_reference

TINY = dict(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256, image_size=64)


@pytest.mark.parametrize("cfg", [DiTConfig(**TINY), DiTConfig(**TINY, use_absolute_position_embeddings=False,
                                                             use_relative_position_bias=True),
                                 DiTConfig(**TINY, use_absolute_position_embeddings=False, use_mask_token=False,
                                           use_shared_relative_position_bias=True, layer_scale_init_value=0.0)])
def test_config_is_inferred_from_an_hf_state_dict(cfg):
    hf_sd = hf_reference.build(cfg.to_dict()).dit.state_dict()     # what a real BeitModel checkpoint holds
    got = checkpoint.infer_config(hf_sd)
    assert got == cfg


def test_base_and_large_shapes_infer_without_materialising_weights():
    for cfg in (dit_base(), dit_large()):
        from layoutdit_b200.synth import state_dict_keys
        meta = {k: torch.empty(s, device="meta") for k, s in state_dict_keys(cfg)}
        assert checkpoint.infer_config(meta) == cfg


@pytest.mark.parametrize("ext", ["pth", "safetensors"])
@pytest.mark.parametrize("prefix", ["", "beit."])
def test_hf_checkpoint_file_roundtrip(tmp_path, ext, prefix):
    cfg = DiTConfig(**TINY)
    sd = make_state_dict(cfg, 3, True)
    on_disk = {prefix + k: v for k, v in sd.items()}
    if prefix:   # heads of the exporting model: must be ignored
        on_disk["lm_head.weight"] = torch.zeros(8, cfg.hidden_size)
        on_disk["layernorm.weight"] = torch.ones(cfg.hidden_size)
    path = str(tmp_path / f"model.{ext}")
    if ext == "safetensors":
        from safetensors.torch import save_file
        save_file(on_disk, path)
    else:
        torch.save(on_disk, path)
    m = checkpoint.build_from_checkpoint(path)
    assert isinstance(m, DiTBackbone) and m.config == cfg
    for k, v in m.dit.state_dict().items():
        assert torch.equal(v, sd[k])
    out = str(tmp_path / f"again.{ext}")
    checkpoint.save_checkpoint(m, out, layout="hf")
    again = checkpoint.read_state_dict(out)
    assert sorted(again.keys()) == sorted(sd.keys()) and all(torch.equal(again[k], sd[k]) for k in sd)   # safetensors sorts keys
    # and the file is loadable by the reference's own model class
    ref = hf_reference.build(cfg.to_dict())
    assert not ref.dit.load_state_dict(again, strict=True).missing_keys


def test_layoutdit_whole_model_checkpoint(tmp_path):
    """R:model.py:90-121 saves the detector's state_dict: backbone under model.backbone.backbone.dit.,
    FPN under model.backbone.fpn., plus heads this module does not own."""
    cfg = DiTConfig(**TINY)
    sd, fsd = make_state_dict(cfg, 4, True), make_fpn_state_dict(cfg.hidden_size, 256, 5, True)
    full = {"model.backbone.backbone.dit." + k: v for k, v in sd.items()}
    full["model.backbone.backbone.dit.encoder.layer.0.attention.attention.relative_position_bias.relative_position_index"] = torch.zeros(3)
    full.update({"model.backbone.fpn." + k: v for k, v in fsd.items()})
    full["model.rpn.head.conv.0.0.weight"] = torch.zeros(256, 256, 3, 3)
    full["model.roi_heads.box_predictor.cls_score.weight"] = torch.zeros(6, 1024)
    path = str(tmp_path / "epoch_3_cpu.pth")
    torch.save(full, path)
    m = checkpoint.build_from_checkpoint(path)
    assert isinstance(m, DiTWithFPN)
    parts = checkpoint.split_checkpoint(full)
    assert parts.dit_prefix == "model.backbone.backbone.dit." and parts.fpn_prefix == "model.backbone.fpn."
    assert sorted(parts.other) == ["model.roi_heads.box_predictor.cls_score.weight", "model.rpn.head.conv.0.0.weight"]
    for k, v in m.backbone.dit.state_dict().items():
        assert torch.equal(v, sd[k])
    for k, v in m.fpn.state_dict().items():
        assert torch.equal(v, fsd[k])
    exported = checkpoint.export_state_dict(m, layout="layoutdit")
    owned = {k: v for k, v in full.items() if not k.startswith(("model.rpn", "model.roi_heads")) and not k.endswith("relative_position_index")}
    assert exported.keys() == owned.keys() and all(torch.equal(exported[k], owned[k]) for k in owned)
    # the reference's own resume call (R:model.py:70) on this file: nn.Module semantics, nothing matches the prefix
    res = DiTBackbone(pretrained=False, config=cfg).dit.load_state_dict(full, strict=False)
    assert len(res.unexpected_keys) == len(full) - 1     # the legacy relative_position_index buffer is dropped silently (HF:674)


def test_errors():
    with pytest.raises(ValueError):
        checkpoint.split_checkpoint({"foo.weight": torch.zeros(1)})
    cfg = DiTConfig(**TINY)
    sd = make_state_dict(cfg, 1, False)
    del sd["encoder.layer.1.output.dense.bias"]
    with pytest.raises(RuntimeError):
        checkpoint.load_checkpoint(DiTBackbone(pretrained=False, config=cfg), sd, strict=True)
    parts = checkpoint.load_checkpoint(DiTBackbone(pretrained=False, config=cfg), sd, strict=False)
    assert parts.dit_prefix == ""


def test_real_dit_checkpoint_layout_without_pooler_loads(tmp_path):
    """microsoft/dit-base|large are BeitForMaskedImageModeling exports: BeitModel(add_pooling_layer=False) under
    ``beit.`` plus top-level ``layernorm.*`` / ``lm_head.*`` and NO ``pooler.*`` (HF:799).  All three entry points
    (ctor, load_checkpoint, build_from_checkpoint) must take that file."""
    cfg = DiTConfig(**TINY)
    sd = make_state_dict(cfg, 5, True)
    mim = {"beit." + k: v for k, v in sd.items() if not k.startswith("pooler.")}
    mim["layernorm.weight"] = torch.ones(cfg.hidden_size)
    mim["layernorm.bias"] = torch.zeros(cfg.hidden_size)
    mim["lm_head.weight"] = torch.zeros(8192, cfg.hidden_size)
    mim["lm_head.bias"] = torch.zeros(8192)
    path = str(tmp_path / "pytorch_model.bin")
    torch.save(mim, path)
    m1 = DiTBackbone(pretrained=False, config=cfg, state_dict=mim)
    m2 = DiTBackbone(pretrained=False, config=cfg)
    parts = checkpoint.load_checkpoint(m2, path)                  # strict=True by default
    m3 = checkpoint.build_from_checkpoint(path)
    assert sorted(parts.other) == ["layernorm.bias", "layernorm.weight", "lm_head.bias", "lm_head.weight"]
    for m in (m1, m2, m3):
        got = m.dit.state_dict()
        for k, v in sd.items():
            if not k.startswith("pooler."):
                assert torch.equal(got[k], v), k
        assert m.pretrained is False
    # a genuinely missing tensor still raises
    broken = {k: v for k, v in mim.items() if "layer.1.output.dense.weight" not in k}
    with pytest.raises(RuntimeError, match="Missing"):
        checkpoint.load_checkpoint(DiTBackbone(pretrained=False, config=cfg), broken)


def test_pretrained_without_weights_warns_and_a_path_is_honoured(tmp_path, monkeypatch):
    """R:dit_backbone.py:27-29 downloads the weights when pretrained=True; offline that must not degrade silently."""
    cfg = DiTConfig(**TINY)
    monkeypatch.delenv("LDIT_PRETRAINED_PATH", raising=False)
    with pytest.warns(RuntimeWarning, match="RANDOM-INIT"):
        m = DiTBackbone(config=cfg)                               # pretrained=True is the reference's default
    assert m.pretrained is True
    with pytest.warns(RuntimeWarning, match="RANDOM-INIT"):
        DiTWithFPN(pretrained=True, config=cfg)
    sd = make_state_dict(cfg, 9, True)
    path = str(tmp_path / "dit.pth")
    torch.save(sd, path)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        a = DiTBackbone(pretrained=path, config=cfg)
        monkeypatch.setenv("LDIT_PRETRAINED_PATH", path)
        b = DiTBackbone(pretrained=True, config=cfg)
        c = DiTBackbone(pretrained=True, config=cfg, state_dict=sd)
    for m in (a, b, c):
        assert m.pretrained is False
        assert torch.equal(m.dit.state_dict()["encoder.layer.0.lambda_1"], sd["encoder.layer.0.lambda_1"])
