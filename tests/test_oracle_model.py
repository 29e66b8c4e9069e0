"""Pins oracle/dav2_oracle.py (the reference's model is an un-vendored external checkout; see the
oracle header): parameter counts, upstream key contract, and agreement with the independent
implementation in `transformers` through a weight remap (fp32, 518x518 -> no pos-embed interpolation)."""
import re

import numpy as np
import pytest
import torch

from oracle import dav2_oracle as O


@pytest.mark.parametrize("enc,total", [("vits", 24785089), ("vitb", 97470785), ("vitl", 335315649)])
def test_param_counts(enc, total):
    cfg = O.MODEL_CONFIGS[enc]
    with torch.device("meta"):
        m = O.DepthAnythingV2(enc, cfg["features"], cfg["out_channels"])
    assert O.count_params(m) == total
    assert O.count_params(m.pretrained) == {"vits": 22056576, "vitb": 86580480, "vitl": 304368640}[enc]


def test_state_dict_key_contract():
    with torch.device("meta"):
        m = O.DepthAnythingV2("vits", 64, [48, 96, 192, 384])
    sd = m.state_dict()
    keys = set(sd)
    for k in ("pretrained.cls_token", "pretrained.pos_embed", "pretrained.mask_token",
              "pretrained.patch_embed.proj.weight", "pretrained.blocks.11.attn.qkv.weight",
              "pretrained.blocks.0.ls1.gamma", "pretrained.blocks.0.mlp.fc2.bias", "pretrained.norm.weight",
              "depth_head.projects.3.bias", "depth_head.resize_layers.0.weight", "depth_head.resize_layers.3.weight",
              "depth_head.scratch.layer4_rn.weight", "depth_head.scratch.refinenet4.resConfUnit1.conv1.weight",
              "depth_head.scratch.refinenet1.out_conv.bias", "depth_head.scratch.output_conv1.weight",
              "depth_head.scratch.output_conv2.0.weight", "depth_head.scratch.output_conv2.2.bias"):
        assert k in keys, k
    assert "depth_head.resize_layers.2.weight" not in keys  # Identity
    assert "depth_head.scratch.layer1_rn.bias" not in keys  # bias=False
    assert tuple(sd["pretrained.pos_embed"].shape) == (1, 1370, 384)
    assert tuple(sd["depth_head.resize_layers.0.weight"].shape) == (48, 48, 4, 4)
    # lightning_model.py:130-140 filters on the substring "pretrained"
    assert all(("pretrained" in k) == k.startswith("pretrained.") for k in keys)


def _to_hf(sd, D):
    out = {}
    for k, v in sd.items():
        if k == "pretrained.cls_token":
            out["backbone.embeddings.cls_token"] = v
        elif k == "pretrained.pos_embed":
            out["backbone.embeddings.position_embeddings"] = v
        elif k == "pretrained.mask_token":
            out["backbone.embeddings.mask_token"] = v
        elif k.startswith("pretrained.patch_embed.proj."):
            out["backbone.embeddings.patch_embeddings.projection." + k.rsplit(".", 1)[1]] = v
        elif k.startswith("pretrained.norm."):
            out["backbone.layernorm." + k.rsplit(".", 1)[1]] = v
        elif k.startswith("pretrained.blocks."):
            m = re.match(r"pretrained\.blocks\.(\d+)\.(.*)", k)
            i, rest = m.group(1), m.group(2)
            p = f"backbone.encoder.layer.{i}."
            if rest.startswith("attn.qkv."):
                wb = rest.rsplit(".", 1)[1]
                for j, name in enumerate(("query", "key", "value")):
                    out[p + f"attention.attention.{name}.{wb}"] = v[j * D:(j + 1) * D]
            elif rest.startswith("attn.proj."):
                out[p + "attention.output.dense." + rest.rsplit(".", 1)[1]] = v
            elif rest == "ls1.gamma":
                out[p + "layer_scale1.lambda1"] = v
            elif rest == "ls2.gamma":
                out[p + "layer_scale2.lambda1"] = v
            else:
                out[p + rest] = v
        elif k.startswith("depth_head.projects."):
            i, wb = k.split(".")[2], k.split(".")[3]
            out[f"neck.reassemble_stage.layers.{i}.projection.{wb}"] = v
        elif k.startswith("depth_head.resize_layers."):
            i, wb = k.split(".")[2], k.split(".")[3]
            out[f"neck.reassemble_stage.layers.{i}.resize.{wb}"] = v
        elif "_rn." in k:
            i = int(re.search(r"layer(\d)_rn", k).group(1)) - 1
            out[f"neck.convs.{i}.weight"] = v
        elif ".refinenet" in k:
            m = re.match(r"depth_head\.scratch\.refinenet(\d)\.(.*)", k)
            j, rest = 4 - int(m.group(1)), m.group(2)
            rest = (rest.replace("out_conv", "projection").replace("resConfUnit", "residual_layer")
                    .replace("conv1", "convolution1").replace("conv2", "convolution2"))
            out[f"neck.fusion_stage.layers.{j}.{rest}"] = v
        elif "output_conv1" in k:
            out["head.conv1." + k.rsplit(".", 1)[1]] = v
        elif "output_conv2.0" in k:
            out["head.conv2." + k.rsplit(".", 1)[1]] = v
        elif "output_conv2.2" in k:
            out["head.conv3." + k.rsplit(".", 1)[1]] = v
        else:
            raise KeyError(k)
    return out


@pytest.mark.slow
def test_oracle_matches_transformers_vits():
    tr = pytest.importorskip("transformers")
    enc = "vits"
    dims = O.ENCODER_DIMS[enc]
    cfg = O.MODEL_CONFIGS[enc]
    bc = tr.Dinov2Config(hidden_size=dims["embed_dim"], num_hidden_layers=dims["depth"],
                         num_attention_heads=dims["num_heads"], image_size=518, patch_size=14,
                         out_indices=[i + 1 for i in O.TAP_LAYERS[enc]], reshape_hidden_states=False,
                         apply_layernorm=True, layer_norm_eps=1e-6)
    hc = tr.DepthAnythingConfig(backbone_config=bc, reassemble_hidden_size=dims["embed_dim"],
                                neck_hidden_sizes=cfg["out_channels"], fusion_hidden_size=cfg["features"],
                                head_hidden_size=32, depth_estimation_type="metric", max_depth=20,
                                reassemble_factors=[4, 2, 1, 0.5])
    hf = tr.DepthAnythingForDepthEstimation(hc).eval()
    ours = O.build_oracle(enc, seed=3)
    missing, unexpected = hf.load_state_dict(_to_hf(ours.state_dict(), dims["embed_dim"]), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    x = O.synthetic_frames(1, 518, 518, seed=7)
    with torch.no_grad():
        a = ours(x)
        b = hf(pixel_values=x).predicted_depth
    assert a.shape == b.shape == (1, 518, 518)
    assert float(a.std()) > 1.0  # non-degenerate
    rel = float((a - b).abs().max() / b.abs().max())
    assert rel < 1e-5, rel


def test_resize_target_examples():
    # SURVEY App. A.3 examples
    assert O.resize_target(475, 475, 518) == (518, 518)
    assert O.resize_target(475, 1000, 518) == (518, 1092)
    assert O.resize_target(1080, 1350, 518) == (518, 644)
    assert O.resize_target(1036, 1036, 518) == (518, 518)


def test_pos_embed_interpolation_shapes():
    m = O.build_oracle("vits", calibrate=False)
    with torch.no_grad():
        p = m.pretrained.interpolate_pos_encoding(74 * 74, 1036, 1036)
        assert p.shape == (1, 1 + 74 * 74, 384)
        q = m.pretrained.interpolate_pos_encoding(37 * 78, 518, 1092)
        assert q.shape == (1, 1 + 37 * 78, 384)
        d = m(O.synthetic_frames(1, 70, 98))
        assert d.shape == (1, 70, 98) and torch.isfinite(d).all()


def test_config1_fixture_pins_oracle_and_geometry(golden_dir):
    """BASELINE configs[0] fixture (scripts/make_golden.py config1): the reference's own frame through the oracle's
    infer_image reproduces the committed depth sample, and the geometry oracle reproduces the points that the verbatim
    depth_to_pointcloud_dav2.py:300-313 arithmetic produced from that depth."""
    import os
    import cv2
    import numpy as np
    from oracle import geometry_oracle as geo
    g = np.load(os.path.join(golden_dir, "config1_vits.npz"))
    crop = cv2.imread(os.path.join(golden_dir, "FrameBuffer_0051_left475.png"))
    assert crop is not None and crop.shape == (475, 475, 3)
    img518 = cv2.resize(crop, (518, 518), interpolation=cv2.INTER_CUBIC)
    oracle = O.build_oracle("vits", seed=0)
    depth = oracle.infer_image(img518, 518)
    s = int(g["stride"])
    assert np.abs(depth[::s, ::s] - g["depth_sub"]).max() < 2e-3  # summation order differs with the thread count only
    np.testing.assert_allclose([depth.min(), depth.max(), depth.mean(), depth.std()], g["depth_stats"], rtol=2e-3)
    # K scaled from 475 to 518 (datasets/UnityCam/cam.txt:1) == SURVEY.md 8d config 1
    np.testing.assert_allclose(g["k4"], [170.1677, 169.8526, 194.7248, 198.2624], atol=1e-3)
    full = np.zeros((518, 518), np.float32)
    full[::s, ::s] = g["depth_sub"]
    pts, valid = geo.backproject(full, tuple(g["k4"]), np.eye(4))
    sub = np.zeros((518, 518), bool)
    sub[::s, ::s] = True
    np.testing.assert_allclose(pts[sub.reshape(-1)], g["points_sub"], rtol=1e-12, atol=1e-15)
