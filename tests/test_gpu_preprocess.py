"""GPU pre-processing: the run.py path batched (upstream image2tensor, OpenCV bicubic) and the
dataset path (data_processing/simcol.py:104-135,161-168: torchvision Resize(BICUBIC, antialias=True) on float tensors),
each against the library call the reference itself makes (cv2 / torchvision, executed here)."""
import numpy as np
import pytest
import torch

from oracle import dav2_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,h,w,size", [(3, 95, 120, 140), (2, 475, 475, 518), (4, 60, 80, 70)])
def test_batched_image2tensor_matches_opencv(B, h, w, size):
    from dav2_b200 import ops
    from dav2_b200.dpt import DepthAnythingV2
    rng = np.random.default_rng(B + h)
    imgs = rng.integers(0, 255, size=(B, h, w, 3), dtype=np.uint8)
    nh, nw = DepthAnythingV2.target_size(h, w, size)
    got = ops.preprocess_bgr_u8(torch.from_numpy(imgs).cuda(), nh, nw).cpu()
    assert got.shape == (B, 3, nh, nw)
    for b in range(B):
        ref, _ = O.image2tensor(imgs[b], size)
        assert float((got[b] - ref[0]).abs().max()) < 2e-5
    one = ops.preprocess_bgr_u8(torch.from_numpy(imgs[1]).cuda(), nh, nw).cpu()   # the un-batched form is the same kernel
    assert torch.equal(one[0], got[1])


@pytest.mark.parametrize("H,W,S", [(475, 475, 518),     # SimCol frames: up-sampling, 4 taps
                                   (600, 720, 518),     # down-sampling: kernel widened by the scale
                                   (518, 518, 518), (100, 37, 70)])
def test_simcol_transforms_match_torchvision(H, W, S):
    """transform_input / transform_output == ToTensor -> Resize((S,S), BICUBIC, antialias=True) [-> Normalize] executed
    with torchvision on the CPU exactly as data_processing/simcol.py:104-135,161-168 does."""
    import torchvision.transforms as T
    from dav2_b200.data_processing import SimColTransforms
    rng = np.random.default_rng(H + W)
    B = 2
    image = rng.integers(0, 255, size=(B, H, W, 3), dtype=np.uint8)
    depth = rng.integers(0, 65535, size=(B, H, W), dtype=np.uint16)
    t_in = T.Compose([T.ToTensor(), T.Resize((S, S), interpolation=T.InterpolationMode.BICUBIC, antialias=True),
                      T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    t_out = T.Compose([T.ToTensor(), T.Resize((S, S), interpolation=T.InterpolationMode.BICUBIC, antialias=True)])
    tr = SimColTransforms(S, "cuda")
    got_i = tr.transform_input(image).cpu()
    got_d = tr.transform_output(depth).cpu()
    assert got_i.shape == (B, 3, S, S) and got_d.shape == (B, 1, S, S)
    for b in range(B):
        ref_i = t_in(image[b].astype(np.float32) / 255.0)                 # simcol.py:161-162
        ref_d = t_out(depth[b].astype(np.float32) / 65535.0)              # simcol.py:163-165
        # fp32 tap sums in a different order (one pass over the 2-D footprint here, two separable passes in ATen):
        # ~1e-6 of the [0, 1] range; the normalised image carries the 1/std = 4.4x factor
        ei, ed = float((got_i[b] - ref_i).abs().max()), float((got_d[b] - ref_d).abs().max())
        print(f"{H}x{W}->{S}: image err {ei:.2e} depth err {ed:.2e}")
        assert ei < 3e-5 and ed < 1e-5, (ei, ed)
    single = tr(image[0], depth[0])
    assert single["image"].shape == (3, S, S) and single["depth"].shape == (1, S, S)
    assert torch.equal(single["image"].cpu(), got_i[0]) and torch.equal(single["depth"].cpu(), got_d[0])


def test_load_item_files(tmp_path):
    from PIL import Image
    from dav2_b200.data_processing import SimColTransforms, load_item
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 255, size=(64, 64, 4), dtype=np.uint8)          # RGBA on disk: the loader keeps [:, :, :3]
    dep = rng.integers(0, 65535, size=(64, 64), dtype=np.uint16)
    Image.fromarray(rgb, "RGBA").save(tmp_path / "FrameBuffer_0000.png")
    Image.fromarray(dep).save(tmp_path / "Depth_0000.png")
    item = load_item(str(tmp_path / "FrameBuffer_0000.png"), str(tmp_path / "Depth_0000.png"), SimColTransforms(70))
    assert item["image"].shape == (3, 70, 70) and item["depth"].shape == (1, 70, 70)
    assert -0.3 <= float(item["depth"].min()) and float(item["depth"].max()) <= 1.3  # [0, 1] plus bicubic over/undershoot


def test_pose_pair_items():
    """data_processing.pose_pair_items (pose_estimation.py:205-311): every frame transformed once, pairs stacked to 8
    channels in the order rgb_i, depth_i, rgb_{i+1}, depth_{i+1}, targets from the absolute poses."""
    from dav2_b200.data_processing import SimColTransforms, pose_pair_items, relative_pose_targets
    rng = np.random.default_rng(2)
    N, H, W, S = 5, 95, 95, 70
    imgs = rng.integers(0, 255, size=(N, H, W, 3), dtype=np.uint8)
    deps = rng.integers(0, 65535, size=(N, H, W), dtype=np.uint16)
    poses = np.concatenate([rng.normal(0, 1, (N, 3)), rng.normal(0, 1, (N, 4))], 1)
    poses[:, 3:] /= np.linalg.norm(poses[:, 3:], axis=1, keepdims=True)
    tr = SimColTransforms(S)
    item = pose_pair_items(imgs, deps, poses, tr)
    assert item["input"].shape == (N - 1, 8, S, S) and item["input"].is_cuda and item["target"].shape == (N - 1, 7)
    for i in range(N - 1):
        one_a, one_b = tr(imgs[i], deps[i]), tr(imgs[i + 1], deps[i + 1])
        want = torch.cat([one_a["image"], one_a["depth"], one_b["image"], one_b["depth"]], 0)
        assert torch.equal(item["input"][i], want)
    assert torch.equal(item["target"].cpu(), relative_pose_targets(poses))
