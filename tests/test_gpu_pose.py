"""GPU parity of the PoseEstimationNet engine (ResNet-18, 8-channel stem, MLP head) and of the config-4
chain: pose net -> compose_poses -> world-frame back-projection, against the CPU oracles."""
import numpy as np
import pytest
import torch

from oracle import geometry_oracle as geo
from oracle import pose_oracle

pytestmark = pytest.mark.gpu


def _build(seed=0, precision="fp16"):
    from dav2_b200.pose_estimation_model import PoseEstimationNet
    ref = pose_oracle.build_pose_oracle(8, seed)
    net = PoseEstimationNet(8, precision=precision)
    sd = {k: v for k, v in ref.state_dict().items()}
    res = net.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return ref, net.cuda().eval()


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 224, 224), (3, 98, 126),
                                   (2, 518, 518)])  # the frame size of BASELINE configs[3] (stacked 518^2 pairs)
def test_pose_net_matches_oracle(B, H, W):
    ref, net = _build()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 8, H, W, generator=g)
    with torch.no_grad():
        want = ref(x)
    got = net(x.cuda()).cpu()
    assert got.shape == (B, 7) and got.dtype == torch.float32
    scale = float(want.abs().max())
    assert scale > 0.05  # non-degenerate
    assert float((got - want).abs().max()) / scale < 1e-2, (got, want)


def test_pose_chain_to_world_cloud():
    """config 4 in miniature: pairs -> relative poses -> compose_poses -> [R|t] rows -> world transform of a depth map."""
    from dav2_b200 import ops
    from dav2_b200.pose_estimation_model import stack_pairs
    ref, net = _build(seed=3)
    g = torch.Generator().manual_seed(8)
    N, H, W = 5, 70, 98
    rgb = torch.randn(N, 3, H, W, generator=g)
    depth = torch.rand(N, 1, H, W, generator=g) * 0.2 + 0.01
    pairs = stack_pairs(rgb, depth)
    assert pairs.shape == (N - 1, 8, H, W)
    rel = net(pairs.cuda())
    with torch.no_grad():
        rel_ref = ref(pairs)
    assert float((rel.cpu() - rel_ref).abs().max()) / float(rel_ref.abs().max()) < 1e-2
    abs7, T12 = ops.compose_poses(rel, None, want_T12=True)
    np.testing.assert_allclose(abs7.cpu().numpy(), geo.compose_poses(rel.cpu().numpy()), rtol=1e-4, atol=1e-5)
    k4 = geo.scale_intrinsics(geo.SIMCOL_K_475, 475, W)
    xyz, valid, counts = ops.backproject(depth[:, 0].cuda().contiguous(), k4, T12)
    for i in (0, N - 1):
        T = np.eye(4)
        T[:3, :4] = T12[i].cpu().numpy().reshape(3, 4)
        pts, v = geo.backproject(depth[i, 0].numpy(), k4, T)
        err = np.linalg.norm(xyz[i].cpu().numpy()[v] - pts[v], axis=1) / np.linalg.norm(pts[v], axis=1)
        assert err.max() < 1e-5


def test_pose_net_contract():
    from dav2_b200._lib import Dav2Error
    from dav2_b200.pose_estimation_model import PoseEstimationNet
    net = PoseEstimationNet(8)
    keys = set(net.state_dict())
    assert {"backbone.conv1.weight", "backbone.bn1.running_var", "backbone.layer2.0.downsample.0.weight", "backbone.fc.bias",
            "pose_head.2.weight", "pose_head.8.bias"} <= keys
    assert tuple(net.backbone.conv1.weight.shape) == (64, 8, 7, 7)
    with pytest.raises(Dav2Error):
        net.eval()(torch.zeros(1, 8, 64, 64))  # CPU tensor: no fallback


def test_reconstruct_sharded_equals_single():
    """Full pass on 6 frames: 2 emulated ranks (run one after the other, relative poses exchanged by hand) must
    reproduce the single-rank trajectory and clouds -- the halo frame makes pair i local to the owner of frame i."""
    from dav2_b200 import reconstruction, weights
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    _, pose = _build(seed=1)
    depth_model = DepthAnythingV2(**MODEL_CONFIGS["vits"])
    weights.randomize_(depth_model, seed=2)
    depth_model = weights.calibrate_(depth_model.cuda().eval())
    probe = depth_model(torch.randn(1, 3, 70, 98, generator=torch.Generator().manual_seed(1)).cuda())
    assert 1.0 < float(probe.std()) < 9.0  # calibrated: neither flat nor saturated
    g = torch.Generator().manual_seed(3)
    frames = torch.randn(6, 3, 70, 98, generator=g)
    k4 = geo.scale_intrinsics(geo.SIMCOL_K_475, 475, 98)
    one = reconstruction.reconstruct(frames, depth_model, pose, k4, scale=0.02, batch=4)
    assert one["abs"].shape == (6, 7) and one["xyz"].shape == (6, 70 * 98, 3)
    np.testing.assert_allclose(one["abs"].cpu().numpy(), geo.compose_poses(one["rel"].cpu().numpy()), rtol=1e-4, atol=1e-6)
    # emulate 2 ranks: local pair poses of each shard must equal the corresponding slice of the single run
    rel_parts = []
    for r in range(2):
        a, b = 3 * r, 3 * r + 3
        hi = min(b + 1, 6)
        sub = reconstruction.reconstruct(frames[a:hi], depth_model, pose, k4, scale=0.02, batch=4)
        n_pairs = min(b, 5) - a
        rel_parts.append(sub["rel"][:n_pairs])
        assert float((sub["depth"][:b - a] - one["depth"][a:b]).abs().max() / one["depth"].abs().max()) < 1e-2
    rel = torch.cat(rel_parts)
    assert float((rel - one["rel"]).abs().max()) < 2e-2 * float(one["rel"].abs().max()) + 1e-4
    s = reconstruction.calculate_scale_factor(one["rel"], one["rel"] * 2.0)
    assert abs(float(s) - 2.0) < 1e-5
