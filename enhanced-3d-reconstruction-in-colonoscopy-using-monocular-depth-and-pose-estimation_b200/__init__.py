"""dav2_b200: B200-native (sm_100a) drop-in for the reference's depth + point-cloud hot path.

Public surface (mirrors the reference's Python call signatures):
  dav2_b200.dpt.DepthAnythingV2                    forward / infer_image / load_state_dict
  dav2_b200.evaluation.compute_errors, compose_poses
  dav2_b200.calculate_metrics.calculate_metrics
  dav2_b200.depth_to_pointcloud.generate_point_cloud, load_camera_intrinsics, load_transformation, PointCloud, ...
  dav2_b200.pose_estimation_model.PoseEstimationNet, dav2_b200.reconstruction.reconstruct, dav2_b200.run.run_frames
  dav2_b200.sharding.frame_range / allreduce_partials / gather_clouds / CloudGather (multi-GPU plumbing)
All compute goes through the C ABI in include/dav2_b200.h (libdav2_b200.so, hand-written CUDA).
There is no CPU or PyTorch fallback: importing works anywhere, computing needs a B200.
"""
__version__ = "0.1.0"
