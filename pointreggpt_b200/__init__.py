"""pointreggpt_b200 -- B200-native (sm_100a) implementation of PointRegGPT's per-pair
data-generation hot path, behind the reference's own Python API.

Everything numerical runs in hand-written CUDA (libprg.so, C ABI in include/prg.h).
"""
from .geometry import (depth2pc_tensor, pc2depth_tensor, pc2depth_ragged, reproject_tensor,
                       point_cloud, point_cloud_batch, intrinsic_transform, param_vector,
                       random_sample_intrinsic, random_sample_pose, num_to_groups,
                       normalize_to_neg_one_to_one, unnormalize_to_zero_to_one,
                       get_mask_from_img_cond, null_image_condition, occlusion_filter,
                       image_condition)

__all__ = [n for n in dir() if not n.startswith("_")]
