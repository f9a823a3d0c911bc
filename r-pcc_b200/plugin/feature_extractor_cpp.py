"""feature_extractor_cpp (cpp_modules.cpp:596-604); only the variant the tools call."""
from ._np import np, _lib, check, ptr, f32, i32, hw


def extract_features_with_segment(range_image, seg_idx, feature_region, segments, sharp_num, less_sharp_num, flat_num):
    """-> (feature_map (H,W) f32, key_point_map (H,W) int32), cpp_modules.cpp:28-121."""
    seg = i32(seg_idx)
    H, W = hw(seg)
    ri = f32(range_image)
    feat = np.empty((H, W), np.float32)
    kp = np.empty((H, W), np.int32)
    check(_lib.lib().rpcc_op_extract_features_with_segment(ptr(ri), ptr(seg), H, W, int(feature_region), int(segments),
                                                           int(sharp_num), int(less_sharp_num), int(flat_num),
                                                           ptr(feat), ptr(kp)))
    return feat, kp
