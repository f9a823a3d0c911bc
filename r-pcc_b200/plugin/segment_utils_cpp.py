"""segment_utils_cpp (cpp_modules.cpp:606-613)."""
from ._np import C, np, _lib, check, ptr, f32, i32, hw


def point_modeling(range_image, seg_idx):
    """mean range per label -> (max_label+1,) f32, cpp_modules.cpp:471-518."""
    seg = i32(seg_idx)
    H, W = hw(seg)
    ri = f32(range_image)
    out = np.empty(256, np.float32)
    K = C.c_int(0)
    check(_lib.lib().rpcc_op_point_modeling(ptr(ri), ptr(seg), H, W, ptr(out), out.size, C.byref(K)))
    return out[:K.value].copy()


def intra_predict(seg_idx, model_param, transform_map):
    """-> (H,W,1) f32, cpp_modules.cpp:248-285."""
    seg = i32(seg_idx)
    H, W = hw(seg)
    mp = f32(model_param)
    tm = f32(transform_map)
    pred = np.empty((H, W, 1), np.float32)
    check(_lib.lib().rpcc_op_intra_predict(ptr(seg), ptr(mp), mp.size // 4, ptr(tm), H, W, ptr(pred)))
    return pred
