"""ctypes binding of the C ABI in include/gossipnet_b200.h.

There is NO fallback: if the shared library is missing the import of any
compute entry point raises, and every non-zero return code raises with the
library's own message.
"""
import ctypes
import os

from gossipnet_b200.build import LIB_PATH

c_int = ctypes.c_int
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    'gn_last_error': [],
    'gn_abi_version': [],
    'gn_sm_count': [],
    'gn_iou_dense': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                     c_void_p, c_void_p],
    'gn_neighbor_count': [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p],
    'gn_exclusive_scan': [c_void_p, c_int, c_void_p, c_void_p],
    'gn_neighbor_fill': [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_int, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p],
    'gn_pair_geometry': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_int, c_int, c_float, c_void_p, c_void_p],
    'gn_pwfeat_prep_bytes': [],
    'gn_pwfeat_mlp_fwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'gn_pwfeat_mlp_fwd_ffma': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    'gn_fc_fwd': [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                  c_int, c_void_p, c_int, c_int, c_void_p],
    'gn_block_gather_concat': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p],
    'gn_segment_max': [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p],
    'gn_block_pair_fwd': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                          c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_void_p, c_void_p],
    'gn_block_pair_fwd_hl': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                             c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_void_p, c_void_p],
    'gn_neighbor_count_masks': [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p],
    'gn_neighbor_fill_masks': [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p],
    'gn_pwfeat_mlp_fwd_bf16': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'gn_block_pair_fwd_pipe_bf16': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p],
    'gn_block_det_fwd_img_bf16': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_int, c_int, c_void_p],
    'gn_frcn_boxes': [c_void_p, c_int, c_float, c_int, c_void_p, c_void_p],
    'gn_predict_collapse': [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'gn_rowdot_fwd': [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    'gn_block_pair_fwd_pipe': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p],
    'gn_block_pair_image_bytes': [],
    'gn_block_det_image_bytes': [],
    'gn_prepare_operands': [c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    'gn_block_det_fwd_img': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_int, c_int, c_void_p],
    'gn_block_pair_ab_image_bytes': [],
    'gn_block_pair_fwd_ab': [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p],
    'gn_block_det_fwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                         c_void_p],
    'gn_block_pair_fwd_ffma': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p, c_void_p],
    'gn_detection_matching_workspace_ints': [c_int],
    'gn_detection_matching': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p],
    'gn_loss_fwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                    c_void_p, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p,
                    c_void_p],
    'gn_relu_mask': [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p],
    'gn_add_inplace': [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p],
    'gn_transpose': [c_void_p, c_int, c_int, c_void_p, c_void_p],
    'gn_fc_bwd_weight': [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                         c_int, c_int, c_void_p],
    'gn_segment_max_bwd': [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                           c_void_p],
    'gn_gather_concat_bwd': [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                             c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'gn_adam_step': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_float,
                     c_float, c_float, c_float, ctypes.c_int64, c_float, c_void_p],
    'gn_momentum_step': [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_float,
                         c_float, c_float, c_void_p],
    'gn_roi_pool_fwd': [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                        c_float, c_void_p, c_void_p, c_void_p],
    'gn_roi_pool_bwd': [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                        c_int, c_float, c_void_p, c_void_p],
    'gn_pwfeat_mlp_fwd_hl': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                             c_void_p],
    'gn_block_pair_tma_image_bytes': [],
    'gn_prepare_pair_tma_image': [c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    'gn_block_pair_fwd_tma': [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                              c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'gn_block_pair_fwd_tma_bf16': [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                   c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'gn_selftest_tma': [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    'gn_block_det_fwd_img_u': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                               c_int, c_int, c_void_p],
    'gn_block_det_fwd_tma': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_int, c_void_p],
    'gn_set_pdl': [c_int],
    'gn_clip_gradients': [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p],
    'gn_prepare_fc_images': [c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    'gn_fc_fwd_tc': [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                     c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    'gn_fc_bwd_weight_tc': [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                            c_void_p, c_int, c_int, c_void_p],
    'gn_selftest_store_bw': [c_void_p, ctypes.c_int64, c_int, c_int, c_void_p],
    'gn_selftest_umma': [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    'gn_selftest_umma_ts': [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    'gn_selftest_umma_rate': [c_int, c_int, c_int, c_int, c_void_p, c_void_p],
}
_RESTYPES = {'gn_last_error': ctypes.c_char_p, 'gn_pwfeat_prep_bytes': ctypes.c_int64,
             'gn_block_pair_image_bytes': ctypes.c_int64,
             'gn_block_det_image_bytes': ctypes.c_int64,
             'gn_block_pair_ab_image_bytes': ctypes.c_int64,
             'gn_block_pair_tma_image_bytes': ctypes.c_int64,
             'gn_detection_matching_workspace_ints': ctypes.c_int64}

_lib = None
# number of C-ABI compute calls made so far (each is one kernel launch of this
# library); bench.py reports the count inside its timed region
CALLS = [0]


class GossipnetError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GossipnetError(
                'CUDA extension %s is missing: run `python -c "import __graft_entry__ as g; '
                'g.build()"` (there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def call(name, *args):
    """Call an int-returning entry point; raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    CALLS[0] += 1
    if rc != 0:
        msg = lib.gn_last_error()
        msg = msg.decode() if msg else ''
        if rc == 1:
            raise ValueError('%s: %s' % (name, msg))
        if rc == 3:
            raise NotImplementedError('%s: %s' % (name, msg))
        raise GossipnetError('%s failed (%d): %s' % (name, rc, msg))
