"""ctypes binding of the C-ABI library (``include/robustcap_b200.h``) — plumbing only.

The library is built in-tree (``robustcap_b200/csrc/librobustcap_b200.so``) by :func:`build`.  There is NO
fallback: if the library is missing or no CUDA device is usable, every compute entry point raises.
"""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO_PATH = os.path.join(CSRC, 'librobustcap_b200.so')
SOURCES = ['rotations.cu', 'kinematics.cu', 'fusion.cu', 'smplify.cu', 'gemm_tc.cu', 'phase_tc.cu', 'seq_tc.cu', 'stream.cu', 'stream2.cu', 'metrics.cu', 'pipeline.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']

_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
vp = ctypes.c_void_p
i64 = ctypes.c_int64
i32 = ctypes.c_int32


class NetConfig(ctypes.Structure):
    """``rc_net_config`` (Net class attributes, net/sig_mp.py:27-45, 91-93)."""
    _fields_ = [('conf_lo', ctypes.c_double), ('conf_hi', ctypes.c_double), ('tran_filter_num', ctypes.c_double),
                ('contact_threshold', ctypes.c_float), ('height_threshold', ctypes.c_float),
                ('distance_threshold', ctypes.c_float), ('use_flat_floor', i32), ('live', i32),
                ('update_vision_freq', i32)]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    deps.append(os.path.join(HERE, '..', 'include', 'robustcap_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    tmp = '%s.%d.tmp' % (SO_PATH, os.getpid())        # several ranks may build at once: private temp file + atomic rename
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get('RC_NVCC_EXTRA', '').split() + ['-o', tmp] + sources()   # RC_NVCC_EXTRA: -D switches of timing experiments
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    os.replace(tmp, SO_PATH)
    return SO_PATH


_SIGS = {
    'rc_version': (ctypes.c_char_p, []),
    'rc_last_error': (ctypes.c_char_p, []),
    'rc_launch_count': (i64, []),
    'rc_r6d_to_rotmat': (i32, [vp, vp, i64, vp]),
    'rc_rotmat_to_r6d': (i32, [vp, vp, i64, vp]),
    'rc_axis_angle_to_rotmat': (i32, [vp, vp, i64, vp]),
    'rc_rotmat_to_axis_angle': (i32, [vp, vp, i64, vp]),
    'rc_batch_rodrigues': (i32, [vp, vp, i64, vp]),
    'rc_quat_to_rotmat': (i32, [vp, vp, i64, vp]),
    'rc_quat_to_axis_angle': (i32, [vp, vp, i64, vp]),
    'rc_axis_angle_to_quat': (i32, [vp, vp, i64, vp]),
    'rc_quat_product': (i32, [vp, vp, vp, i64, vp]),
    'rc_tree_fk_R': (i32, [vp, vp, vp, i32, i64, vp]),
    'rc_tree_ik_R': (i32, [vp, vp, vp, i32, i64, vp]),
    'rc_tree_fk_T': (i32, [vp, vp, vp, i32, i64, vp]),
    'rc_tree_ik_T': (i32, [vp, vp, vp, i32, i64, vp]),
    'rc_tree_bone_to_joint': (i32, [vp, vp, vp, i32, i64, vp]),
    'rc_tree_joint_to_bone': (i32, [vp, vp, vp, i32, i64, vp]),
    'rc_model_create': (i32, [ctypes.POINTER(vp), vp, vp, vp, i32, vp, vp]),
    'rc_model_destroy': (None, [vp]),
    'rc_model_forward_kinematics': (i32, [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp]),
    'rc_model_keypoints': (i32, [vp, vp, vp, i64, vp, vp, vp]),
    'rc_net_default_config': (None, [ctypes.POINTER(NetConfig), i32]),
    'rc_net_create': (i32, [ctypes.POINTER(vp), vp, ctypes.POINTER(NetConfig)]),
    'rc_net_destroy': (None, [vp]),
    'rc_net_set_config': (i32, [vp, ctypes.POINTER(NetConfig)]),
    'rc_net_set_tensor': (i32, [vp, ctypes.c_char_p, vp, i64]),
    'rc_net_finalize': (i32, [vp]),
    'rc_net_weight_bytes': (i64, [vp]),
    'rc_net_set_gemm_mode': (i32, [vp, i32]),
    'rc_state_create': (i32, [ctypes.POINTER(vp), vp, i32]),
    'rc_state_destroy': (None, [vp]),
    'rc_state_reset': (i32, [vp, vp]),
    'rc_state_set_gravity': (i32, [vp, vp, vp]),
    'rc_forward_step': (i32, [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp]),
    'rc_forward_online': (i32, [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]),
    'rc_forward_sequence': (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, vp]),
    'rc_forward_sequence_host': (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    'rc_state_debug_output': (i32, [vp, i32, vp, vp]),
    'rc_state_debug_phase_trace': (i32, [vp, i32, i32, vp, i32, vp]),
    'rc_state_debug_lstm': (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    'rc_smplify_create': (i32, [ctypes.POINTER(vp), vp, vp, vp, vp, i32]),
    'rc_smplify_destroy': (None, [vp]),
    'rc_smplify_loss_grad': (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    'rc_smplify_run': (i32, [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, i32, ctypes.c_float, vp, vp, vp, vp]),
    'rc_metrics_mpjpe': (i32, [vp, vp, i32, vp, vp, i64, i32, vp, vp]),
    'rc_pack_inputs': (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_float, ctypes.c_float, vp, vp, vp, vp, vp, vp]),
    'rc_synthesize_imu': (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, vp, vp, vp, vp]),
    'rc_live_parse_frame': (i32, [ctypes.c_char_p, i32, vp, vp, vp, vp]),
    'rc_live_format_pose': (i32, [vp, vp, ctypes.c_char_p, i32]),
    'rc_live_parse_imu_packet': (i32, [ctypes.c_char_p, i32, i32, vp, vp, vp]),
    'rc_state_set_branch_log': (i32, [vp, vp]),
    'rc_net_set_seq_options': (i32, [vp, i32, i32]),
    'rc_state_debug_seq_stats': (i32, [vp, vp, i32]),
    'rc_profile_enable': (i32, [vp, i32]),
    'rc_profile_collect': (i32, [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double)]),
}
_OPTIONAL_SIGS = {}


def exported_symbols():
    """Names ``include/robustcap_b200.h`` declares (parsed from the header)."""
    import re
    hdr = open(os.path.join(HERE, '..', 'include', 'robustcap_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(rc_[a-z0-9_]+)\s*\(', hdr)))


def load():
    """Load the shared library (no CUDA call is made here). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError('robustcap_b200: %s is missing — run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(there is no CPU fallback)' % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in list(_SIGS.items()) + list(_OPTIONAL_SIGS.items()):
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in _SIGS:
                raise
            continue
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError('robustcap_b200 native call failed (%d): %s' % (rc, load().rc_last_error().decode()))


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('robustcap_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def dptr(t):
    """Device pointer of a contiguous float32/int32 CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return t.data_ptr()


def hptr(t):
    if t is None:
        return None
    assert (not t.is_cuda) and t.is_contiguous(), 'expected a contiguous CPU tensor'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
