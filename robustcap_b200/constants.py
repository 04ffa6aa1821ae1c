"""Index tables and constants of the hot path (bit-exact parity items).

Reference: ``config.py:97-101`` (``vel_scale``, ``mp_mask``, ``vi_mask``, ``ji_mask``), the SMPL
kinematic tree ``articulate/model.py:38-39`` (== ``config.py:64-78`` first 24 entries) and the
MediaPipe-33 synthesis rule ``net/sig_mp.py:287-299`` / ``utils.py:129-135``.
"""

VEL_SCALE = 3
FPS = 60
NUM_JOINTS = 24
NUM_IMU = 6
NUM_KP = 33
NUM_VERTS = 6890

# SMPL parent table (parent[0] is the base; the pickle stores 2**32-1 there).
SMPL_PARENT = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# config.py:99-101
MP_MASK = [332, 2809, 2800, 455, 6260, 3634, 3621, 583, 4071, 45, 3557, 1873, 4123, 1652, 5177, 2235, 5670,
           2673, 6133, 2319, 5782, 2746, 6191, 3138, 6528, 1176, 4662, 3381, 6727, 3387, 6787, 3226, 6624]
VI_MASK = [1961, 5424, 1176, 4662, 411, 3021]
JI_MASK = [18, 19, 4, 5, 15, 0]

# MediaPipe rows overwritten by SMPL joints: row -> joint  (net/sig_mp.py:295-298)
MP_JOINT_OVERRIDE = {11: 16, 12: 17, 13: 18, 14: 19, 15: 20, 16: 21, 23: 1, 24: 2, 25: 4, 26: 5, 27: 7, 28: 8}

# temporal_smplify.py:92-94
SMPLIFY_IGNORED_KP = [1, 2, 3, 4, 5, 6, 7, 8, 9, 31, 32]
SMPLIFY_IGNORED_KP_HEAD = [31, 32]


def mp_source_table():
    """For each of the 33 synthesised keypoints: (is_joint, index).

    ``is_joint == 1`` -> take SMPL joint ``index``; else skin mesh vertex ``index``.
    """
    out = []
    for r in range(NUM_KP):
        if r in MP_JOINT_OVERRIDE:
            out.append((1, MP_JOINT_OVERRIDE[r]))
        else:
            out.append((0, MP_MASK[r]))
    return out


# Sub-network table of Net (net/sig_mp.py:52-81): name -> (input, hidden, output)
NET_DIMS = {
    'rnn2': (72, 512, 69),
    'rnn3': (141, 512, 3),
    'rnn4': (171, 1280, 69),
    'rnn6': (240, 1024, 3),
    'rnn7': (141, 512, 144),
    'rnn8': (141, 512, 2),
}
NET_ORDER = ['rnn2', 'rnn3', 'rnn4', 'rnn6', 'rnn7', 'rnn8']
INIT_NET_DIMS = (69, 512, 1024, 2048)  # rnn2.init_net (rnn.py:195-201)
