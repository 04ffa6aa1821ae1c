r"""``from config import *`` — the constants of the reference's config.py that the hot path and evaluate.py read."""
from robustcap_b200.constants import MP_MASK as mp_mask, VI_MASK as vi_mask, JI_MASK as ji_mask  # noqa: F401

vel_scale = 3
tran_offset = [0, 0.25, 5]


class paths:
    smpl_file = 'models/SMPL_male.pkl'
    smpl_file_f = 'models/SMPL_female.pkl'
    aist_dir = 'data/dataset_work/AIST/'
    aist_tip_dir = 'data/dataset_work/AIST/tip'
    amass_dir = 'data/dataset_work/AMASS/'
    totalcapture_dir = 'data/dataset_work/TotalCapture/'
    pw3d_dir = 'data/dataset_work/3DPW/'
    pw3d_tip_dir = 'data/dataset_work/3DPW/tip'
    pw3d_pip_dir = 'data/dataset_work/3DPW/pip'
    offline_dir = 'data/dataset_work/live/'
    live_dir = 'data/sig_asia/live'
    weight_dir = 'data/weights/'
    j_regressor_dir = 'data/dataset_work/J_regressor_h36m.npy'
