r"""``import utils`` — the one function of the reference's utils.py on the hot path (utils.py:129-135)."""
import torch
from robustcap_b200.constants import MP_MASK

mp_mask = torch.tensor(MP_MASK)


def sync_mp3d_from_smpl(vert, joint):
    syn_3d = vert[:, mp_mask.to(vert.device)]
    syn_3d[:, 11:17] = joint[:, 16:22].clone()
    syn_3d[:, 23:25] = joint[:, 1:3].clone()
    syn_3d[:, 25:27] = joint[:, 4:6].clone()
    syn_3d[:, 27:29] = joint[:, 7:9].clone()
    return syn_3d


from robustcap_b200.metrics import reconstruction_error  # noqa: E402,F401  (utils.py:195-203, used by evaluate.cal_mpjpe)
