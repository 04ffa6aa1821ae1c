r"""``from net.sig_mp import Net`` — the B200 implementation behind the reference module name (net/sig_mp.py:23-299)."""
from robustcap_b200.net import Net, get_bbox_scale, sync_mp3d  # noqa: F401
