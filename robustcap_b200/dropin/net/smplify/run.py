r"""``from net.smplify.run import smplify_runner`` (net/smplify/run.py:6-35)."""
from robustcap_b200.smplify import smplify_runner  # noqa: F401
