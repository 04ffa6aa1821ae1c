from robustcap_b200.math import *  # noqa: F401,F403
from robustcap_b200.math import __all__  # noqa: F401
