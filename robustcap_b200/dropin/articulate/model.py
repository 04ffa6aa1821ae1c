from robustcap_b200.model import ParametricModel  # noqa: F401
