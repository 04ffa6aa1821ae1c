r"""``import articulate as art`` — drop-in for the reference package (articulate/__init__.py:5-10), hot path only."""
from robustcap_b200.model import ParametricModel
from robustcap_b200 import math
from .evaluator import PositionErrorEvaluator

__all__ = ['ParametricModel', 'math', 'PositionErrorEvaluator']
