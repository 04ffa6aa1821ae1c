r"""The one evaluator evaluate.py uses on the path next to the hot path (articulate/evaluator.py: PositionErrorEvaluator).
Metric code, host torch ops; kept so that evaluate.py imports resolve."""
import torch


class PositionErrorEvaluator:
    def __init__(self, dimension=3, align_joint=None):
        self.dimension = dimension
        self.align_joint = align_joint

    def __call__(self, p, t):
        p, t = p.view(p.shape[0], -1, self.dimension), t.view(t.shape[0], -1, self.dimension)
        if self.align_joint is not None:
            offset = t[:, self.align_joint] - p[:, self.align_joint]
            if not isinstance(self.align_joint, (list, tuple)):
                offset = offset.unsqueeze(1)
            else:
                offset = offset.mean(dim=1, keepdim=True)
            p = p + offset
        return (p - t).norm(dim=2)
