"""BPR loss  -logsigmoid(pos - neg)  (divrec/losses/log_sigmoid_difference_loss.py:6-14).

Reference quirk kept on purpose: its constructor forwards `*kwargs` (the keys) instead of `**kwargs`,
so `reduction=` / `reduce=` never reach ScoreWithReduction and the loss is ALWAYS a mean.

pair_wise() runs divrec_bpr_loss_fwd / _bwd (elementwise kernels); when the scores come from
MatrixFactorization the train loop uses the fused gather + loss + scatter-add path instead
(divrec.train.bpr_step), which never materialises the gathered rows.
"""
import torch

from .. import _native
from .base_losses import PairWiseLoss


class _NegLogSigmoidDiff(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positives, negatives):
        ctx.save_for_backward(positives, negatives)
        return _native.bpr_loss_fwd(positives, negatives)

    @staticmethod
    def backward(ctx, g):
        positives, negatives = ctx.saved_tensors
        gp, gn = _native.bpr_loss_bwd(positives, negatives, g.contiguous())
        return gp, gn


class LogSigmoidDifferenceLoss(PairWiseLoss):
    def __init__(self, *args, **kwargs):
        PairWiseLoss.__init__(self, *args, *kwargs)

    def pair_wise(self, positives: torch.Tensor, negatives: torch.Tensor) -> torch.Tensor:
        return _NegLogSigmoidDiff.apply(positives, negatives)
