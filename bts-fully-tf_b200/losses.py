"""si_log_loss of the reference (bts.py:27-41) on the fused decoder-tail kernels.

`si_log_loss_wrapper(dataset)` keeps the reference's name, argument and error behaviour (an unknown
dataset fails the same assertion) and returns `si_log_loss(y_true, y_pred)`; the body is one kernel
forward and one backward behind the C ABI instead of ~10 TF ops (greater, two boolean_masks, two
logs, square, two means, sqrt) and their autodiff.

`depth_silog(logit, y_true, max_depth, dataset)` is the fused form the new decoder uses: it also
takes over the sigmoid of the last Conv2D and the `depth_est` Lambda (bts_decoder.py:102-103), so
the full-resolution logit is read once and depth_est written once.
"""
from . import ops

GT_TH = {"nyu": 0.1, "kitti": 1.0, "matterport": 0.1}     # bts.py:28


def si_log_loss_wrapper(dataset):
    assert dataset in GT_TH                                # bts.py:40

    def si_log_loss(y_true, y_pred):
        return ops.si_log_loss(y_true, y_pred, GT_TH[dataset])

    return si_log_loss


def depth_silog(logit, y_true, max_depth, dataset, workspace=None):
    """Returns (depth_est, loss); gradient flows from `loss` to `logit`.  workspace: see ops.depth_silog."""
    assert dataset in GT_TH
    return ops.depth_silog(logit, y_true, max_depth, GT_TH[dataset], workspace)
