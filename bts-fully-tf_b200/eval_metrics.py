"""Eval metrics of the reference (custom_eval_metrics.py:21-88) from ONE fused pass.

`metrics_list_factory(args)` keeps the reference's signature and returns the same list of nine
callables, in the same order and with the same `__name__`s (Keras uses them as metric names):
[silog, abs_rel, log10, rmse, sq_rel, rmse_log, d1, d2, d3], each `(y_true, y_pred) -> scalar`.
The reference runs `pre_eval` and a masked reduction per metric (nine passes over both maps); here
the first metric asked for a given (y_true, y_pred) pair launches one kernel that produces all
nine, and the other eight read the cached result.

`args` needs `min_depth_eval` and `max_depth_eval` (bts_eval.py / bts_train.py argparse names).  The
crop flags (`garg_crop`, `eigen_crop`) are accepted and ignored exactly as in the reference, whose
`ground_truth_mask` helper is defined but never called (custom_eval_metrics.py:27-37).
"""
from . import ops


class _FusedMetrics:
    """Cache of the last (y_true, y_pred) pair's metric vector.  The pair is held by strong reference and compared by
    IDENTITY plus torch's version counters: a freed batch whose addresses the caching allocator hands to the next batch
    can never alias the cached entry (keys made of data_ptr() / id() can).  Kernels that write a tensor through a raw
    pointer (out= buffers, CUDA-graph replays) do not bump the version counter -- call `invalidate()` after such a
    write, or use `all_metrics()` which never caches."""

    def __init__(self, min_depth_eval, max_depth_eval):
        self.lo, self.hi = float(min_depth_eval), float(max_depth_eval)
        self.invalidate()

    def invalidate(self):
        self._yt = self._yp = self._value = None
        self._versions = None

    def __call__(self, y_true, y_pred):
        versions = (y_true._version, y_pred._version)
        if not (self._yt is y_true and self._yp is y_pred and self._versions == versions):
            self._value = ops.eval_metrics(y_true, y_pred, self.lo, self.hi)
            self._yt, self._yp, self._versions = y_true, y_pred, versions
        return self._value


def metrics_list_factory(args):
    fused = _FusedMetrics(args.min_depth_eval, args.max_depth_eval)

    def make(index, name):
        def metric(y_true, y_pred):
            return fused(y_true, y_pred)[index]
        metric.__name__ = name
        return metric

    by_name = {name: make(i, name) for i, name in enumerate(ops.METRIC_NAMES)}
    by_name["silog"].invalidate = fused.invalidate          # reachable from the list for raw-pointer writers
    return [by_name[n] for n in ("silog", "abs_rel", "log10", "rmse", "sq_rel", "rmse_log", "d1", "d2", "d3")]


def all_metrics(y_true, y_pred, min_depth_eval, max_depth_eval):
    """dict name -> 0-d device tensor, plus 'n_valid'."""
    v = ops.eval_metrics(y_true, y_pred, min_depth_eval, max_depth_eval)
    out = {name: v[i] for i, name in enumerate(ops.METRIC_NAMES)}
    out["n_valid"] = v[9]
    return out
