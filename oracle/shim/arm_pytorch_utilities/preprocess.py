"""preprocess restated: Transformer / Compose / PytorchTransformer and the per-dim affine scalers
(util.py:408-423, push_main.py:66-73, peg_main.py:60-67)."""
import abc
import copy
import torch


class Transformer(abc.ABC):
    """Operates on (XU, Y, labels) datasets."""

    def __init__(self):
        self.fitted = False

    def fit(self, XU, Y, labels=None):
        self._fit_impl(XU, Y, labels)
        self.fitted = True

    @abc.abstractmethod
    def _fit_impl(self, XU, Y, labels):
        pass

    @abc.abstractmethod
    def transform(self, XU, Y, labels=None):
        pass

    @abc.abstractmethod
    def transform_x(self, XU):
        pass

    @abc.abstractmethod
    def transform_y(self, Y):
        pass

    @abc.abstractmethod
    def invert_transform(self, Y, X=None):
        pass

    @abc.abstractmethod
    def invert_x(self, X):
        pass

    def update_data_config(self, config):
        pass


class NoTransform(Transformer):
    def _fit_impl(self, XU, Y, labels):
        pass

    def transform(self, XU, Y, labels=None):
        return XU, Y, labels

    def transform_x(self, XU):
        return XU

    def transform_y(self, Y):
        return Y

    def invert_transform(self, Y, X=None):
        return Y

    def invert_x(self, X):
        return X


class Compose(Transformer):
    """Apply in order; invert in reverse order."""

    def __init__(self, transforms):
        self.transforms = transforms
        super().__init__()

    def _fit_impl(self, XU, Y, labels):
        for t in self.transforms:
            t.fit(XU, Y, labels)
            XU, Y, labels = t.transform(XU, Y, labels)

    def transform(self, XU, Y, labels=None):
        for t in self.transforms:
            XU, Y, labels = t.transform(XU, Y, labels)
        return XU, Y, labels

    def transform_x(self, XU):
        for t in self.transforms:
            XU = t.transform_x(XU)
        return XU

    def transform_y(self, Y):
        for t in self.transforms:
            Y = t.transform_y(Y)
        return Y

    def invert_transform(self, Y, X=None):
        # every stage is handed the same (original-space) X: on the MPPI path stage 1 is the identity on
        # the input side, so this is also "X as it was at that stage's input" for the invariant stage
        for t in reversed(self.transforms):
            Y = t.invert_transform(Y, X)
        return Y

    def invert_x(self, X):
        for t in reversed(self.transforms):
            X = t.invert_x(X)
        return X

    def update_data_config(self, config):
        for t in self.transforms:
            t.update_data_config(config)


class SingleTransformer(abc.ABC):
    """Column-wise transformer of one array."""

    @abc.abstractmethod
    def fit(self, X):
        pass

    @abc.abstractmethod
    def transform(self, X):
        pass

    @abc.abstractmethod
    def inverse_transform(self, X):
        pass

    def data_dim_change(self):
        return 0, 0


class NullSingleTransformer(SingleTransformer):
    def fit(self, X):
        pass

    def transform(self, X):
        return X

    def inverse_transform(self, X):
        return X


def _quantile_rows(X, q):
    """Per-column q-quantile by order statistic (k-th smallest, k = round(q*(N-1)))."""
    n = X.shape[0]
    k = int(round(q * (n - 1)))
    return torch.sort(X, dim=0).values[k]


class MinMaxScaler(SingleTransformer):
    def __init__(self, feature_range=(0, 1)):
        self.feature_range = feature_range
        self._scale, self._min = None, None

    def _bounds(self, X):
        return X.min(dim=0).values, X.max(dim=0).values

    def fit(self, X):
        lo, hi = self._bounds(X)
        rng = hi - lo
        rng = torch.where(rng == 0, torch.ones_like(rng), rng)
        n = X.shape[1]
        fr_lo = torch.zeros(n, dtype=X.dtype, device=X.device)
        fr_hi = torch.ones(n, dtype=X.dtype, device=X.device)
        fr = self.feature_range
        if isinstance(fr[0], (list, tuple)) or torch.is_tensor(fr[0]):
            # per-dim ranges for the LEADING dims; remaining dims keep (0,1)  [SURVEY.md §7 #3, E]
            m = len(fr[0])
            fr_lo[:m] = torch.as_tensor(fr[0], dtype=X.dtype)
            fr_hi[:m] = torch.as_tensor(fr[1], dtype=X.dtype)
        else:
            fr_lo[:] = fr[0]
            fr_hi[:] = fr[1]
        self._scale = (fr_hi - fr_lo) / rng
        self._min = fr_lo - lo * self._scale

    def transform(self, X):
        return X * self._scale + self._min

    def inverse_transform(self, X):
        return (X - self._min) / self._scale


class RobustMinMaxScaler(MinMaxScaler):
    """Min-max on the [p, 1-p] quantiles instead of the extremes (p = 0.025)."""

    def __init__(self, dims_share_scale=None, feature_range=(0, 1), percentile=0.025):
        super().__init__(feature_range=feature_range)
        self.percentile = percentile
        self.dims_share_scale = dims_share_scale

    def _bounds(self, X):
        return _quantile_rows(X, self.percentile), _quantile_rows(X, 1 - self.percentile)


class PytorchTransformer(Transformer):
    """method_x on the XU columns, method_y on the Y columns (defaults to a copy of method_x's class)."""

    def __init__(self, method_x: SingleTransformer, method_y: SingleTransformer = None):
        self.method_x = method_x
        self.method_y = method_y if method_y is not None else copy.deepcopy(method_x)
        super().__init__()

    def _fit_impl(self, XU, Y, labels):
        self.method_x.fit(XU)
        self.method_y.fit(Y)

    def transform(self, XU, Y, labels=None):
        return self.method_x.transform(XU), self.method_y.transform(Y), labels

    def transform_x(self, XU):
        return self.method_x.transform(XU)

    def transform_y(self, Y):
        return self.method_y.transform(Y)

    def invert_transform(self, Y, X=None):
        return self.method_y.inverse_transform(Y)

    def invert_x(self, X):
        return self.method_x.inverse_transform(X)


class DatasetPreprocessor:
    """What `ds.preprocessor` is: a thin holder with `.tsf` (hybrid_model.py:29, model.py:157-164,200)."""

    def __init__(self, tsf: Transformer):
        self.tsf = tsf

    def transform_x(self, XU):
        return self.tsf.transform_x(XU)

    def transform_y(self, Y):
        return self.tsf.transform_y(Y)

    def invert_transform(self, Y, X=None):
        return self.tsf.invert_transform(Y, X)

    def invert_x(self, X):
        return self.tsf.invert_x(X)

    def update_data_config(self, config):
        return self.tsf.update_data_config(config)
