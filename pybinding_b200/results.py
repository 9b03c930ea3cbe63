"""Result containers: the shape contract of `pybinding.results.Series` without the plotting
(reference: pybinding/results.py:162-213)."""
import numpy as np

__all__ = ["Series"]


class Series:
    """A series of data points determined by a common relation, i.e. y = f(x)

    Attributes
    ----------
    variable : array_like
        Independent variable for which the data was computed.
    data : array_like
        An array of values which were computed as a function of `variable`.
        It can be 1D or 2D. In the latter case each column represents the result
        of a different function applied to the same `variable` input.
    labels : dict
        Plot labels: 'variable', 'data', 'title' and 'columns'.
    """

    def __init__(self, variable, data, labels=None):
        self.variable = np.atleast_1d(variable)
        self.data = np.atleast_1d(data)
        self.labels = dict(variable="x", data="y", columns="")
        self.labels.update(labels or {})

    def with_data(self, data):
        """Return a copy of this result object with different data"""
        return Series(self.variable.copy(), data, dict(self.labels))

    def reduced(self):
        """Return a copy where the data is summed over the columns"""
        return self.with_data(self.data.sum(axis=1)) if self.data.ndim == 2 else self.with_data(self.data)

    def __repr__(self):
        return "Series(variable.shape={}, data.shape={})".format(self.variable.shape, self.data.shape)
