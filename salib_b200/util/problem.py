"""``ProblemSpec`` facade for the Sobol' path (reference: src/SALib/util/problem.py:37-652).

Only the boundary the hot path sits behind is mirrored: construction checks, the ``samples`` /
``results`` / ``analysis`` properties, ``sample`` / ``evaluate`` / ``analyze`` chaining with any callable,
and the bound shortcuts ``sample_sobol``, ``sample_saltelli`` and ``analyze_sobol``
(the reference discovers them by walking its packages, :576-598).  Process-pool variants
(``evaluate_parallel``, ``evaluate_distributed``, ``analyze_parallel``) are replaced by GPU sharding:
``nprocs`` is accepted and ignored.

Arrays may stay on the device: ``sample_sobol(..., device=True)`` stores a CUDA tensor, ``evaluate``
hands it to the model unchanged, and ``analyze_sobol`` consumes a CUDA tensor of results directly.
"""
from functools import partial

import numpy as np

from .results import ResultDict


def _rows(a):
    return a.shape[0]


class ProblemSpec(dict):
    """Dictionary-like SALib problem specification with a fluent sample/evaluate/analyze interface."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert "names" in self, "Names not defined"
        assert "bounds" in self, "Bounds not defined"
        assert len(self["bounds"]) == len(self["names"]), (
            f"Number of bounds do not match number of names\n"
            f"Number of names: {len(self['names'])} | {self['names']}\n"
            f"Number of bounds: {len(self['bounds'])}")
        self._samples = None
        self._results = None
        self._analysis = None
        self["num_vars"] = len(self["names"])
        if "groups" not in self or not self["groups"]:
            self["groups"] = self["names"]
        elif len(self["groups"]) != self["num_vars"]:
            raise ValueError("Number of entries in 'groups' should be the same as in 'names'")

    # -- properties ------------------------------------------------------------------------
    @property
    def samples(self):
        return self._samples

    @samples.setter
    def samples(self, vals):
        cols = vals.shape[1]
        if cols != self["num_vars"]:
            raise ValueError("Mismatched sample size: Expected {} cols, got {}".format(self["num_vars"], cols))
        self._samples = vals
        self._results = None

    @property
    def results(self):
        return self._results

    @results.setter
    def results(self, vals):
        cols = 1 if len(vals.shape) == 1 else vals.shape[1]
        if self.get("outputs") is None:
            self["outputs"] = ["Y"] if cols == 1 else [f"Y{i}" for i in range(1, cols + 1)]
        elif cols != len(self["outputs"]):
            raise ValueError("Mismatched sample size: Expected {} cols, got {}".format(self["outputs"], cols))
        self._results = vals

    @property
    def analysis(self):
        return self._analysis

    # -- fluent interface ------------------------------------------------------------------
    def sample(self, func, *args, **kwargs):
        """``self.samples = func(self, *args, **kwargs)``; clears results and analysis."""
        self._analysis = None
        self._results = None
        self._samples = func(self, *args, **kwargs)
        return self

    def set_samples(self, samples):
        self.samples = samples
        return self

    def set_results(self, results):
        if self.samples is not None:
            assert _rows(self.samples) == _rows(results), (
                "Provided result array does not match existing number of existing samples!")
        self.results = results
        return self

    def evaluate(self, func, *args, **kwargs):
        """``self.results = func(self.samples, *args, **kwargs)`` (``nprocs`` accepted, ignored)."""
        kwargs.pop("nprocs", None)
        self.results = func(self._samples, *args, **kwargs)
        return self

    def analyze(self, func, *args, **kwargs):
        """Run ``func(self, [X=samples,] Y=results, ...)``; one ResultDict per output column."""
        if self["num_vars"] == 1 or (self["groups"] and len("groups") == 1):
            raise ValueError(
                "There is only a single parameter or group defined. There is "
                "no point in conducting sensitivity analysis as any and all"
                "effect(s) will be mapped to the single parameter/group.")
        kwargs.pop("nprocs", None)
        if self._results is None:
            raise RuntimeError("Model not yet evaluated")
        code = getattr(func, "__code__", None)
        if code is not None and "X" in code.co_varnames:
            func = partial(func, *args, X=self._samples, **kwargs)
        else:
            func = partial(func, *args, **kwargs)
        if self.get("outputs") is None:
            if len(self._results.shape) == 1:
                self["outputs"] = ["Y"]
            else:
                self["outputs"] = [f"Y{i}" for i in range(1, self._results.shape[1] + 1)]
        if len(self["outputs"]) > 1:
            self._analysis = {out: func(self, Y=self._results[:, i]) for i, out in enumerate(self["outputs"])}
        else:
            self._analysis = func(self, Y=self._results)
        return self

    # -- Sobol' shortcuts (the reference binds these by name at construction, :567-598) ---------
    def sample_sobol(self, *args, device=False, **kwargs):
        from ..sample import sobol

        return self.sample(sobol.sample_device if device else sobol.sample, *args, **kwargs)

    def sample_saltelli(self, *args, **kwargs):
        from ..sample import saltelli

        return self.sample(saltelli.sample, *args, **kwargs)

    def analyze_sobol(self, *args, **kwargs):
        """``analyze(sobol.analyze, ...)``; with several outputs the columns are analysed in one pass that
        shares the resample multiplicities (same results as per-column calls with the same ``seed``)."""
        from ..analyze import sobol

        res = self._results
        if res is None or len(res.shape) == 1 or res.shape[1] == 1 or args:
            return self.analyze(sobol.analyze, *args, **kwargs)
        if self["num_vars"] == 1:
            return self.analyze(sobol.analyze, *args, **kwargs)  # raises the reference's ValueError
        kwargs.pop("nprocs", None)
        for ignored in ("parallel", "n_processors"):
            kwargs.pop(ignored, None)
        show = kwargs.pop("print_to_console", False)
        if self.get("outputs") is None:
            self["outputs"] = [f"Y{i}" for i in range(1, res.shape[1] + 1)]
        outs = sobol.analyze_outputs(self, [res[:, i] for i in range(res.shape[1])], **kwargs)
        self._analysis = dict(zip(self["outputs"], outs))
        if show:
            for S in outs:
                for df in S.to_df():
                    print(df)
        return self

    def sample_finite_diff(self, *args, **kwargs):
        from ..sample import finite_diff

        return self.sample(finite_diff.sample, *args, **kwargs)

    def analyze_dgsm(self, *args, **kwargs):
        from ..analyze import dgsm

        return self.analyze(dgsm.analyze, *args, **kwargs)

    def analyze_discrepancy(self, *args, **kwargs):
        from ..analyze import discrepancy

        return self.analyze(discrepancy.analyze, *args, **kwargs)

    # -- views -----------------------------------------------------------------------------
    def to_df(self):
        res = self._analysis
        if isinstance(res, ResultDict):
            return res.to_df()
        if isinstance(res, dict):
            return [r.to_df() for r in res.values()]
        raise RuntimeError("Analysis not yet conducted")

    def __repr__(self):
        return str(self)

    def __str__(self):
        parts = []
        if self._samples is not None:
            shp = tuple(self._samples.shape)
            nr, nx = shp if len(shp) > 1 else (shp[0], 1)
            parts.append(f"Samples:\n\t{nx} parameters: {self['names']}\n\t{nr} samples\n")
        if self._results is not None:
            shp = tuple(self._results.shape)
            nr, ny = shp if len(shp) > 1 else (shp[0], 1)
            parts.append(f"Outputs:\n\t{ny} outputs: {self['outputs']}\n\t{nr} evaluations\n")
        if self._analysis is not None:
            txt = "Analysis:\n"
            res = self._analysis
            items = [(None, res)] if isinstance(res, ResultDict) else list(res.items())
            for name, r in items:
                if name is not None:
                    txt += f"{name}:\n"
                dfs = r.to_df()
                for df in dfs if isinstance(dfs, (list, tuple)) else [dfs]:
                    txt += f"{df}\n"
            parts.append(txt)
        return "".join(parts) or "ProblemSpec does not currently contain any samples, evaluations or results."
