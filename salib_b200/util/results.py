"""Result container (reference: src/SALib/util/results.py:5-38)."""
import pandas as pd


class ResultDict(dict):
    """Dictionary of sensitivity indices.  ``analyze`` attaches ``.problem`` and a bound ``to_df``."""

    def to_df(self):
        """All entries except ``names`` as one DataFrame indexed by ``names``."""
        cols = {k: v for k, v in self.items() if k != "names"}
        return pd.DataFrame(cols, index=self["names"])

    def plot(self, ax=None):
        """Bar charts of the DataFrame view (matplotlib is an optional dependency)."""
        import matplotlib.pyplot as plt  # noqa: F401  (ImportError tells the user what is missing)

        dfs = self.to_df()
        if not isinstance(dfs, (list, tuple)):
            dfs = [dfs]
        if ax is None:
            _, ax = plt.subplots(1, len(dfs))
        axes = [ax] if len(dfs) == 1 and not isinstance(ax, (list, tuple, __import__("numpy").ndarray)) else ax
        for a, df in zip(axes, dfs):
            conf_cols = [c for c in df.columns if c.endswith("_conf")]
            val_cols = [c for c in df.columns if not c.endswith("_conf")]
            err = df[conf_cols].rename(columns=lambda c: c[: -len("_conf")]) if conf_cols else None
            df[val_cols].plot(kind="bar", yerr=err, ax=a)
        return ax
