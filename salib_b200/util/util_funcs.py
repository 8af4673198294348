"""Parameter-file reader (reference: src/SALib/util/util_funcs.py:33-115).

File format, one parameter per line, columns separated by a sniffed delimiter::

    name  lower_bound  upper_bound  [group|NA]  [dist]

Lines whose name starts with ``#`` are comments.  Returns the SALib problem dict."""
import csv
from typing import Dict, Optional

__all__ = ["read_param_file", "handle_seed"]

_FIELDS = ("name", "lower_bound", "upper_bound", "group", "dist")


def read_param_file(filename: str, delimiter: Optional[str] = None) -> Dict:
    names, bounds, groups, dists = [], [], [], []
    with open(filename, "r") as fh:
        dialect = csv.Sniffer().sniff(fh.read(1024), delimiters=delimiter)
        fh.seek(0)
        for row in csv.DictReader(fh, fieldnames=list(_FIELDS), dialect=dialect):
            name = row["name"]
            if name.strip().startswith("#"):
                continue
            names.append(name)
            bounds.append([float(row["lower_bound"]), float(row["upper_bound"])])
            grp = row["group"]
            groups.append(name if grp is None or grp == "NA" else grp)
            dists.append("unif" if row["dist"] is None else row["dist"])
    if groups == names:
        groups = None
    elif len(set(groups)) == 1:
        raise ValueError(
            """Only one group defined, results will not be
            meaningful"""
        )
    if all(d == "unif" for d in dists):
        dists = None  # uniform everywhere: plain bounds scaling
    return {"names": names, "bounds": bounds, "num_vars": len(names), "groups": groups, "dists": dists}


def handle_seed(seed):
    """Seed-like -> ``numpy.random.Generator`` (reference util_funcs.py:184-202): a Generator is spawned, ints /
    None / BitGenerator / SeedSequence / int sequences seed a new one, anything else is a ValueError."""
    import numpy as np

    if isinstance(seed, np.random.Generator):
        return seed.spawn(1)[0]
    if isinstance(seed, (np.random.BitGenerator, np.random.SeedSequence, int, type(None), list, tuple, np.ndarray)):
        return np.random.default_rng(seed)
    raise ValueError(f"Unsupported seed type: {type(seed)}")
