"""Cell-level -> sample-level metadata, the ``cna.ut.obs_to_sample`` helper of the reference
(``src/cna/utils/multisample.py:4-11``).

Host-only pandas bookkeeping that sits just before the hot path in the reference's demo notebook
(``samplem = cna.ut.obs_to_sample(d, ['case', 'male', 'batch'], 'id')``); provided so that a notebook
written against ``cna`` runs unchanged on ``import cna_b200 as cna``.
"""
import pandas as pd


def obs_to_sample(d, columns, sid_name, aggregate="mean"):
    """One row per sample id, in order of first appearance in ``d.obs[sid_name]``, holding the
    requested obs columns aggregated within each sample (``aggregate`` as accepted by pandas'
    ``GroupBy.aggregate``).  The index carries no name, like the reference's result."""
    wanted = [columns] if isinstance(columns, str) else list(columns)
    first_seen = pd.unique(d.obs[sid_name])
    per_sample = d.obs.groupby(sid_name)[wanted].aggregate(aggregate)
    out = per_sample.reindex(first_seen)
    out.index.name = None
    out.columns.name = None
    return out
