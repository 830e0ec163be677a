"""Drop-in for the reference's `snippets_save` module (snippets_save.py:18-31): the `cov_vv.csv` hand-off in pandas'
`to_csv` / `read_csv` layout (header row of column numbers, index column dropped on load), written and parsed without
pandas.  The implementations live in `cov_producer`."""
from .cov_producer import load_cov_vv, save_cov_vv  # noqa: F401
