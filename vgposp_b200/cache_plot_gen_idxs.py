"""Drop-in for the reference's `cache_plot_gen_idxs` module (cache_plot_gen_idxs.py:9-32): the [N, 3] int32 table of
grid indices of every covariance row (index = i0 I1 I2 + i1 I2 + i2) that `placement_algorithm_xyz_cov_idxs.csv`
holds.  The implementation lives in `cov_producer`."""
from .cov_producer import gen_idxs  # noqa: F401
