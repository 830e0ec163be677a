"""Host side of the rows either side of the placement path: the empirical covariance that
feeds the greedy in the reference, and the on-disk hand-off formats around it.

    reference                                                        here
    ---------------------------------------------------------------  ------------------------------------------
    gp_functions.create_cov_matrix(...)            :1019-1057         create_cov_matrix (one batched encoder call,
    main_architecture_2.py:322-494 (graph while-loops, one (i, j)       one centring pass + SYRK on the device:
      pair per iteration, tfp.stats.covariance :431)                    empirical_cov)
    ..._sampledistribution.py:355-421 decay filter                   cov_taper
    cache_plot_gen_idxs.gen_idxs                   :9-32              gen_idxs / xyz_cov_idxs
    snippets_save.load_cov_vv / save_cov_vv        :18-31             load_cov_vv / save_cov_vv
    main_architecture_2.py:754-769 CSV dumps                         save_placement_csvs

The numeric work (centred Gram, taper) runs on the device through libvgposp.so; the CSV layer is pandas-compatible
text I/O (index column first, header row of column numbers) written without pandas so that a file produced by either
side loads on the other.
"""
import numpy as np

from . import _ffi
from ._ffi import call

DEVICE = 0


# --------------------------------------------------------------------------------------------------
# empirical covariance (f-1)
# --------------------------------------------------------------------------------------------------
def empirical_cov(samples, as_device=False):
    """cov[i, j] = np.cov(samples[i], samples[j], bias=True)[0, 1] for every pair of rows of `samples` [n, S]."""
    m = samples if isinstance(samples, _ffi.DeviceArray) else \
        _ffi.DeviceArray.from_host(np.ascontiguousarray(samples, dtype=np.float64), DEVICE)
    n, s = m.shape
    ld = n + (n % 2)
    out = _ffi.DeviceArray((n, ld), np.float64, DEVICE)
    call("vgp_empirical_cov", DEVICE, m.ptr, n, s, s, out.ptr, ld, None)
    return out if as_device else out.to_host()[:, :n]


def cov_taper(cov_vv, xyz_idxs, beta, cutoff=0.01):
    """The reference's local-kernel decay filter (main_architecture_2_sampledistribution.py:376-394):
    cov[i, j] * exp(-(beta delta_ij)^2 / (2 pi)), factor set to 0 below `cutoff`; delta = distance of the grid
    indices `xyz_idxs` [n, 3].  (The reference's own `tf.cond` at :409-412 swaps its branches and always stores
    zero; this is the filter its `decay_fn` defines.)"""
    a = np.ascontiguousarray(cov_vv, dtype=np.float64)
    n = a.shape[0]
    d = _ffi.DeviceArray.from_host(a, DEVICE)
    idx = _ffi.DeviceArray.from_host(np.ascontiguousarray(xyz_idxs, dtype=np.int32), DEVICE)
    call("vgp_cov_taper", DEVICE, d.ptr, n, n, idx.ptr, float(beta), float(cutoff), None)
    return d.to_host()


def gen_idxs(input_splits, output_file=None):
    """xyz_cov_idxs [N, 3] int32 with line = I2 I1 i0 + I2 i1 + i2 (cache_plot_gen_idxs.py:9-32)."""
    i0, i1, i2 = (int(v) for v in input_splits)
    g = np.indices((i0, i1, i2), dtype=np.int32).reshape(3, -1).T.copy()
    if output_file:
        write_indexed_csv(output_file, g)
    return g


def create_cov_matrix(minmax_x, minmax_y, minmax_z, minmax_pressure, minmax_temperature, SPATIAL_COVER,
                      SPATIAL_COV_PR_TEMP, encoder, sess=None):
    """gp_functions.py:1019-1057: covariance between the SPATIAL_COVER^3 grid locations of the tracer predicted over a
    SPATIAL_COV_PR_TEMP^2 grid of (pressure, temperature) values.

    `encoder(points [S, 5]) -> [S]` maps (i0, i1, i2, pressure, temperature) rows to predicted tracer values (the
    reference samples its VAE encoder point by point inside get_tracers_for_coordloc, :911-928; any predictor fits,
    e.g. `lambda p: vgp_at(p).mean()`).  Location index = i0 + i1 I0 + i2 I0 I1 (:1041-1046).  The reference's
    O(N^2 S) double loop becomes N encoder calls and one device SYRK."""
    n1 = int(SPATIAL_COVER)
    linsp_p = np.linspace(minmax_pressure[0], minmax_pressure[1], SPATIAL_COV_PR_TEMP).reshape(-1, 1)
    linsp_t = np.linspace(minmax_temperature[0], minmax_temperature[1], SPATIAL_COV_PR_TEMP).reshape(-1, 1)
    X, Y = np.meshgrid(linsp_p, linsp_t)
    grid_pt = np.array([X.flatten(), Y.flatten()]).T                     # :1027-1028
    s = grid_pt.shape[0]
    n = n1 ** 3
    tracers = np.empty((n, s))
    for i2 in range(n1):
        for i1 in range(n1):
            for i0 in range(n1):
                pts = np.empty((s, 5))
                pts[:, 0], pts[:, 1], pts[:, 2] = i0, i1, i2
                pts[:, 3:] = grid_pt
                tracers[i0 + i1 * n1 + i2 * n1 * n1] = np.asarray(encoder(pts), dtype=np.float64).reshape(-1)
    return empirical_cov(tracers)


# --------------------------------------------------------------------------------------------------
# CSV hand-off (f-3): pandas' to_csv / read_csv layout, without pandas
# --------------------------------------------------------------------------------------------------
def write_indexed_csv(file_name, array):
    """`pd.DataFrame(array).to_csv(file_name)`: header `,0,1,...`, each row prefixed by its index; floats in repr
    form (round-trip exact, as pandas writes them)."""
    a = np.asarray(array)
    if a.ndim == 1:
        a = a[:, None]
    with open(file_name, "w", encoding="utf-8") as fh:
        fh.write("," + ",".join(str(j) for j in range(a.shape[1])) + "\n")
        integer = np.issubdtype(a.dtype, np.integer)
        for i, row in enumerate(a):
            fh.write(str(i) + "," + ",".join(str(int(v)) if integer else repr(float(v)) for v in row) + "\n")


def read_indexed_csv(file_name, dtype=np.float64):
    """`np.array(pd.read_csv(file_name).iloc[:, 1:])`: skip the header row and the index column."""
    with open(file_name, "r", encoding="utf-8") as fh:
        header = fh.readline()
        ncols = len(header.rstrip("\n").split(",")) - 1
        rows = [line.rstrip("\n").split(",")[1:] for line in fh if line.strip()]
    out = np.array(rows, dtype=np.float64).reshape(len(rows), ncols)
    return out.astype(dtype) if dtype != np.float64 else out


def save_cov_vv(cov_vv, file_name="cov_vv.csv"):
    """snippets_save.py:27-31."""
    write_indexed_csv(file_name, np.asarray(cov_vv, dtype=np.float64))


def load_cov_vv(file_name="cov_vv.csv"):
    """snippets_save.py:18-24."""
    return read_indexed_csv(file_name)


def save_placement_csvs(directory, cov_vv, xyz_cov_idxs, delta_cached_iters, selection_idxs):
    """The four files main_architecture_2.py:754-769 leaves for the plotting scripts."""
    import os
    write_indexed_csv(os.path.join(directory, "cov_vv_small.csv"), np.asarray(cov_vv, dtype=np.float64))
    write_indexed_csv(os.path.join(directory, "placement_algorithm_xyz_cov_idxs.csv"),
                      np.asarray(xyz_cov_idxs, dtype=np.int32))
    write_indexed_csv(os.path.join(directory, "placement_algorithm_cache.csv"),
                      np.asarray(delta_cached_iters, dtype=np.float64))
    write_indexed_csv(os.path.join(directory, "placement_algorithm_selection_idxs.csv"),
                      np.asarray(selection_idxs, dtype=np.int64))
