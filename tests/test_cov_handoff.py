"""CPU tests of the on-disk hand-off formats around the placement path: files written by
pandas load here, files written here load in pandas -- `cov_vv.csv`, `placement_algorithm_*.csv`."""
import os

import numpy as np
import pytest

from vgposp_b200 import cov_producer as cp


def test_cov_vv_csv_round_trip(tmp_path):
    dim = 11
    u = np.random.default_rng(0).normal(1, 1, size=(dim, dim))
    cov = u @ u.T + 0.001 * np.eye(dim)                            # snippets_save.py:36-40
    f = str(tmp_path / "test_cov_vv.csv")
    cp.save_cov_vv(cov, f)
    back = cp.load_cov_vv(f)
    assert np.array_equal(back, cov)                               # repr floats: exact round trip


def test_csv_interoperates_with_pandas(tmp_path):
    pd = pytest.importorskip("pandas")
    cov = np.random.default_rng(1).standard_normal((6, 6))
    ours, theirs = str(tmp_path / "ours.csv"), str(tmp_path / "theirs.csv")
    cp.save_cov_vv(cov, ours)
    pd.DataFrame(cov).to_csv(theirs)                               # snippets_save.save_cov_vv
    assert open(ours).read() == open(theirs).read()                # byte-identical files
    assert np.array_equal(cp.load_cov_vv(theirs), cov)
    df = pd.read_csv(ours, encoding="utf-8", engine="c")           # snippets_save.load_cov_vv
    np.testing.assert_allclose(np.array(df.iloc[:, 1:]), cov, rtol=1e-13)      # pandas' default float parser is not exact


def test_gen_idxs_matches_reference_loop(tmp_path):
    splits = [3, 4, 2]
    i0n, i1n, i2n = splits
    want = np.zeros([i0n * i1n * i2n, 3], dtype=np.int32)
    for i0 in range(i0n):                                          # cache_plot_gen_idxs.py:24-28
        for i1 in range(i1n):
            for i2 in range(i2n):
                want[i2n * i1n * i0 + i2n * i1 + i2, :] = [i0, i1, i2]
    f = str(tmp_path / "idx.csv")
    got = cp.gen_idxs(splits, f)
    assert np.array_equal(got, want) and got.dtype == np.int32
    assert np.array_equal(cp.read_indexed_csv(f, np.int32), want)


def test_save_placement_csvs(tmp_path):
    cov = np.eye(4)
    idx = cp.gen_idxs([2, 2, 1])
    cache = np.random.default_rng(2).standard_normal((4, 3))
    sel = np.array([2, 0, 3])
    cp.save_placement_csvs(str(tmp_path), cov, idx, cache, sel)
    names = sorted(os.listdir(tmp_path))
    assert names == ["cov_vv_small.csv", "placement_algorithm_cache.csv", "placement_algorithm_selection_idxs.csv",
                     "placement_algorithm_xyz_cov_idxs.csv"]                   # main_architecture_2.py:754-769
    assert np.array_equal(cp.read_indexed_csv(str(tmp_path / "placement_algorithm_cache.csv")), cache)
    assert np.array_equal(cp.read_indexed_csv(str(tmp_path / "placement_algorithm_selection_idxs.csv"), np.int64)[:, 0], sel)


def test_host_helpers_of_gp_functions():
    import vgposp_b200.gp_functions as gpf
    px, py = np.linspace(0, 1, 4), np.linspace(2, 3, 3)
    h = np.array(np.meshgrid(px, py, sparse=False)).swapaxes(0, -1).reshape(-1, 2)          # gp_functions.py:272-277
    assert np.array_equal(gpf.create_meshgrid(px, py), h)
    assert np.array_equal(gpf.slice_grid_xyz(1, 2, 0, [0., 1., 2.], [3., 4., 5.], [6., 7.]), [1., 5., 6.])
    idx = cp.gen_idxs([3, 3, 3])
    sel = [5, 0, 26, 13, 7, 1, 2, 9]
    assert np.array_equal(gpf.py_get_coord_idxs(sel, idx), idx[sel[:7]])
    c = np.ones((2, 3))
    out = gpf.denormalize_coord(c)
    assert out is c and c[0, 0] == pytest.approx(0.0007434639347162126 * 3000000 + 0.0018159087825037148)


def test_reference_module_names_resolve():
    """`import snippets_save`, `import cache_plot_gen_idxs` in the reference's scripts map to modules of the same name."""
    from vgposp_b200 import cache_plot_gen_idxs, cov_producer, snippets_save
    assert snippets_save.load_cov_vv is cov_producer.load_cov_vv
    assert snippets_save.save_cov_vv is cov_producer.save_cov_vv
    assert cache_plot_gen_idxs.gen_idxs is cov_producer.gen_idxs
