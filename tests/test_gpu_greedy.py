"""GPU parity: greedy MI placement through the C-ABI and the drop-in module, against the golden vectors from the
reference's own code, the CPU oracle on seeded inputs, and size-independent properties at larger n."""
import contextlib
import ctypes
import io

import numpy as np
import pytest

from oracle import greedy_oracle as go
from vgposp_b200 import _ffi, greedy
import vgposp_b200.placement_algorithm2 as alg2

pytestmark = pytest.mark.gpu
D = 0


def cloud_cov(n, seed, nugget=1e-2):
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / max(n, 1000)) ** (1 / 3)          # keeps cond(Sigma) n-independent
    d = x[:, None, :] - x[None, :, :]
    return np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + nugget * np.eye(n)


def test_golden_selections_bit_exact(golden, golden_name):
    cov, k = golden.cov(golden_name), golden.cases[golden_name]["k"]
    sel, scores, steps, secs = greedy.place_single(cov, k, D, want_step_scores=True)
    assert [int(s) for s in sel] == golden.selection(golden_name)          # bit-exact index parity
    ref = golden.step_scores(golden_name)
    if ref is not None:
        tol = 1e-10 if golden_name.startswith(("fixture", "expquad")) else 1e-7
        np.testing.assert_allclose(steps, ref, rtol=tol, equal_nan=True)
        np.testing.assert_allclose(scores, [np.nanmax(r) for r in ref], rtol=tol)


def test_dropin_api_types_and_prints(golden):
    cov = alg2.cov_vv_4x4()
    a1 = alg2.placement_algorithm_1(cov, 4)
    assert a1 == [2, 1, 3, 0] and all(isinstance(v, np.int64) for v in a1) and isinstance(a1, list)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        a2 = alg2.placement_algorithm_2(cov, 4)
    assert a2 == a1
    assert buf.getvalue().splitlines() == golden.cases["fixture4x4"]["alg2_stdout"]     # same print trace
    name = "expquad_n100"
    with contextlib.redirect_stdout(io.StringIO()) as out:
        sel = alg2.placement_algorithm_2(golden.cov(name), golden.cases[name]["k"])
    assert sel == golden.selection(name)
    assert out.getvalue().splitlines() == golden.cases[name]["alg2_stdout"]


def test_dropin_nominator_denominator_argmax(golden):
    cov = golden.cov("expquad_n50")
    A, A_bar = [44, 35], [v for v in range(50) if v not in (44, 35)]
    for y in (0, 7, 29):
        np.testing.assert_allclose(alg2.nominator(y, A, cov), go.literal_nominator(y, A, cov), rtol=1e-10)
        np.testing.assert_allclose(alg2.denominator(y, A_bar, cov), go.literal_denominator(y, A_bar, cov), rtol=1e-10)
    assert alg2.nominator(3, [], cov).shape == (1, 1) and alg2.nominator(3, [], cov)[0, 0] == cov[3, 3]
    y_st, delta = alg2.argmax_(A, A_bar, np.arange(50), cov)
    assert y_st == 29                                                      # third golden pick
    np.testing.assert_allclose(delta, np.nanmax(golden.step_scores("expquad_n50")[2]), rtol=1e-10)
    # arbitrary (non-greedy) A goes through the per-candidate route
    y2, d2 = alg2.argmax_([1, 2], [v for v in range(50) if v not in (1, 2)], np.arange(50), cov)
    lit = max((go.literal_delta(y, [1, 2], [v for v in range(50) if v not in (1, 2)], cov), -y) for y in range(50)
              if y not in (1, 2))
    assert y2 == -lit[1] and abs(d2 - lit[0]) < 1e-10 * lit[0]
    inv = alg2.call_pinv(cov)
    np.testing.assert_allclose(inv @ cov, np.eye(50), atol=1e-11)


def test_dropin_error_behaviour():
    cov = alg2.cov_vv_4x4()
    with pytest.raises(ValueError, match="not in list"):                  # reference: A_bar.remove(-1), :144
        alg2.placement_algorithm_1(cov, 5)
    assert alg2.placement_algorithm_1(cov, 0) == []
    # rank-deficient PSD input: the pseudo-inverse path gives what the reference's pinv gives (all deltas 0 -> [0, 1])
    assert alg2.placement_algorithm_1(np.ones((6, 6)), 2) == [0, 1]
    with pytest.raises(np.linalg.LinAlgError):                             # indefinite: neither path applies
        alg2.placement_algorithm_1(np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), 2)
    with pytest.raises(ValueError):
        alg2.placement_algorithm_1(np.ones((3, 4)), 1)


@pytest.mark.parametrize("n,k,seed", [(257, 12, 1), (1000, 10, 2), (2000, 25, 3), (3000, 6, 4)])
def test_matches_cpu_oracle_on_seeded_clouds(n, k, seed):
    cov = cloud_cov(n, seed)
    want_sel, want_scores, want_steps, gaps = go.incremental_greedy(cov, k, return_all_scores=True)
    sel, scores, steps, _ = greedy.place_single(cov, k, D, want_step_scores=True)
    assert gaps.min() > 1e-12, "near-tie in the test input"
    assert [int(s) for s in sel] == want_sel
    np.testing.assert_allclose(scores, want_scores, rtol=1e-9)
    np.testing.assert_allclose(steps, want_steps, rtol=1e-9, equal_nan=True)


def test_tf_graph_compat_mode():
    cov = cloud_cov(300, 9)
    want = go.incremental_greedy(cov, 6, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH, return_all_scores=True)
    sel, scores, steps, _ = greedy.place_single(cov, 6, D, small=greedy.GUARD_TF_GRAPH, jitter=greedy.JITTER_TF_GRAPH,
                                                want_step_scores=True)
    assert [int(s) for s in sel] == want[0]
    np.testing.assert_allclose(steps, want[2], rtol=1e-9, equal_nan=True)


def test_non_contiguous_and_float32_inputs(golden):
    cov = golden.cov("expquad_n100")
    big = np.zeros((100, 130))
    big[:, :100] = cov
    sel, *_ = greedy.place_single(big[:, :100], 8, D)                      # row stride 130
    assert [int(s) for s in sel] == golden.selection("expquad_n100")
    sel32, *_ = greedy.place_single(np.asfortranarray(cov), 8, D)          # column-major input
    assert [int(s) for s in sel32] == golden.selection("expquad_n100")


def run_shards(cov, k, world):
    """G shards on ONE device, driven through the step-wise C-ABI exactly as G ranks would be; the two
    all-gathers are device-to-device copies into the gathered buffers."""
    n = cov.shape[0]
    prec = go.spd_inverse(cov)                      # test input only: the precision panels are given
    bounds = greedy.shard_bounds(n, world)
    stride = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
    stride += stride % 2
    shards = []
    for g in range(world):
        s = greedy.GreedyShard(n, bounds[g], bounds[g + 1], k, D)
        s.load_cov_host(cov)
        s.load_prec_host(prec)
        s.reset()
        shards.append(s)
    recs = _ffi.DeviceArray((world, 4), np.float64, D)
    segs = _ffi.DeviceArray((world, 2, stride), np.float64, D)
    for _ in range(k):
        for g, s in enumerate(shards):
            s.local_best(recs.ptr + g * 32)
        for s in shards:
            s.select(recs.ptr, world)
        for g, s in enumerate(shards):
            s.segments(segs.ptr + g * 2 * stride * 8, stride)
        for s in shards:
            s.apply(segs.ptr, stride, bounds)
    out = [s.results() for s in shards]
    for s in shards:
        s.close()
    return out


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_steps_equal_single_shard(world):
    cov = cloud_cov(700, 11)
    k = 9
    single_sel, single_scores, _, _ = greedy.place_single(cov, k, D)
    out = run_shards(cov, k, world)
    for sel, scores in out:
        np.testing.assert_array_equal(sel, single_sel)
        np.testing.assert_allclose(scores, single_scores, rtol=1e-10)      # P from LAPACK here vs device inverse
    for sel, scores in out[1:]:
        np.testing.assert_array_equal(scores, out[0][1])                  # every rank sees identical bits


def test_properties_at_size(golden):
    """Size-independent properties at n = 6000 (the CPU oracle is not run here): selections are distinct and
    in range, winning scores are non-increasing (delta is non-increasing in A), each winning score is the
    maximum of its step's score vector, and P stays symmetric with exact zero rows for selected points."""
    n, k = 6000, 20
    x = np.random.default_rng(42).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    shard = greedy.GreedyShard(n, 0, n, k, D)
    xd = _ffi.DeviceArray.from_host(x, D)
    shard.build_cov_expquad(xd.ptr, 3, 1.0, ls, 1e-2)
    shard.factor()
    shard.record_scores(True)
    shard.run(k)
    sel, scores = shard.results()
    steps = shard.step_scores()
    assert len(set(sel.tolist())) == k and sel.min() >= 0 and sel.max() < n
    assert np.all(np.diff(scores) <= 1e-12 * scores[:-1])
    for t in range(k):
        assert scores[t] == np.nanmax(steps[t]) and sel[t] == int(np.nanargmax(steps[t]))
        assert np.isnan(steps[t][sel[:t]]).all()
    p = _ffi.DeviceArray((shard.n_pad, shard.ld), np.float64, D, ptr=shard.prec_ptr, owner=shard).to_host()[:n, :n]
    assert np.array_equal(p, p.T)
    assert not p[sel].any() and not p[:, sel].any()
    # the remaining block of P is the inverse of Sigma restricted to the unselected set
    rest = np.setdiff1d(np.arange(n), sel)[:400]
    cov = _ffi.DeviceArray((shard.n_pad, shard.ld), np.float64, D, ptr=shard.cov_ptr, owner=shard).to_host()[:n, :n]
    keep = np.setdiff1d(np.arange(n), sel)
    resid = (p[np.ix_(rest, keep)] @ cov[np.ix_(keep, rest)]) - np.eye(len(rest))
    assert np.abs(resid).max() < 1e-9
    shard.close()


def test_save_restore_precision_replays_identically():
    cov = cloud_cov(900, 5)
    shard = greedy.GreedyShard(900, 0, 900, 8, D)
    shard.load_cov_host(cov)
    shard.factor()
    shard.save_precision()
    shard.run(8)
    first = shard.results()
    shard.restore_precision()
    shard.run(8)
    second = shard.results()
    np.testing.assert_array_equal(first[0], second[0])
    np.testing.assert_array_equal(first[1], second[1])
    assert shard.launch_count() > 8 * 4
    shard.close()


def run_shards_peer(cov, k, world, chunk=1):
    """G shards on ONE device, each on its own stream, exchanging through the peer-memory mailboxes
    (same-process pointers): the kernels of one shard wait for the flags the other shards' kernels set."""
    n = cov.shape[0]
    prec = go.spd_inverse(cov)
    bounds = greedy.shard_bounds(n, world)
    shards, streams = [], []
    for g in range(world):
        st = _ffi.c_vp()
        _ffi.call("vgp_stream_create", D, ctypes.byref(st))
        streams.append(st)
        s = greedy.GreedyShard(n, bounds[g], bounds[g + 1], k, D, stream=st)
        s.load_cov_host(cov)
        s.load_prec_host(prec)
        s.reset()
        s.sync()
        shards.append(s)
    boxes = [s.comm_create(g, world, bounds)[1] for g, s in enumerate(shards)]
    for s in shards:
        s.comm_connect_pointers(boxes)
    done = 0
    while done < k:                      # interleaved so that no stream's queue has to drain for another to start
        step = min(chunk, k - done)
        for s in shards:
            s.run_peer(step)
        done += step
    out = []
    for s in shards:
        s.comm_status()
        out.append(s.results())
    for s, st in zip(shards, streams):
        s.close()
        _ffi.call("vgp_stream_destroy", D, st)
    return out


@pytest.mark.parametrize("world,chunk", [(2, 1), (3, 2), (4, 1)])
def test_peer_exchange_equals_single_shard(world, chunk):
    cov = cloud_cov(700, 11)
    k = 9
    single_sel, single_scores, _, _ = greedy.place_single(cov, k, D)
    out = run_shards_peer(cov, k, world, chunk)
    ref = run_shards(cov, k, world)
    for (sel, scores), (rsel, rscores) in zip(out, ref):
        np.testing.assert_array_equal(sel, single_sel)
        np.testing.assert_array_equal(scores, rscores)                    # same bits as the all-gather protocol


@pytest.mark.timeout(300)
def test_peer_exchange_two_processes_ipc():
    """Two ranks, one process per GPU, mailboxes mapped through CUDA IPC (needs two devices)."""
    import os
    import subprocess
    import sys
    cnt = _ffi.c_int(0)
    _ffi.call("vgp_device_count", ctypes.byref(cnt))
    if cnt.value < 2:
        pytest.skip("needs two CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "peer_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=root)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PEER_WORKER_OK" in res.stdout
