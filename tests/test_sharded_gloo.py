"""CPU, world_size 2 over gloo: the sharded per-selection protocol (vgposp_b200.greedy.ShardedGreedy) with an
oracle-backed engine in place of the CUDA shard.  Checks the host logic of the N>1 path: shard bounds, the two
all-gathers, the winner rule across ranks, segment packing -- against the golden selections."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import greedy_oracle as go  # noqa: E402
from vgposp_b200.greedy import ShardedGreedy, shard_bounds  # noqa: E402


class OracleEngine:
    """Engine interface of ShardedGreedy on top of oracle.greedy_oracle.IncrementalState (test double)."""

    def __init__(self, cov, prec, c0, c1):
        self.state = go.IncrementalState(cov[:, c0:c1], prec[:, c0:c1], c0)
        self.n = cov.shape[0]
        self.wfull = np.zeros((0, self.n))
        self.cur = None
        self.selection, self.scores = [], []

    def local_best(self, rec):
        score, idx, num = self.state.local_best()
        rec[0], rec[2], rec[3] = score, num, 0.0
        rec.view(torch.int64)[1] = idx

    def select(self, recs, nrecords):
        r = recs.view(nrecords, 4)
        cands = [(float(r[g, 0]), int(r[g].view(torch.int64)[1]), float(r[g, 2])) for g in range(nrecords)]
        self.cur = go.pick_winner(cands)
        self.selection.append(self.cur[1])
        self.scores.append(self.cur[0])

    def segments(self, seg, stride):
        score, y, num_y = self.cur
        w, p = self.state.segments(y, num_y, self.wfull[:, y])
        seg.zero_()
        seg[:len(w)] = torch.from_numpy(w)
        seg[stride:stride + len(p)] = torch.from_numpy(p)

    def apply(self, gathered, stride, bounds):
        g = gathered.view(len(bounds) - 1, 2, stride).numpy()
        w_full = np.concatenate([g[r, 0, :bounds[r + 1] - bounds[r]] for r in range(len(bounds) - 1)])
        p_full = np.concatenate([g[r, 1, :bounds[r + 1] - bounds[r]] for r in range(len(bounds) - 1)])
        y = self.cur[1]
        st = self.state
        st.apply(y, w_full[st.c0:st.c0 + st.nloc].copy(), p_full)
        self.wfull = np.vstack([self.wfull, w_full[None, :]])


def _worker(rank, world, port, name, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import Golden
        cov = Golden().cov(name)
        n = cov.shape[0]
        prec = go.spd_inverse(cov)
        b = shard_bounds(n, world)
        eng = OracleEngine(cov, prec, b[rank], b[rank + 1])
        sg = ShardedGreedy(eng, n, rank, world,
                           make_buffer=lambda m: torch.zeros(m, dtype=torch.float64),
                           all_gather=lambda dst, src: dist.all_gather_into_tensor(dst, src))
        sg.run(k)
        out[rank] = (eng.selection, eng.scores)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["fixture4x4", "legacy_random_n20", "expquad_n100"])
def test_two_rank_protocol_matches_reference(golden, name):
    world = 2
    k = golden.cases[name]["k"]
    port = 29500 + (os.getpid() + hash(name)) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, name, k, out), nprocs=world, join=True)
        results = dict(out)
    assert results[0][0] == results[1][0] == golden.selection(name)
    single = go.incremental_greedy(golden.cov(name), k)
    np.testing.assert_array_equal(results[0][1], single[1])     # scores identical to the unsharded run


def test_single_rank_protocol_degenerates_to_plain_loop(golden):
    cov = golden.cov("expquad_n50")
    n = cov.shape[0]
    eng = OracleEngine(cov, go.spd_inverse(cov), 0, n)
    sg = ShardedGreedy(eng, n, 0, 1, make_buffer=lambda m: torch.zeros(m, dtype=torch.float64), all_gather=None)
    sg.run(5)
    assert eng.selection == golden.selection("expquad_n50")
