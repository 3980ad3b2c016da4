"""World-size-2 (and 3) gloo tests of the node-partitioned path on CPU.

The collectives, the partitioning, the row exchange and the pipelined convergence loop of
bigclam_apachespark_b200/dist.py are exercised with a test-side engine that implements the
multi-GPU engine interface with the CPU oracle (allowed here: tests/ may call oracle/).  The same
DistBigClam code drives the CUDA contexts on the GPU box (tests/test_gpu_dist.py, bench.py --gpus N).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO, random_graph


class OracleEngine:
    """Engine interface of dist.py on top of oracle.step (owned rows via node_mask)."""

    def __init__(self, O, rp, col, F0, sumF, k, lo, hi):
        self.O, self.rp, self.col, self.k, self.ld = O, rp, col, k, k
        self.P = O.make_params(k)
        n = len(rp) - 1
        self.F = [torch.from_numpy(F0.copy()), torch.zeros(n, k, dtype=torch.float64)]
        self.sumF = [torch.from_numpy(np.array(sumF, copy=True)), torch.zeros(k, dtype=torch.float64)]
        self.cur = 0
        self.mask = np.zeros(n, dtype=np.uint8)
        self.mask[lo:hi] = 1
        self.part = torch.zeros(2 * k + 2, dtype=torch.float64)

    def state(self):
        return self.F[self.cur], self.F[self.cur ^ 1], self.sumF[self.cur]

    def step_local(self):
        F, s = self.F[self.cur].numpy(), self.sumF[self.cur].numpy()
        r = self.O.step(self.rp, self.col, F, s, self.P, node_mask=self.mask, want_pre=True)
        own = self.mask.astype(bool)
        upd = (r.accepted >= 0) & own
        self._changed = torch.from_numpy(np.nonzero(upd)[0].astype(np.int64))
        self.F[self.cur ^ 1][torch.from_numpy(own)] = torch.from_numpy(r.F[own])
        self.part.zero_()
        self.part[: self.k] = torch.from_numpy((F[upd] - r.F[upd]).sum(axis=0))
        self.part[2 * self.k] = float(r.llh_u[own].sum())
        self.part[2 * self.k + 1] = float(upd.sum())
        return self.part

    def llh_local(self):
        F, s = self.F[self.cur].numpy(), self.sumF[self.cur].numpy()
        _, per = self.O.llh(self.rp, self.col, F, s, self.P, per_node=True)
        self.part.zero_()
        self.part[2 * self.k] = float(per[self.mask.astype(bool)].sum())
        return self.part

    def finish_local(self, sync=True):
        llh, nupd = float(self.part[2 * self.k]), int(round(float(self.part[2 * self.k + 1])))
        s = self.sumF[self.cur]
        self.sumF[self.cur ^ 1].copy_(s - self.part[: self.k] if nupd > 0 else s)
        self.cur ^= 1
        return llh, nupd

    def rollback(self):
        self.cur ^= 1

    def changed_owned(self):
        return self._changed


def _worker(rank, world, port, variant, exchange, out):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from bigclam_apachespark_b200.dist import DistBigClam, partition_by_nnz
    n, k = 300, 6
    rp, col = random_graph(n, 6, seed=17, hub=50)
    rng = np.random.default_rng(17)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.4)
    sumF = O.colsum(F0)
    bounds = partition_by_nnz(rp, world)
    eng = OracleEngine(O, rp, col, F0, sumF, k, int(bounds[rank]), int(bounds[rank + 1]))
    d = DistBigClam(eng, rp, rank, world, bounds, exchange=exchange)
    # three single calls
    llhs = [d.backtrackingLineSearchs() for _ in range(3)]
    F3 = eng.state()[0].numpy().copy()
    # then the pipelined loop from the initial state
    eng2 = OracleEngine(O, rp, col, F0, sumF, k, int(bounds[rank]), int(bounds[rank + 1]))
    d2 = DistBigClam(eng2, rp, rank, world, bounds, exchange=exchange)
    ret, calls, trace = d2.run(variant=variant)
    if rank == 0:
        np.savez(out, llhs=np.array(llhs), F3=F3, ret=ret, calls=calls, trace=np.array(trace),
                 Fend=eng2.state()[0].numpy(), sumFend=eng2.state()[2].numpy(), bounds=bounds)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,variant,exchange", [(2, 4, "full"), (3, 2, "delta"), (2, 3, "delta")])
def test_partitioned_path_equals_single_process(tmp_path, oracle, world, variant, exchange):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), variant, exchange, out), nprocs=world, join=True)
    z = np.load(out)
    n, k = 300, 6
    rp, col = random_graph(n, 6, seed=17, hub=50)
    rng = np.random.default_rng(17)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.4)
    sumF = oracle.colsum(F0)
    P = oracle.make_params(k)
    F, s = F0, sumF
    for it in range(3):
        r = oracle.step(rp, col, F, s, P)
        assert abs(z["llhs"][it] - r.llh) <= 1e-11 * abs(r.llh)
        F, s = r.F, r.sumF
    assert np.allclose(z["F3"], F, rtol=1e-12, atol=1e-14)          # property (ii): partition independent
    Fo, so, llho, callso, tro = oracle.run(rp, col, F0, sumF, P, variant=variant)
    assert int(z["calls"]) == callso
    assert np.allclose(z["trace"], tro, rtol=1e-11)
    assert abs(float(z["ret"]) - llho) <= 1e-11 * abs(llho)
    err = np.abs(z["Fend"] - Fo).max() / np.abs(Fo).max()
    assert err <= 1e-7, err                    # fp-reduction noise compounds over the loop; contract is 1e-4
    assert np.allclose(z["sumFend"], so, rtol=1e-7)
    b = z["bounds"]
    assert b[0] == 0 and b[-1] == n and (np.diff(b) > 0).all()


def test_partition_by_nnz_balance(graphs):
    from bigclam_apachespark_b200.dist import partition_by_nnz
    rp, col, _ = graphs.load_npz_graph("com-amazon")
    for world in (1, 2, 4, 8):
        b = partition_by_nnz(rp, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(rp) - 1
        w = np.diff(rp[b]) + np.diff(b)
        assert w.max() <= 1.02 * w.mean() + 600


def test_deal_all_by_degree_balances_edges_and_covers_every_node():
    """The owned-node rule of the node-partitioned path (dist.deal_all_by_degree == bigclam_multi_create): every node
    exactly once, neighbour-list entries per rank level even on a skewed graph."""
    from bigclam_apachespark_b200 import graphs as G
    from bigclam_apachespark_b200.dist import deal_all_by_degree
    rp, col = G.rmat_graph(20000, 150000, seed=7)
    deg = np.diff(rp)
    for world in (2, 3, 8):
        deal = deal_all_by_degree(rp, world)
        allv = np.concatenate(deal)
        assert len(allv) == len(deg) and len(np.unique(allv)) == len(deg)
        loads = np.array([deg[x].sum() for x in deal])
        assert loads.max() - loads.min() <= max(deg.max(), 0.002 * loads.mean())
        assert max(len(x) for x in deal) - min(len(x) for x in deal) <= 0.02 * len(deg) / world + 4
