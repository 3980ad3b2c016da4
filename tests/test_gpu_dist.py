"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): the node-partitioned CUDA path over
NCCL must equal the single-GPU path / the oracle to fp-reduction noise (SURVEY §8c property ii)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO, random_graph

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, exchange, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from bigclam_apachespark_b200 import BigClam
    from bigclam_apachespark_b200.dist import CudaEngine, DistBigClam, deal_all_by_degree, partition_by_nnz
    from oracle import oracle as O
    n, k = 3000, 40
    rp, col = random_graph(n, 8, seed=5, hub=300)
    rng = np.random.default_rng(5)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.2)
    sumF = O.colsum(F0)
    sparse = exchange.startswith("p2p-sparse")  # sparse rows of F, pushed by the step kernel (regions of the pools)
    if sparse:
        # default: the fused device-side collective (bigclam_xchg_*); "-nccl": the all-reduce goes through NCCL instead
        os.environ["BIGCLAM_NCCL_ALLREDUCE"] = "1" if exchange.endswith("-nccl") else "0"
        exchange = "p2p"
    b = BigClam(device=rank, record_accepted=True, sparse_rows=sparse)
    b.set_graph(rp, col).set_K(k)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    b.set_F(F0, sumF=sumF)
    bounds = partition_by_nnz(rp, world)
    deal = deal_all_by_degree(rp, world)
    nodes = deal[rank] if exchange == "p2p" else None
    counts = [len(x) for x in deal]
    d = DistBigClam(CudaEngine(b, int(bounds[rank]), int(bounds[rank + 1]), nodes=nodes, owned_counts=counts, rank=rank),
                    rp, rank, world, bounds, exchange=exchange)
    llhs = [d.backtrackingLineSearchs() for _ in range(3)]
    F3, s3 = b.F, b.sumF
    b.set_F(F0, sumF=sumF)
    d._need_sync = True
    d._prev_changed = None
    if exchange == "p2p":
        d.e.mark_all_changed()
    ret, calls, trace = d.run(variant=4, max_outer=40)
    if rank == 0:
        np.savez(out, llhs=np.array(llhs), F3=F3, s3=s3, ret=ret, calls=calls, trace=np.array(trace), Fend=b.F)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "delta", "full", "p2p-sparse", "p2p-sparse-nccl"])
def test_two_gpu_partitioned_equals_oracle(tmp_path, oracle, exchange):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, port, exchange, out), nprocs=world, join=True)
    z = np.load(out)
    n, k = 3000, 40
    rp, col = random_graph(n, 8, seed=5, hub=300)
    rng = np.random.default_rng(5)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.2)
    sumF = oracle.colsum(F0)
    P = oracle.make_params(k)
    F, s = F0, sumF
    for it in range(3):
        r = oracle.step(rp, col, F, s, P)
        assert abs(z["llhs"][it] - r.llh) <= 1e-9 * abs(r.llh)
        F, s = r.F, r.sumF
    assert np.abs(z["F3"] - F).max() <= 1e-7 * np.abs(F).max()
    assert np.allclose(z["s3"], s, rtol=1e-9)
    Fo, so, llho, callso, tro = oracle.run(rp, col, F0, sumF, P, variant=4, max_outer=40)
    assert int(z["calls"]) == callso
    assert np.allclose(z["trace"], tro, rtol=1e-7)
    assert np.abs(z["Fend"] - Fo).max() <= 1e-4 * np.abs(Fo).max()
