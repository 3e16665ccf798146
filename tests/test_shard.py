"""Multi-rank host logic on CPU: 2 processes, gloo backend (SURVEY 8(e) / tier note 5).  The
partition + single-gather path is exercised with the oracle standing in for the GPU minimiser."""
import os
import socket

import numpy as np
import pytest

from better_flow_b200 import shard, synth


def test_partition_is_a_block_cyclic_cover():
    for n in (0, 1, 7, 64, 65, 297):
        for world in (1, 2, 4, 8):
            parts = [shard.partition(n, world, r) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import port as oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = synth.make_stream(240, 180, 1.5e6, 0.06, seed=13)
    slices = synth.cut_slices(st, 0.01)

    def minimise(batch):
        return [dict(oracle.minimize(s.fr_x, s.fr_y, s.t_ns, max_iter=4, accum_mode=1), n_events=len(s.fr_x), flags=0)
                for s in batch]

    rec = shard.run_sharded(slices, minimise, dist, block=1)
    q.put((rank, rec))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_shard_and_gather(oracle_port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-process answer
    st = synth.make_stream(240, 180, 1.5e6, 0.06, seed=13)
    slices = synth.cut_slices(st, 0.01)
    want = np.stack([oracle_port.minimize(s.fr_x, s.fr_y, s.t_ns, max_iter=4, accum_mode=1)["model"] for s in slices])
    for rank in (0, 1):
        rec = got[rank]
        assert rec.shape[0] == len(slices)
        assert list(rec[:, 0].astype(int)) == list(range(len(slices)))
        assert np.array_equal(rec[:, 5:16], want)
        assert np.all(rec[:, 2] == 5)


def test_native_owner_rule_equals_python_partition():
    """bf_multi_owner (C ABI, used by the C++ front end's --gpus) deals slices exactly like shard.partition."""
    import better_flow_b200 as bf
    lib = bf.load()
    for n in (1, 7, 64, 297):
        for world in (1, 2, 4, 8):
            for block in (1, 4, 5):
                owners = [lib.bf_multi_owner(k, world, block) for k in range(n)]
                for r in range(world):
                    assert [k for k, o in enumerate(owners) if o == r] == shard.partition(n, world, r, block)
    assert lib.bf_multi_owner(-1, 2, 4) == -1 and lib.bf_multi_owner(3, 0, 4) == -1


def test_multi_front_fails_loudly_without_devices():
    import better_flow_b200 as bf
    lib = bf.load()
    if lib.bf_device_count() >= 2:
        pytest.skip("two CUDA devices are present")
    with pytest.raises(bf.BfError) as e:
        bf.MultiContext(2, 180, 240, 3, 1 << 16, 8)
    assert "no CPU fallback" in str(e.value)


def _gather_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nbytes = 3 * 160                                   # three bf_slice_result records
    rg = shard.RecordGather(dist, torch, nbytes, slots=4, device="cpu")
    ok = True

    def batch(step):                                   # what rank `r` leaves on its device after step `s`
        return lambda r: torch.arange(nbytes, dtype=torch.int64).add(17 * r + 5 * step).remainder(251).to(torch.uint8)

    ok &= rg.flush() is None                           # nothing to exchange yet: no collective
    for step in range(3):
        rg.snapshot(batch(step)(rank))
    got = rg.flush()
    ok &= rg.gathers == 1 and tuple(got.shape) == (world, 3, nbytes)
    for r in range(world):
        for step in range(3):
            ok &= bool(torch.equal(got[r, step], batch(step)(r)))
    # more batches than slots: the ring is flushed when it is full, then once more at the end
    for step in range(6):
        rg.snapshot(batch(10 + step)(rank))
    got = rg.flush()
    ok &= rg.gathers == 3 and tuple(got.shape) == (world, 2, nbytes)
    for r in range(world):
        ok &= bool(torch.equal(got[r, 0], batch(14)(r))) and bool(torch.equal(got[r, 1], batch(15)(r)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_deferred_record_gather_two_ranks():
    """bench.py at N > 1: per-step device snapshots of the result records, ONE all_gather for all of them."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
