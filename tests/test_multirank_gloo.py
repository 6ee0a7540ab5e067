"""World-size-2 (gloo, CPU) tests of the single-box scheduler's host logic: output-block
partition (no exchange) and contracted-index partition (one reduce-scatter of partials).
The local contraction is the NumPy plan interpreter (tests/plan_interp.py) standing in for the
CUDA launch; what is under test is ownership, padding and the exchange step."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import plan_interp as pi
        from rosnet_b200.engine import scheduler as sch
        from rosnet_b200.engine.plan import plan_tensordot

        rng = np.random.default_rng(5)
        # C4-like: C = tensordot(U, V, [(1,), (1,)]), contracted block axis of extent 4, 3x3 output blocks
        U = rng.standard_normal((12, 16)) + 1j * rng.standard_normal((12, 16))
        V = rng.standard_normal((9, 16)) + 1j * rng.standard_normal((9, 16))
        ub, vb = (4, 4), (3, 4)
        want = np.tensordot(U, V, [(1,), (1,)])
        comm = sch.TorchComm()

        # ---- partition 2: contracted block index split, reduce-scatter of partials ----
        ks = sch.split_contracted_axis((3, 4), 1, world, rank)
        Ul, Vl = U[:, ks.start * 4:ks.stop * 4], V[:, ks.start * 4:ks.stop * 4]
        fa, ga = pi.block_major(Ul, ub)
        fb, gb = pi.block_major(Vl, vb)
        plan = plan_tensordot(ga, ub, gb, vb, np.complex128, [(1,), (1,)])
        partial = pi.run_plan(plan, fa, ub, fb, vb)
        chunk = sch.padded_chunk(plan.n_out, world)
        padded = np.zeros(world * chunk * plan.out_block_elems, dtype=partial.dtype)
        padded[: partial.size] = partial
        owned = sch.contract_k_partition(plan, torch.from_numpy(padded), comm, lambda n: torch.zeros(n, dtype=torch.complex128))
        got_blocks = owned.data.numpy().reshape(chunk, *plan.out_blockshape)[: owned.count]
        ok_k = True
        for blk, idx in zip(got_blocks, owned.block_indices()):
            sl = tuple(slice(i * s, (i + 1) * s) for i, s in zip(idx, plan.out_blockshape))
            ok_k &= bool(np.allclose(blk, want[sl], rtol=1e-13, atol=1e-13))
        counts_k = (owned.first, owned.count)

        # ---- partition 1: output blocks dealt to ranks, no exchange ----
        fa, ga = pi.block_major(U, ub)
        fb, gb = pi.block_major(V, vb)
        plan = plan_tensordot(ga, ub, gb, vb, np.complex128, [(1,), (1,)])
        full = pi.run_plan(plan, fa, ub, fb, vb).reshape(plan.n_out, -1)
        owned = sch.contract_output_partition(plan, lambda first, n: full[first:first + n], rank, world)
        ok_o = True
        for blk, idx in zip(owned.data, owned.block_indices()):
            sl = tuple(slice(i * s, (i + 1) * s) for i, s in zip(idx, plan.out_blockshape))
            ok_o &= bool(np.allclose(blk.reshape(plan.out_blockshape), want[sl], rtol=1e-13, atol=1e-13))
        # every output block is owned exactly once across ranks
        mine = torch.zeros(plan.n_out, dtype=torch.int64)
        mine[owned.first:owned.first + owned.count] = 1
        dist.all_reduce(mine)
        # slices (partition 3): round robin + scalar all-reduce
        mine_slices = sch.round_robin(7, world, rank)
        s = torch.tensor([float(sum(mine_slices))], dtype=torch.float64)
        comm.all_reduce_sum(s)
        out_q.put((rank, ok_k, counts_k, ok_o, bool((mine == 1).all()), float(s[0])))
    finally:
        dist.destroy_process_group()


def test_world2_partitions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    for rank, ok_k, counts_k, ok_o, once, ssum in results:
        assert ok_k and ok_o and once
        assert ssum == float(sum(range(7)))
    # 9 output blocks over 2 ranks, padded chunk 5: rank 0 owns 0..4, rank 1 owns 5..8
    assert results[0][2] == (0, 5) and results[1][2] == (5, 4)


def test_block_range_and_round_robin():
    from rosnet_b200.engine import scheduler as sch

    for n in (1, 7, 16, 64):
        for world in (1, 2, 4, 8):
            spans = [sch.block_range(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            assert sorted(sum((sch.round_robin(n, world, r) for r in range(world)), [])) == list(range(n))
    assert sch.padded_chunk(9, 2) == 5


def test_choose_tier_and_ksplit_plan_host_logic():
    """What the scheduler behind np.tensordot decides for replicated operands, and the k-split plan every rank runs:
    the per-rank partials (NumPy plan interpreter as the local executor) add up to the whole contraction."""
    import plan_interp as pi
    from rosnet_b200.engine import scheduler as sch
    from rosnet_b200.engine.plan import plan_tensordot

    assert sch.choose_tier(n_out=64, n_pairs=8, k=2048, world=8, threshold_k=256) == "output"
    assert sch.choose_tier(n_out=4, n_pairs=8, k=2048, world=8, threshold_k=256) == "ksplit"
    assert sch.choose_tier(n_out=4, n_pairs=8, k=128, world=8, threshold_k=256) == "local"      # k below tuning.threshold_k
    assert sch.choose_tier(n_out=4, n_pairs=4, k=2048, world=8, threshold_k=256) == "local"     # fewer k-blocks than GPUs
    assert sch.choose_tier(n_out=1, n_pairs=64, k=4096, world=1, threshold_k=256) == "local"
    rng = np.random.default_rng(9)
    a, b = rng.standard_normal((8, 24)), rng.standard_normal((24, 6))
    ab, bb = (8, 4), (4, 6)
    fa, ga = pi.block_major(a, ab)
    fb, gb = pi.block_major(b, bb)
    plan = plan_tensordot(ga, ab, gb, bb, np.float64, [(1,), (0,)])
    assert (plan.n_out, plan.n_pairs) == (1, 6)
    for world in (2, 3, 4):
        total, seen = 0, []
        for rank in range(world):
            mine = sch.ksplit_plan(plan, world, rank)
            seen += list(mine.pair_a[0])
            total = total + pi.run_plan(mine, fa, ab, fb, bb)
        assert sorted(seen) == sorted(plan.pair_a[0])          # every contracted block used exactly once
        assert np.allclose(total.reshape(8, 6), a @ b, rtol=1e-13, atol=1e-13)
    with pytest.raises(ValueError):
        sch.ksplit_plan(plan, 8, 7)                            # more ranks than contracted blocks
