"""World-size-2 tests of the single-box scheduler ON GPUs (NCCL, one process per GPU): the collective tier with both
exchanges (NCCL reduce-scatter and the fused GEMM + peer-memory epilogue) checked on EVERY owned block, the tiers the
scheduler picks for replicated operands behind ``np.tensordot``, ``ShardedBlockArray.gather`` and the slice tier of
``tn.contract_network(comm=...)`` against the one-GPU result.  Skipped on boxes with fewer than two GPUs
(``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu``)."""
import os
import sys
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(x, ref):
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(np.asarray(ref).ravel()) or 1.0))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    report = {}
    try:
        import rosnet_b200 as rn
        from rosnet_b200 import tn as T
        from rosnet_b200.engine import scheduler

        ctx = scheduler.init()
        rng = np.random.default_rng(7)      # same stream on every rank: replicated operands

        def rand(shape, dtype):
            x = rng.standard_normal(shape)
            if np.dtype(dtype).kind == "c":
                x = x + 1j * rng.standard_normal(shape)
            return x.astype(dtype)

        # ---- tier 2: both operands sharded along the contracted block axis; every owned block is checked ----
        for dtype, tol, exchanges in (("complex128", 1e-12, ("nccl", "fused")), ("float64", 1e-12, ("nccl", "fused")), ("complex64", 1e-5, ("nccl",))):
            u, v = rand((768, 1024), dtype), rand((520, 1024), dtype)
            ub, vb = (256, 256), (260, 256)
            kloc = 1024 // world
            A = rn.ShardedBlockArray.from_local_slab(rn.array(u[:, rank * kloc:(rank + 1) * kloc], blockshape=ub, inner="cuda"), 1, ctx)
            B = rn.ShardedBlockArray.from_local_slab(rn.array(v[:, rank * kloc:(rank + 1) * kloc], blockshape=vb, inner="cuda"), 1, ctx)
            want = np.tensordot(u.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64),
                                v.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64), [(1,), (1,)])
            for ex in exchanges:
                ctx.exchange = ex
                C = np.tensordot(A, B, [(1,), (1,)])
                assert isinstance(C, rn.ShardedBlockArray) and C.grid == (3, 2) and C.blockshape == (256, 260)
                owned = C.owned()
                worst = 0.0
                for local, (gi, gj) in enumerate(C.owned_indices()):
                    blk = np.asarray(owned.data.flat[local])
                    worst = max(worst, _rel(blk, want[gi * 256:(gi + 1) * 256, gj * 260:(gj + 1) * 260]))
                assert worst <= tol, (dtype, ex, worst)
                report[f"ksplit_{dtype}_{ex}"] = (C.first, C.count, worst)
                full = np.asarray(C)            # gather: the whole array on every rank
                assert _rel(full, want) <= tol
            ctx.exchange = "auto"

        # ---- replicated operands behind np.tensordot: the scheduler picks the tier ----
        a, b = rand((512, 384), "float64"), rand((384, 256), "float64")
        A, B = rn.array(a, blockshape=(128, 128), inner="cuda"), rn.array(b, blockshape=(128, 128), inner="cuda")
        C = np.tensordot(A, B, [(1,), (0,)])                       # 4 x 2 output blocks >= 2 ranks: output partition
        assert isinstance(C, rn.ShardedBlockArray) and ctx.stats["output"] >= 1
        assert _rel(np.asarray(C), a @ b) <= 1e-12
        report["output_tier"] = (C.first, C.count)
        with rn.tuning.configure(threshold_k=64):
            A1, B1 = rn.array(a, blockshape=(512, 128), inner="cuda"), rn.array(b, blockshape=(128, 256), inner="cuda")
            before = ctx.stats["ksplit"]
            C = np.tensordot(A1, B1, [(1,), (0,)])                 # 1 output block < 2 ranks, 3 k-blocks of 128 >= threshold: k-split
            assert isinstance(C, rn.ShardedBlockArray) and ctx.stats["ksplit"] == before + 1
            assert _rel(np.asarray(C), a @ b) <= 1e-12
        with rn.tuning.configure(threshold_k=4096):
            C = np.tensordot(A1, B1, [(1,), (0,)])                 # k below the threshold: stays local, plain BlockArray
            assert isinstance(C, rn.BlockArray)
            assert _rel(np.asarray(C), a @ b) <= 1e-12
        with rn.tuning.configure(n_gpus=1):
            assert isinstance(np.tensordot(A, B, [(1,), (0,)]), rn.BlockArray)

        # ---- tier 1 with a sharded left operand (free axis) ----
        S = rn.ShardedBlockArray.shard(A, 0, ctx)
        C = np.tensordot(S, B, [(1,), (0,)])
        assert isinstance(C, rn.ShardedBlockArray) and C.axis == 0
        assert _rel(np.asarray(C), a @ b) <= 1e-12

        # ---- slices of a tensor network dealt over the ranks: all-reduced amplitude == one-rank amplitude ----
        net = T.rand_equation(14, 3, n_out=1, d_min=2, d_max=3, seed=5)
        net.random_arrays("complex128", seed=11)
        path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
        _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
        sliced = T.find_slices(net.inputs, net.output, net.sizes, path, max(largest // 8, 2))
        assert sliced
        tot, st = T.contract_network(net, path, sliced=sliced, slice_mode="loop", comm=ctx)
        with rn.tuning.configure(n_gpus=1):
            one, _ = T.contract_network(net, path, sliced=sliced, slice_mode="loop")
        assert _rel(tot, np.asarray(one)) <= 1e-12, (tot, one)
        assert st["n_slices_all_ranks"] > st["n_slices"] >= 1
        report["slices"] = (st["n_slices"], st["n_slices_all_ranks"])
        # the same with a SEARCHED tree: every rank runs the (deterministic) search itself, finds the same tree and the same slices
        from rosnet_b200.tn import contract as C_

        old_limit = C_.SEARCH_TARGET_BYTES
        C_.SEARCH_TARGET_BYTES = 16 * max(largest // 8, 2)
        try:
            tot_s, st_s = T.contract_network(net, path="search", comm=ctx)
        finally:
            C_.SEARCH_TARGET_BYTES = old_limit
        assert _rel(tot_s, np.asarray(one)) <= 1e-12, (tot_s, one)
        report["searched_slices"] = (st_s.get("n_slices"), st_s.get("n_slices_all_ranks"))
        # slicing-as-blocks with the context active: the pairwise tensordots are dealt by the scheduler itself
        blocks, _ = T.contract_network(net, path, sliced=sliced, slice_mode="blocks")
        assert _rel(np.asarray(blocks), np.asarray(one)) <= 1e-12
        scheduler.shutdown()
        q.put((rank, "ok", report))
    except Exception:
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def test_world2_scheduler_on_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for rank, status, payload in results:
        assert status == "ok", f"rank {rank}:\n{payload}"
    # 6 output blocks over 2 ranks: rank 0 owns blocks 0..2, rank 1 owns 3..5
    assert results[0][2]["ksplit_complex128_nccl"][:2] == (0, 3) and results[1][2]["ksplit_complex128_nccl"][:2] == (3, 3)
