"""
Data-parallel classifier training (heybuddy_b200/dp.py): the sharded step must equal the single-device step on the
concatenated batch.  CPU: two gloo ranks drive distributed_train_step with the oracle classifier.  GPU: the split C-ABI
(select / backward / grads_copy / adam) against the fused hb_mlp_train_step, and two emulated ranks against one.
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(seed, n=96):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 16, 96)).astype(np.float32)
    y = (rng.random(n) < 0.3).astype(np.int64)
    x[y == 1] += 0.4
    return x, y


WORKER = textwrap.dedent("""
    import sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
    from dp_oracle_engine import OracleEngine
    from test_dp_training import _batch
    from heybuddy_b200.dp import distributed_train_step, shard_batch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = OracleEngine()
    packed = OracleEngine()
    single = OracleEngine()
    for step in range(3):
        x, y = _batch(100 + step)
        xs, ys = shard_batch(torch.from_numpy(x), torch.from_numpy(y), rank, world)
        _, stats = distributed_train_step(eng, xs, ys, 1e-3, negative_weight=0.7, high_loss_threshold=1e-4, min_selected=16, one_collective=False)
        # the one-collective form (unnormalised sums, ONE all-reduce, division afterwards) gives the same global statistics
        _, stats1 = distributed_train_step(packed, xs, ys, 1e-3, negative_weight=0.7, high_loss_threshold=1e-4, min_selected=16, one_collective=True)
        assert float(stats1[1]) == float(stats[1]) and float(stats1[2]) == float(stats[2])
        assert abs(float(stats1[0]) - float(stats[0])) < 1e-5 * max(1.0, abs(float(stats[0])))
        # the same global batch on one "device"
        _, st1 = single.dp_select(torch.from_numpy(x), torch.from_numpy(y), 1e-4)
        st1 = single.dp_backward(st1[1:2].clone(), 0.7, 1e-4, 16)
        single.dp_adam(1e-3, st1)
        assert float(stats[1]) == float(st1[1]), (stats, st1)
        assert abs(float(stats[0]) - float(st1[0])) < 1e-5 * max(1.0, abs(float(st1[0])))
    # the all-reduced gradients travel as f32 (like NCCL on the GPU): agreement to f32 rounding of the update
    d = np.abs(eng.flat_params() - single.flat_params()).max()
    assert d < 2e-6, d
    d1 = np.abs(packed.flat_params() - single.flat_params()).max()
    assert d1 < 5e-6, d1
    gathered = [None] * world
    dist.all_gather_object(gathered, float(np.abs(eng.flat_params()).sum()))
    assert len(set(gathered)) == 1, gathered          # replicas identical on every rank
    if rank == 0:
        print("DP OK", d)
    dist.destroy_process_group()
""")


def test_two_rank_dp_step_gloo(tmp_path):
    """world_size 2 over gloo: select -> all-reduce n -> backward -> all-reduce grads -> Adam == the single-process step."""
    script = tmp_path / "dp_worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "DP OK" in res.stdout


def test_one_collective_is_chosen_only_for_engines_that_have_it(monkeypatch):
    """dp.py picks the packed protocol when the engine offers dp_local_step / dp_apply and no parity mode is requested."""
    import sys as _sys

    _sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dp_oracle_engine import OracleEngine

    from heybuddy_b200 import dp

    class TwoCollectiveOnly:
        dp_select = dp_backward = dp_grads = dp_adam = None

    monkeypatch.delenv("HB_MLP_STAGED", raising=False)
    monkeypatch.delenv("HB_MLP_FMA", raising=False)
    assert dp._one_collective(OracleEngine()) and not dp._one_collective(TwoCollectiveOnly())
    monkeypatch.setenv("HB_MLP_STAGED", "1")
    assert not dp._one_collective(OracleEngine())
    monkeypatch.delenv("HB_MLP_STAGED")
    monkeypatch.setenv("HB_MLP_FMA", "1")
    assert not dp._one_collective(OracleEngine())


def test_shard_batch_covers_every_row_once():
    import torch

    from heybuddy_b200.dp import shard_batch

    x, y = torch.arange(103).reshape(103, 1), torch.arange(103)
    for world in (1, 2, 3, 8):
        parts = [shard_batch(x, y, r, world)[1] for r in range(world)]
        assert torch.equal(torch.cat(parts), y) and max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@pytest.mark.gpu
def test_split_step_equals_fused_step(cuda_device):
    import torch

    from heybuddy_b200.dp import distributed_train_step
    from heybuddy_b200.wakeword import WakeWordMLPModel

    fused, split = WakeWordMLPModel(device_id=0, seed=5), WakeWordMLPModel(device_id=0, seed=5)
    for step in range(4):
        x, y = _batch(200 + step, n=256)
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        _, s1 = fused.train_step(xd, yd, 1e-3, 0.7, 1e-4, min_selected=16)
        _, s2 = distributed_train_step(split, xd, yd, 1e-3, 0.7, 1e-4, min_selected=16)     # no process group: single rank
        assert torch.equal(s1[1:], s2[1:])                                   # selected rows, stepped, high-loss rate: exact
        assert abs(float(s1[0]) - float(s2[0])) <= 1e-6 * abs(float(s1[0]))  # the loss is summed with atomics: order varies
    a, b = fused.state_dict(), split.state_dict()
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.gpu
def test_two_emulated_ranks_equal_one(cuda_device):
    """Two replicas, half the batch each, reductions done by hand on the device == one replica on the whole batch."""
    import torch

    from heybuddy_b200.dp import shard_batch
    from heybuddy_b200.wakeword import WakeWordMLPModel

    one = WakeWordMLPModel(device_id=0, seed=5)
    ranks = [WakeWordMLPModel(device_id=0, seed=5) for _ in range(2)]
    for step in range(4):
        x, y = _batch(300 + step, n=256)
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        _, s_one = one.train_step(xd, yd, 1e-3, 0.7, 1e-4, min_selected=16)
        sel = [m.dp_select(*shard_batch(xd, yd, r, 2), 1e-4)[1] for r, m in enumerate(ranks)]
        n_total = (sel[0][1:2] + sel[1][1:2]).contiguous()
        stats = [m.dp_backward(n_total, 0.7, 1e-4, 16) for m in ranks]
        g = ranks[0].dp_grads() + ranks[1].dp_grads()
        for m, st in zip(ranks, stats):
            m.dp_grads(g, to_model=True)
            m.dp_adam(1e-3, st)
        assert float(n_total) == float(s_one[1])
        assert abs(float(stats[0][0] + stats[1][0]) - float(s_one[0])) <= 1e-5 * abs(float(s_one[0]))
    p1, pa, pb = one.state_dict(), ranks[0].state_dict(), ranks[1].state_dict()
    diffs = []
    for k in p1:
        assert torch.equal(pa[k], pb[k]), k                                    # replicas stay identical
        diffs.append((p1[k] - pa[k]).abs().reshape(-1))
    d = torch.cat(diffs)
    # f32 summation order only; Adam normalises the update, so a gradient element near zero may move by up to 2 lr per step
    assert float((d <= 2e-5).float().mean()) >= 0.999 and float(d.max()) <= 4 * 2 * 1e-3, (float(d.max()), float((d <= 2e-5).float().mean()))
