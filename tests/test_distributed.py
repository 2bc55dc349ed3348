"""CPU tests of the N>1 host logic with the gloo backend (world_size 2): fields shard
by index, each rank accumulates its own per-strain rows, one all-reduce (sum) of the
[S, 8] accumulator gives every rank the global result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cell_image_analysis_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_scores(field):
    rng = np.random.default_rng(1000 + field)
    n = int(rng.integers(3, 9))
    return dict(strain=field % 3, mse=rng.random(n).astype(np.float32), mae=rng.random(n).astype(np.float32),
                pc=rng.choice([-1, 1], n), pm=rng.choice([-1, 1], n))


def _accumulate(fields, n_strains):
    acc = np.zeros((n_strains, D.ACC_COLS))
    for f in fields:
        s = _fake_scores(f)
        a = acc[s["strain"]]
        m, e = s["mse"].astype(np.float64), s["mae"].astype(np.float64)
        a[0] += len(m); a[1] += (s["pc"] == -1).sum(); a[2] += (s["pm"] == -1).sum()
        a[3] += m.sum(); a[4] += (m * m).sum(); a[5] += e.sum(); a[6] += (e * e).sum()
    return acc


def _worker(rank, world, port, n_fields, n_strains, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.shard_fields(n_fields, rank, world)
    acc = torch.from_numpy(_accumulate(mine, n_strains))
    D.allreduce_strain_acc(acc)
    q.put((rank, mine, acc.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    n_fields, n_strains, world = 11, 3, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_fields, n_strains, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _accumulate(range(n_fields), n_strains)
    seen = sorted(f for _, mine, _ in got for f in mine)
    assert seen == list(range(n_fields))                    # every field on exactly one rank
    for _, _, acc in got:
        assert np.array_equal(acc[:, :3], ref[:, :3])       # counts exact
        np.testing.assert_allclose(acc, ref, rtol=1e-13)    # sums within fp64 rounding


def test_strain_summary_matches_numpy():
    acc = _accumulate(range(9), 3)
    for s in range(3):
        mse = np.concatenate([_fake_scores(f)["mse"] for f in range(9) if f % 3 == s]).astype(np.float64)
        pc = np.concatenate([_fake_scores(f)["pc"] for f in range(9) if f % 3 == s])
        r = D.strain_summary(acc[s], sample_name=f"s{s}")
        assert r["total_cells"] == len(mse)
        assert abs(r["mean_mse"] - mse.mean()) < 1e-14 and abs(r["std_mse"] - mse.std()) < 1e-12
        assert r["conservative_anomaly_rate"] == (pc == -1).sum() / len(pc)
    assert D.strain_summary(np.zeros(8)) is None


def test_shard_fields_round_robin():
    assert D.shard_fields(10, 1, 4) == [1, 5, 9]
    assert sum(len(D.shard_fields(1000, r, 8)) for r in range(8)) == 1000


def test_allreduce_without_group_is_identity():
    a = torch.arange(16, dtype=torch.float64).reshape(2, 8)
    assert torch.equal(D.allreduce_strain_acc(a.clone()), a)


# ---- ShardedScreen: sharding, all-reduce and host-side row assembly with a NumPy scorer ----
class _FakeScorer:
    """One 'cell' per label present in a field; scores derived from the label id and the field's
    first pixel, so that results identify (field, label) uniquely."""

    def score(self, fields, strains, n_strains):
        acc = np.zeros((n_strains, D.ACC_COLS))
        out = []
        for f, s in zip(fields, strains):
            g, l = f() if callable(f) else f
            labs = np.unique(l[l > 0])
            k = float(g[0, 0])
            rows = dict(label=labs.astype(np.int32), mse=(labs * 0.001 + k).astype(np.float32),
                        mae=(labs * 0.002 + k).astype(np.float32), dec_cons=labs * 0.5 - k, dec_mod=labs * 0.25 - k,
                        pred_cons=np.where(labs % 2 == 0, -1, 1).astype(np.int8),
                        pred_mod=np.where(labs % 3 == 0, -1, 1).astype(np.int8))
            D.accumulate_rows(acc, s, rows)
            out.append(rows)
        return torch.from_numpy(acc), out


def _fake_field(i):
    g = np.full((8, 8), i, np.uint16)
    l = np.zeros((8, 8), np.int32)
    for k in range(1 + i % 4):
        l[k, k] = 3 * k + 1 + (i % 2)
    return g, l


def _sharded_worker(rank, world, port, n_fields, n_strains, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    strain = [i % n_strains for i in range(n_fields)]
    loaded = []

    def load(i):
        loaded.append(i)
        return _fake_field(i)
    acc, rows = D.ShardedScreen(_FakeScorer()).screen(load, n_fields, strain, n_strains)
    q.put((rank, loaded, acc, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_screen_two_ranks_equals_one():
    n_fields, n_strains, world = 13, 3, 2
    strain = [i % n_strains for i in range(n_fields)]
    acc1, rows1 = D.ShardedScreen(_FakeScorer()).screen(_fake_field, n_fields, strain, n_strains)   # no group: 1 rank
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n_fields, n_strains, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == list(range(0, n_fields, 2)) and got[1][1] == list(range(1, n_fields, 2))   # only its own fields
    for _, _, acc, _ in got:
        assert np.array_equal(acc[:, :3], acc1[:, :3])                   # counts exact on every rank
        np.testing.assert_allclose(acc, acc1, rtol=1e-13)                # sums within fp64 rounding
    assert got[1][3] is None                                             # rows are assembled on rank 0 only
    rows2 = got[0][3]
    for k in rows1:
        assert np.array_equal(rows1[k], rows2[k]), k                     # same rows, same (field, label) order
    assert np.all(np.diff(rows2["field"]) >= 0)
    same = np.diff(rows2["field"]) == 0
    assert np.all(np.diff(rows2["label"])[same] > 0)
