"""Multi-GPU screening through the product API: fields sharded over
ranks, one NCCL all-reduce of the per-strain accumulator, per-cell rows assembled on rank 0 in
reference order.  N-GPU results must equal the 1-GPU results: counts exactly, sums within fp64
rounding, rows identical.  (The 2-rank test needs two GPUs: `gpurun --gpus 2`; skipped otherwise.)"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_FIELDS, N_STRAINS = 12, 3


def _fields():
    from cell_image_analysis_b200 import synth
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    return [synth.make_field(40 + i, H, W, n, lo, hi, lu) for i in range(N_FIELDS)]


def _strains():
    return [i * N_STRAINS // N_FIELDS for i in range(N_FIELDS)]      # contiguous blocks, like folders


@pytest.fixture(scope="module")
def single(model_dir):
    from cell_image_analysis_b200.screening import ProductionMutantScreening
    s = ProductionMutantScreening(model_dir, segmenter=lambda ch: None, device=0, precision=1)
    acc, rows = s.screen_fields_sharded(_fields(), _strains(), N_STRAINS, chunk_fields=4)
    yield s, acc, rows
    s.engine.close()


def test_sharded_call_equals_per_field_dropin(single):
    """world size 1: the sharded call returns what the reference-shaped per-file loop returns."""
    s, acc, rows = single
    from cell_image_analysis_b200.distributed import strain_summary
    fields, strains = _fields(), _strains()
    assert rows["label"].size > 50
    off = 0
    per_strain = {k: [] for k in range(N_STRAINS)}
    for i, (g, l) in enumerate(fields):
        cells, _stats, rec = s.extract_quality_cells_from_labels(g, l, return_regions=True)
        n = len(cells)
        assert np.array_equal(rows["label"][off:off + n], rec["label"]) and np.all(rows["field"][off:off + n] == i)
        r = s.compute_anomaly_scores(cells)
        assert np.array_equal(rows["mse"][off:off + n], r["reconstruction_mse"])
        assert np.array_equal(rows["dec_cons"][off:off + n], -r["conservative_scores"])
        assert np.array_equal(rows["pred_mod"][off:off + n], r["moderate_predictions"])
        per_strain[strains[i]].append(r)
        off += n
    assert off == rows["label"].size
    for k in range(N_STRAINS):
        mse = np.concatenate([r["reconstruction_mse"] for r in per_strain[k]])
        pc = np.concatenate([r["conservative_predictions"] for r in per_strain[k]])
        summ = strain_summary(acc[k])
        assert summ["total_cells"] == len(mse)
        assert summ["conservative_anomaly_rate"] == np.sum(pc == -1) / len(pc)             # det:151
        assert abs(summ["mean_mse"] - np.mean(mse.astype(np.float64))) <= 1e-12           # det:206
        assert abs(summ["std_mse"] - np.std(mse.astype(np.float64))) <= 1e-9


def _worker(rank, world, port, model_dir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cell_image_analysis_b200.screening import ProductionMutantScreening
    s = ProductionMutantScreening(model_dir, segmenter=lambda ch: None, device=rank, precision=1)
    acc, rows = s.screen_fields_sharded(_fields(), _strains(), N_STRAINS, chunk_fields=4)
    q.put((rank, acc, rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpus_equal_one(single, model_dir):
    import torch.multiprocessing as mp
    _s, acc1, rows1 = single
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, model_dir, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for _, acc, _ in got:
        assert np.array_equal(acc[:, :3], acc1[:, :3])                    # counts exact on every rank
        np.testing.assert_allclose(acc, acc1, rtol=1e-12)                 # fp64 sums: order of addition only
    rows2 = got[0][2]
    assert got[1][2] is None
    for k in rows1:
        assert np.array_equal(rows1[k], rows2[k]), k
    print(f"2 GPUs == 1 GPU: {rows2['label'].size} cells, per-strain counts {acc1[:, 0].astype(int).tolist()}")
