"""Multi-GPU plan for a screen: fields are independent, so they shard
by index across ranks with no data-path collective; the only exchange is one
all-reduce (sum) of the per-strain accumulator [S, 8] at the end
(improved_detection.py:151-152, 202-211 are per-strain reductions).

One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is the
plumbing.  Accumulator row: {n_cells, n_conservative_anomalies, n_moderate_anomalies,
sum mse, sum mse^2, sum mae, sum mae^2, 0}, float64.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

ACC_COLS = 8


def shard_fields(n_fields: int, rank: int, world: int):
    """Field i -> rank i mod world (round robin keeps per-strain load even)."""
    return list(range(rank, n_fields, world))


def allreduce_strain_acc(acc: torch.Tensor) -> torch.Tensor:
    """Sum the [S, 8] float64 accumulator over ranks (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def strain_summary(acc_row, sample_name=None, files_processed=None):
    """Per-strain result dict (det:202-212) from an accumulator row.  std is the
    population std (ddof=0) via sqrt(E[x^2] - E[x]^2) in float64."""
    n, nc, nm, s1, s2, a1, a2 = [float(v) for v in acc_row[:7]]
    if n <= 0:
        return None
    mean_mse, mean_mae = s1 / n, a1 / n
    out = {
        "total_cells": int(round(n)),
        "conservative_anomaly_rate": nc / n,
        "moderate_anomaly_rate": nm / n,
        "mean_mse": mean_mse,
        "std_mse": math.sqrt(max(s2 / n - mean_mse * mean_mse, 0.0)),
        "mean_mae": mean_mae,
        "std_mae": math.sqrt(max(a2 / n - mean_mae * mean_mae, 0.0)),
    }
    if sample_name is not None:
        out = {"sample_name": sample_name, **out}
    if files_processed is not None:
        out["files_processed"] = files_processed
    return out
