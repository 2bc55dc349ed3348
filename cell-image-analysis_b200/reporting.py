"""Result writers of the screening run: the two CSV files and the
text report the reference writes in ``save_and_visualize_results`` /
``generate_screening_report`` (improved_detection.py:246-261, 351-403).  File names, column
order, the report's section headings and its 15 % / 25 % / 10 % thresholds are the on-disk
contract; plots (det:263-349, matplotlib/seaborn) are presentation only and not produced.
"""
from __future__ import annotations

import os
from datetime import datetime

import pandas as pd

SUMMARY_CSV = "screening_summary.csv"
DETAILED_CSV = "detailed_cell_results.csv"
REPORT_TXT = "mutant_screening_report.txt"

HIGH_CONSERVATIVE = 0.15
HIGH_MODERATE = 0.25
NORMAL_CONSERVATIVE = 0.10


def save_results(results: dict, detailed_results: list, output_dir: str) -> pd.DataFrame:
    """det:249-255: per-strain summary (index = strain name) and per-cell rows."""
    os.makedirs(output_dir, exist_ok=True)
    summary = pd.DataFrame.from_dict(results, orient="index")
    summary.to_csv(os.path.join(output_dir, SUMMARY_CSV))
    pd.DataFrame(detailed_results).to_csv(os.path.join(output_dir, DETAILED_CSV), index=False)
    return summary


def screening_report_lines(summary: pd.DataFrame, now: datetime | None = None) -> list:
    """det:355-403 as a list of lines (without trailing newlines)."""
    stamp = (now or datetime.now()).strftime("%Y-%m-%d %H:%M:%S")
    rule = "-" * 80
    out = ["=== MUTANT SCREENING REPORT (IMPROVED MODEL) ===", "", f"Generated: {stamp}", "",
           "MODEL PERFORMANCE BASELINE:",
           "- Conservative model: ~5% anomaly rate for normal cells",
           "- Moderate model: ~10% anomaly rate for normal cells", "",
           "SCREENING RESULTS:", rule,
           f"{'Sample':<20} {'Cells':<8} {'Conservative':<12} {'Moderate':<12} {'Mean MSE':<12}", rule]
    for name, row in summary.iterrows():
        out.append(f"{name:<20} {row['total_cells']:<8} {row['conservative_anomaly_rate'] * 100:>8.1f}% "
                   f"{row['moderate_anomaly_rate'] * 100:>10.1f}% {row['mean_mse']:>10.6f}")
    out += ["", "ANOMALY ANALYSIS:"]
    groups = (
        ("HIGH ANOMALY CANDIDATES (Conservative >15%):", "conservative_anomaly_rate",
         summary["conservative_anomaly_rate"] > HIGH_CONSERVATIVE),
        ("HIGH ANOMALY CANDIDATES (Moderate >25%):", "moderate_anomaly_rate",
         summary["moderate_anomaly_rate"] > HIGH_MODERATE),
        ("NORMAL-LEVEL SAMPLES (Conservative ≤10%):", "conservative_anomaly_rate",
         summary["conservative_anomaly_rate"] <= NORMAL_CONSERVATIVE),
    )
    for title, col, mask in groups:
        sel = summary[mask]
        if sel.empty:
            continue
        out += ["", title]
        out += [f"- {name}: {row[col] * 100:.1f}%" for name, row in sel.iterrows()]
    out += ["", "", "RECOMMENDATIONS:",
            "1. Focus on samples with Conservative >15% for detailed analysis",
            "2. Samples with Conservative ≤10% are likely normal phenotype",
            "3. Consider morphological analysis for high-anomaly candidates",
            "4. Validate results with independent experimental methods"]
    return out


def write_report(summary: pd.DataFrame, output_dir: str, now: datetime | None = None) -> str:
    path = os.path.join(output_dir, REPORT_TXT)
    with open(path, "w") as f:
        f.write("\n".join(screening_report_lines(summary, now)) + "\n")
    return path


def save_and_report(results: dict, detailed_results: list, output_dir: str) -> pd.DataFrame:
    """det:246-261 minus the figures."""
    summary = save_results(results, detailed_results, output_dir)
    if len(summary):
        write_report(summary, output_dir)
    return summary
