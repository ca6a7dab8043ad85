"""tests/golden/sweep_dense_three_matrices_t500_seed777.csv: the CSV the reference's own main() (oracle/_ref/QKD_LDPC_ref) writes for a
directory holding the three shipped dense matrices (rates 0.34 / 0.5 / 0.57 -> three QBER presets, seven sweep points in all).
Row order follows the directory iteration order of the machine that ran it; the test compares rows per (matrix, QBER).
Run where /root/reference is present:   python tests/golden/make_multi_matrix.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import DENSE, GOLD, REFERENCE_ROOT, base_config, run_ref_main  # noqa: E402

PRESETS = [{"code_rate": 0.34, "QBER_begin": 0.17, "QBER_end": 0.51, "QBER_step": 0.17},
           {"code_rate": 0.5, "QBER_begin": 0.1, "QBER_end": 0.3, "QBER_step": 0.1},
           {"code_rate": 0.58, "QBER_begin": 0.15, "QBER_end": 0.35, "QBER_step": 0.1}]

if __name__ == "__main__":
    cfg = base_config(trials_number=500, use_dense_matrices=True, code_rate_QBER_parameters=PRESETS)
    csv = run_ref_main(cfg, [REFERENCE_ROOT / "dense_matrices" / f for f in DENSE.values()], dense=True)
    (GOLD / "sweep_dense_three_matrices_t500_seed777.csv").write_text(csv)
    print(csv)
