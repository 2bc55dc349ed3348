"""Per-phase cycle counts of the accumulating conv kernels (needs a -DCIA_ACC_TIMING build)."""
import sys
import torch
sys.path.insert(0, '.')
from cell_image_analysis_b200.screening import Engine
from cell_image_analysis_b200.artifacts import load_model_dir

eng = Engine()
eng.load_artifacts(load_model_dir('tests/golden/model_dir'))
n = 1024 * 4
x = torch.rand((n, 64, 64), dtype=torch.float32, device='cuda')
for _ in range(2):
    eng.cae_forward(x, n)
torch.cuda.synchronize()
