"""Two NCCL ranks on one box: the sharded paths of SURVEY.md 8(e) against their single-GPU results.
Skipped when fewer than two GPUs are visible (the per-round 1-GPU test box); run with `gpurun --gpus 2`."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_paths_equal_single_gpu(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "_nccl_worker.py"), str(tmp_path)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(2)]
    for r in res:
        assert r["world"] == 2
        for key in ("topk_exact", "topk_prefilter", "topk_inplace_exact", "topk_inplace_prefilter", "ranks_square",
                    "ranks_grouped", "eval_square", "eval_grouped", "identify_replicated", "identify_sharded"):
            assert r[key] is True, (r["rank"], key)
        assert r["cca_sigma"] <= 1e-9 and r["cca_UV"] <= 1e-9 and r["cca_means"] <= 1e-12, r
    assert res[0]["refine_unchanged"] is True and res[0]["refine_diff"] <= 1e-6, res[0]
