"""Multi-GPU parity of the training path (SyncBatchNorm statistics, bucketed overlapped gradient all-reduce, captured
as one CUDA graph): 2 ranks must reproduce one process on the concatenated batch.  Needs >= 2 GPUs (skipped on the
single-GPU box; run with `gpurun --gpus 2`).  The host logic of the same code is covered on CPU with gloo
(tests/test_ddp_gloo.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_two_gpus_reproduce_single_process_training_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "mp_train_check.py")],
                       capture_output=True, text=True, timeout=540, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("MPCHECK ")]
    assert lines, (r.stdout[-2000:], r.stderr[-3000:])
    res = json.loads(lines[-1][8:])
    print(res)
    assert r.returncode == 0 and res["all_ranks_ok"], res
    assert res["regions_launched_in_backward"] and all(res["regions_launched_in_backward"])
