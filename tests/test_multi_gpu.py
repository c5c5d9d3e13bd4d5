"""Row-slab (multi-GPU) path: host logic on CPU with gloo (world_size 2 and 3),
device parity on >= 2 GPUs with NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "mp_slab_check.py")


def _torchrun(nproc, mode, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), SCRIPT, mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_host_logic_gloo(world):
    r = _torchrun(world, "gloo", 29611 + world)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_slab_parity(mode):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    r = _torchrun(world, mode, 29641 + (mode == "peer"), timeout=1200)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("PARITY")]
    print("\n".join(lines))
    keep = os.path.join(ROOT, "gpurun_out")          # kept record of the run (copied to profiles/ by hand)
    if os.path.isdir(keep):
        with open(os.path.join(keep, f"multi_gpu_parity_{mode}_world{world}.log"), "w") as fh:
            fh.write("\n".join(lines) + "\n")
    want = 10 if mode == "peer" else 9              # slab cases + the driver case
    assert r.returncode == 0 and "FAIL" not in r.stdout and len(lines) == want, r.stdout[-3000:] + r.stderr[-3000:]
    if mode == "peer":
        assert sum("peer_kernel=True" in ln for ln in lines) >= 7   # the equal-slab cases really ran the NVLink kernels
