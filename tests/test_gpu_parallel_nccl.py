"""N > 1 on real hardware: two ranks, one per GPU, NCCL.  The batch is sharded with e2e_tts_b200.parallel, every rank
synthesises its rows with the CUDA path, rank 0 gathers the waveforms and checks them bit for bit against the whole
batch synthesised on its own GPU (utterances are independent and the kernels deterministic, so a wrong split / offset /
row order in the gather cannot hide).  Also the int16 PCM variant (travels as bytes).  Skips below two GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["E2E_ROOT"])
import torch, torch.distributed as dist
import e2e_tts_b200 as pkg
from e2e_tts_b200 import parallel, synthetic as sy
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
voc = pkg.HifiGan(sy.DEFAULT_CONFIG)
voc.load_state_dict(sy.make_state_dict(sy.DEFAULT_CONFIG, 3, "strong"))
voc = voc.eval().to(dev)
ok = True
for B in (5, 8):
    mel = sy.mel_like(B, 40, 11).to(dev)
    got = parallel.synthesize_sharded(voc, mel, dst=0)
    got16 = parallel.synthesize_sharded(voc.forward_pcm16, mel, dst=0)
    if rank == 0:
        with torch.no_grad():
            want = voc(mel).squeeze(1)
            want16 = voc.forward_pcm16(mel)
        ok = ok and got.shape == want.shape and torch.equal(got, want) and torch.equal(got16, want16)
    else:
        ok = ok and got is None and got16 is None
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
print("NCCL_GATHER_OK" if flag.item() == 1 else "NCCL_GATHER_MISMATCH", flush=True)
sys.exit(0 if flag.item() == 1 else 1)
'''


@pytest.mark.gpu
def test_two_rank_nccl_sharded_synthesis_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "nccl_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, E2E_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_GATHER_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
