"""The host-buffer path under torchrun, every rank on its own GPU at the same time, with the per-chunk timeline of two
ranks (SARLACC_DEBUG_TIMING): python -m torch.distributed.run --nproc-per-node N tools/e2e_probe_n.py [nreads] [reps]"""
import os
import sys
import time

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
if rank in (0, world - 1) and len(sys.argv) > 3:
    os.environ["SARLACC_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from sarlacc_b200 import native, synth, _lib, ReadSet  # noqa: E402

torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
front, back, widths, _ = synth.mock_windows_device(n, A1, A2, seed=2000, first_index=rank * n, device=local)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
front = ReadSet(pin(front.seq_pool), front.seq_off, pin(front.qual_pool), front.qual_off, front.names)
back = ReadSet(pin(back.seq_pool), back.seq_off, pin(back.qual_pool), back.qual_off, back.names)
enc = native.phred_encoding()
w = widths.astype(np.int32)
_lib.lib.sarlacc_set_devices((_lib.C.c_int * 1)(local), 1)
keep = {}
for r in range(reps + 1):
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    native.adaptor_align_windows(front, back, enc, 5.0, 1.0, A1, A2, ([16, 42], [28, 46]), ([], []), read_width=w, reuse=keep)
    dt = time.perf_counter() - t0
    if r:
        print("rank %d: adaptor_align_windows %.1f ms  %.2f M reads/s  %s" % (rank, dt * 1e3, n / dt / 1e6, native.last_pair_timing()), flush=True)
if world > 1:
    dist.destroy_process_group()
