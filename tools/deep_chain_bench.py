"""BASELINE.json configs[4]: N = 65536 deep multiplicative chain with RNS-limb sharding and an NCCL all-gather at the
key-switch ModUp, 1..8 B200 of one box (one process per GPU):
    repeat D times:  x = x *** x;  x = rotate(x, 1)
Parameters: 30 x 55-bit + one 56-bit prime from SEAL's get_primes rule (SEAL has no BFVDefault above 32768),
t = 786433, only the relin key and the Galois key for step 1 (one key-switching key is 975 MiB).

  python tools/deep_chain_bench.py [--depth 8] [--n 65536] [--limbs 31]
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/deep_chain_bench.py
Rank 0 prints one JSON line: ops/s (mul+relin and rotate each count one op), device ms per (mul+relin, rotate) pair as
max over ranks, bytes all-gathered per pair, and the decrypted-slot check.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory  # noqa: E402
from abc_b200.sharding import max_over_ranks  # noqa: E402


def get_primes(N, bits, count, skip=()):
    from tools.op_microbench import seal_primes
    return seal_primes(N, bits, count, skip)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--limbs", type=int, default=0, help="k, including the special prime (0: 25 = 24 data limbs, divisible by 2/4/8 ranks)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--profile", action="store_true", help="rank 0: per-kernel device time of one more chain")
    ap.add_argument("--out", default="", help="rank 0: also write the JSON line to this file")
    args = ap.parse_args()
    run(args)


def run(args, emit=True):
    """Returns the result dict on rank 0 (None elsewhere)."""
    if not args.limbs:
        args.limbs = 25
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("gloo")
    N = args.n
    data = get_primes(N, 55, args.limbs - 1)
    primes = data + get_primes(N, 56, 1, skip=data)
    f = CudaCiphertextFactory(N, primes=primes, device=local, seed=4673838, galois_steps=[1])
    if world > 1:
        ident = [f.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        f.comm_init(rank, world, ident[0])
    L = f.L
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2, N)                 # 0/1 slots: x*x = x, so the expected result is a pure rotation
    x0 = f.createCiphertext(bits)

    def chain(x):
        for _ in range(args.depth):
            x = x.multiply(x)
            x.rotateRowsInplace(1)
        return x

    y = chain(x0)                                 # warm-up + correctness
    got = f.decryptCiphertext(y)
    half = N // 2
    want = np.concatenate([np.roll(bits[:half], -args.depth), np.roll(bits[half:], -args.depth)])
    ok = bool(np.array_equal(got, want))
    times = []
    for _ in range(args.reps):
        f.sync()
        if dist:
            dist.barrier()
        f.timer_start()
        y = chain(x0)
        times.append(f.timer_stop())
    (ms,) = max_over_ranks([min(times)], dist)
    lo, hi = f.owned_limbs()
    b0, n0 = f.comm_stats()
    chain(x0)
    f.sync()
    b1, n1 = f.comm_stats()
    prof = None
    if args.profile:                               # every rank runs the profiled chain (it contains collectives)
        f.profile_enable(True)
        chain(x0)
        prof = sorted(f.profile(), key=lambda r: -r["ms"])
        f.profile_enable(False)
    result = None
    if rank == 0:
        pair_ms = ms / args.depth
        gathered = (b1 - b0) / args.depth          # all-gather payload this rank received per (mul+relin, rotate) pair
        line = json.dumps({
            "workload": "deep chain: %d x (x = x***x; x = rotate(x,1)), BFV N=%d k=%d (%dx55-bit + one 56-bit prime), batch 1" % (
                args.depth, N, f.k, f.k - 1),
            "n_gpus": world, "ops_per_s": 2 * args.depth / (ms * 1e-3), "ms_per_pair": pair_ms, "scaling": "strong",
            "limbs_per_rank": hi - lo, "allgather_bytes_received_per_pair_per_rank": int(gathered),
            "nccl_collectives_per_pair": (n1 - n0) / args.depth,
            "allgather_gbs_if_serial": None if not gathered else gathered / (pair_ms * 1e-3) / 1e9,
            "decrypt_check": "ok" if ok else "FAILED", "gpu_launches_rank0": f.launch_count(),
            "kernels_rank0_one_chain": prof})
        result = json.loads(line)
        if emit:
            print(line, flush=True)
        if args.out:
            open(args.out, "w").write(line + "\n")
    assert ok, "deep chain result mismatch on rank %d" % rank
    f.close()
    if dist:
        dist.barrier()
        if emit:
            dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
