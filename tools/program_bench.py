"""BASELINE.json configs[3]: N independent encrypted program instances (default 10 000) sharded across the GPUs of one
box with no collectives.  Programs are the hand-written batched forms of SURVEY.md 8(d): BoxBlur and GxKernel on a
64x64 image, HammingDistance and L2Distance on 4096-vectors, BFV N=8192 (SEAL defaults), keys shared.
Every instance is created from host slots, run, decrypted (timed) and CHECKED against the plain evaluation (after the clock).

  python tools/program_bench.py [--instances 10000] [--batch 250] [--programs boxblur,gx,hamming,l2]
  python -m torch.distributed.run --nproc-per-node G ... tools/program_bench.py ...
One JSON line per program on rank 0: instances/s over all ranks (max-over-ranks device+host time), strong scaling.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory  # noqa: E402
from abc_b200.sharding import instance_range, max_over_ranks  # noqa: E402

N_POLY, ROW, IMG = 8192, 4096, 64
SEED = 4673838
BOX = [[1, 1, 1], [1, 1, 1], [1, 1, 1]]
GX = [[1, 2, 1], [0, 0, 0], [-1, -2, -1]]          # weightMatrix of test/end-to-end/GxKernelTest.cpp:22


def stencil_program(img, w):
    """acc = sum over non-zero taps of weight * rotate(img, dx*64+dy); ciphertext on the left of every op."""
    acc = None
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            weight = w[dx + 1][dy + 1]
            if weight == 0:
                continue
            k = dx * IMG + dy
            r = img.rotateRows(k) if k else img.clone()
            if abs(weight) != 1:
                r.multiplyPlainInplace([abs(weight)])
            if acc is None:
                acc = r if weight > 0 else r.negate()
            elif weight > 0:
                acc.addInplace(r)
            else:
                acc.subtractInplace(r)
    return acc


def stencil_plain(img, w):
    out = np.zeros_like(img)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            if w[dx + 1][dy + 1]:
                out += w[dx + 1][dy + 1] * np.roll(img, -(dx * IMG + dy), axis=1)
    return out


def ladder_program(x, y):
    d = x.subtract(y)
    s = d.multiply(d)
    k = ROW // 2
    while k >= 1:
        s.addInplace(s.rotateRows(k))
        k //= 2
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=500)
    ap.add_argument("--programs", default="boxblur,gx,hamming,l2")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = instance_range(args.instances, world, rank)
    B = args.batch
    f = CudaCiphertextFactory(N_POLY, device=local, batch=B, seed=SEED)
    t = f.t

    def centre(v):
        v = v % t
        return np.where(v > t // 2, v - t, v)

    for prog in args.programs.split(","):
        stencil = prog in ("boxblur", "gx")
        w = BOX if prog == "boxblur" else GX
        hi_val = 2 if prog == "hamming" else 1025
        # synthetic inputs of every batch are made before the clock starts and the plain evaluation is compared after it
        # stops: the timed region is host slots -> createCiphertext -> program -> decryptCiphertext -> host, nothing else
        starts = list(range(lo, hi, B))
        ins, outs = [], []
        for start in starts:
            rng = np.random.default_rng(SEED + start)          # instance ids seed the inputs
            if stencil:
                ins.append((rng.integers(0, 1025, size=(B, ROW), dtype=np.int64),))
            else:
                ins.append((rng.integers(0, hi_val, size=(B, ROW), dtype=np.int64),
                            rng.integers(0, hi_val, size=(B, ROW), dtype=np.int64)))
        if dist:
            dist.barrier()
        f.sync()
        t0 = time.perf_counter()
        l0 = f.launch_count()
        for arrs in ins:
            if stencil:
                outs.append(f.decryptCiphertext(stencil_program(f.createCiphertext(arrs[0]), w)))
            else:
                outs.append(f.decryptCiphertext(ladder_program(f.createCiphertext(arrs[0]), f.createCiphertext(arrs[1]))))
        f.sync()
        dt = time.perf_counter() - t0
        checked = 0
        for start, arrs, out in zip(starts, ins, outs):
            n = min(B, hi - start)
            if stencil:
                assert np.array_equal(out[:n, :ROW], centre(stencil_plain(arrs[0], w))[:n]), prog
            else:
                assert np.array_equal(out[:n, 0], centre(((arrs[0] - arrs[1]) ** 2).sum(axis=1))[:n]), prog
            checked += n
        (dtmax,) = max_over_ranks([dt], dist, "cuda" if dist else "cpu")
        if rank == 0:
            print(json.dumps({"program": prog, "instances": args.instances, "n_gpus": world, "batch": B,
                              "instances_per_s": args.instances / dtmax, "seconds": dtmax, "scaling": "strong",
                              "checked_on_rank0": checked, "gpu_launches_rank0": f.launch_count() - l0,
                              "includes": "host slots -> createCiphertext, program, decryptCiphertext -> host; every "
                                          "instance compared with the plain evaluation after the clock stops"}),
                  flush=True)
    f.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
