"""BASELINE.json configs[3]: 10 000 independent encrypted BoxBlur / GxKernel instances (64 x 64 image in one row of an
N = 8192 ciphertext) sharded across 1/2/4/8 B200 of one box, no data-path collective.

  python tools/stencil10k_bench.py [--instances 10000] [--batch 625]
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/stencil10k_bench.py

Every rank owns a contiguous range of the instances (abc_b200.sharding.instance_range), runs them in lock-step batches
through CudaCiphertextFactory with HOST buffers on both sides (createCiphertext from pinned slots ... decryptCiphertext to
pinned slots, the D2H of one batch under the next batch's kernels), and checks EVERY pixel of EVERY instance against the
reference's plain functions (naiveBoxBlur / naiveGxKernel: test/end-to-end/BoxBlurTest.cpp:23-43, GxKernelTest.cpp:20-44;
their wrap-around index is a cyclic rotation of the 4096-slot row).  The program is the batched canonical form
the RuntimeVisitor runs (abc_driver.cpp: stencilProgram): one rotateRows per non-zero tap, |weight| > 1 by multiplyPlain,
add / subtract into the accumulator.  Rank 0 prints one JSON line per program: instances/s end to end (wall clock between
barriers, max over ranks) and key switches/s."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertext, CudaCiphertextFactory  # noqa: E402
from abc_b200.sharding import instance_range, max_over_ranks  # noqa: E402

SIZE, N_POLY, SEED = 64, 8192, 4673838
BOX = ((1, 1, 1), (1, 1, 1), (1, 1, 1))
GX = ((1, 2, 1), (0, 0, 0), (-1, -2, -1))      # weightMatrix of GxKernelTest.cpp:22


def taps(w):
    return [(dx * SIZE + dy, w[dx + 1][dy + 1]) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if w[dx + 1][dy + 1]]


def naf_weight(k):
    """Non-zero digits of the non-adjacent form of |k|: the key switches rotateRows(k) costs with power-of-two Galois keys
    (SEAL's rotate_internal; SURVEY.md 8d: +-1, +-64 one key switch, +-63, +-65 two)."""
    k, n = abs(k), 0
    while k:
        if k & 1:
            n += 1
            k -= 2 - (k & 3)
        k >>= 1
    return n


def key_switches(w):
    return sum(naf_weight(k) for k, _ in taps(w))


def program(img, w):
    acc = None
    for k, weight in taps(w):
        r = img.rotateRows(k) if k else img.clone()
        if abs(weight) != 1:
            r = r.multiplyPlain([abs(weight)])
        if acc is None and weight > 0:
            acc = r
        elif acc is None:
            acc = r.negate()
        elif weight > 0:
            acc.addInplace(r)
        else:
            acc.subtractInplace(r)
    return acc


def expected(imgs, w):
    out = np.zeros_like(imgs)
    for k, weight in taps(w):
        out += weight * np.roll(imgs, -k, axis=1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=625, help="instances per lock-step batch (10000 = 16 x 625)")
    ap.add_argument("--passes", type=int, default=5, help="timed passes over all instances (median reported)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")     # timings only: there is no data-path collective
    lo, hi = instance_range(args.instances, world, rank)
    B = args.batch
    n_batches = (hi - lo + B - 1) // B
    f = CudaCiphertextFactory(N_POLY, device=local, batch=B, seed=SEED)
    lib, C = f._lib, __import__("ctypes")
    n = SIZE * SIZE
    # instance i's image: seeded by its global index (any sharding computes the same 10 000 results); the last batch of a
    # rank is padded with copies of its last instance
    rng_imgs = np.stack([np.random.default_rng(SEED + i).integers(0, 1025, size=n, dtype=np.int64) for i in range(lo, hi)])
    pad = n_batches * B - (hi - lo)
    if pad:
        rng_imgs = np.concatenate([rng_imgs, np.repeat(rng_imgs[-1:], pad, axis=0)])
    h_in = torch.from_numpy(rng_imgs).pin_memory()
    h_out = [torch.empty((B, N_POLY), dtype=torch.int64).pin_memory() for _ in range(2)]
    lines = []
    for name, w in (("BoxBlur", BOX), ("GxKernel", GX)):
        want = expected(rng_imgs, w)
        bad = 0

        def run_batch(b):
            h = C.c_void_p()
            f._ck(lib.abc_encode_encrypt(f._h, h_in[b * B:(b + 1) * B].data_ptr(), n, 0, C.byref(h)))
            img = CudaCiphertext(f, h)
            res = program(img, w)
            f._ck(lib.abc_decrypt_decode_async(f._h, res._h, h_out[b & 1].data_ptr()))
            return res

        # pass 1, untimed: every pixel of every instance is compared (the timed pass reuses its two staging buffers, and
        # comparing 20 MB per batch on the host would otherwise be what is measured); it is also the warm-up
        for b in range(n_batches):
            run_batch(b); f.sync()
            got = h_out[b & 1][:, :n].numpy()
            bad += int((got != want[b * B:(b + 1) * B]).any(axis=1).sum())
        if dist is not None:
            dist.barrier()
        # timed passes: the same batches back to back, the D2H of batch b under the kernels of batch b + 1; a pass over
        # 10 000 instances is a fraction of a second, so several passes are timed and the MEDIAN (of the per-pass maximum
        # over ranks) is reported, every pass listed
        walls, devs = [], []
        for _ in range(args.passes):
            if dist is not None:
                dist.barrier()
            launches0, ks0 = f.launch_count(), f.key_switch_count()
            t0 = time.perf_counter()
            f.timer_start()
            for b in range(n_batches):
                run_batch(b)
            f.sync()                           # the last copies are inside the timed region
            dev_ms = f.timer_stop()
            wall = time.perf_counter() - t0
            launches = f.launch_count() - launches0
            ks_per_batch = (f.key_switch_count() - ks0) / n_batches
            w_, d_ = max_over_ranks([wall, dev_ms * 1e-3], dist)
            walls.append(w_); devs.append(d_)
        last = h_out[(n_batches - 1) & 1][:, :n].numpy()
        bad += int((last != want[(n_batches - 1) * B:n_batches * B]).any(axis=1).sum())
        order = sorted(range(len(walls)), key=lambda i: walls[i])
        wall_max, dev_max = walls[order[len(order) // 2]], devs[order[len(order) // 2]]
        bad_total = bad
        if dist is not None:
            t = torch.tensor([bad], dtype=torch.int64)
            dist.all_reduce(t)
            bad_total = int(t.item())
        ks_naf = key_switches(w)
        ks = ks_per_batch                      # executed: one batched call = one key switch of every instance (the
                                               # rotation-prefix cache shares the first step of the NAF chains)
        if rank == 0:
            lines.append({
                "workload": "%d independent encrypted %s instances (64x64 image, BFV N=8192), sharded over %d GPU(s), no collectives" % (args.instances, name, world),
                "n_gpus": world, "instances": args.instances, "batch": B, "batches_per_rank": n_batches,
                "instances_per_s_e2e": args.instances / wall_max, "wall_s": wall_max, "device_s_max_over_ranks": dev_max, "wall_s_every_pass": [round(v, 5) for v in walls],
                "key_switches_per_instance": ks, "key_switches_per_instance_naf": ks_naf,
                "key_switches_per_s": args.instances * ks / wall_max,
                "h2d_bytes_per_instance": n * 8, "d2h_bytes_per_instance": N_POLY * 8,
                "gpu_launches_rank0": launches,
                "every_pixel_of_every_instance_checked": bad_total == 0, "mismatching_instances": bad_total})
            print(json.dumps(lines[-1]), flush=True)
    f.close()
    if rank == 0 and args.out:
        json.dump(lines, open(args.out, "w"), indent=1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
