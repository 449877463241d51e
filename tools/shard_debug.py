"""Limb-sharded vs single-device results, op by op (no oracle: sizes the CPU oracle is too slow for).
  python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 --master-port P tools/shard_debug.py [N] [k]
Every rank builds a sharded factory (NCCL communicator) and an unsharded one with the same seed on its own GPU, runs the
same ops on both and compares the limbs it owns after every op.  Prints one verdict line per op on rank 0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from abc_b200 import CudaCiphertextFactory  # noqa: E402
from tools.op_microbench import seal_primes  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    data = seal_primes(N, 55, k - 1)
    primes = data + seal_primes(N, 56, 1, skip=data)
    seed = 4673838
    fs = CudaCiphertextFactory(N, primes=primes, device=local, seed=seed, galois_steps=[1])
    f1 = CudaCiphertextFactory(N, primes=primes, device=local, seed=seed, galois_steps=[1])
    ident = [fs.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    fs.comm_init(rank, world, ident[0])
    lo, hi = fs.owned_limbs()
    rng = np.random.default_rng(3)
    d = rng.integers(0, 2, N)
    fs.set_encrypt_nonce(1); f1.set_encrypt_nonce(1)
    xs, x1 = fs.createCiphertext(d), f1.createCiphertext(d)
    bad = []

    def check(name, a, b):
        ga, gb = a.export()[0][:, lo:hi], b.export()[0][:, lo:hi]
        ok = np.array_equal(ga, gb)
        res = torch.tensor([0 if ok else 1])
        dist.all_reduce(res)
        if rank == 0:
            print("%-28s %s" % (name, "ok" if res.item() == 0 else "DIFFERS on %d rank(s)" % res.item()), flush=True)
        if not ok:
            diff = np.argwhere(ga != gb)
            print("  rank %d %s: %d words differ; first (poly, limb, coeff) = %s" % (rank, name, len(diff), diff[0] if len(diff) else None), flush=True)
            bad.append(name)

    check("encrypt", xs, x1)
    rs, r1 = xs.rotateRows(1), x1.rotateRows(1)
    check("rotate(1)", rs, r1)
    ms, m1 = xs.multiply(xs), x1.multiply(x1)
    check("square", ms, m1)
    ys, y1 = fs.createCiphertext(1 - d), f1.createCiphertext(1 - d)
    ps, p1 = xs.multiply(ys), x1.multiply(y1)
    check("multiply(x, y)", ps, p1)
    ss, s1 = ps.add(rs), p1.add(r1)
    check("add", ss, s1)
    qs, q1 = ss.rotateRows(1), s1.rotateRows(1)
    check("rotate(1) of the sum", qs, q1)
    # the deep chain, enqueued back to back (no host synchronisation between the ops)
    for depth in (1, 2, 4):
        cs, c1 = xs, x1
        for _ in range(depth):
            cs = cs.multiply(cs); cs.rotateRowsInplace(1)
            c1 = c1.multiply(c1); c1.rotateRowsInplace(1)
        check("chain depth %d" % depth, cs, c1)
    ok = np.array_equal(fs.decryptCiphertext(qs), f1.decryptCiphertext(q1))
    res = torch.tensor([0 if ok else 1]); dist.all_reduce(res)
    if rank == 0:
        print("%-28s %s" % ("decrypt", "ok" if res.item() == 0 else "DIFFERS"), flush=True)
        print("shard_debug %s: N=%d k=%d world=%d" % ("ok" if not bad and res.item() == 0 else "FAILED", N, k, world), flush=True)
    fs.close(); f1.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
