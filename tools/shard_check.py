"""Limb-sharded key switch across GPUs, parity against the oracle (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_check.py [N]
Every rank owns a contiguous range of RNS limbs of every ciphertext; rotate all-gathers c1 in front of ModUp (NCCL),
mul+relin all-gathers its operands; owned limbs must equal the single-device / oracle result bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from abc_b200 import CudaCiphertextFactory  # noqa: E402
from oracle.bfv_oracle import Oracle  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")           # only to share the NCCL unique id and to agree on the verdict
    seed = 4673838
    f = CudaCiphertextFactory(N, device=local, seed=seed)
    ident = [f.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    f.comm_init(rank, world, ident[0])
    lo, hi = f.owned_limbs()
    o = Oracle(N, seed=seed)
    rng = np.random.default_rng(1)
    da, db = rng.integers(0, 1025, N), rng.integers(0, 1025, N)
    f.set_encrypt_nonce(1)
    a, b = f.createCiphertext(da), f.createCiphertext(db)
    a_w, b_w = o.encrypt_slots(da, 1), o.encrypt_slots(db, 2)

    def own(x):
        return np.ascontiguousarray(x[:, lo:hi])

    def check(name, ct, want):
        got = ct.export()[0]
        assert np.array_equal(own(got), own(want)), "rank %d: %s differs on owned limbs [%d,%d)" % (rank, name, lo, hi)

    check("encrypt", a, a_w)
    s_w = o.add(a_w, b_w); s = a.add(b); check("add", s, s_w)
    d_w = o.sub(s_w, b_w); d = s.subtract(b); check("sub", d, d_w)
    m_w = o.mul_relin(d_w, b_w); m = d.multiply(b); check("mul+relin", m, m_w)
    r_w = o.rotate_rows(m_w, 1); r = m.rotateRows(1); check("rotate(1)", r, r_w)
    r2_w = o.rotate_rows(r_w, -24); r2 = r.rotateRows(-24); check("rotate(-24) NAF", r2, r2_w)
    p_w = o.multiply_plain(r2_w, o.encode(o.expand([3, -2]))); p = r2.multiplyPlain([3, -2]); check("mul plain", p, p_w)
    q_w = o.add_plain(p_w, o.encode(o.expand([7]))); q = p.addPlain([7]); check("add plain", q, q_w)
    sq_w = o.mul_relin(q_w, q_w); sq = q.multiply(q); check("square", sq, sq_w)
    full = sq.clone().allgather().export()[0]
    assert np.array_equal(full, sq_w), "rank %d: all-gathered ciphertext differs" % rank
    assert np.array_equal(f.decryptCiphertext(sq), o.decrypt_slots(sq_w)), "rank %d: decrypt differs" % rank
    ok = torch.tensor([1])
    dist.all_reduce(ok)
    if rank == 0:
        print("shard_check ok: N=%d world=%d, every rank bit-exact on its limbs; launches(rank0)=%d" % (N, world, f.launch_count()))
    f.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
