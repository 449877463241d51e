#!/usr/bin/env python
"""bench.py — BFV mul+relin & rotate throughput on B200 (BASELINE.json metric), one process per GPU.

Workload (BASELINE.json configs[1]): the batched L2Distance / HammingDistance program of SURVEY.md 8(d) at
N=8192 (SEAL BFVDefault, k=5), driven through CudaCiphertextFactory / the C ABI:
    d = x --- y;  s = d *** d;  s = s +++ rotate(s, 2048); ... ; s = s +++ rotate(s, 1);
= 1 sub, 1 mul+relin, 12 rotateRows (one key switch each), 12 adds per instance, result in slot 0.
A "step" runs the program on `--batch` independent encrypted instances per GPU (keys shared).
value = (mul+relin + rotate ops of all ranks) / device time; inputs are resident ciphertexts whose
working set exceeds L2.  e2e = the same through createCiphertext(host slots) ... decryptCiphertext(host).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POLY = 8192
N_VEC = 4096                       # vector length of the distance program (one batching row)
ROT_STEPS = [2048, 1024, 512, 256, 128, 64, 32, 16, 8, 4, 2, 1]
OPS_PER_INSTANCE = 1 + len(ROT_STEPS)   # mul+relin + rotates (the ops the metric counts)
SEED = 4673838                     # the reference's RAND_SEED (test/end-to-end/BoxBlurTest.cpp:111)
METRIC = "bfv_mul_relin_and_rotate_ops_per_s"
UNIT = "ops/s"
WORKLOAD = ("l2distance_batched: BFV N=8192 k=5 t=1032193, n=4096; per instance 1 sub + 1 mul+relin + "
            "12 rotateRows(2^j) + 12 add, through CudaCiphertextFactory")


def config_dict(batch, world):
    """The `config` object of the JSON line: both arms print exactly this (the reference arm times a bounded sample of the
    same workload on the host cores and says so in cpu_baseline.sample)."""
    L, row = 4, N_POLY * 8   # BFVDefault(8192): k = 5 primes, L = 4 data limbs
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "instances_total": batch * world,
            "parallelism": "independent instances sharded across GPUs, no collectives",
            "l2": "inputs+intermediates exceed L2 (%.0f MiB of ciphertext per operand)" % (batch * 2 * L * row / 2**20)}


def synth_inputs(batch, rank):
    """i.i.d. ints in [0,1024] (the reference's test distribution, BoxBlurTest.cpp:123), seeded per rank."""
    rng = np.random.default_rng(SEED + rank)
    return (rng.integers(0, 1025, size=(batch, N_VEC), dtype=np.int64),
            rng.integers(0, 1025, size=(batch, N_VEC), dtype=np.int64))


def expected_slot0(x, y):
    return ((x - y) ** 2).sum(axis=1)


# ------------------------------------------------------------------------------------------------ GPU arm
def program_gpu(x, y):
    d = x.subtract(y)
    s = d.multiply(d)
    for k in ROT_STEPS:
        r = s.rotateRows(k)
        s.addInplace(r)
    return s


class NvmlSampler:
    """SM clock and throttle reasons of ONE device through NVML, sampled every 5 ms from a thread (nvidia-smi needs longer
    to start than a timed region lasts, and on 8-GPU boxes produced no sample at all)."""

    def __init__(self, device):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
        self.sm, self.reasons, self.run = [], 0, False
        self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)

    def start(self):
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reasons |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self.run = False
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        return {"sm_mhz": int(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(k for k, v in names.items() if self.reasons & v), "samples": len(self.sm)}


def make_sampler(device):
    try:
        return NvmlSampler(device)
    except Exception:
        return ClockSampler(device)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=10)   # make sure no nvidia-smi query overlaps what is timed next
            except Exception:
                self.proc.kill()
        sm = [int(r[1]) for r in self.rows if len(r) >= 7 and r[1].isdigit()]
        mx = [int(r[2]) for r in self.rows if len(r) >= 7 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].startswith("Active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def pin_to_gpu_local_cores(device):
    """Bind this process to the CPU cores NVML reports as local to `device` (the host leg of the end-to-end path —
    pinned-buffer copies, the ctypes call sequence — otherwise wanders across sockets when 8 ranks share a box).
    Returns the number of cores bound, or 0 when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def ncu_pipes_for(kernel, batch):
    """pipe utilisation of `kernel` from the same committed ncu capture (profiles/traffic.json), else None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
        if rec and rec["batch"] == batch and rec["N"] == N_POLY and "fp64_pipe_pct_of_peak" in rec:
            return {"fp64_pipe_pct_of_peak": rec["fp64_pipe_pct_of_peak"], "issue_active_pct": rec["issue_active_pct"],
                    "l1_data_pipe_pct": rec["l1_data_pipe_pct"], "dram_pct_of_peak": rec["dram_pct_of_peak"],
                    "source": rec["source"]}
    except Exception:
        pass
    return None


def traffic_for(kernel, batch):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), else None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
        return rec["bytes_per_launch"] if rec and rec["batch"] == batch and rec["N"] == N_POLY else None
    except Exception:
        return None


REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "abc_ref_driver")


def run_ref_driver(cores, per_core, steps, warmup):
    """The CPU arm through the reference's own interpreter: oracle/_ref/abc_ref_driver = Parser + TypeCheckingVisitor +
    RuntimeVisitor of /root/reference (compiled unchanged into libabc_ref.a) over OracleCiphertextFactory (oracle/cpp), one
    walk per instance, `cores` forked workers.  Returns its JSON line, or None when the binary is not built."""
    if not os.path.exists(REF_DRIVER):
        return None
    p = subprocess.run([REF_DRIVER, str(N_POLY), str(cores), str(per_core), str(steps), str(warmup)],
                       capture_output=True, text=True, timeout=1200)
    if p.returncode != 0:
        raise RuntimeError("abc_ref_driver failed: " + p.stdout[-500:] + p.stderr[-500:])
    return json.loads(p.stdout.strip().splitlines()[-1])


def cpu_baseline_port(cores, seconds_budget=12.0):
    """The CPU baseline on the box's host cores, bounded sample (about `seconds_budget` s of wall time): through the
    reference's RuntimeVisitor when the driver is built (encrypt x, y + program + decrypt per instance), else the oracle
    driven from Python threads."""
    cal = run_ref_driver(cores, 1, 1, 0)
    if cal is not None:
        per = max(1, int(seconds_budget / max(cal["ms_per_step"] * 1e-3, 1e-3)))
        r = run_ref_driver(cores, per, 1, 0)
        return {"value": r["ops_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d instances on %d host processes (%.1f s): encrypt x, y + program + decrypt through the reference's "
                          "RuntimeVisitor over OracleCiphertextFactory (oracle/cpp; SEAL-3.6.5 restatement, SEAL itself not "
                          "installable); result check %s" % (r["instances_per_step"], cores, r["ms_per_step"] * 1e-3, r["result_check"])}
    return cpu_baseline_python_threads(cores, seconds_budget)


def cpu_baseline_python_threads(cores, seconds_budget=12.0):
    """The oracle (SEAL-3.6.5 restatement) running the same program on host cores; bounded sample."""
    from oracle.bfv_oracle import Oracle
    o = Oracle(N_POLY, seed=SEED)
    x, y = synth_inputs(cores * 8, 0)
    cts = [(o.encrypt_slots(x[i], 2 * i), o.encrypt_slots(y[i], 2 * i + 1)) for i in range(len(x))]

    def prog(i):
        d = o.sub(*cts[i % len(cts)])       # the encrypted inputs are reused round-robin: same work per instance
        s = o.mul_relin(d, d)
        for k in ROT_STEPS:
            s = o.add(s, o.rotate_rows(s, k))
        return s

    t0 = time.perf_counter()
    prog(0)
    per = time.perf_counter() - t0
    n = int(cores * max(1, int(seconds_budget / max(per, 1e-3))))   # ~seconds_budget of wall time on `cores` threads
    out = [None]

    def work(tid):
        for i in range(tid, n, cores):
            r = prog(i)
            if i == 0:
                out[0] = r

    th = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    w0 = int(expected_slot0(x[:1], y[:1])[0] % o.t)
    ok = int(o.decrypt_slots(out[0])[0]) == (w0 - o.t if w0 > o.t // 2 else w0)
    return {"value": n * OPS_PER_INSTANCE / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d instances of the same program on %d host threads (%.1f s); SEAL-3.6.5 restatement "
                      "(oracle/), SEAL itself not installable; result check %s" % (n, cores, dt, "ok" if ok else "FAILED")}


def per_op_numbers(f, x, y, B, reps=30):
    """BASELINE.json's metric names the ops separately: mul+relin, rotateRows and add alone at the bench batch, device-resident,
    each timed `reps` times between two synchronisations (CUDA events on the library stream): ops/s from the median,
    p50 / p99 of the batched launch-sequence latency.  rotateRows is timed eagerly (the default defers its key switch to
    the consumer) through abc_rotate_rows_add's sibling with no addend: ABC_EAGER_ROTATE semantics via rotate + export-free add."""
    lib = f._lib
    out = f.allocCiphertext()
    ops = {"mul_relin": lambda: f._ck(lib.abc_mul_relin(f._h, out._h, x._h, y._h)),
           # rotate_rows_add runs the key switch immediately (the addend rides in its ModDown): one rotation + one add
           "rotate_add": lambda: f._ck(lib.abc_rotate_rows_add(f._h, out._h, x._h, 1, y._h)),
           "add": lambda: f._ck(lib.abc_add(f._h, out._h, x._h, y._h))}
    res = {}
    for name, fn in ops.items():
        fn(); fn()
        f.sync()
        ts = []
        for _ in range(reps):
            f.timer_start()
            fn()
            ts.append(f.timer_stop())
        ts.sort()
        p50, p99 = ts[len(ts) // 2], ts[min(len(ts) - 1, int(0.99 * len(ts)))]
        res[name] = {"ops_per_s": B / (p50 * 1e-3), "p50_ms": round(p50, 4), "p99_ms": round(p99, 4)}
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from abc_b200 import CudaCiphertextFactory

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    host_cores_bound = 0 if os.environ.get("ABC_BENCH_NO_PIN") else pin_to_gpu_local_cores(local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator comes up; stdout carries the ONE JSON line, so the
        # process-level descriptor points at stderr while NCCL initialises
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    f = CudaCiphertextFactory(N_POLY, device=local, batch=B, seed=SEED)   # same keys on every rank (same seed)
    xs, ys = synth_inputs(B, rank)
    x, y = f.createCiphertext(xs), f.createCiphertext(ys)
    want0 = expected_slot0(xs, ys) % f.t
    want0 = np.where(want0 > f.t // 2, want0 - f.t, want0)

    # ---- device-resident timing
    for _ in range(args.warmup):
        s = program_gpu(x, y)
    f.sync()
    got = f.decryptCiphertext(s)
    got0 = got[:, 0] if B > 1 else got[:1]
    assert np.array_equal(got0, want0), "program result mismatch"
    sampler = make_sampler(local)
    sampler.start()
    for _ in range(3):                     # let the sampler come up while the GPU is under the same load
        program_gpu(x, y)
    f.sync()
    barrier()
    l0 = f.launch_count()
    f.timer_start()
    for _ in range(args.steps):
        s = program_gpu(x, y)
    ms = f.timer_stop()
    launches = f.launch_count() - l0
    barrier()
    clocks = sampler.stop()
    if world > 1:   # every rank samples its own device: report the slowest median and the union of the reasons
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        meds = [c["sm_mhz"] for c in allc if c["sm_mhz"] is not None]
        clocks = {"sm_mhz": min(meds) if meds else None, "sm_max_mhz": clocks["sm_max_mhz"],
                  "reasons": sorted({r for c in allc for r in c["reasons"]}), "samples": sum(c["samples"] for c in allc),
                  "sm_mhz_per_rank": [c["sm_mhz"] for c in allc]}

    # ---- end to end through the factory with host buffers (pinned), copies inside the timed region
    hx = torch.from_numpy(xs).pin_memory()
    hy = torch.from_numpy(ys).pin_memory()
    houts = [torch.empty((B, N_POLY), dtype=torch.int64).pin_memory() for _ in range(2)]
    lib, C = f._lib, __import__("ctypes")
    from abc_b200 import CudaCiphertext

    host_t = {"encode_encrypt_x": 0.0, "encode_encrypt_y": 0.0, "program": 0.0, "decrypt_async": 0.0, "release_handles": 0.0}

    def e2e_step(i):
        """createCiphertext(x), createCiphertext(y) from pinned host slots; the program; decryptCiphertext to pinned host
        memory.  The decrypted slots of step i leave on the library's D2H stream while step i + 1 computes
        (abc_decrypt_decode_async); every copy is inside the timed region, which ends with abc_sync.  host_t accumulates the
        HOST time of each part of the call sequence (enqueue only)."""
        t0 = time.perf_counter()
        hx_ct, hy_ct = C.c_void_p(), C.c_void_p()
        f._ck(lib.abc_encode_encrypt(f._h, hx.data_ptr(), N_VEC, 0, C.byref(hx_ct)))
        t1 = time.perf_counter()
        f._ck(lib.abc_encode_encrypt(f._h, hy.data_ptr(), N_VEC, 0, C.byref(hy_ct)))
        t2 = time.perf_counter()
        cx, cy = CudaCiphertext(f, hx_ct), CudaCiphertext(f, hy_ct)
        r = program_gpu(cx, cy)
        t3 = time.perf_counter()
        f._ck(lib.abc_decrypt_decode_async(f._h, r._h, houts[i & 1].data_ptr()))
        t4 = time.perf_counter()
        del cx, cy, r
        t5 = time.perf_counter()
        for k, v in zip(host_t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            host_t[k] += v

    for i in range(max(2, args.warmup)):
        e2e_step(i)
    f.sync()
    assert np.array_equal(houts[0][:, 0].numpy(), want0) and np.array_equal(houts[1][:, 0].numpy(), want0), "e2e result mismatch"
    barrier()
    t0 = time.perf_counter()
    f.timer_start()
    t_enq = 0.0
    for k in host_t:
        host_t[k] = 0.0
    for i in range(args.steps):
        te = time.perf_counter()
        e2e_step(i)
        t_enq += time.perf_counter() - te
    f.sync()                                   # the last steps' D2H copies are part of the timed region
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = f.timer_stop()
    assert np.array_equal(houts[(args.steps - 1) & 1][:, 0].numpy(), want0), "e2e result mismatch (timed region)"
    barrier()

    # ---- where an end-to-end step spends its time: one more step, phase by phase, DEVICE time of each phase (CUDA events
    # on the library stream, a synchronisation after each phase: no overlap), the copies alone through torch
    def phase(fn):
        f.sync()
        f.timer_start()
        out = fn()
        return f.timer_stop(), out

    def enc_both():
        hx_ct, hy_ct = C.c_void_p(), C.c_void_p()
        f._ck(lib.abc_encode_encrypt(f._h, hx.data_ptr(), N_VEC, 0, C.byref(hx_ct)))
        f._ck(lib.abc_encode_encrypt(f._h, hy.data_ptr(), N_VEC, 0, C.byref(hy_ct)))
        return CudaCiphertext(f, hx_ct), CudaCiphertext(f, hy_ct)

    def copy_ms(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    d_in = torch.empty((B, N_VEC), dtype=torch.int64, device="cuda")
    d_res = torch.empty((B, N_POLY), dtype=torch.int64, device="cuda")
    rounds = []
    for _ in range(4):   # the first round pays one-time costs (first torch copy, pool growth); the minimum of the rest is
        # reported (a phase that starts on an idle GPU is now and then measured at half speed for its first milliseconds)
        ms_h2d = copy_ms(lambda: (d_in.copy_(hx, non_blocking=True), d_in.copy_(hy, non_blocking=True)))
        ms_enc, (cx, cy) = phase(enc_both)
        ms_prog, res = phase(lambda: program_gpu(cx, cy))
        ms_dec, _ = phase(lambda: f._ck(lib.abc_decrypt_decode_async(f._h, res._h, houts[0].data_ptr())))
        f.sync()
        ms_d2h = copy_ms(lambda: houts[1].copy_(d_res, non_blocking=True))
        rounds.append((ms_h2d, ms_enc, ms_prog, ms_dec, ms_d2h))
    ms_h2d, ms_enc, ms_prog, ms_dec, ms_d2h = (min(r[i] for r in rounds[1:]) for i in range(5))
    breakdown = {"h2d_copy_alone": round(ms_h2d, 3), "encode_encrypt_incl_h2d": round(ms_enc, 3), "program": round(ms_prog, 3),
                 "decrypt_decode_kernels": round(ms_dec, 3), "d2h_copy_alone": round(ms_d2h, 3),
                 "note": "device time of each phase run alone (CUDA events, a synchronisation after each phase: no overlap); in the "
                         "timed region the H2D of y runs under the encryption of x and the D2H of step i under the kernels of step i+1"}
    del cx, cy, res

    e2e_per_rank = [round(e2e_ms / args.steps, 3)]
    if world > 1:
        allv = [None] * world
        dist.all_gather_object(allv, e2e_per_rank[0])
        e2e_per_rank = allv
    # max over ranks
    from abc_b200.sharding import max_over_ranks
    ms, e2e_ms, e2e_wall = max_over_ranks([ms, e2e_ms, e2e_wall], dist if world > 1 else None, "cuda")
    e2e_dev = e2e_ms
    e2e_ms = max(e2e_ms, e2e_wall)

    line = None
    if rank == 0:
        # ---- per-kernel device time of one more step (CUDA events on the library's stream, outside the timed region)
        f.profile_enable(True)
        program_gpu(x, y)
        prof = f.profile()
        f.profile_enable(False)
        tot = sum(r["ms"] for r in prof)
        top = max(prof, key=lambda r: r["ms"])
        L, k = f.L, f.k
        row = N_POLY * 8
        # algorithmic bytes per launch of each kernel family (DESIGN.md "kernels")
        nb = L + 1
        # rows per polynomial of the FP64 BEHZ path: L + ceil((32 + bits(t) + bits(Q) + 8) / 43) (lib.cu build_tables)
        qbits = sum(int(p).bit_length() for p in f.primes[:L])
        W2 = L + (32 + int(f.t).bit_length() + qbits + 8 + 42) // 43
        alg = {"ks_modup_ntt": B * (L + k * L) * row, "ks_inner": B * (k * L + 2 * k) * row + 2 * k * L * row,
               "ks_intt_special": B * 4 * row, "ks_intt_moddown": B * (2 * L + 2 + 2 * L + 2 * L) * row,
               # fused tail of a rotation's key switch: T rows in, sigma(c0) and the addend in, sum out, + key
               "ks_inner_intt_moddown": B * (k * L + L + 2 * L + 2 * L) * row + 2 * k * L * row,
               # whole key switch of a rotation in one grid (chained ModUp + tail rows, or the single-launch kernel): c1 and
               # sigma(c0) in, the addend ciphertext in, the sum out, + key; the ModUp block T is not algorithmic traffic
               "ks_chain": B * (L + L + 2 * L + 2 * L) * row + 2 * k * L * row,
               "ks_fused": B * (L + L + 2 * L + 2 * L) * row + 2 * k * L * row,
               "ks_chain_relin": B * (L + 2 * L + 2 * L) * row + 2 * k * L * row,   # relinearisation: c2 in, (c0, c1) in, sum out
               "behz_ntt_q": B * 4 * L * row, "behz_ntt_bsk": B * 4 * nb * row,      # d *** d: squaring path, 2 of 4 polys
               "behz_intt_q": B * 6 * L * row, "behz_intt_bsk": B * 6 * nb * row,
               "behz_ntt": B * 4 * W2 * row, "behz_intt": B * 6 * W2 * row,
               "behz_tensor_intt": B * (4 + 3) * W2 * row,   # d *** d: 2 operand rows read by 3 products (1 + 2 + 1 row reads), 3 out
               "behz_lift": B * 2 * (L + 2 * L + 1) * row, "behz_tensor": B * 7 * (2 * L + 1) * row,
               "behz_scale": B * 3 * (3 * L + 1) * row, "add": B * 6 * L * row, "sub": B * 6 * L * row}
        peaks, how = measured_peaks()
        per_launch_ms = top["ms"] / top["launches"]
        ach = alg.get(top["kernel"], 0) / (per_launch_ms * 1e-3) / 1e9
        bf_peak = f.measure_butterfly_peak()
        imad, iadd = f.measure_int_peak()
        logn = N_POLY.bit_length() - 1
        # limb-pipeline rows per launch and the arithmetic class they run (ntt.cuh): the key-level primes' class, or
        # Shoup (class 0) for the 61-bit Bsk rows of the BEHZ product
        arq = f.ntt_arith_class()
        ntt_rows = {"ks_modup_ntt": (k * L, arq), "ks_intt_special": (2, arq),
                    "ks_intt_moddown": (2 * L + (0 if any(r["kernel"] == "ks_intt_special" for r in prof) else 2), arq),
                    "ks_inner_intt_moddown": (2 * L + 2, arq), "ks_chain": (k * L + 2 * L + 2, arq),
                    "ks_fused": (k * L + 2 * L + 2, arq), "ks_chain_relin": (k * L + 2 * L + 2, arq), "behz_ntt_q": (2 * L, arq), "behz_ntt_bsk": (2 * nb, 0),
                    "behz_intt_q": (3 * L, arq), "behz_intt_bsk": (3 * nb, 0),
                    # FP64 BEHZ (behz_f64.cuh): q rows + the sub-2^45 auxiliary rows in one launch each, squaring path
                    "behz_ntt": (2 * W2, arq), "behz_intt": (3 * W2, arq), "behz_tensor_intt": (3 * W2, arq)}
        bf_peaks = {arq: bf_peak, 0: f.measure_butterfly_peak(0)}
        ntt_ms = sum(r["ms"] for r in prof if r["kernel"] in ntt_rows)
        bf_per_row = (N_POLY // 2) * logn
        ntt_bf = sum(r["launches"] * ntt_rows[r["kernel"]][0] for r in prof if r["kernel"] in ntt_rows) * B * bf_per_row
        # time the same butterflies would take register-resident, class by class
        ideal_ms = sum(r["launches"] * ntt_rows[r["kernel"]][0] * B * bf_per_row / bf_peaks[ntt_rows[r["kernel"]][1]] * 1e3
                       for r in prof if r["kernel"] in ntt_rows)
        top_rows = ntt_rows.get(top["kernel"], (0, arq))[0]
        top_bf = top_rows * B * bf_per_row                                  # butterflies of one launch of the dominant kernel
        top_inner = (2 * k * L * N_POLY * B) if top["kernel"].startswith("ks_") else 0   # key inner-product modular products
        t_hbm = alg.get(top["kernel"], 0) / (peaks["hbm_gbs"] * 1e9)
        t_pipe = (top_bf + 0.75 * top_inner) / bf_peak
        bound = "fp64" if t_pipe >= t_hbm else "hbm"
        roofline = {"kernel": top["kernel"], "bound": bound,
                    "achieved": (top_bf + 0.75 * top_inner) / (per_launch_ms * 1e-3) / 1e9 if bound == "fp64" else ach,
                    "peak": bf_peak / 1e9 if bound == "fp64" else peaks["hbm_gbs"],
                    "unit": "G butterfly-equivalents/s (8 FP64-pipe instructions each)" if bound == "fp64" else "GB/s",
                    "frac": max(t_pipe, t_hbm) / (per_launch_ms * 1e-3),
                    "frac_butterflies_only": top_bf / bf_peak / (per_launch_ms * 1e-3),
                    "ms_per_launch": per_launch_ms, "t_roof_ms": max(t_pipe, t_hbm) * 1e3,
                    "work_per_launch": {"ntt_butterflies": top_bf, "inner_product_modmuls": top_inner},
                    "traffic": traffic_for(top["kernel"], B), "algorithmic_bytes_per_launch": alg.get(top["kernel"], 0),
                    "share_of_step": top["ms"] / tot, "ncu": ncu_pipes_for(top["kernel"], B),
                    "peak_source": "k_peak_butterfly_f64, measured in this run" if bound == "fp64" else how,
                    "hbm_view": {"achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                                 "peak_source": how, "note": "the non-binding roof: algorithmic bytes / launch time"}}
        ops_total = B * OPS_PER_INSTANCE * world
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        cpu = cpu_baseline_port(cores) if world == 1 and not args.no_cpu else None
        per_op = per_op_numbers(f, x, y, B)
        line = {
            "metric": METRIC, "value": ops_total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_dict(B, world),
            "e2e": {"value": ops_total * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": 2 * B * N_VEC * 8, "d2h_bytes_per_step": B * N_POLY * 8,
                    "device_ms_per_step": e2e_dev / args.steps, "wall_ms_per_step": e2e_wall / args.steps,
                    "frac_of_device_resident": (ms / args.steps) / (e2e_ms / args.steps),
                    "breakdown_ms": breakdown, "host_cores_bound_to_gpu": host_cores_bound, "device_ms_per_step_per_rank": e2e_per_rank,
                    "host_enqueue_ms_per_step": round(t_enq * 1e3 / args.steps, 3),
                    "host_enqueue_ms_by_call": {k: round(v * 1e3 / args.steps, 3) for k, v in host_t.items()},
                    "includes": "createCiphertext(x), createCiphertext(y) from pinned host slots, program, decryptCiphertext to pinned host "
                                "memory; the D2H of step i overlaps step i+1 (abc_decrypt_decode_async), the region ends with abc_sync"},
            "gpu_launches": launches,
            "clocks": clocks,
            # t_roof = max(algorithmic bytes / HBM, work / pipe peak).  For the key switch the FP64 pipe binds:
            # work = the launch's NTT butterflies (8 FP64 instructions each) + its key inner-product modular products
            # (6 FP64 instructions = 0.75 butterfly each), peak = register-resident butterflies/s measured in this run.
            "roofline": roofline,
            "int_roofline": {"unit": "64-bit modular NTT butterflies/s (all limb-pipeline kernels of the step)",
                             "achieved": ntt_bf / (ntt_ms * 1e-3), "peak": bf_peak,
                             "frac": ideal_ms / ntt_ms, "ntt_share_of_step": ntt_ms / tot,
                             "arith_class": arq, "peak_shoup_class": bf_peaks[0],
                             "frac_definition": "time the step's butterflies take register-resident at their class's peak "
                                                "(FP64-pipe class for the q rows, Shoup class for the 61-bit Bsk rows) / "
                                                "measured time of the limb-pipeline kernels",
                             "imad_per_s": imad, "iadd_lop_per_s": iadd,
                             "peak_source": "register-resident butterflies of the same arithmetic class, no memory "
                                            "traffic, measured in this run (k_peak_butterfly); IMAD / IADD+LOP issue "
                                            "rates from k_peak_imad / k_peak_iadd"},
            "kernels": [{"kernel": r["kernel"], "launches": r["launches"], "ms": round(r["ms"], 4),
                         "share": round(r["ms"] / tot, 4)} for r in sorted(prof, key=lambda r: -r["ms"])],
            "per_op": dict(per_op, programs_per_s=B * world * args.steps / (ms * 1e-3)),
        }
        if cpu:
            line["cpu_baseline"] = cpu
    f.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (SEAL is not installable here, so the oracle port),
    all host threads, same end-to-end pipeline as our `e2e`: encrypt x, y; program; decrypt."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    n = cores * args.ref_instances_per_core
    r = run_ref_driver(cores, args.ref_instances_per_core, args.steps, args.warmup)
    if r is not None:
        v, dt_ms = r["ops_per_s"], r["ms_per_step"]
        assert r["result_check"] == "ok", "reference arm result mismatch"
        sample = ("%d instances per step on %d host processes: encrypt x,y + program + decrypt, one RuntimeVisitor walk per "
                  "instance over OracleCiphertextFactory" % (n, cores))
    else:
        v, dt_ms, sample = reference_python_threads(args, cores, n)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(args.batch, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; SEAL-3.6.5 restatement (oracle/), SEAL itself not installable here"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def reference_python_threads(args, cores, n):
    """Fallback of the reference arm when abc_ref_driver is not built: the oracle driven from Python threads."""
    from oracle.bfv_oracle import Oracle
    o = Oracle(N_POLY, seed=SEED)
    xs, ys = synth_inputs(n, 0)
    res = np.zeros(n, dtype=np.int64)

    def one(i, nonce):
        x, y = o.encrypt_slots(xs[i], nonce), o.encrypt_slots(ys[i], nonce + 1)
        s = o.sub(x, y)
        s = o.mul_relin(s, s)
        for k in ROT_STEPS:
            s = o.add(s, o.rotate_rows(s, k))
        res[i] = o.decrypt_slots(s)[0]

    def step(base):
        def work(tid):
            for i in range(tid, n, cores):
                one(i, base + 2 * i)
        th = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
        [t.start() for t in th]
        [t.join() for t in th]

    for w in range(args.warmup):
        step(w * 2 * n)
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        step((args.warmup + s_) * 2 * n)
    dt = time.perf_counter() - t0
    want = expected_slot0(xs, ys) % o.t
    want = np.where(want > o.t // 2, want - o.t, want)
    assert np.array_equal(res, want), "reference arm result mismatch"
    sample = "%d instances per step on %d host threads: encrypt x,y + program + decrypt" % (n, cores)
    return n * OPS_PER_INSTANCE * args.steps / dt, dt / args.steps * 1e3, sample


def run_deep_chain(args):
    """BASELINE.json configs[4] through the same launch contract: one (mul+relin, rotate) x depth chain per step on ONE
    ciphertext whose RNS limbs are sharded over the ranks (NCCL all-gather in front of the key-switch ModUp, column
    all-gathers around the base conversions).  Strong scaling: the work is fixed, value = ops of the chain / device time
    (max over ranks)."""
    from types import SimpleNamespace
    from tools import deep_chain_bench
    world = int(os.environ.get("WORLD_SIZE", 1))
    a = SimpleNamespace(depth=args.depth, n=65536, limbs=0, reps=max(1, args.steps), profile=True, out="")
    r = deep_chain_bench.run(a, emit=False)
    if r is not None:
        print(json.dumps({
            "metric": METRIC, "value": r["ops_per_s"], "unit": UNIT, "n_gpus": world, "steps": a.reps, "warmup": 1,
            "ms_per_step": r["ms_per_pair"] * args.depth, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": {"workload": r["workload"], "parallelism": "RNS limbs sharded over the ranks"},
            "ms_per_pair": r["ms_per_pair"], "limbs_per_rank": r["limbs_per_rank"],
            "allgather_bytes_received_per_pair_per_rank": r["allgather_bytes_received_per_pair_per_rank"],
            "nccl_collectives_per_pair": r["nccl_collectives_per_pair"], "decrypt_check": r["decrypt_check"],
            "gpu_launches": r["gpu_launches_rank0"], "kernels": r.get("kernels_rank0_one_chain")}))
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=592, help="independent instances per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-instances-per-core", type=int, default=8)
    ap.add_argument("--workload", default="l2distance", choices=["l2distance", "deep_chain", "stencil10k"],
                    help="deep_chain: BASELINE.json configs[4], N=65536 multiplicative chain, RNS limbs sharded over the ranks; "
                         "stencil10k: configs[3], 10 000 BoxBlur / GxKernel instances sharded over the ranks (tools/stencil10k_bench.py, "
                         "one JSON line per program)")
    ap.add_argument("--depth", type=int, default=8, help="deep_chain: (mul+relin, rotate) pairs per step")
    args = ap.parse_args()
    if args.workload == "deep_chain":
        return run_deep_chain(args)
    if args.workload == "stencil10k":
        from tools import stencil10k_bench
        sys.argv = [sys.argv[0], "--passes", str(max(1, args.steps))]
        return stencil10k_bench.main()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
