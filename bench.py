#!/usr/bin/env python3
"""bench.py — MSM throughput (scalar.point terms / second) on BASELINE config C2:
curve25519 / ristretto255 MSM, random 252-bit scalars, n = 2^20 terms per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path

A "step" is one complete MSM (one call of the hot path) over the synthetic batch.
  value  device-resident: scalars + generators already in HBM in the ABI layout; timed with CUDA
         events on the library stream (the stream every kernel is launched on).
  e2e    the same metric through the reference-facing C ABI with HOST (pinned) buffers: H2D of the
         192 B/term inputs and D2H of the 32-byte commitment are inside the timed region.
N > 1 (torchrun, one process per GPU): weak scaling — every rank owns a 2^20-term generator-range
shard of one N*2^20-term MSM, computes its partial point, and the only exchange is an NCCL
all-gather of the N partial points followed by N-1 point additions.
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

LOGN = 20
BYTES_PER_TERM = 192  # 160 B generator + 32 B scalar, read once at ABI width
METRIC = "MSM throughput (scalar*point terms/sec), ristretto255, n=2^20 per GPU"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.dev = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_scalars(n, seed):
    """Random scalars < 2^252 (uniform bytes, top byte & 0x0f — SURVEY §8d C2)."""
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    s[:, 31] &= 0x0F
    return s


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------------
_REF_CACHE = {}


def _ref_worker(args):
    """One single-threaded reference MSM of n terms; returns the seconds spent inside the MSM call
    (input generation is outside the timed span)."""
    n, seed, use_ref = args
    if use_ref:
        from oracle import refcpu as o
    else:
        from oracle import port as o
    if n not in _REF_CACHE:
        _REF_CACHE[n] = o.ristretto_generators(n, 0)
    g = _REF_CACHE[n]
    s = make_scalars(n, seed)
    t = time.perf_counter()
    o.commit(0, [(s, 0)], g)
    return time.perf_counter() - t


def cpu_baseline_sample(n_sample, procs):
    """`procs` independent reference MSMs of n_sample terms running concurrently (the reference cpu
    backend is single-threaded per call: README.md:89-92). Returns (terms/s, kind)."""
    import multiprocessing as mp
    from oracle import refcpu
    use_ref = refcpu.available()
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_ref_worker, [(n_sample, 7 + i, use_ref) for i in range(procs)])  # warm caches
        times = pool.map(_ref_worker, [(n_sample, 100 + i, use_ref) for i in range(procs)])
    return procs * n_sample / max(times), ("reference" if use_ref else "port")


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = 1 << 14  # per process and step: keeps a 20-step run within ~2 minutes on a 200-core host
    from oracle import refcpu
    use_ref = refcpu.available()
    kind = "reference" if use_ref else "port"
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(_ref_worker, [(n_sample, i, use_ref) for i in range(cores)])
        wall = 0.0
        for k in range(args.steps):
            wall += max(pool.map(_ref_worker, [(n_sample, 1000 * k + i, use_ref)
                                               for i in range(cores)]))
    terms = args.steps * cores * n_sample
    value = terms / wall
    sample = (f"{cores} concurrent single-threaded MSMs of n=2^14 ristretto terms per step "
              f"(the workload's 2^20-term column is out of reach of a bounded CPU run; the "
              f"reference cpu backend is serial per call)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "terms/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (radix-2^51 limbs, integer)", "data": "synthetic",
        "config": {"workload": "C2: ristretto255 MSM, 252-bit scalars, n=2^20 per GPU",
                   "reference_sample_terms_per_step": cores * n_sample},
        "cpu_baseline": {"value": value, "unit": "terms/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "terms/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------
def run_cuda(args, rank, local_rank, world):
    import torch
    import blitzar_b200 as bb

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    assert bb.sxt_init(device=local_rank) == 0

    n = 1 << LOGN
    curve = 0
    pb = bb.point_bytes(curve)
    # synthetic inputs: rank r owns generators g(r*n .. (r+1)*n) and its own scalar shard
    gens_host = torch.empty((n, 160), dtype=torch.uint8).pin_memory()
    scal_host = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    gens_host.numpy()[:] = bb.get_generators(n, rank * n)
    scal_host.numpy()[:] = make_scalars(n, 12345 + rank)
    out_host = torch.empty((64,), dtype=torch.uint8).pin_memory()
    d_gens = torch.empty((n, 160), dtype=torch.uint8, device="cuda")
    d_scal = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    d_gens.copy_(gens_host)
    d_scal.copy_(scal_host)
    d_partial = torch.zeros((pb,), dtype=torch.uint8, device="cuda")
    d_all = torch.zeros((world, pb), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((64,), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    # the library's stream as a torch stream: NCCL collectives issued under it are ordered against the
    # engine's kernels on the device, without host synchronisation
    lib_stream = torch.cuda.ExternalStream(bb.stream_ptr(), device=torch.device("cuda", local_rank))

    def barrier():
        bb.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """inputs resident in HBM -> commitment (N=1) or partial + all-gather + combine (N>1)"""
        if world == 1:
            bb.commit_device(curve, [(n, 32, 0)], [d_scal.data_ptr()], d_gens.data_ptr(),
                             d_out.data_ptr(), None)
        else:
            bb.commit_device(curve, [(n, 32, 0)], [d_scal.data_ptr()], d_gens.data_ptr(), None,
                             d_partial.data_ptr())
            with torch.cuda.stream(lib_stream):
                dist.all_gather_into_tensor(d_all.view(-1), d_partial)
            bb.combine_partials_device(curve, d_out.data_ptr(), d_all.data_ptr(), world, 1)

    def step_e2e():
        """host (pinned) buffers in, host result out"""
        if world == 1:
            out = bb.compute_pedersen_commitments(curve, [(scal_host.numpy(), 0)], gens_host.numpy())
            return out
        import ctypes as C
        L = bb.lib()
        L.b200_memcpy_h2d(C.c_void_p(d_gens.data_ptr()), C.c_void_p(gens_host.data_ptr()),
                          C.c_uint64(n * 160))
        L.b200_memcpy_h2d(C.c_void_p(d_scal.data_ptr()), C.c_void_p(scal_host.data_ptr()),
                          C.c_uint64(n * 32))
        step_device()
        L.b200_memcpy_d2h(C.c_void_p(out_host.data_ptr()), C.c_void_p(d_out.data_ptr()),
                          C.c_uint64(32))
        return out_host.numpy()[:32].copy()

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    bb.profile_accumulate(True)
    bb.profile_read()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = bb.launch_count()
    e0, e1 = bb.Event(), bb.Event()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    ms = e0.elapsed_ms(e1)
    barrier()
    launches = bb.launch_count() - launches0
    acc_ms, acc_launches = bb.profile_read()
    bb.profile_accumulate(False)
    dev_result = d_out.cpu().numpy()[:32].copy()

    # ---- end-to-end timing ------------------------------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_result = step_e2e()
    bb.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None  # sampled across both timed regions

    if dist is not None:
        t = torch.tensor([ms, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1])
    assert np.array_equal(np.asarray(e2e_result).reshape(-1)[:32], dev_result), \
        "device-resident and end-to-end paths disagree"

    if rank == 0:
        total_terms = world * n * args.steps
        value = total_terms / (ms * 1e-3)
        e2e_value = total_terms / t_e2e
        peak, peak_src = read_peaks()
        acc_avg_ms = acc_ms / max(1, acc_launches)
        achieved = BYTES_PER_TERM * n / (acc_avg_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r01_accumulate_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # CPU baseline beside it: bounded sample on this box's host cores (N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            n_sample = 1 << 17
            v, kind = cpu_baseline_sample(n_sample, cores)
            cpu = {"value": v, "unit": "terms/s", "cores": cores, "kind": kind,
                   "sample": f"one reference cpu-backend MSM of n=2^17 ristretto terms (same "
                             f"generator / scalar distribution as the workload), 1 thread — the "
                             f"reference cpu backend is serial"}
        line = {
            "metric": METRIC, "value": value, "unit": "terms/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (integer modular arithmetic)", "data": "synthetic",
            "config": {"workload": "C2: ristretto255 MSM, random 252-bit scalars, n=2^20 per GPU "
                                   "(generator-range shard per rank)",
                       "global_terms": world * n, "curve": "curve25519/ristretto255",
                       "cache": "inputs (192 MiB per step) exceed the 126 MB L2",
                       "parallelism": f"generator-range x{world}"},
            "e2e": {"value": e2e_value, "unit": "terms/s", "h2d_bytes_per_step": n * 192,
                    "d2h_bytes_per_step": 32, "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k_run<AccumulateBody<Ed25519,true>> (level-1 bucket accumulation)",
                         "kernel_ms": acc_avg_ms, "kernel_share_of_step": acc_avg_ms / (ms / args.steps),
                         "algorithmic_bytes_per_launch": BYTES_PER_TERM * n,
                         "note": "integer-ALU bound (about 16 point additions of 8-9 field "
                                 "multiplications per term); see DESIGN.md",
                         "secondary": imad_roofline(n, acc_avg_ms, clocks)},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def imad_roofline(n, kernel_ms, clocks):
    """The bound that actually limits the accumulation kernel: the 32x32->64 multiplier. One term
    has 16 signed 16-bit digits (252-bit scalars), i.e. 16 bucket entries; every entry except the
    first of a bucket run is one cached-form addition = 8 field multiplications = 8 x (64 + 8)
    IMAD.WIDE.U32 (schoolbook 8x8 limbs + the 2^256 = 38 fold). Peak = 29.2 lane-ops/clk/SM measured
    on B200 for the multiply-accumulate-with-carry form this kernel issues (tests/micro/pipes.cu;
    plain IMAD.WIDE 23.0, IMAD.HI 24.6) x 148 SMs x the SM clock sampled during the run."""
    windows, nbuckets = 16, 1 << 15
    entries = windows * n * (1.0 - 2.0 ** -16)  # zero digits are skipped
    runs = windows * nbuckets * (1.0 - (1.0 - 1.0 / nbuckets) ** (entries / windows))
    imads = (entries - runs) * 8 * 72
    mhz = clocks.get("sm_mhz") or 1965.0
    peak = 29.2 * 148 * mhz * 1e6
    achieved = imads / (kernel_ms * 1e-3)
    return {"bound": "imad_wide", "achieved": achieved, "peak": peak, "unit": "IMAD.WIDE lane-ops/s",
            "frac": achieved / peak, "imad_wide_per_launch": imads}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
