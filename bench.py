#!/usr/bin/env python
"""Benchmark of the GANMF hot path on B200 (driver contract: see the round brief).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference's arithmetic on the host CPU

Workload (BASELINE.json configs[3]): synthetic 138 000 x 27 000 implicit matrix at 0.5 % density,
GANMF --user, k=250, emb_dim=1024, batch 1024 PER GPU (weak scaling: every GPU owns 138 000 users).
A "step" = one D update + one G update on one 1024-row minibatch (each row gets one D pass and one G
pass, which is how BASELINE.md turns epochs into user-rows/s).
value  = rows/s with the CSR, ids and weights already resident in HBM.
e2e    = the same through ganmf_train_epoch() (the call GANMF.fit makes): host ids in, losses out.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG4 = dict(name="cfg4: synthetic 138000x27000 d=0.5% GANMF-u k=250 E=1024 B=1024/GPU", users=138000, items=27000,
            density=0.005, k=250, E=1024, B=1024)
HP = dict(d_lr=1e-4, g_lr=1e-4, d_reg=1e-4, g_reg=0.0, m=10.0, alpha=0.01)


def workload(name, world=1):
    """cfg4 (default, the workload of record: 138 000 users PER GPU) or --workload cfg5 (BASELINE.json configs[4]:
    2 000 000 x 200 000 at 0.1 %, the users split over the ranks)."""
    if name == "cfg5":
        per = 2000000 // max(world, 1)
        return dict(name="cfg5: synthetic 2000000x200000 d=0.1%% GANMF-u k=250 E=1024 B=1024/GPU, %d users per GPU" % per,
                    users=per, items=200000, density=0.001, k=250, E=1024, B=1024)
    return CFG4


def synthetic_urm(n_users, n_items, density, seed):
    """Implicit interaction matrix: round(density*n_items) distinct uniform items per user."""
    rs = np.random.RandomState(seed)
    per = max(1, int(round(density * n_items)))
    cols = rs.randint(0, n_items, size=(n_users, per)).astype(np.int32)
    cols.sort(axis=1)
    keep = np.ones_like(cols, dtype=bool)
    keep[:, 1:] = cols[:, 1:] != cols[:, :-1]
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(keep.sum(1), out=indptr[1:])
    m = sps.csr_matrix((np.ones(int(indptr[-1]), np.float32), cols[keep], indptr.astype(np.int32)),
                       shape=(n_users, n_items))
    m.has_sorted_indices = True
    return m


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (NVML every ~2 ms; nvidia-smi as a fallback)."""
    SMI_Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.sm, self.bits, self.max_mhz, self.stop, self.src = index, [], 0, None, False, "nvml"
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(reasons(h))
                time.sleep(0.002)
            self.names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                          "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                          "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        except Exception:
            self.src = "nvidia-smi"
            self.smi_reasons = set()
            while not self.stop:
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.SMI_Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout
                    r = [x.strip() for x in out.strip().split(",")]
                    self.sm.append(float(r[0]))
                    self.max_mhz = float(r[1])
                    for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4),
                                      ("sw_power_cap", 5)):
                        if r[col].lower().startswith("active"):
                            self.smi_reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        time.sleep(0.01)
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if self.src == "nvml":
            reasons = [n for n, bit in getattr(self, "names", {}).items() if self.bits & bit]
        else:
            reasons = sorted(getattr(self, "smi_reasons", ()))
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": self.src}


def gemm_algorithmic_bytes(c, fused_adam=False):
    """Operand + output bytes (each read/written once) of the 13 tensor-core GEMMs of one D+G step pair.
    fused_adam (single GPU): the dWd / dWe epilogues do not write the gradient but read and write theta, m, v
    in place (24 B per parameter instead of 4)."""
    B, I, k, E = c["B"], c["items"], c["k"], c["E"]
    f = 4.0
    g = lambda M, N, K, extra=0: f * (M * K + N * K + M * N * (1 + extra))
    wg = 5 if fused_adam else 0                   # 6 tensors moved instead of 1
    d = g(B, I, k) + g(2 * B, E, I) + g(2 * B, I, E, 1) + g(E, I, 2 * B, wg) + g(2 * B, E, I) + g(I, E, 2 * B, wg)
    gs = g(B, I, k) + g(2 * B, E, I) + g(B, I, E, 1) + g(B, E, I, 2) + g(B, I, E, 1) + g(I, k, B) + g(B, k, I)
    return d + gs


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms must use all host cores."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def cpu_eval_baseline(c, n_users=768):
    """The reference's evaluation path on the host: NumPy P[u].V^T scorer + the oracle port of
    BaseRecommender.recommend / EvaluatorHoldout (per-user Python loop, as in the reference), top-10."""
    from oracle import eval_oracle as eo
    rs = np.random.RandomState(7)
    train = synthetic_urm(n_users, c["items"], c["density"], 99)
    test = synthetic_urm(n_users, c["items"], c["density"] / 4, 100)
    test = sps.csr_matrix(test - test.multiply(train))
    test.eliminate_zeros()
    lim = np.sqrt(6.0 / (c["items"] + c["k"]))
    P = rs.uniform(-lim, lim, (n_users, c["k"])).astype(np.float32)
    V = rs.uniform(-lim, lim, (c["items"], c["k"])).astype(np.float32)
    t0 = time.perf_counter()
    _, n_eval = eo.evaluate(lambda u: P[u] @ V.T, train, test, [10], promotion="legacy")
    dt = time.perf_counter() - t0
    return {"value": n_eval / dt, "unit": "users/s", "sample": "%d users x %d items, cutoff 10 (NumPy scorer + oracle "
            "port of recommend()/EvaluatorHoldout: single-threaded per-user loop as in the reference)" % (n_eval, c["items"])}


def flops_per_row(c):
    return c["items"] * (8 * c["k"] + 30 * c["E"])


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU path for this metric: its TF-1.12 graph restated in NumPy (TensorFlow cannot
    be installed here), all host threads, same shapes and hyper-parameters."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import train_oracle as to
    cores = use_all_host_threads()
    c = workload(args.workload)
    B = 256                                                   # bounded sample: 256-row minibatches
    n_users = 4096                                            # user-factor rows kept small: P is not the cost driver
    urm = synthetic_urm(n_users, c["items"], c["density"], 1337)
    p0 = to.init_ganmf_params(n_users, c["items"], c["k"], c["E"], seed=1234)
    orc = to.GanmfOracle(p0, HP["d_lr"], HP["g_lr"], dtype=np.float32)
    rs = np.random.RandomState(1337)

    def step():
        ids = rs.permutation(n_users)[:B]
        R = to.csr_rows_to_dense(urm, ids)
        orc.d_step(ids, R, d_reg=HP["d_reg"], m=HP["m"])
        orc.g_step(ids, R, g_reg=HP["g_reg"], recon_coefficient=HP["alpha"])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    sample = "NumPy restatement of the TF graph, %d steps of B=%d rows at %d items, k=250, E=1024 " \
             "(user-factor table cut to %d rows)" % (args.steps, B, c["items"], n_users)
    line = {"impl": "reference", "metric": "GANMF-u train user-rows/s", "value": v, "unit": "rows/s", "n_gpus": 0,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["name"]},
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "eval": dict(cpu_eval_baseline(c), metric="top-10 eval users/s (score -> seen mask -> top-10 -> metric sums)")}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-users", type=int, default=8192)
    ap.add_argument("--quick", action="store_true", help="A/B runs: training throughput + GEMM roofline only")
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"],
                    help="cfg4 = workload of record (default); cfg5 = 2M x 200k split over the ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ganmf_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        from ganmf_b200.parallel import init_nccl
        init_nccl(local_rank)
    from ganmf_b200 import _lib as L
    from ganmf_b200 import build as _build
    _build.build()
    from ganmf_b200.engine import Engine
    from ganmf_b200.parallel import DataParallelTrainer

    c = workload(args.workload, world)
    K, W, B = args.steps, args.warmup, c["B"]
    urm = synthetic_urm(c["users"], c["items"], c["density"], 1337 + rank)      # this rank's user shard
    eng = Engine(L.KIND_GANMF, c["users"], c["items"], c["k"], emb_dim=c["E"], max_batch=B, device=local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_csr(L.CSR_SEEN, urm, with_data=False)
    eng.init_params(1234)                                     # same seed on every rank: replicated D and V
    trainer = DataParallelTrainer(eng, world) if world > 1 else None
    rs = np.random.RandomState(1337 + rank)
    # One whole epoch first (135 D + 135 G updates, untimed): every user row then carries Adam moments and
    # the timed rows owe the deferred zero-gradient optimiser steps a steady-state epoch gives them (the
    # user-factor optimiser is lazy, kernels.cuh K6b -- on untouched rows it would have nothing to replay).
    warm = rs.permutation(c["users"]).astype(np.int32)
    if trainer:
        trainer.train_epoch(warm, B, 1, 1, HP)
    else:
        eng.train_epoch(warm, B, 1, 1, HP["d_lr"], HP["g_lr"], HP["d_reg"], HP["g_reg"], HP["m"], HP["alpha"])
    n_ids = (W + K) * B
    perm = rs.permutation(c["users"])[:n_ids].astype(np.int32)
    eng.upload_ids(perm)

    def run_steps(first, count, slot0=0):
        """reference schedule on `count` batches: all D updates, then all G updates, same batches"""
        for i in range(count):
            off = (first + i) * B
            if trainer:
                trainer.d_step(off, B, HP["d_lr"], HP["d_reg"], HP["m"], slot0 + i)
            else:
                eng.d_step(off, B, HP["d_lr"], HP["d_reg"], HP["m"], loss_slot=slot0 + i)
        for i in range(count):
            off = (first + i) * B
            if trainer:
                trainer.g_step(off, B, HP["g_lr"], HP["g_reg"], HP["alpha"], slot0 + count + i)
            else:
                eng.g_step(off, B, HP["g_lr"], HP["g_reg"], HP["alpha"], loss_slot=slot0 + count + i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput --------------------------------------------------------
    run_steps(0, W)
    barrier()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = os.environ.get("GANMF_BENCH_PROFILER_RANGE") == "1"     # ncu --profile-from-start off: timed steps only
    with ClockSampler(local_rank) as clk:
        barrier()
        if prof_range:
            torch.cuda.profiler.start()
        e0.record()
        run_steps(W, K)
        e1.record()
        barrier()
        if prof_range:
            torch.cuda.profiler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - launches0
    value = world * B * K / (ms * 1e-3)
    losses = eng.read_losses(2 * K)
    if not np.all(np.isfinite(losses)):
        raise SystemExit("non-finite losses in the timed region: %r" % losses[:8])

    # ---- live roofline of the dominant kernel (tcgen05 GEMM), CUDA events around every launch --
    eng.profile(True)
    run_steps(W, K)
    rec_ms, rec_shape = eng.profile_records()
    if rank == 0 and os.environ.get("GANMF_BENCH_GEMM_TABLE"):
        agg = {}
        for t, (M_, N_, K_, S_) in zip(rec_ms, rec_shape):
            a = agg.setdefault((int(M_), int(N_), int(K_), int(S_)), [0, 0.0])
            a[0] += 1
            a[1] += t
        with open(os.environ["GANMF_BENCH_GEMM_TABLE"], "w") as f:
            f.write("# tcgen05 GEMM launches inside the timed steps (CUDA events around each launch)\n")
            f.write("#     M      N      K splits  calls   ms/call   TFLOP/s\n")
            for (M_, N_, K_, S_), (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%7d %6d %6d %6d %6d %9.4f %9.1f\n" % (M_, N_, K_, S_, cnt, tot / cnt,
                                                            2.0 * M_ * N_ * K_ / (tot / cnt * 1e-3) / 1e12))
    # the two weight-gradient GEMMs of the D step (dWd, dWe) carry the fused Adam update in their epilogue on
    # the single-GPU path; report the GEMM rate with and without them
    pure = [(t, sh) for t, sh in zip(rec_ms, rec_shape)
            if not (world == 1 and int(sh[2]) == 2 * B and int(sh[0]) * int(sh[1]) == c["items"] * c["E"])]
    pure_ms = float(sum(t for t, _ in pure))
    pure_flops = float(sum(2.0 * int(sh[0]) * int(sh[1]) * int(sh[2]) for _, sh in pure))
    gemm_ms, gemm_flops, gemm_launches = eng.profile_read()
    eng.profile(False)
    pk = peaks()
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_tc_gemm_dram_traffic.json")
    if os.path.exists(tpath):                                 # from the committed `ncu --set full` capture
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_launch_mean"], "profiles/r01_tc_gemm_dram_traffic.json: " + tj["note"]
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {"kernel": "tc_gemm_kernel (tcgen05 kind::tf32, TMA, TMEM; CTA pairs cta_group::2 on the many-tile GEMMs)", "bound": "tensor", "achieved": achieved,
                "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch_mean": gemm_algorithmic_bytes(c, fused_adam=(world == 1)) / 13.0,
                "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json); kind::tf32 issues at half the "
                "bf16 rate, so frac <= 0.5 by construction", "frac_of_tf32_rate": 2 * achieved / pk["tf_sust"],
                "achieved_excl_fused_adam_gemms": pure_flops / (pure_ms * 1e-3) / 1e12 if pure_ms > 0 else None,
                "note": "on 1 GPU the dWd/dWe GEMM epilogues also run TF-Adam on Wd/We in place (24 B/param of HBM "
                        "traffic inside those 2 of the 13 launches), which lowers their FLOP rate but removes the "
                        "separate optimiser pass",
                "gemm_launches_per_step": gemm_launches / K, "gemm_share_of_step": gemm_ms / (ms if world == 1 else
                                                                                             max(ms, 1e-9)),
                "step_algorithmic_tflops": flops_per_row(c) * B * K / (ms * 1e-3) / 1e12}

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": "GANMF-u train user-rows/s", "value": value, "unit": "rows/s", "n_gpus": world,
                              "steps": K, "warmup": W, "ms_per_step": ms / K, "quick": True, "clocks": clk.summary(),
                              "gemm_tflops": achieved, "gemm_share_of_step": roofline["gemm_share_of_step"],
                              "env": {k_: v_ for k_, v_ in os.environ.items() if k_.startswith("GANMF_")}}))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- end to end through the public call (host ids in, losses out) -------------------------
    e2e = None
    if world == 1:
        hp = (B, 1, 1, HP["d_lr"], HP["g_lr"], HP["d_reg"], HP["g_reg"], HP["m"], HP["alpha"])
        pin = torch.from_numpy(rs.permutation(c["users"])[:K * B].astype(np.int32)).pin_memory()
        eng.train_epoch(pin.numpy()[:W * B], *hp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dl, gl = eng.train_epoch(pin.numpy(), *hp)             # H2D ids + all steps + D2H losses + sync
        dt = time.perf_counter() - t0
        e2e = {"value": B * K / dt, "unit": "rows/s", "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 8,
               "api": "ganmf_train_epoch (GANMF.fit's per-epoch call): pinned host row ids in, per-step losses out"}
    else:
        e2e = trainer.e2e_epoch(rs, c["users"], K, B, HP)
        e2e["value"] = world * B * K / (max_over_ranks(e2e.pop("ms")) * 1e-3)

    # ---- evaluator: score -> seen mask -> top-10 -> metric sums, users sharded by rank ---------
    n_eval = min(args.eval_users, c["users"])
    test = synthetic_urm(c["users"], c["items"], c["density"] / 4, 4242 + rank)
    test = sps.csr_matrix(test - test.multiply(urm))
    test.eliminate_zeros()
    test.sort_indices()
    eng.set_test(test, urm)
    users = np.flatnonzero(np.diff(test.indptr) > 0)[:n_eval].astype(np.int32)
    eng.evaluate(users, [10], remove_seen=True, want_counts=False)       # warm-up: sizes the device buffers
    ev_s = 1e30
    for _ in range(3):                                                    # steady state, best of 3 whole calls
        barrier()
        t0 = time.perf_counter()
        sums, _ = eng.evaluate(users, [10], remove_seen=True, want_counts=False)
        torch.cuda.synchronize()
        ev_s = min(ev_s, time.perf_counter() - t0)
    ev_ms = max_over_ranks(ev_s * 1e3)
    eval_users_s = world * len(users) / (ev_ms * 1e-3)
    eval_info = {"metric": "top-10 eval users/s (score -> seen mask -> top-10 -> metric sums)", "value": eval_users_s,
                 "unit": "users/s", "users": int(len(users)) * world, "hbm_frac_4I_bytes_per_user":
                 eval_users_s / world * 4 * c["items"] / 1e9 / pk["hbm"], "e2e": True,
                 "precision_at_10": float(sums[0, 0] / max(len(users), 1))}

    # ---- HBM-bound kernels timed alone (CUDA events, inputs larger than L2): achieved GB/s vs measured copy peak
    hbm_kernels = None
    if rank == 0:
        hbm_kernels = hbm_kernel_rooflines(eng, torch, L, c, pk)

    line = {"metric": "GANMF-u train user-rows/s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (fp32 storage, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": c["name"], "state": "after one full untimed epoch (all 138000 rows hold Adam moments)",
                       "l2": "per-step working set (weights 221 MB + activations) exceeds the "
                       "126 MB L2", "parallelism": "dp%d over users; D and item factors replicated, NCCL reduce-scatter / sharded Adam / all-gather" %
                       world if world > 1 else "single GPU"},
            "gpu_launches": int(launches), "clocks": clk.summary(), "e2e": e2e, "roofline": roofline,
            "eval": eval_info, "hbm_kernels": hbm_kernels,
            "loss_last": [float(losses[K - 1]), float(losses[2 * K - 1])]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(c)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def hbm_kernel_rooflines(eng, torch, L, c, pk):
    """top-k (4*I bytes/user), fused Adam (28 B/param) and CSR gather (4*B*ld written) on their own."""
    out = {}

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    n, I = 8192, c["items"]
    ld = (I + 31) // 32 * 32
    sc = torch.randn((n, ld), device="cuda", dtype=torch.float32)                      # 885 MB >> L2
    idx = torch.empty((n, 10), device="cuda", dtype=torch.int32)
    val = torch.empty((n, 10), device="cuda", dtype=torch.float32)
    t = timed(lambda: L.check(eng.lib.ganmf_k_topk(eng.ctx, sc.data_ptr(), ld, n, I, 10, idx.data_ptr(),
                                                   val.data_ptr())))
    gbs = n * I * 4 / t / 1e9
    out["topk_rows_kernel(K=10)"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                                     "frac": gbs / pk["hbm"], "users_per_s": n / t, "bytes_per_user": 4 * I}
    del sc
    npar = 2 * I * c["E"]                                                               # the two D kernels
    th, m, v, g = (torch.zeros(npar, device="cuda", dtype=torch.float32) for _ in range(4))
    t = timed(lambda: L.check(eng.lib.ganmf_k_adam(eng.ctx, th.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(),
                                                   npar, 1e-4, 1e-4)))
    gbs = 28.0 * npar / t / 1e9
    out["fused_adam_kernel"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                                "frac": gbs / pk["hbm"], "bytes_per_param": 28}
    del th, m, v, g
    B = c["B"]
    dst = torch.empty((B, ld), device="cuda", dtype=torch.float32)
    t = timed(lambda: L.check(eng.lib.ganmf_k_csr_gather_dense(eng.ctx, 0, B, dst.data_ptr(), ld)), reps=20)
    gbs = 4.0 * B * ld / t / 1e9
    out["csr_gather_dense_kernel"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                                      "frac": gbs / pk["hbm"], "note": "110 MB output stays in the 126 MB L2"}
    return out


def cpu_baseline(c):
    """Oracle port of the reference graph on the host cores, bounded sample (about 10-30 s)."""
    from oracle import train_oracle as to
    cores = use_all_host_threads()
    B, n_users = 256, 4096
    urm = synthetic_urm(n_users, c["items"], c["density"], 1337)
    orc = to.GanmfOracle(to.init_ganmf_params(n_users, c["items"], c["k"], c["E"], seed=1234), HP["d_lr"], HP["g_lr"],
                         dtype=np.float32)
    rs = np.random.RandomState(0)

    def step():
        ids = rs.permutation(n_users)[:B]
        R = to.csr_rows_to_dense(urm, ids)
        orc.d_step(ids, R, d_reg=HP["d_reg"], m=HP["m"])
        orc.g_step(ids, R, g_reg=HP["g_reg"], recon_coefficient=HP["alpha"])
    step()
    t0 = time.perf_counter()
    n = 0
    while n < 3 or (time.perf_counter() - t0 < 12 and n < 40):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": B * n / dt, "unit": "rows/s", "cores": cores, "kind": "port",
            "sample": "%d D+G steps of B=%d rows, %d items, k=250, E=1024 (NumPy/BLAS, all host threads; "
                      "user-factor table cut to %d rows)" % (n, B, c["items"], n_users),
            "eval": cpu_eval_baseline(c)}


if __name__ == "__main__":
    sys.exit(main())
