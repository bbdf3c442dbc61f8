#!/usr/bin/env python
"""bench.py - training examples/sec of the relation-autoencoder hot path on B200 (BASELINE.json's metric).

One "step" = one fused forward/backward/update pass of the hot path over one batch of synthetic input
(SURVEY 8d).  Default workload: BASELINE.json configs[1] (synthetic NYT-scale AC: K=100, d=30, ~30 features/example,
1M feature vocab, 500k entities, batch 4096, 5 negatives); ``--workload T|cfg3|cfg4|cfg5|cfg1`` select the others.

  python bench.py --gpus N --steps K --warmup W            our arm  (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's CPU path (NumPy float64 port of
                                                           the Theano graph with the reference's DENSE AdaGrad sweep;
                                                           Theano itself is not installable) on the box's host cores

Prints ONE JSON line (rank 0).  `value` = whole-job examples/sec with inputs resident in HBM; `e2e` = the same through
Engine.train(batch_index, neg1, neg2) with HOST negatives copied in and the cost read back every step (the reference's
func['train'] call, OieInduction.py:189).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from relation_autoencoder_b200 import synthetic as SY  # noqa: E402

METRIC = "AC-model train examples/sec"
UNIT = "examples/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return dict(hbm_gbs=float(pk["hbm_gbs"]), bf16_burst=float(pk["bf16_tflops"]),
                    bf16_sustained=float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")   # B200_PROFILING.md:17-20


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md:70-72): NVML polled from a thread
    every ~2 ms (nvidia-smi -lms as a fallback, which is too slow for short regions)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, device_index=0, period_s=0.002):
        self.dev = device_index
        self.period = period_s
        self.sm, self.mask = [], 0
        self.max_sm = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        self._h = None
        self._smi = None

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        self._nvml = pynvml
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.dev).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.dev)

    def start(self):
        try:
            self._h = self._handle()
            self.max_sm = float(self._nvml.nvmlDeviceGetMaxClockInfo(self._h, self._nvml.NVML_CLOCK_SM))
            self._t = threading.Thread(target=self._poll, daemon=True)
            self._t.start()
        except Exception:
            self._h = None
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                self._smi = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                              "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self._t = threading.Thread(target=self._read_smi, daemon=True)
                self._t.start()
            except Exception:
                self._smi = None

    def _poll(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            except Exception:
                pass
            time.sleep(self.period)

    def _read_smi(self):
        for line in self._smi.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.sm.append(float(f[0]))
                self.max_sm = float(f[1])
                self.mask |= int(f[2], 16)
            except Exception:
                continue

    def stop(self):
        self._stop.set()
        if self._smi is not None:
            time.sleep(0.05)
            self._smi.terminate()
        if self._t is not None:
            self._t.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "samples": 0, "reasons": ["no clock samples"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": self.max_sm,
                "samples": len(self.sm), "reasons": sorted(name for name, bit in self.REASONS if self.mask & bit)}


def _needed_examples(wl, steps, warmup, extra):
    return (steps + warmup + extra) * wl["B"]


def _make_inputs(wl, n_examples, rank=0, uniform=False):
    data = SY.make_dataset(n_examples, wl["F"], wl["N"], wl["fbar"], seed=1234 + rank, uniform=uniform)
    rng = np.random.RandomState(2)
    params = SY.init_params(rng, wl["model"], wl["F"], wl["K"], wl["N"], wl["d"])
    neg1, neg2 = SY.draw_negatives(rng, data.neg_cum, data.n, wl["S"])
    return data, params, neg1, neg2


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's CPU path (the ONLY place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(wl, data, params32, neg1, neg2, steps, warmup, budget_s, sparse_rows=False):
    from oracle import rae_oracle as O
    p = {k: v.astype(np.float64) for k, v in params32.items()}
    om = O.OracleModel(wl["model"], p, K=wl["K"], d=wl["d"], S=wl["S"], B=wl["B"], lr=0.1, l1=0.0, l2=wl.get("l2", 0.0),
                       alpha=wl.get("alpha", 1.0), sparse_rows=sparse_rows)
    om.bind_split("train", data.indptr, data.indices, data.args1, data.args2)
    B = wl["B"]
    nb = data.n // B
    times = []
    t_begin = time.perf_counter()
    done = 0
    for s in range(warmup + steps):
        b = s % nb
        t0 = time.perf_counter()
        om.train(b, neg1[:, b * B:(b + 1) * B], neg2[:, b * B:(b + 1) * B])
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
            done += 1
        if time.perf_counter() - t_begin > budget_s and done >= 1:
            break
    sec = float(np.mean(times))
    return B / sec, sec, done


def _host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    n_ex = min(wl["N_train"], max(4, min(steps + warmup, 8)) * wl["B"])
    data, params, neg1, neg2 = _make_inputs(wl, n_ex)
    eps, sec, done = cpu_reference_steps(wl, data, params, neg1, neg2, steps, min(warmup, 1), budget_s=args.cpu_budget)
    cores = _host_threads()
    sample = ("%d of %d requested steps of batch %d on the same synthetic workload; NumPy float64 port of the Theano graph "
              "with the reference's dense AdaGrad sweep over every parameter (Optimizers.py:29-32); Theano is not "
              "installable offline" % (done, steps, wl["B"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": eps, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(wl, args, world),
        "cpu_baseline": {"value": eps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": eps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _config(wl, args, world):
    return {"workload": "%s: %s" % (args.workload, wl["desc"]), "decoder": wl["model"], "K": wl["K"], "d": wl["d"],
            "S": wl["S"], "batch_per_gpu": wl["B"], "global_batch": wl["B"] * world, "fbar": wl["fbar"], "F": wl["F"],
            "N": wl["N"], "N_train_nominal": wl["N_train"], "optimizer": "adagrad", "lr": 0.1,
            "l2": wl.get("l2", 0.0), "alpha": wl.get("alpha", 1.0), "feature_distribution": "zipf(1.0)" if not args.uniform else "uniform",
            "parallelism": "dp%d" % world,
            "l2_cache_policy": "inputs larger than L2: W+acc %.0f MB, A+acc %.0f MB, a different batch every step"
                               % (2 * 4e-6 * wl["F"] * wl["K"], 2 * 4e-6 * wl["N"] * wl["d"])}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args, wl, rank, world, local_rank):
    # libraries (NCCL's version banner, ...) must not write to stdout: only the JSON line does
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args, wl, rank, world, local_rank)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line))
        sys.stdout.flush()


def _run_ours(args, wl, rank, world, local_rank):
    import torch

    from relation_autoencoder_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    steps, warmup = args.steps, args.warmup
    prof_steps = 5
    B = wl["B"]
    n_ex = min(wl["N_train"], _needed_examples(wl, 2 * steps, 2 * warmup, prof_steps))
    n_ex = (n_ex // B) * B
    data, params, neg1, neg2 = _make_inputs(wl, n_ex, rank=rank, uniform=args.uniform)
    nb = data.n // B

    if world > 1:
        from relation_autoencoder_b200.dist import DistributedEngine
        eng = DistributedEngine(wl["model"], wl["K"], wl["d"], wl["S"], B, wl["F"], wl["N"], wl["N_train"], lr=0.1,
                                l2=wl.get("l2", 0.0), alpha=wl.get("alpha", 1.0), device=local_rank, rank=rank, world=world)
    else:
        eng = Engine(wl["model"], wl["K"], wl["d"], wl["S"], B, wl["F"], wl["N"], wl["N_train"], lr=0.1,
                     l2=wl.get("l2", 0.0), alpha=wl.get("alpha", 1.0), device=local_rank)
    eng.set_params_numpy(params)
    del params
    eng.bind_split("train", data.indptr, data.indices, data.args1, data.args2)
    eng.bind_epoch_negatives(neg1, neg2)
    pinned1 = torch.from_numpy(neg1).pin_memory()
    pinned2 = torch.from_numpy(neg2).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, first_batch):
        for s in range(warmup):
            fn((first_batch + s) % nb)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for s in range(steps):
            fn((first_batch + warmup + s) % nb)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    # ---- (1) device-resident throughput ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(lambda b: eng.train_device(b, want_cost=False), 0)
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()
    # ---- (2) end to end through the reference-facing call: host negatives in, cost out, every step ----
    n1np, n2np = pinned1.numpy(), pinned2.numpy()

    def e2e_step(b):
        eng.train(b, n1np[:, b * B:(b + 1) * B], n2np[:, b * B:(b + 1) * B])
    ms_e2e, wall_e2e = timed(e2e_step, (warmup + steps) % nb)
    ms_e2e = max(ms_e2e, wall_e2e * 1e3)       # host-side staging counts
    # ---- (3) per-phase device times for the roofline of the dominant kernel ----
    phase_ms = None
    dist_phase_ms = None
    if world > 1:
        eng.set_profiling(True)
        for s in range(20):
            eng.train_device((2 * (warmup + steps) + s) % nb, want_cost=False)
        dist_phase_ms = eng.phase_times_ms()
        eng.set_profiling(False)
    if world == 1:
        eng.set_profiling(True)
        acc = {}
        for s in range(prof_steps):
            eng.train_device((2 * (warmup + steps) + s) % nb, want_cost=False)
            for k, v in eng.phase_times_ms().items():
                acc[k] = acc.get(k, 0.0) + v / prof_steps
        eng.set_profiling(False)
        phase_ms = acc
        st = eng.stats()

    p_cpu = eng.get_params_numpy() if (world == 1 and not args.no_cpu_baseline) else None
    eng.close()                      # collective for N > 1 (nobody unmaps peer memory while a peer may still read it)
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return None
    pk = _peaks()
    ms_step = ms_dev / steps
    value = world * B * steps / (ms_dev * 1e-3)
    e2e_val = world * B * steps / (ms_e2e * 1e-3)
    alg_bytes = st["algorithmic_bytes"]
    nnz = st["nnz"]
    flops = SY.algorithmic_flops(wl["model"], wl["K"], wl["d"], wl["S"], B, nnz)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _config(wl, args, world), "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 2 * wl["S"] * B * 4, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / steps,
                "api": "Engine.train(batch_index, neg1, neg2) == func['train'] (OieInduction.py:189): dataset bound on the device once, "
                       "host negatives copied per step, cost read back per step"},
        "gpu_launches": int(st["kernel_launches"]) * steps,
        "step_stats": {"nnz": int(nnz), "unique_w_rows": int(st["unique_w_rows"]), "unique_e_rows": int(st["unique_e_rows"]),
                       "entity_occ": int(st["entity_occ"]), "kernel_launches_per_step": int(st["kernel_launches"]),
                       "tensor_path": int(st["tensor_path"]), "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops},
        "step_roofline": {"bound": "hbm", "achieved": alg_bytes / (ms_step * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": alg_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"], "of": pk["source"],
                          "note": "whole step: SURVEY 8(d) algorithmic bytes / device time per step"},
    }
    if dist_phase_ms:
        line["dist_phase_ms"] = {k: round(v, 5) for k, v in dist_phase_ms.items()}
    if phase_ms is not None:
        line["phase_ms"] = {k: round(v, 5) for k, v in phase_ms.items()}
        line["roofline"] = _dominant_roofline(wl, st, phase_ms, pk, args.workload)
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = min(data.n // B, 3) * B
        sub = SY.SyntheticData(data.indptr[: n_cpu + 1], data.indices[: data.indptr[n_cpu]], data.args1[:n_cpu],
                               data.args2[:n_cpu], data.neg_cum, n_cpu)
        eps, sec, done = cpu_reference_steps(wl, sub, p_cpu, neg1[:, :n_cpu], neg2[:, :n_cpu], steps=6, warmup=1,
                                             budget_s=args.cpu_budget)
        eps_sp, sec_sp, done_sp = cpu_reference_steps(wl, sub, p_cpu, neg1[:, :n_cpu], neg2[:, :n_cpu], steps=6, warmup=1,
                                                      budget_s=args.cpu_budget / 2, sparse_rows=True)
        line["cpu_baseline"] = {
            "value": eps, "unit": UNIT, "cores": _host_threads(), "kind": "port",
            "sample": "%d steps of batch %d of the same workload, NumPy float64 port of the reference graph with its dense "
                      "AdaGrad sweep (Theano not installable offline); %.2f s/step" % (done, B, sec),
            "sparse_row_variant": {"value": eps_sp, "sec_per_step": sec_sp,
                                   "note": "same math, touched rows only (what a tuned CPU port would do)"}}
    return line


PHASE_KERNELS = {      # phase -> the kernel it times (rae.h: one phase = one kernel of the step plus small helpers)
    "encoder_forward": "k_encoder_forward_v4", "entity_sort": "cub::DeviceRadixSort (entity occurrences)",
    "operand_prep": "k_tc_prep_c + k_tc_prep_qt", "contract_forward": "k_tc_bilinear", "score": "k_score",
    "entity_update": "k_rows_chunk<1> + k_entity_long2", "w_update": "k_rows_chunk<0> + k_w_long2",
    "contract_recompute": "k_tc_bilinear", "contract_dq": "k_tc_dq", "backward_finish": "k_tc_bwd_finish",
    "contract_dc": "k_tc_dc", "dense_finalize": "k_dense_finalize", "cost": "k_cost", "dense_apply": "k_dense_apply",
}


def _traffic(workload, phase):
    """DRAM bytes per launch of the phase's kernel from the committed `ncu --set full` capture (profiles/r01_traffic.json,
    made by profiles/make_traffic.py); None when that kernel was not captured for this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return t.get(workload, {}).get(phase, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def _dominant_roofline(wl, st, phase_ms, pk, workload):
    """Roofline of the kernel with the largest share of the step: SURVEY 8(d) algorithmic bytes / flops of one launch
    divided by its duration (CUDA events around the kernel on its stream, averaged over the profiled steps)."""
    K, d, S, B = wl["K"], wl["d"], wl["S"], wl["B"]
    nnz, UW, UE = st["nnz"], st["unique_w_rows"], st["unique_e_rows"]
    hasM = wl["model"] in ("rescal", "rescal+sp")
    hasSP = wl["model"] in ("sp", "rescal+sp")
    units = (d if hasM else 0) + (2 if hasSP else 0)
    bytes_of = {
        "encoder_forward": 4.0 * nnz * K + 4.0 * nnz + 8.0 * B * K,
        "w_update": 16.0 * UW * K + 8.0 * nnz + 4.0 * nnz * K,
        "entity_update": 16.0 * UE * (d + 1) + 8.0 * (2 + 2 * S) * B + 4.0 * (2 + 2 * S) * B * d,
        "score": 4.0 * (2 * S) * B * (d + 1) + 4.0 * 2 * S * B,
        "dense_apply": 16.0 * (units * d * K + K),
        "dense_finalize": 8.0 * (units * d * K + K),
    }
    gemm = 2.0 * B * units * d * K
    flops_of = {"contract_forward": gemm, "contract_recompute": gemm, "contract_dq": gemm, "contract_dc": gemm}
    cand = {k: v for k, v in phase_ms.items() if k in bytes_of or k in flops_of}
    name = max(cand, key=lambda k: cand[k])
    t = phase_ms[name] * 1e-3
    total = sum(phase_ms.values())
    common = {"kernel": PHASE_KERNELS.get(name, name), "phase": name, "ms_per_launch": phase_ms[name],
              "share_of_step": phase_ms[name] / total, "of": pk["source"], "traffic": _traffic(workload, name)}
    if name in flops_of:
        ach = flops_of[name] / t / 1e12
        # fp32-accurate contraction (3 TF32 MMAs per product); the denominator is the measured bf16 cuBLAS peak of
        # MEASURED_PEAKS.json (no TF32 figure there; dense TF32 is nominally half of it) - stated, not hidden
        return dict(common, bound="tensor", achieved=ach, peak=pk["bf16_sustained"], unit="TFLOP/s", frac=ach / pk["bf16_sustained"],
                    algorithmic_flops_per_launch=flops_of[name],
                    note="algorithmic flops 2*B*(d+2)*d*K per launch / measured bf16 cuBLAS peak (sustained); the kernel issues 3x "
                         "that in TF32")
    ach = bytes_of[name] / t / 1e9
    return dict(common, bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                algorithmic_bytes_per_launch=bytes_of[name])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(SY.WORKLOADS), default="cfg2")
    ap.add_argument("--uniform", action="store_true", help="uniform instead of Zipf feature ids")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=45.0, help="seconds of CPU work allowed for the CPU legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1 and args.impl == "ours":
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 "
                         "bench.py --gpus %d ..." % (args.gpus, args.gpus))
    wl = dict(SY.WORKLOADS[args.workload])
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
