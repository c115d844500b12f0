#!/usr/bin/env python
"""bench.py -- throughput of the QFA hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--precision P]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic spectra.

  headline at N = 1:  sdss100k_predict (BASELINE.json configs[1]) -- batched log-likelihood + posterior continuum
      inference of 100 000 SDSS-shaped spectra (Npix 1913 / Nb 720) with the pretrained Nh = 8 model.
  headline at N > 1:  sdss_train (configs[3]) -- the likelihood + gradient train step, data parallel with ONE NCCL
      all-reduce of the packed accumulation buffer per step, so that the driver's 1 -> 8 curve sees the collective
      (predict / scoring shard with no collective at all: their efficiency is 1.0 by construction).  The N = 1 value
      of the SAME workload is `also.sdss_train.value` of the N = 1 line.
  `also` (same line, every entry with its own clocks / e2e / cpu_baseline / launch count):
      sdss_train, l32_train (configs[4]: Npix 1000, Nh 32, 30 % masked), desi_score (configs[2]: NLL-only scoring + on-device
      top-k), sdss_train_b8192 and sdss_train_b500 (configs[3] at 8192 spectra per GPU and at the reference's global batch of
      500: the WHOLE step -- shuffled gather + delta, accumulate, all-reduce, Adam+clip -- replayed as one CUDA graph),
      sdss_train_tf32x3 (3xTF32 operands).

`value` is spectra/s with the inputs resident in HBM; `e2e` is the same metric through the reference-shaped Python API
with HOST (pinned) buffers, H2D/D2H copies inside the timed region.  `gpu_launches` is counted by the library
(qfa_launch_count), not assumed.  `--impl reference` times the UNMODIFIED reference (baseline/_ref, pip --target install)
when it is present, else the dense CPU port in oracle/, on a bounded sample of the same workload, honouring --steps/--warmup.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--precision", default="mixed", choices=["mixed", "tf32", "tf32x3", "fp32", "fp64"])
    ap.add_argument("--spectra", type=int, default=0, help="spectra per GPU per step (0 = workload default)")
    ap.add_argument("--also", default="default", help="comma list of secondary workloads, 'default' or 'none'")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1 train steps: the library's one-shot peer-memory all-reduce (default) or ncclAllReduce")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# name: grid, Nh, kind, spectra per GPU per step, CPU sample (spectra per CPU step), BASELINE.json config, extras
WORKLOADS = {
    "sdss100k_predict": dict(grid="sdss", Nh=8, kind="predict", nspec=100_000, sample=192, baseline_config=1),
    "sdss_train": dict(grid="sdss", Nh=8, kind="train", nspec=71_040, sample=64, baseline_config=3),   # 148 SMs x 4 x 120
    "l32_train": dict(grid="l32", Nh=32, kind="train", nspec=65_536, sample=256, baseline_config=4),
    "desi_score": dict(grid="desi", Nh=8, kind="score", nspec=32_768, sample=4, baseline_config=2),
    "l32_predict": dict(grid="l32", Nh=32, kind="predict", nspec=65_536, sample=256, baseline_config=4),
    "sdss_train_b8192": dict(grid="sdss", Nh=8, kind="graph_train", nspec=8192, sample=64, baseline_config=3),
    "sdss_train_b500": dict(grid="sdss", Nh=8, kind="graph_train", nspec=500, sample=64, baseline_config=3,
                            global_batch=500),
    "sdss_train_tf32x3": dict(grid="sdss", Nh=8, kind="train", nspec=71_040, sample=64, baseline_config=3,
                              precision="tf32x3"),
    "sdss100k_predict_tf32x3": dict(grid="sdss", Nh=8, kind="predict", nspec=100_000, sample=192, baseline_config=1,
                                    precision="tf32x3"),
}
DEFAULT_ALSO = {1: ["sdss_train", "l32_train", "desi_score", "l32_predict", "sdss_train_b8192", "sdss_train_b500",
                    "sdss_train_tf32x3", "sdss100k_predict_tf32x3"],
                0: ["l32_train", "sdss_train_b8192", "sdss_train_b500"]}     # 0 = any N > 1


def bytes_per_spectrum(kind, P, Nb, Nh):
    """Algorithmic HBM bytes per spectrum."""
    if kind == "predict":
        return 17 * P + 4 * Nb + 4 * Nh * Nh + 4 * Nh + 4
    return 9 * P + 4 * Nb + 4


def committed_traffic(kind, grid_name, Nh, nspec):
    """DRAM bytes per launch of the dominant kernel, SCALED from the committed `ncu --set full` capture of the same kernel
    (profiles/r2_traffic.json, else r1): dram__bytes_read.sum + dram__bytes_write.sum per spectrum of that capture x spectra."""
    for f in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", f)))
            k = "train" if kind in ("train", "graph_train") else kind
            return t[f"{k}_{grid_name}_nh{Nh}"]["dram_bytes_per_spectrum"] * nspec, f"profiles/{f} (committed ncu capture, scaled)"
        except Exception:
            continue
    return None, None


def model_and_params(grid_name, Nh):
    from qfa_b200 import synth
    grid = synth.GRIDS[grid_name]
    gold = os.path.join(ROOT, "tests", "golden")
    if grid_name == "sdss" and Nh == 8:      # pretrained reference model (as load_from_npz holds it: c0 <- beta)
        k = np.load(os.path.join(gold, "kat_sdss.npz"))
        P = {key: torch.tensor(k["param_" + key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "beta")}
        P["c0"] = P["beta"].clone()
        mu = torch.tensor(k["param_mu"])
    elif grid_name == "desi" and Nh == 8:
        k = np.load(os.path.join(gold, "desi_params.npz"))
        P = {key: torch.tensor(k[key], dtype=torch.float32) for key in ("F", "Psi", "omega", "tau0", "c0", "beta")}
        mu = torch.tensor(k["mu"])
    else:
        P, mu = synth.smooth_random_params(grid, Nh, seed=1237)
    return grid, P, mu


def workload_config(name, nspec, precision, world=1):
    w = WORKLOADS[name]
    from qfa_b200 import synth
    grid = synth.GRIDS[w["grid"]]
    kind = w["kind"]
    cfg = {"workload": name, "baseline_config": "BASELINE.json configs[%d]" % w["baseline_config"],
           "kind": "train" if kind == "graph_train" else kind, "grid": w["grid"], "Npix": grid.Npix, "Nb": grid.Nb,
           "Nh": w["Nh"], "spectra_per_gpu_per_step": nspec, "precision": w.get("precision", precision),
           "l2": "inputs per step exceed L2 (%.0f MB)" % (nspec * (9 * grid.Npix + 4 * grid.Nb) / 1e6)
           if nspec * (9 * grid.Npix + 4 * grid.Nb) > 126e6 else
           "a NEW batch is gathered from the HBM-resident data set every step (shuffled rows; the data set exceeds L2)"}
    if kind == "graph_train":
        cfg["step"] = "one CUDA-graph replay: gather+delta of a shuffled batch, accumulate, all-reduce, Adam+clip"
        cfg["global_batch"] = w.get("global_batch", nspec * world)
    return cfg


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed regions are ~10 ms, far shorter than nvidia-smi's
    sampling period, so NVML is polled directly from a thread (every ~1 ms)."""

    def __init__(self, index):
        self.rows, self.ok, self.stop_flag = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.ok = False

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.time(), sm, rs))
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.t.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if not rows:                                   # region shorter than one poll: nearest samples around it
            rows = sorted(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:3]
        sm = [r[1] for r in rows]
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm)}


_REAL_STDOUT = None


def capture_stdout():
    """File descriptor 1 is pointed at stderr for the whole run, so that nothing a library writes to stdout (NCCL banner,
    warnings) can precede the JSON line; emit() writes that line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the CPUs next to its GPU.  Round 1
    ran all 8 ranks on NUMA node 0 (`CPU Affinity 0-31`), so every rank's H2D/D2H crossed the same host memory
    controller and the end-to-end numbers did not scale (SCALE_r01: 0.22 at N = 8)."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * wi + b for wi, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        info.update(gpu_cpus=len(cpus), allowed_before=len(allowed))
        try:
            os.sched_setaffinity(0, cpus)              # the GPU's node, even if the launcher pinned us elsewhere
            info.update(bound=True, cpus=len(cpus))
        except OSError:
            both = cpus & allowed
            if both:
                os.sched_setaffinity(0, both)
                info.update(bound=True, cpus=len(both))
        try:
            info["numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" %
                                         pynvml.nvmlDeviceGetPciInfo(h).busId.decode().lower()[-12:]).read())
        except Exception:
            pass
    except Exception as e:           # binding is an optimisation, never a failure
        info["error"] = str(e)[:80]
    return info


def dist_setup(n):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # NCCL prints its version banner on STDOUT (NCCL_DEBUG=VERSION, also from an nccl.conf): stdout carries the one JSON
        # line only -- see emit()
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def timed(fn, steps, warmup, world):
    """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks."""
    for _ in range(warmup):
        fn()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms, t0, t1


# ----------------------------------------------------------------------------- CPU arm
def cpu_run(name, steps, warmup):
    """The reference's CPU implementation of the workload on all host cores, on a bounded sample (`sample` spectra per step)
    of the same synthetic data: the UNMODIFIED reference (pip --target install under baseline/_ref, or $QFA_REF) through
    its own public API when it is importable -- QFA.forward (model.py:74-105) for the train step,
    prediction_for_single_spectra (model.py:160-180) per spectrum for predict / scoring, exactly the loops of
    model.py:210-214 and main.py:94-95 -- else the dense port oracle/qfa_dense.py."""
    from qfa_b200 import synth
    w = WORKLOADS[name]
    grid_name, Nh, kind, sample = w["grid"], w["Nh"], w["kind"], w["sample"]
    grid, P, mu = model_and_params(grid_name, Nh)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = dict(mask_iid=0.15, run_len=(40, 160)) if grid_name == "l32" else {}
    d = synth.make_spectra(P, mu, grid, sample, seed=1234, **kw)
    from oracle import ref_loader
    root = ref_loader.find_reference()
    train = kind in ("train", "graph_train")
    if root is not None:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            RQFA, _, _, _ = ref_loader.load_reference(root=root)
        cpu = torch.device("cpu")
        m = RQFA(grid.Nb, grid.Nr, Nh, cpu)
        m.F, m.Psi, m.omega = P["F"].clone(), P["Psi"].clone(), P["omega"].clone()
        m.tau0, m.c0, m.beta = P["tau0"].clone(), P["c0"].clone(), P["beta"].clone()
        m.mu = mu.clone()
        kind_s = "reference"

        def step():
            if train:
                m.forward(d["delta"], d["error"], d["zabs"], d["mask"])
            else:
                for b in range(sample):
                    m.prediction_for_single_spectra(d["flux"][b], d["error"][b], d["zabs"][b], d["mask"][b])
    else:
        from oracle import qfa_dense
        kind_s = "port"

        def step():
            if train:
                qfa_dense.forward(P, d["delta"], d["error"], d["zabs"], d["mask"], grid.Nb)
            else:
                for b in range(sample):
                    qfa_dense.predict_single(P, mu, d["flux"][b], d["error"][b], d["zabs"][b], d["mask"][b], grid.Nb)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    desc = (f"{sample} spectra/step of the {name} workload (seed 1234), {steps} step(s) after {warmup} warm-up; "
            + ("unmodified reference from " + os.path.relpath(root, ROOT) if root else "dense port oracle/qfa_dense.py"))
    return {"value": sample * steps / dt, "unit": "spectra/s", "cores": cores, "kind": kind_s, "sample": desc}, dt / steps * 1e3


def headline_for(n_gpus, requested):
    if requested != "auto":
        return requested
    return "sdss100k_predict" if n_gpus <= 1 else "sdss_train"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = headline_for(args.gpus, args.workload)
    w = WORKLOADS[name]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, ms = cpu_run(name, steps, warmup)
    line = {"impl": "reference", "metric": "spectra/sec", "value": cb["value"], "unit": "spectra/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(name, args.spectra or w["nspec"], args.precision, max(1, args.gpus)),
                           **({"allreduce": "none (the CPU reference is one host process: no exchange step)"}
                              if args.gpus > 1 and w["kind"] in ("train", "graph_train") else {})),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def kernel_names(kind, Nh, precision):
    if precision == "tf32x3":
        return ("k_tc_gram_x3<TRAIN> (tcgen05, 3xTF32) + k_grad<float> (CUDA cores) + k_reduce + k_adam"
                if kind in ("train", "graph_train") else "k_tc_gram_x3<PREDICT> (tcgen05, 3xTF32 Gram GEMM + double solve + continuum GEMM)")
    tc = precision in ("mixed", "tf32", "tf32x3")
    if kind in ("train", "graph_train"):
        if tc and Nh <= 8:
            return "k_tc_gram<TRAIN> + k_tc_grad (tcgen05) + k_reduce + k_adam"
        if tc and Nh <= 32:
            return "k_tc_gram32 + k_solve32 + k_tc_grad32 (tcgen05) + k_reduce + k_adam"
        return "k_gram_solve + k_grad (CUDA cores) + k_reduce + k_adam"
    if tc and Nh <= 8:
        return "k_tc_gram<PREDICT> (tcgen05 Gram GEMM + solve + continuum GEMM)"
    if tc and Nh <= 32:
        return "3 x k_tc_gram32<PRED> (tcgen05, 3xTF32 Grams) + k_solve32<PRED> + k_out32 (tcgen05)"
    return "k_gram_solve (CUDA cores)"


def measure_workload(name, args, rank, world, peaks, want_e2e, want_cpu, nspec_override=0):
    from qfa_b200 import QFA, Adam, step_scheduler, synth, DeviceDataloader, _lib
    w = WORKLOADS[name]
    grid_name, Nh, kind, nspec = w["grid"], w["Nh"], w["kind"], w["nspec"]
    precision = w.get("precision", args.precision)
    if nspec_override:
        nspec = nspec_override
    if kind == "graph_train" and "global_batch" in w:
        nspec = max(1, w["global_batch"] // world)          # strong scaling of the reference's batch of 500
    grid, P, mu = model_and_params(grid_name, Nh)
    dev = torch.device("cuda", torch.cuda.current_device())
    Pn = {k: v.numpy() for k, v in P.items()}
    m = QFA(grid.Nb, grid.Nr, Nh, dev, model_params=Pn, precision=precision)
    m.mu = mu
    if world > 1:
        m.enable_data_parallel(peer_allreduce=args.allreduce == "peer")
    kw = dict(mask_iid=0.15, run_len=(40, 160)) if grid_name == "l32" else {}
    ood = 0.01 if kind == "score" else 0.0
    L = _lib.lib()
    res = {}
    steps = args.steps
    if kind == "graph_train":
        # the data set: enough fresh rows for every timed step, at most ~1.6 GB; the cursor wraps by an in-stream reset
        rows = min(nspec * (steps + args.warmup), max(nspec, 81_920))
        d = synth.make_spectra(P, mu, grid, rows, seed=1234 + rank, device=dev, **kw)
        wav = grid.wav()
        dl = DeviceDataloader(d["flux"], d["error"], d["zqso"], d["mask"], wav, batch_size=nspec, device=dev,
                              shuffle=True, seed=5)      # rank-local loader: nspec rows per rank and step
        dl.data_size = rows * world
        del d
        opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
        p0 = m._params.clone()
        dl.rewind()
        g = m.capture_train_step(opt, dl, max(1, rows // nspec))
        per_epoch = rows // nspec
        state = {"n": 0}

        def step():
            if state["n"] % per_epoch == 0:
                dl._cursor.zero_()
            g.replay()
            state["n"] += 1
        launches_per_step = m.graph_launches_per_step
        X = E = Z = M = None
    else:
        d = synth.make_spectra(P, mu, grid, nspec, seed=1234 + rank, device=dev, ood_frac=ood, **kw)
        X = d["delta"] if kind == "train" else d["flux"]
        E, Z, M = d["error"], d["zabs"], d["mask"].view(torch.uint8)
    if kind == "train":
        opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
        p0 = m._params.clone()

        def step():
            acc = m.accumulate(X, E, Z, M, zero=True)
            m._allreduce(acc)
            opt.update_from_acc(m, acc)
    elif kind in ("predict", "score"):
        outs = {"nll": torch.empty(nspec, device=dev)}
        if kind == "predict":
            outs.update(hmean=torch.empty(nspec, Nh, device=dev), hcov=torch.empty(nspec, Nh, Nh, device=dev),
                        cont=torch.empty(nspec, grid.Npix, device=dev), unc=torch.empty(nspec, grid.Npix, device=dev))
        if precision == "fp64":
            outs = {k: v.double() for k, v in outs.items()}
        if kind == "score":
            def step():
                m.predict_into(X, E, Z, M, outs)
                res["ood"] = m.ood_select(outs["nll"].float(), threshold=None, k=256)     # on-device top-k of the NLLs
        else:
            def step():
                m.predict_into(X, E, Z, M, outs)
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    n0 = L.qfa_launch_count()
    ms, t0, t1 = timed(step, steps, args.warmup, world)
    n1 = L.qfa_launch_count()
    clocks = sampler.stop(t0, t1) if sampler else None
    if kind == "graph_train":
        launches = launches_per_step * steps
        launch_note = "kernel nodes of the captured step graph (counted by the library at capture) x replays"
    else:
        launches = int((n1 - n0) * steps // (steps + args.warmup))
        launch_note = "counted by the library (qfa_launch_count) over the timed steps"
    if kind in ("train", "graph_train"):
        m._params.copy_(p0)
    value = world * nspec * steps / (ms * 1e-3)
    bps = bytes_per_spectrum("predict" if kind == "predict" else "train", grid.Npix, grid.Nb, Nh)
    achieved = nspec * steps * bps / (ms * 1e-3) / 1e9        # per GPU, GB/s of algorithmic bytes
    traffic, tsrc = committed_traffic(kind, grid_name, Nh, nspec)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": tsrc,
            "kernel": kernel_names(kind, Nh, precision), "bytes_per_spectrum": bps, "peak_source": peaks["source"]}
    if Nh > 8 and kind == "predict":
        # tensor-bound like the Nh = 32 train step: algorithmic FLOPs per spectrum n*[H(H+1) + 4H + 20] + P*[H(H+1) + 2H] + 2H^3
        #  against the TF32 tensor peak
        n_un = float(M.float().sum() / nspec)
        flops = n_un * (Nh * (Nh + 1) + 4 * Nh + 20) + grid.Npix * (Nh * (Nh + 1) + 2 * Nh) + 2 * Nh ** 3
        tf = nspec * steps * flops / (ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": tf, "peak": peaks["tf32_tflops"], "unit": "TFLOP/s",
                "frac": tf / peaks["tf32_tflops"], "traffic": traffic, "traffic_source": tsrc,
                "kernel": kernel_names(kind, Nh, precision), "flops_per_spectrum": flops, "peak_source": peaks["tf32_source"],
                "note": "the Grams are evaluated as 3xTF32 (three tensor passes) to hold the 1e-3 continuum bar at cond(M) ~ 1e3"}
    if Nh > 8 and kind == "train":
        # Nh = 32 is tensor-bound (AI ~ 330 FLOP/B): algorithmic FLOPs per spectrum
        # n*[H(H+1)(2+r_b) + 2H^2 + 8H + 30] + 3H^3 against the TF32 tensor peak (= half the measured bf16 peak)
        n_un = float(M.float().sum() / nspec)
        r_b = float(M[:, :grid.Nb].float().sum() / max(1.0, float(M.float().sum())))
        flops = n_un * (Nh * (Nh + 1) * (2 + r_b) + 2 * Nh * Nh + 8 * Nh + 30) + 3 * Nh ** 3
        tf = nspec * steps * flops / (ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": tf, "peak": peaks["tf32_tflops"], "unit": "TFLOP/s",
                "frac": tf / peaks["tf32_tflops"], "traffic": traffic, "traffic_source": tsrc,
                "kernel": kernel_names(kind, Nh, precision), "flops_per_spectrum": flops, "peak_source": peaks["tf32_source"]}
    res.update(value=value, ms_per_step=ms / steps, clocks=clocks, gpu_launches=launches, gpu_launches_how=launch_note,
               roofline=roof, config=dict(workload_config(name, nspec, args.precision, world),
                                          **({"allreduce": m.allreduce_kind} if world > 1 and kind in ("train", "graph_train") else {})),
               dtype={"fp64": "f64", "fp32": "f32"}.get(precision, "tf32"))
    if "ood" in res:
        res["ood"] = {"top_k": 256, "top_nll_max": float(res["ood"]["top_val"][0]), "top_nll_min": float(res["ood"]["top_val"][-1])}
    if want_e2e:
        esteps = max(2, steps // 2)
        if kind == "graph_train":
            # end to end at the reference's batch size: host batch -> H2D -> forward -> all-reduce -> update -> loss D2H
            hb = synth.make_spectra(P, mu, grid, nspec, seed=99 + rank, **kw)
            hX, hE, hZ, hM = (hb[k].pin_memory() for k in ("delta", "error", "zabs", "mask"))
            opt2 = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3, weight_decay=0.1)
            hl = torch.empty(1, dtype=torch.float64).pin_memory()

            def estep():
                dX, dE, dZ, dM = (t.to(dev, non_blocking=True) for t in (hX, hE, hZ, hM))
                acc = m.accumulate(dX, dE, dZ, dM, zero=True)
                m._allreduce(acc)
                opt2.update_from_acc(m, acc)
                hl.copy_(m._loss_from_acc(acc).view(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            h2d = sum(t.numel() * t.element_size() for t in (hX, hE, hZ, hM))
            d2h = 8
        elif kind == "train":
            hX, hE, hZ, hM = (t.cpu().pin_memory() for t in (X, E, Z, M))

            def estep():
                m.forward_host(hX, hE, hZ, hM)
            h2d = sum(t.numel() * t.element_size() for t in (hX, hE, hZ, hM))
            d2h = (m.Nparams + 1) * 4
        else:
            hX, hE, hZ, hM = (t.cpu().pin_memory() for t in (X, E, Z, M))
            hout = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outs.items()}

            def estep():
                m.predict_host(hX, hE, hZ, hM, out=hout, want=tuple(outs))
            h2d = sum(t.numel() * t.element_size() for t in (hX, hE, hZ, hM))
            d2h = sum(t.numel() * t.element_size() for t in hout.values())
        ems, _, _ = timed(estep, esteps, 1, world)
        if kind in ("train", "graph_train"):
            m._params.copy_(p0)
        res["e2e"] = {"value": world * nspec * esteps / (ems * 1e-3), "unit": "spectra/s",
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": esteps}
        del hX, hE, hZ, hM
    if want_cpu and rank == 0:
        res["cpu_baseline"], _ = cpu_run(name, 1, 1)
    torch.cuda.empty_cache()
    return res


def dp_parity_check(rank, world):
    """Inside the N > 1 run: the all-reduced accumulation buffer is bit-identical on every rank, and equals the buffer a
    single rank computes over the SAME global batch (counts exactly, sums to float round-off of the summation order)."""
    import torch.distributed as dist
    from qfa_b200 import QFA, synth
    grid, P, mu = model_and_params("sdss", 8)
    dev = torch.device("cuda", torch.cuda.current_device())
    m = QFA(grid.Nb, grid.Nr, 8, dev, model_params={k: v.numpy() for k, v in P.items()}, precision="tf32")
    m.enable_data_parallel()
    B = 1536
    d = synth.make_spectra(P, mu, grid, B, seed=777 + rank, device=dev)
    ins = [d["delta"], d["error"], d["zabs"], d["mask"].view(torch.uint8)]
    loc = m.accumulate(*ins, zero=True).clone()
    acc = loc.clone()
    m._allreduce(acc)                                  # the exchange the train step really uses (peer one-shot kernel, or NCCL)
    nccl = loc.clone()
    dist.all_reduce(nccl, op=dist.ReduceOp.SUM)
    vs_nccl = float((acc - nccl).abs().max() / nccl.abs().max())
    gathered = [torch.empty_like(acc) for _ in range(world)]
    dist.all_gather(gathered, acc)
    identical = all(torch.equal(g.view(torch.int32), gathered[0].view(torch.int32)) for g in gathered)
    full = []
    for t in ins:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        full.append(torch.cat(parts))
    single = QFA(grid.Nb, grid.Nr, 8, dev, model_params={k: v.numpy() for k, v in P.items()}, precision="tf32")
    acc1 = single.accumulate(*full, zero=True)
    n = m.Nparams
    cnt_equal = bool(torch.equal(acc[n:n + grid.Npix + 3], acc1[n:n + grid.Npix + 3])) and float(acc[n + grid.Npix + 4]) == B * world
    rel = float((acc[:n] - acc1[:n]).abs().max() / acc1[:n].abs().max())
    rel_nll = abs(float(acc[n + grid.Npix + 3]) - float(acc1[n + grid.Npix + 3])) / abs(float(acc1[n + grid.Npix + 3]))
    ok = identical and cnt_equal and rel < 1e-4 and rel_nll < 1e-5 and vs_nccl < 1e-5
    return {"ok": bool(ok), "allreduce": m.allreduce_kind, "vs_nccl_all_reduce_max_rel": vs_nccl, "bit_identical_across_ranks": bool(identical), "counts_equal_single_rank": cnt_equal,
            "grad_sums_vs_single_rank_max_rel": rel, "nll_sum_vs_single_rank_rel": rel_nll,
            "global_batch": B * world, "note": "dp_parity: ok" if ok else "dp_parity: FAILED"}


def main():
    args = parse()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the QFA kernels have no CPU fallback"})
        sys.exit(1)
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if local == 0:
        from qfa_b200 import _lib
        _lib.build()                      # no-op when libqfa_b200.so is up to date
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    peaks = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)", "tf32_tflops": 1590.0 / 2,
             "tf32_source": "half of the fallback bf16 peak (B200_PROFILING.md)"}
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pj):
        mp = json.load(open(pj))
        peaks = {"hbm_gbs": float(mp["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)",
                 "tf32_tflops": float(mp.get("bf16_tflops_sustained", mp.get("bf16_tflops", 1590.0))) / 2,
                 "tf32_source": "half of the measured sustained bf16 peak (MEASURED_PEAKS.json)"}
    head = headline_for(world, args.workload)
    dp = dp_parity_check(rank, world) if world > 1 else None
    if dp is not None and not dp["ok"]:
        if rank == 0:
            emit({"error": "data-parallel parity check failed", "dp_parity": dp})
        sys.exit(2)
    main_res = measure_workload(head, args, rank, world, peaks, not args.no_e2e, not args.no_cpu_baseline, args.spectra)
    also = {}
    names = []
    if not args.no_also and args.also != "none":
        names = DEFAULT_ALSO[1 if world == 1 else 0] if args.also == "default" else [s for s in args.also.split(",") if s]
    for name in names:
        if name == head:
            continue
        r = measure_workload(name, args, rank, world, peaks, not args.no_e2e, not args.no_cpu_baseline and world == 1)
        also[name] = {k: r[k] for k in ("value", "ms_per_step", "roofline", "config", "gpu_launches", "gpu_launches_how",
                                         "clocks", "dtype", "e2e", "cpu_baseline", "ood") if k in r}
        also[name]["unit"] = "spectra/s"
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank == 0:
        line = {"metric": "spectra/sec", "value": main_res["value"], "unit": "spectra/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": main_res["dtype"], "data": "synthetic",
                "config": main_res["config"], "clocks": main_res["clocks"], "gpu_launches": main_res["gpu_launches"],
                "gpu_launches_how": main_res["gpu_launches_how"], "roofline": main_res["roofline"], "numa": numa}
        if world > 1:
            line["scaling_note"] = ("headline at N > 1 is the data-parallel train step (one all-reduce of the accumulator per step: "
                                    "dp_parity.allreduce says which kernel); its N = 1 "
                                    "value is also.sdss_train.value of the N = 1 line, NOT that line's predict headline")
            line["dp_parity"] = dp
        for k in ("e2e", "cpu_baseline"):
            if k in main_res:
                line[k] = main_res[k]
        if also:
            line["also"] = also
        emit(line)
    if world > 1:
        import gc
        import threading
        import torch.distributed as dist
        # the line is out; tear the process group down, but never let a stuck communicator teardown hang the run
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.destroy_process_group()
        t.cancel()


if __name__ == "__main__":
    main()
