#!/usr/bin/env python
"""bench.py -- throughput of the QFA hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic spectra:

  workload sdss100k_predict (default; BASELINE.json configs[1]): batched log-likelihood +
      posterior continuum inference of 100 000 SDSS-shaped spectra (Npix 1913 / Nb 720) with the
      pretrained Nh=8 model, per GPU (weak scaling: every rank owns its own 100k spectra, no
      collective on the data path).
  `also` (reported inside the same line): the likelihood+gradient train step of configs[3]
      (SDSS shape, Nh=8) and configs[4] (Npix 1000, Nh=32, 30 % masked), data-parallel with ONE
      all-reduce of the packed accumulation buffer per step.

`value` is spectra/s with the inputs resident in HBM; `e2e` is the same metric through the
reference-shaped Python API with HOST (pinned) buffers, H2D/D2H copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
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
    ap.add_argument("--workload", default="sdss100k_predict",
                    choices=["sdss100k_predict", "sdss_train", "l32_train", "desi_score"])
    ap.add_argument("--precision", default="mixed", choices=["mixed", "fp32", "fp64"])
    ap.add_argument("--spectra", type=int, default=0, help="spectra per GPU per step (0 = workload default)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary (train-step) measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


WORKLOADS = {
    # name: (grid, Nh, kind, default spectra per GPU, cpu sample)
    "sdss100k_predict": ("sdss", 8, "predict", 100_000, 192),
    "sdss_train": ("sdss", 8, "train", 71_040, 64),        # 148 SMs x 4 waves x 120-spectra tiles
    "l32_train": ("l32", 32, "train", 65_536, 256),
    "desi_score": ("desi", 8, "score", 32_768, 4),
}


def measured_traffic(kind, grid_name, Nh, nspec):
    """DRAM bytes per launch of the dominant kernel, scaled from the committed `ncu --set full` capture
    (profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per spectrum of that capture)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return t[f"{kind}_{grid_name}_nh{Nh}"]["dram_bytes_per_spectrum"] * nspec
    except Exception:
        return None


def bytes_per_spectrum(kind, P, Nb, Nh):
    """Algorithmic HBM bytes per spectrum."""
    if kind == "predict":
        return 17 * P + 4 * Nb + 4 * Nh * Nh + 4 * Nh + 4
    return 9 * P + 4 * Nb + 4


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


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region of the default run is ~10 ms, far
    shorter than nvidia-smi's sampling period, so NVML is polled directly from a thread (every ~1 ms)."""

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


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_reference_run(workload, steps, warmup, sample):
    """Times the dense CPU port of the reference algorithm (oracle/qfa_dense.py -- the reference itself
    is pure Python and does not travel to the GPU box) on all host cores."""
    from oracle import qfa_dense
    from qfa_b200 import synth
    grid_name, Nh, kind, _, _ = WORKLOADS[workload]
    grid, P, mu = model_and_params(grid_name, Nh)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = dict(mask_iid=0.15, run_len=(40, 160)) if grid_name == "l32" else {}
    d = synth.make_spectra(P, mu, grid, sample, seed=1234, **kw)

    def step():
        if kind == "train":
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
    return sample * steps / dt, dt / steps * 1e3, cores, f"{sample} spectra/step of the {workload} workload (seed 1234)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid_name, Nh, kind, nspec, sample = WORKLOADS[args.workload]
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    v, ms, cores, desc = cpu_reference_run(args.workload, steps, warmup, sample)
    line = {"impl": "reference", "metric": "spectra/sec", "value": v, "unit": "spectra/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "kind": kind, "grid": grid_name, "Nh": Nh},
            "cpu_baseline": {"value": v, "unit": "spectra/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def measure_workload(name, args, rank, world, peaks, want_e2e, want_cpu, nspec_override=0):
    from qfa_b200 import QFA, Adam, step_scheduler, synth
    grid_name, Nh, kind, nspec, sample = WORKLOADS[name]
    if nspec_override:
        nspec = nspec_override
    grid, P, mu = model_and_params(grid_name, Nh)
    dev = torch.device("cuda", torch.cuda.current_device())
    Pn = {k: v.numpy() for k, v in P.items()}
    m = QFA(grid.Nb, grid.Nr, Nh, dev, model_params=Pn, precision=args.precision)
    m.mu = mu
    if world > 1:
        m.enable_data_parallel()
    kw = dict(mask_iid=0.15, run_len=(40, 160)) if grid_name == "l32" else {}
    ood = 0.01 if kind == "score" else 0.0
    d = synth.make_spectra(P, mu, grid, nspec, seed=1234 + rank, device=dev, ood_frac=ood, **kw)
    X = d["delta"] if kind == "train" else d["flux"]
    E, Z, M = d["error"], d["zabs"], d["mask"].view(torch.uint8)
    launches = [0]
    res = {}
    if kind == "train":
        opt = Adam(params=m.parameters, device=dev, scheduler=step_scheduler(0.9, 10), learning_rate=1e-3,
                   weight_decay=0.1)
        p0 = m._params.clone()

        def step():
            acc = m.accumulate(X, E, Z, M, zero=True)
            m._allreduce(acc)
            opt.update_from_acc(m, acc)
        if Nh <= 8 and args.precision == "mixed":
            # k_tc_build_images, k_tc_gram<TRAIN>, k_tc_grad, k_reduce, k_adam
            launches_per_step = 5
            kernel_name = "k_tc_gram<TRAIN> + k_tc_grad (tcgen05) + k_reduce + k_adam"
        elif 8 < Nh <= 32 and args.precision == "mixed":
            # k_tc_build_images32, k_tc_gram32, k_solve32, k_tc_grad32, k_reduce, k_adam
            launches_per_step = 6
            kernel_name = "k_tc_gram32 (tcgen05, 3 passes) + k_solve32 + k_tc_grad32 (tcgen05) + k_reduce + k_adam"
        else:
            n_sub = -(-nspec // max(1, min(nspec, (48 << 20) // (9 * grid.Npix + 4 * grid.Nb))))
            launches_per_step = 2 * n_sub + 1 + 1
            kernel_name = "k_gram_solve + k_grad (CUDA cores) + k_reduce + k_adam"
    else:
        want = ("nll", "hmean", "hcov", "cont", "unc") if kind == "predict" else ("nll",)
        outs = {"nll": torch.empty(nspec, device=dev)}
        if kind == "predict":
            outs.update(hmean=torch.empty(nspec, Nh, device=dev), hcov=torch.empty(nspec, Nh, Nh, device=dev),
                        cont=torch.empty(nspec, grid.Npix, device=dev), unc=torch.empty(nspec, grid.Npix, device=dev))
        if args.precision == "fp64":
            outs = {k: v.double() for k, v in outs.items()}

        def step():
            m.predict_into(X, E, Z, M, outs)
        if Nh <= 8 and args.precision == "mixed":
            launches_per_step = 2          # k_tc_build_images, k_tc_gram<PREDICT>
            kernel_name = "k_tc_gram<PREDICT> (tcgen05 Gram GEMM + solve + continuum GEMM)"
        else:
            launches_per_step = 1
            kernel_name = "k_gram_solve (CUDA cores)"
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    ms, t0, t1 = timed(step, args.steps, args.warmup, world)
    clocks = sampler.stop(t0, t1) if sampler else None
    if kind == "train":
        m._params.copy_(p0)
    value = world * nspec * args.steps / (ms * 1e-3)
    bps = bytes_per_spectrum("predict" if kind == "predict" else "train", grid.Npix, grid.Nb, Nh)
    achieved = nspec * args.steps * bps / (ms * 1e-3) / 1e9        # per GPU, GB/s of algorithmic bytes
    roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": measured_traffic(kind, grid_name, Nh, nspec),
            "kernel": kernel_name, "bytes_per_spectrum": bps, "peak_source": peaks["source"]}
    if Nh > 8 and kind == "train":
        # Nh = 32 is tensor-bound (AI ~ 330 FLOP/B): algorithmic FLOPs per spectrum
        # n*[H(H+1)(2+r_b) + 2H^2 + 8H + 30] + 3H^3 against the TF32 tensor peak (= half the measured bf16 peak)
        n_un = float(M.float().sum() / nspec)
        r_b = float(M[:, :grid.Nb].float().sum() / max(1.0, float(M.float().sum())))
        flops = n_un * (Nh * (Nh + 1) * (2 + r_b) + 2 * Nh * Nh + 8 * Nh + 30) + 3 * Nh ** 3
        tf = nspec * args.steps * flops / (ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": tf, "peak": peaks["tf32_tflops"], "unit": "TFLOP/s",
                "frac": tf / peaks["tf32_tflops"], "traffic": None, "kernel": kernel_name,
                "flops_per_spectrum": flops, "peak_source": peaks["tf32_source"]}
    res.update(value=value, ms_per_step=ms / args.steps, clocks=clocks, gpu_launches=launches_per_step * args.steps,
               roofline=roof,
               config={"workload": name, "kind": kind, "grid": grid_name, "Npix": grid.Npix, "Nb": grid.Nb, "Nh": Nh,
                       "spectra_per_gpu_per_step": nspec, "precision": args.precision,
                       "l2": "inputs per step exceed L2 (%.0f MB)" % (nspec * (9 * grid.Npix + 4 * grid.Nb) / 1e6)})
    if want_e2e:
        hX, hE, hZ, hM = (t.cpu().pin_memory() for t in (X, E, Z, M))
        if kind == "train":
            def estep():
                m.forward_host(hX, hE, hZ, hM)
            h2d = sum(t.numel() * t.element_size() for t in (hX, hE, hZ, hM))
            d2h = (m.Nparams + 1) * 4
        else:
            hout = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outs.items()}

            def estep():
                m.predict_host(hX, hE, hZ, hM, out=hout, want=tuple(outs))
            h2d = sum(t.numel() * t.element_size() for t in (hX, hE, hZ, hM))
            d2h = sum(t.numel() * t.element_size() for t in hout.values())
        ems, _, _ = timed(estep, max(2, args.steps // 2), 1, world)
        res["e2e"] = {"value": world * nspec * max(2, args.steps // 2) / (ems * 1e-3), "unit": "spectra/s",
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        del hX, hE, hZ, hM
    if want_cpu and rank == 0:
        v, cms, cores, desc = cpu_reference_run(name, 1, 1, sample)
        res["cpu_baseline"] = {"value": v, "unit": "spectra/s", "cores": cores, "kind": "port", "sample": desc}
    del d, X, E, Z, M
    torch.cuda.empty_cache()
    return res


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
    main_res = measure_workload(args.workload, args, rank, world, peaks, not args.no_e2e, not args.no_cpu_baseline,
                                args.spectra)
    also = {}
    if not args.no_also:
        for name in ("sdss_train", "l32_train"):
            if name == args.workload:
                continue
            r = measure_workload(name, args, rank, world, peaks, False, False)
            also[name] = {"value": r["value"], "unit": "spectra/s", "ms_per_step": r["ms_per_step"],
                          "roofline": r["roofline"], "config": r["config"], "gpu_launches": r["gpu_launches"]}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank == 0:
        line = {"metric": "spectra/sec", "value": main_res["value"], "unit": "spectra/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
                "config": main_res["config"], "clocks": main_res["clocks"], "gpu_launches": main_res["gpu_launches"],
                "roofline": main_res["roofline"]}
        if "e2e" in main_res:
            line["e2e"] = main_res["e2e"]
        if "cpu_baseline" in main_res:
            line["cpu_baseline"] = main_res["cpu_baseline"]
        if also:
            line["also"] = also
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
