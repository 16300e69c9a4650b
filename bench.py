#!/usr/bin/env python
"""bench.py — images/sec of the generator hot path on B200 (BASELINE.json's metric, on BASELINE.json's configs).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl ours|reference|gpu_reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workloads (one "step" = one pass over one per-GPU batch of synthetic inputs; per-GPU batch fixed = weak scaling):
  picnet_ref    (default, headline) configs[0]: PICNet-ref `ReferenceFill` forward, 256^2, batch 4, fp32 I/O
  refpsp        configs[2]: RefpSp `pSp` forward (IR-SE50 encoder x2 + attention + StyleGAN2-1024), batch 8, bf16 operands
  train_picnet  configs[3]: train_reference_fill.py's step — G forward, GANOptimizer (D forward x2, VGG perceptual / style /
                contextual losses, both backwards, both Adam steps), batch 4 per GPU, NCCL gradient all-reduce when N > 1
  train_psp     configs[4]: train_psp.py's step with --train_decoder --use_ref --use_attention --randomize_noise, batch 2 per
                GPU, bf16 operands, NCCL gradient all-reduce when N > 1
The default run prints ONE JSON line for `picnet_ref` and carries the other three as full records under "records"
(--no-extras skips them), so the driver's 1/2/4/8-GPU runs also time the training steps with their collectives.

Keys of a record:
  value         whole-job img/s, inputs resident in HBM, CUDA events per step, L2 flushed between steps, max over ranks
  e2e           the same through the public call from pinned HOST buffers: H2D of the step's inputs + forward (+ loss /
                backward / optimizer) + D2H of the result inside the timed region
  roofline      the dominant kernel of the step, timed per launch with CUDA events on its launch stream (fmi_profile_*) in an
                eager pass of the same step; `kernels` lists every timed kernel with its share of the step
  cpu_baseline  (N = 1) the UNMODIFIED reference (baseline/_ref, a verbatim copy) on the host cores, bounded sample
  gpu_reference the UNMODIFIED reference on the same B200 (its own formulation: ATen / cuBLAS / cuDNN + its two CUDA ops)
`--impl reference`: the reference arm — the unmodified reference module(s) of the workload on the host cores, same `config`.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "picnet_ref": dict(batch=4, dtype="tf32", cfg="configs[0]: PICNet ref-guided generator (ReferenceFill) forward, 256x256, "
                       "use_att=1, decoder_img_f=256 decoder_z_nc=256, random-init, synthetic masked / reference / mask inputs"),
    "refpsp": dict(batch=8, dtype="bf16", cfg="configs[2]: RefpSp inference (pSp: IR-SE50 GradualStyleEncoder on source + "
                   "reference, attention1/2, StyleGAN2 1024x1024 decoder, face_pool to 256x256), randomize_noise=False"),
    "train_picnet": dict(batch=4, dtype="tf32", cfg="configs[3]: PICNet reference-fill GAN train step (train_reference_fill.py: "
                         "generator forward + GANOptimizer: D forward x2, VGG16 perceptual/style/contextual losses, G and D "
                         "backward, Adam x2), 256x256, random-init VGG16"),
    "train_psp": dict(batch=2, dtype="bf16", cfg="configs[4]: RefpSp decoder + attention training step (train_psp.py "
                      "--train_decoder 1 --use_ref --use_attention --randomize_noise 1: pSp forward 1024x1024, pSpLoss "
                      "(masked L2 + LPIPS-alex + VGG style/contextual), backward, Adam), random-init loss networks"),
}


def workload_config(name: str, n_gpus: int) -> dict:
    """The `config` object — identical in both arms (the driver compares them)."""
    w = WORKLOADS[name]
    return {"workload": w["cfg"], "name": name, "per_gpu_batch": w["batch"], "global_batch": w["batch"] * n_gpus,
            "image": "256x256 inputs" + (", 1024x1024 synthesis" if "psp" in name else ", 1024x1024 decoder output pooled to 256x256"),
            "l2": "a 256 MiB buffer is overwritten between timed steps (L2 = 126 MB); per-step activations are > 1 GB anyway",
            "parallelism": f"batch-sharded x{n_gpus}" + (", NCCL gradient all-reduce" if name.startswith("train") and n_gpus > 1
                                                          else ", no collective")}


# ------------------------------------------------------------------------------------------------------------ inputs
def make_inputs(name: str, batch: int, seed: int):
    """U[0,1) images (PICNet) / U[-1,1) (pSp), binary mask = lower-face rectangle (+ Bernoulli(0.3) speckle)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    src = torch.rand(batch, 3, 256, 256, generator=g)
    ref = torch.rand(batch, 3, 256, 256, generator=g)
    gt = torch.rand(batch, 3, 256, 256, generator=g)
    mask = (torch.rand(batch, 256, 256, generator=g) < 0.3).float() if "picnet" in name else torch.zeros(batch, 256, 256)
    mask[:, 128:230, 50:206] = 1.0
    if "psp" in name:
        src, ref, gt = src * 2 - 1, ref * 2 - 1, gt * 2 - 1
    return src, ref, gt, mask


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
        return {"sm_mhz": sorted(load)[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d.get("hbm_gbs", 6456.8), "bf16_tflops": d.get("bf16_tflops", 1667.1),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1403.9), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_sustained": 1400.0,
            "source": "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"}


# ------------------------------------------------------------------------------------------------------------ harness
class Harness:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py: no CUDA device — the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1 and not dist.is_initialized():
            import datetime
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(minutes=30))
            dist.barrier(device_ids=[self.local_rank])
        self.steps, self.warmup = args.steps, max(3, args.warmup)
        self._flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        from face_mask_inpaint_b200 import _lib
        self.lib = _lib.load()
        _lib.check(self.lib.fmi_device_check(), "fmi_device_check")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps=None, warmup=None) -> float:
        """W untimed steps, then exactly K steps, each between its own pair of CUDA events with the L2 overwritten before it;
        barrier + synchronize on both sides; mean step time, max over ranks."""
        steps = steps or self.steps
        for _ in range(self.warmup if warmup is None else warmup):
            step()
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            self._flush.fill_(1)
            e0.record()
            step()
            e1.record()
        self.barrier()
        return self.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs) / steps)

    # ---- per-kernel records of an eager pass (fmi_profile_dump) -> kernel table + roofline of the dominant kernel
    def profile(self, step, steps, tensor_peak_tf, peaks):
        lib = self.lib
        for k in range(lib.fmi_profile_kinds()):     # drop anything recorded earlier
            n = ctypes.c_int(0)
            lib.fmi_profile_dump(k, None, None, None, 0, ctypes.byref(n))
        step()
        torch.cuda.synchronize()
        lib.fmi_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        lib.fmi_profile_enable(0)
        eager_ms = e0.elapsed_time(e1) / steps
        cap = 1 << 16
        ms_b, fl_b, by_b = (ctypes.c_double * cap)(), (ctypes.c_double * cap)(), (ctypes.c_double * cap)()
        kernels = []
        for k in range(lib.fmi_profile_kinds()):
            n = ctypes.c_int(0)
            lib.fmi_profile_dump(k, ms_b, fl_b, by_b, cap, ctypes.byref(n))
            if n.value == 0:
                continue
            recs = [(ms_b[i], fl_b[i], by_b[i]) for i in range(n.value)]
            t = sum(r[0] for r in recs) * 1e-3
            fl, by = sum(r[1] for r in recs), sum(r[2] for r in recs)
            sol = sum(max(r[2] / (peaks["hbm_gbs"] * 1e9), r[1] / (tensor_peak_tf * 1e12)) for r in recs)
            hbm_sol = sum(r[2] / (peaks["hbm_gbs"] * 1e9) for r in recs
                          if r[2] / (peaks["hbm_gbs"] * 1e9) >= r[1] / (tensor_peak_tf * 1e12))
            groups = {}
            for ms, f, b in recs:
                g = groups.setdefault((f, b), [0, 0.0])
                g[0] += 1
                g[1] += ms
            top = sorted(groups.items(), key=lambda kv: -kv[1][1])[:4]
            kernels.append({
                "kernel": lib.fmi_profile_kind_name(k).decode(), "launches_per_step": n.value / steps,
                "ms_per_step": t * 1e3 / steps, "algorithmic_gflop_per_step": fl / steps / 1e9,
                "algorithmic_mb_per_step": by / steps / 1e6, "tflops": fl / t / 1e12 if t > 0 else 0.0,
                "gbs": by / t / 1e9 if t > 0 else 0.0, "sol_frac": sol / t if t > 0 else 0.0,
                "hbm_bound_share_of_sol": hbm_sol / sol if sol > 0 else 0.0,
                "top_launches": [{"count_per_step": c / steps, "avg_us": 1e3 * ms / c, "gflop": f / 1e9, "mb": b / 1e6,
                                  "tflops": f / (ms / c * 1e-3) / 1e12, "gbs": b / (ms / c * 1e-3) / 1e9,
                                  "sol_frac": max(b / (peaks["hbm_gbs"] * 1e9), f / (tensor_peak_tf * 1e12)) / (ms / c * 1e-3)}
                                 for (f, b), (c, ms) in top]})
        timed_ms = sum(k["ms_per_step"] for k in kernels)
        for k in kernels:
            k["share_of_timed_kernels"] = k["ms_per_step"] / timed_ms if timed_ms else 0.0
            k["share_of_eager_step"] = k["ms_per_step"] / eager_ms if eager_ms else 0.0
        kernels.sort(key=lambda k: -k["ms_per_step"])
        return kernels, eager_ms


_TRAFFIC_FILES = {"picnet_ref": "r02_traffic_picnet_b4.json", "refpsp": "r02_traffic_refpsp_b8.json"}
_TRAFFIC_NAMES = {"conv_gemm": ("conv_gemm: ", "modconv_gemm_kernel"), "conv_gemm_ir": ("conv_gemm_ir: ",), "attn_fwd": ("attn_fwd2_kernel",),
                  "norm_act": ("norm_act_kernel",), "out_conv_tanh": ("out_conv_tanh_kernel",), "instnorm_stats": ("instnorm_stats_kernel",)}


def ncu_traffic(workload, kernel):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of all launches of `kernel` in ONE step of `workload`, from the
    committed ncu launch list of the same forward (tools/gpu_launchlists.sh -> tools/dram_summary.py -> profiles/r02_traffic_*.json;
    same per-GPU batch as the bench). None when no capture of that workload / kernel is committed."""
    f = ROOT / "profiles" / _TRAFFIC_FILES.get(workload, "")
    if not f.is_file():
        return None
    try:
        rows = json.loads(f.read_text())["kernels"]
    except (ValueError, KeyError):
        return None
    for pat in _TRAFFIC_NAMES.get(kernel, ()):
        hit = [r for r in rows if (r["kernel"].startswith(pat) if pat.endswith(": ") else pat in r["kernel"])]
        if hit:
            return {"bytes_per_step": sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in hit),
                    "read_bytes_per_step": sum(r["dram_read_bytes"] for r in hit),
                    "write_bytes_per_step": sum(r["dram_write_bytes"] for r in hit),
                    "launches_per_step": sum(r["launches"] for r in hit), "source": "profiles/" + f.name}
    return None


def roofline_of(kernels, peaks, tensor_peak_tf, operand, traffic=None, workload=None):
    """The dominant kernel (largest summed duration per step) on its roofline. A kernel with heterogeneous launches (the
    implicit-GEMM conv serves HBM-bound 1024^2 layers and tensor-bound 64^2 layers) is put on the roofline that bounds most
    of its speed-of-light time; `sol_frac` = sum over launches of max(bytes / HBM, flops / tensor) / measured time."""
    if not kernels:
        return None
    k = kernels[0]
    hbm = k["hbm_bound_share_of_sol"] >= 0.5
    out = {"bound": "hbm" if hbm else "tensor", "kernel": k["kernel"], "achieved": k["gbs"] if hbm else k["tflops"],
           "peak": peaks["hbm_gbs"] if hbm else tensor_peak_tf, "unit": "GB/s" if hbm else "TFLOP/s",
           "traffic": traffic, "sol_frac": k["sol_frac"], "ms_per_step": k["ms_per_step"],
           "launches_per_step": k["launches_per_step"], "share_of_timed_kernels": k["share_of_timed_kernels"],
           "algorithmic_bytes_per_step": k["algorithmic_mb_per_step"] * 1e6,
           "algorithmic_flops_per_step": k["algorithmic_gflop_per_step"] * 1e9,
           "peak_source": f"{peaks['source']}: HBM copy {peaks['hbm_gbs']} GB/s; tensor {tensor_peak_tf:.0f} TFLOP/s = burst bf16 "
                          f"{peaks['bf16_tflops']}" + (" / 2 (kind::tf32 MMAs run at half the bf16 rate)" if operand == "tf32" else ""),
           "how": "CUDA events around every launch on its own stream during an eager pass of the same step (the timed `value` "
                  "replays a CUDA graph of the same launches); algorithmic bytes = inputs once + outputs once + weights once"}
    out["frac"] = out["achieved"] / out["peak"]
    t = ncu_traffic(workload, k["kernel"]) if traffic is None and workload else None
    if t:
        out["traffic"] = t["bytes_per_step"]
        out["traffic_detail"] = t
    return out


# ------------------------------------------------------------------------------------------------------------ ours: forward
def _forward_record(h: Harness, name, net, call_args, kw_tensors, kw_other, dtype, what):
    """`call_args` / `kw_tensors`: HOST tensors (positional / keyword) of one step; `kw_other`: the non-tensor keywords."""
    from face_mask_inpaint_b200.graphs import CapturedForward
    w = WORKLOADS[name]
    b = w["batch"]
    peaks = measured_peaks()
    tensor_peak = peaks["bf16_tflops"] / (2.0 if dtype == "tf32" else 1.0)
    with torch.no_grad():
        host_in = [t.pin_memory() for t in list(call_args) + list(kw_tensors.values())]   # CapturedForward's flatten order
        dev_in = [t.to(h.dev) for t in call_args]
        call_kwargs = {**{k: t.to(h.dev) for k, t in kw_tensors.items()}, **kw_other}
        for _ in range(2):
            out = net(*dev_in, **call_kwargs)
        torch.cuda.synchronize()
        n0 = h.lib.fmi_kernel_launch_count()
        out = net(*dev_in, **call_kwargs)
        torch.cuda.synchronize()
        launches = h.lib.fmi_kernel_launch_count() - n0
        kernels, eager_ms = h.profile(lambda: net(*dev_in, **call_kwargs), min(h.steps, 5), tensor_peak, peaks)
        fwd = CapturedForward(net, *dev_in, **call_kwargs)
        sampler = ClockSampler(h.local_rank) if h.rank == 0 else None
        ms = h.timed(fwd.replay)
        # end to end: pinned host inputs -> the graph's static inputs, one replay, result -> pinned host
        out_dev = fwd.replay()
        out_dev = out_dev[0] if isinstance(out_dev, (tuple, list)) else out_dev
        host_out = torch.empty(out_dev.shape, dtype=out_dev.dtype, pin_memory=True)
        stat = fwd.static_inputs

        def e2e_step():
            for dst, src in zip(stat, host_in):
                dst.copy_(src, non_blocking=True)
            fwd.replay()
            host_out.copy_(out_dev, non_blocking=True)

        ms_e2e = h.timed(e2e_step)
        clocks = sampler.stop() if sampler else None
        if not torch.isfinite(host_out).all():
            raise RuntimeError(f"{name}: non-finite output")
    rec = {"metric": "images/sec", "value": h.world * b / (ms * 1e-3), "unit": "img/s", "n_gpus": h.world, "steps": h.steps,
           "warmup": h.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": dtype, "data": "synthetic", "config": workload_config(name, h.world), "clocks": clocks,
           "e2e": {"value": h.world * b / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_in),
                   "d2h_bytes_per_step": host_out.numel() * host_out.element_size(),
                   "api": "pinned host tensors -> static inputs of graphs.CapturedForward(model.forward) -> pinned host image"},
           "gpu_launches": int(launches) * h.steps,
           "launch": f"one CUDA-graph replay per forward capturing {int(launches)} sm_100a kernel launches of this package",
           "eager_ms_per_step": eager_ms, "what": what,
           "roofline": roofline_of(kernels, peaks, tensor_peak, dtype, workload=name), "kernels": kernels}
    del fwd
    return rec


def _parity(h: Harness, name, mirror, call, ref_ops, tol, strict_too=False, full_call=None):
    """Whole-model parity of THIS run's model: the UNMODIFIED reference class (baseline/_ref) on the same GPU with the same
    weights (strict state_dict load), inputs and RNG seed, run in strict fp32 (TF32 off everywhere) = `want`; this package's
    forward = `got`. Also the reference's own default GPU execution (cuDNN TF32 allowed, PyTorch's default) against `want`, i.e.
    the error the reference's users have on this GPU. Every forward starts from a deep copy (SpectralNorm u / v advance)."""
    import copy
    from baseline import reference as R
    if not R.available():
        return {"unavailable": "baseline/_ref is missing"}
    try:
        R.import_unpatched(ref_ops)
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):      # the reference's constructors print; stdout carries the JSON line only
            ref = (R.reference_fill() if name == "picnet_ref" else R.psp(output_size=1024)).eval()
        ref.load_state_dict(mirror.state_dict(), strict=True)
        ref = ref.to(h.dev)
        if name == "refpsp":
            ref.latent_avg = ref.latent_avg.to(h.dev)
        a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        rel = lambda x, y: ((x.double() - y.double()).abs().max() / y.double().abs().max()).item()
        with torch.no_grad():
            torch.manual_seed(5)
            ref_default = call(copy.deepcopy(ref), True)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            try:
                torch.manual_seed(5)
                want = call(copy.deepcopy(ref), True)
                if strict_too:      # this package under the same switch: conv blocks with 3xTF32 split operands
                    torch.manual_seed(5)
                    got_strict = call(copy.deepcopy(mirror), False)
                    if full_call is not None:     # the reference's strict-fp32 GPU run timed on the workload's full batch
                        r2 = copy.deepcopy(ref)
                        for _ in range(2):
                            full_call(r2, True)
                        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                        ev[0].record()
                        for _ in range(5):
                            full_call(r2, True)
                        ev[1].record()
                        torch.cuda.synchronize()
                        ref_fp32_ms = ev[0].elapsed_time(ev[1]) / 5
                        del r2
            finally:
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
            n0 = h.lib.fmi_kernel_launch_count()
            torch.manual_seed(5)
            got = call(copy.deepcopy(mirror), False)
            n1 = h.lib.fmi_kernel_launch_count()
        out = {"ours_vs_reference_fp32": rel(got, want), "reference_default_gpu_vs_reference_fp32": rel(ref_default, want),
               "tolerance": tol, "kernel_launches_of_the_checked_forward": int(n1 - n0),
               "how": "max|a-b|/max|b| of the output image; reference = the unmodified class from baseline/_ref on this GPU, same weights "
                      "(strict state_dict load), inputs and RNG seed; fp32 = cuDNN / cuBLAS TF32 switched off"}
        if strict_too:
            out["ours_strict_fp32_vs_reference_fp32"] = rel(got_strict, want)
            if full_call is not None:
                out["reference_fp32_gpu_ms_per_step"] = ref_fp32_ms
        del ref
        return out
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": f"{type(ex).__name__}: {str(ex)[:300]}"}


def _strict_fp32_forward(h: Harness, name, net, dev_in):
    """The same forward under the strict-fp32 contract (torch.backends.cudnn.allow_tf32 = False, what `parity.want` is computed
    with): this package's conv blocks switch to error-compensated 3xTF32 operands (ops.tf32_split, fmi_tf32_split3)."""
    from face_mask_inpaint_b200.graphs import CapturedForward
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            for _ in range(2):
                net(*dev_in)
            torch.cuda.synchronize()
            n0 = h.lib.fmi_kernel_launch_count()
            net(*dev_in)
            torch.cuda.synchronize()
            launches = h.lib.fmi_kernel_launch_count() - n0
            fwd = CapturedForward(net, *dev_in)
            ms = h.timed(fwd.replay)
            del fwd
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
    bsz = WORKLOADS[name]["batch"]
    return {"ms_per_step": ms, "value": h.world * bsz / (ms * 1e-3), "unit": "img/s", "kernel_launches_per_step": int(launches),
            "precision": "fp32 I/O; conv blocks with error-compensated 3xTF32 operands (x = hi + lo, three exact-product terms per "
                         "GEMM, fp32 accumulation) — selected by torch.backends.cudnn.allow_tf32 = False or FMI_PRECISION=tf32x3"}


def ours_picnet_ref(h: Harness):
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    torch.manual_seed(7)
    net = build_picnet_ref().eval().to(h.dev)
    with torch.no_grad():
        net.decoder.attn1.gamma.fill_(1.0)     # init 0 would switch the attention branch off (SURVEY §7 'random init hides bugs')
    src, ref, _, mask = make_inputs("picnet_ref", WORKLOADS["picnet_ref"]["batch"], 1000 + h.rank)
    parity = None
    if h.rank == 0:
        s_d, r_d, m_d = src[:2].to(h.dev), ref[:2].to(h.dev), mask[:2].to(h.dev)
        f_d = [t.to(h.dev) for t in (src, ref, mask)]
        parity = _parity(h, "picnet_ref", net, lambda m, is_ref: m(s_d, r_d, src_mask=m_d) if is_ref else m(s_d, r_d, m_d), "none", 1e-3,
                         strict_too=True, full_call=lambda m, is_ref: m(f_d[0], f_d[1], src_mask=f_d[2]))
    rec = _forward_record(h, "picnet_ref", net, [src, ref, mask], {}, {}, "tf32",
                          "ReferenceFill.forward (modules/model.py:81-112): 2 ResEncoders, ExampleGuidedAttention@32^2, ResGenerator "
                          "with Auto_Attn@128^2 up to 1024^2, AdaptiveAvgPool to 256^2 — every conv block, both attentions, the "
                          "compositing and the pooling on this package's sm_100a kernels")
    rec["precision"] = ("fp32 I/O; TF32 tensor-core operands where the reference's own GPU run has them (cuDNN allow_tf32 default) "
                        "and in the attention (hi/lo split logits), fp32 accumulation / softmax / InstanceNorm")
    rec["parity"] = parity
    rec["strict_fp32"] = _strict_fp32_forward(h, "picnet_ref", net, [t.to(h.dev) for t in (src, ref, mask)])
    return rec


def ours_refpsp(h: Harness):
    from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts
    prev = os.environ.get("FMI_PRECISION")
    os.environ["FMI_PRECISION"] = "bf16"
    try:
        torch.manual_seed(11)
        net = pSp(refpsp_opts(output_size=1024)).eval().to(h.dev)
        x, ref, _, mask = make_inputs("refpsp", WORKLOADS["refpsp"]["batch"], 2000 + h.rank)
        parity = None
        if h.rank == 0:
            x_d, r_d, m_d = x[:2].to(h.dev), ref[:2].to(h.dev), mask[:2].to(h.dev)
            parity = _parity(h, "refpsp", net, lambda m, is_ref: m(x_d, ref=r_d, src_mask=m_d, resize=True, randomize_noise=False),
                             "cuda", 2e-2)
        rec = _forward_record(h, "refpsp", net, [x], dict(ref=ref, src_mask=mask), dict(resize=True, randomize_noise=False), "bf16", "pSp.forward (modules/psp/psp.py:72-130): IR-SE50 GradualStyleEncoder on source + reference "
                              "(one 2N batch), attention1/2, masked blend, 18 map2style heads, StyleGAN2-1024 decoder, face_pool")
        rec["precision"] = "fp32 I/O; bf16 tensor-core operands (FMI_PRECISION=bf16), fp32 accumulation"
        rec["parity"] = parity
        return rec
    finally:
        if prev is None:
            os.environ.pop("FMI_PRECISION", None)
        else:
            os.environ["FMI_PRECISION"] = prev


# ------------------------------------------------------------------------------------------------------------ ours: training
def _patched_reference():
    """The reference's own training harness (GANOptimizer, pSpLoss, define_d, ReferenceFill / pSp classes) from the verbatim copy
    baseline/_ref with this package's drop-ins installed over it — what `python -m face_mask_inpaint_b200.run train_*.py` runs."""
    from baseline import reference as R
    if not R.available():
        raise FileNotFoundError("baseline/_ref is missing (made by __graft_entry__.build() where /root/reference exists)")
    from face_mask_inpaint_b200 import patch
    from face_mask_inpaint_b200.offline import stub_pretrained
    patch.install(str(R.REF))
    stub_pretrained()
    return R


def _train_record(h: Harness, name, step_dev, step_host, params, dtype, what, extra, timed_step=None):
    """`step_dev`: the eager step (launch counting, per-kernel event profile); `timed_step`: what `value` times when the step was
    captured into a CUDA graph (the same launches replayed), else the eager step itself."""
    peaks = measured_peaks()
    tensor_peak = peaks["bf16_tflops"] / (2.0 if dtype == "tf32" else 1.0)
    b = WORKLOADS[name]["batch"]
    for _ in range(2):
        step_dev()
    torch.cuda.synchronize()
    n0 = h.lib.fmi_kernel_launch_count()
    step_dev()
    torch.cuda.synchronize()
    launches = h.lib.fmi_kernel_launch_count() - n0
    sampler = ClockSampler(h.local_rank) if h.rank == 0 else None
    ms = h.timed(timed_step or step_dev)
    ms_e2e, h2d, d2h = step_host(h)
    clocks = sampler.stop() if sampler else None
    kernels, eager_ms = h.profile(step_dev, min(h.steps, 3), tensor_peak, peaks)
    nparam = sum(p.numel() for p in params)
    rec = {"metric": "images/sec", "value": h.world * b / (ms * 1e-3), "unit": "img/s", "n_gpus": h.world, "steps": h.steps,
           "warmup": h.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": dtype, "data": "synthetic", "config": workload_config(name, h.world), "clocks": clocks,
           "e2e": {"value": h.world * b / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "api": "pinned host batch -> device, the reference's own train-step calls over the "
                   "installed drop-ins, loss scalars -> host"},
           "gpu_launches": int(launches) * h.steps,
           "launch": (f"one CUDA-graph replay per train step (graphs.CapturedStep) capturing {int(launches)} sm_100a kernel launches of "
                      f"this package; eager step {eager_ms:.1f} ms" if timed_step is not None else
                      f"eager: {int(launches)} sm_100a kernel launches of this package per step"),
           "what": what, "trainable_params": nparam, "allreduce_bytes_per_step": nparam * 4 if h.world > 1 else 0,
           "roofline": roofline_of(kernels, peaks, tensor_peak, dtype, workload=name), "kernels": kernels}
    rec.update(extra)
    return rec


def ours_train_picnet(h: Harness):
    R = _patched_reference()
    from face_mask_inpaint_b200 import dist as fdist
    from modules.loss import GANOptimizer
    torch.manual_seed(21)
    G = R.reference_fill().to(h.dev)
    D = R.discriminator().to(h.dev)
    fdist.broadcast_module_state(G)
    fdist.broadcast_module_state(D)
    graph = os.environ.get("FMI_TRAIN_GRAPH", "1") != "0"
    optG, optD = (torch.optim.Adam(m.parameters(), lr=1e-5, capturable=graph) for m in (G, D))
    nb = 0
    if h.world > 1:
        rg = fdist.GradientAllReducer([p for p in G.parameters() if p.requires_grad]).attach(optG)
        rd = fdist.GradientAllReducer([p for p in D.parameters() if p.requires_grad]).attach(optD)
        nb = len(rg.buckets) + len(rd.buckets)
    gan = GANOptimizer(optD, optG).to(h.dev)
    G.train()
    D.train()
    torch.manual_seed(100 + h.rank)
    host = [t.pin_memory() for t in make_inputs("train_picnet", WORKLOADS["train_picnet"]["batch"], 3000 + h.rank)]
    dev = [t.to(h.dev) for t in host]
    losses_host = torch.empty(5, pin_memory=True)

    def run(src, ref, gt, mask):      # train_reference_fill.py:342-346
        gen = G(src, ref, src_mask=mask)
        return gan(D, src, gt, ref, gen, mask)

    def step_host(hh):
        def st():
            d = [t.to(hh.dev, non_blocking=True) for t in host]
            ls = run(*d)
            losses_host.copy_(torch.stack([l.detach().float().reshape(()) for l in ls]), non_blocking=True)
        ms = hh.timed(st)
        return ms, sum(t.numel() * t.element_size() for t in host), 20

    captured = None
    if graph:
        # the whole step (generator forward, losses, both backwards, both Adam steps[, the NCCL gradient all-reduces]) as ONE
        # CUDA graph (graphs.CapturedStep); a step that cannot be captured is timed eagerly and says so
        from face_mask_inpaint_b200.graphs import CapturedStep
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            captured = CapturedStep(run, *dev, modules=(G, D))
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"[bench] train_picnet: CUDA-graph capture failed ({type(ex).__name__}: {str(ex)[:200]}); eager step\n")
            captured = None
            torch.cuda.synchronize()

    def step_host_graph(hh):
        stat = captured.static_inputs

        def st():
            for dst, src in zip(stat, host):
                dst.copy_(src, non_blocking=True)
            ls = captured.replay()
            losses_host.copy_(torch.stack([l.detach().float().reshape(()) for l in ls]), non_blocking=True)
        ms = hh.timed(st)
        return ms, sum(t.numel() * t.element_size() for t in host), 20

    rec = _train_record(h, "train_picnet", lambda: run(*dev), step_host_graph if captured else step_host,
                        [p for m in (G, D) for p in m.parameters() if p.requires_grad], "tf32",
                        "train_reference_fill.py:342-346 — ReferenceFill forward (train mode) + GANOptimizer.__call__ "
                        "(modules/loss.py:120-134) over the installed drop-ins; conv blocks, attention, SpectralNorm forward and "
                        "backward on the sm_100a kernels",
                        {"buckets": nb}, timed_step=captured.replay if captured else None)
    if not torch.isfinite(losses_host).all():
        raise RuntimeError("train_picnet: non-finite losses")
    if h.world > 1:     # replicas must still agree after all those steps (each rank trained on its own batch)
        rec["replicas_in_sync"] = fdist.replicas_in_sync([p for m in (G, D) for p in m.parameters() if p.requires_grad])
        if not rec["replicas_in_sync"]:
            sys.stderr.write("[bench] train_picnet: replicas DIVERGED — the gradient all-reduce did not run in every step\n")
    return rec


def ours_train_psp(h: Harness):
    from argparse import Namespace
    prev = os.environ.get("FMI_PRECISION")
    os.environ["FMI_PRECISION"] = "bf16"
    try:
        R = _patched_reference()
        from face_mask_inpaint_b200 import dist as fdist
        from modules.psp.criteria import pSpLoss
        torch.manual_seed(31)
        G = R.psp(output_size=1024, use_attention=1, train_decoder=1).to(h.dev)
        G.latent_avg = G.latent_avg.to(h.dev)
        fdist.broadcast_module_state(G)
        params = list(G.encoder.parameters()) + list(G.decoder.parameters())      # train_psp.py:287-289
        opt = torch.optim.Adam(params, lr=1e-5)
        nb = 0
        if h.world > 1:
            red = fdist.GradientAllReducer([p for p in params if p.requires_grad]).attach(opt)
            nb = len(red.buckets)
        largs = Namespace(id_lambda=0, lpips_lambda=0.8, l2_lambda=1.0, style_lambda=250.0, lpips_lambda_ref=0, l2_lambda_ref=0,
                          cx_lambda=1.0, w_norm_lambda=0, start_from_latent_avg=1)
        loss_fn = pSpLoss(largs).to(h.dev)
        G.train()
        torch.manual_seed(200 + h.rank)
        host = [t.pin_memory() for t in make_inputs("train_psp", WORKLOADS["train_psp"]["batch"], 4000 + h.rank)]
        dev = [t.to(h.dev) for t in host]
        loss_host = torch.empty(1, pin_memory=True)

        def run(src, ref, gt, mask):      # train_psp.py:307-335
            gen, latent = G(src, ref=ref, src_mask=mask, return_latents=True, randomize_noise=1)
            loss, _, _ = loss_fn(src, gt, gen, latent, latent_avg=G.latent_avg, ref=ref, mask=mask)
            if fdist.all_ranks_finite(loss):
                opt.zero_grad()
                loss.backward()
                opt.step()
            return loss

        def step_host(hh):
            def st():
                d = [t.to(hh.dev, non_blocking=True) for t in host]
                loss_host.copy_(run(*d).detach().float().reshape(1), non_blocking=True)
            ms = hh.timed(st)
            return ms, sum(t.numel() * t.element_size() for t in host), 4

        rec = _train_record(h, "train_psp", lambda: run(*dev), step_host, [p for p in params if p.requires_grad], "bf16",
                            "train_psp.py:307-335 — pSp forward (train mode, encoder + decoder trainable), pSpLoss, backward, Adam over "
                            "the installed drop-ins; StyleGAN2 decoder + attention forward/backward on the sm_100a kernels",
                            {"buckets": nb})
        if not torch.isfinite(loss_host).all():
            raise RuntimeError("train_psp: non-finite loss")
        if h.world > 1:
            rec["replicas_in_sync"] = fdist.replicas_in_sync([p for p in params if p.requires_grad])
            if not rec["replicas_in_sync"]:
                sys.stderr.write("[bench] train_psp: replicas DIVERGED — the gradient all-reduce did not run in every step\n")
        return rec
    finally:
        if prev is None:
            os.environ.pop("FMI_PRECISION", None)
        else:
            os.environ["FMI_PRECISION"] = prev


OURS = {"picnet_ref": ours_picnet_ref, "refpsp": ours_refpsp, "train_picnet": ours_train_picnet, "train_psp": ours_train_psp}


# ------------------------------------------------------------------------------------------------------------ the unmodified reference
def _reference_step(name, device, batch):
    """(step function, description) running the UNMODIFIED reference for `name` on `device` with `batch` images per step."""
    from baseline import reference as R
    cuda = device.type == "cuda"
    R.import_unpatched("none" if "picnet" in name else ("cuda" if cuda else "cpu"))
    src, ref, gt, mask = (t.to(device) for t in make_inputs(name, batch, 1000))
    if name == "picnet_ref":
        torch.manual_seed(7)
        net = R.reference_fill().eval().to(device)
        with torch.no_grad():
            net.decoder.attn1.gamma.fill_(1.0)

        def step():
            with torch.no_grad():
                return net(src, ref, src_mask=mask)
        return step, "modules/model.py ReferenceFill.forward, unmodified", "reference"
    if name == "refpsp":
        torch.manual_seed(11)
        net = R.psp(output_size=1024).eval().to(device)
        net.latent_avg = net.latent_avg.to(device)

        def step():
            with torch.no_grad():
                return net(src, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
        ops = "its own compiled CUDA ops (oracle/_ref)" if cuda else \
            "upfirdn2d = its own upfirdn2d_native, fused_leaky_relu restated (the reference has no CPU path for these two ops)"
        return step, f"modules/psp/psp.py pSp.forward, unmodified, fp32; {ops}", "reference" if cuda else "port"
    if name == "train_picnet":
        from modules.loss import GANOptimizer
        torch.manual_seed(21)
        G, D = R.reference_fill().to(device), R.discriminator().to(device)
        optG, optD = torch.optim.Adam(G.parameters(), lr=1e-5), torch.optim.Adam(D.parameters(), lr=1e-5)
        gan = GANOptimizer(optD, optG).to(device)
        G.train()
        D.train()

        def step():
            gen = G(src, ref, src_mask=mask)
            return gan(D, src, gt, ref, gen, mask)
        return step, "train_reference_fill.py:342-346 with the unmodified modules, random-init VGG16", "reference"
    if name == "train_psp":
        if not cuda:
            raise RuntimeError("the reference's LPIPS hard-codes .to('cuda') (modules/psp/criteria/lpips/lpips.py:24-27): no CPU path")
        from argparse import Namespace
        from modules.psp.criteria import pSpLoss
        torch.manual_seed(31)
        G = R.psp(output_size=1024, use_attention=1, train_decoder=1).to(device)
        G.latent_avg = G.latent_avg.to(device)
        params = list(G.encoder.parameters()) + list(G.decoder.parameters())
        opt = torch.optim.Adam(params, lr=1e-5)
        loss_fn = pSpLoss(Namespace(id_lambda=0, lpips_lambda=0.8, l2_lambda=1.0, style_lambda=250.0, lpips_lambda_ref=0,
                                    l2_lambda_ref=0, cx_lambda=1.0, w_norm_lambda=0, start_from_latent_avg=1)).to(device)
        G.train()

        def step():
            gen, latent = G(src, ref=ref, src_mask=mask, return_latents=True, randomize_noise=1)
            loss, _, _ = loss_fn(src, gt, gen, latent, latent_avg=G.latent_avg, ref=ref, mask=mask)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        return step, "train_psp.py:307-335 with the unmodified modules (fp32: the reference's CUDA ops reject bf16)", "reference"
    raise ValueError(name)


def _oracle_port_cpu(args, cores):
    """Fallback of the reference arm when baseline/_ref is absent: oracle/ref_ops.py (the CPU restatement of the reference's
    algorithm) on the pieces of one PICNet-ref image that dominate its forward — ExampleGuidedAttention @32^2, Auto_Attn C=256
    @128^2 (58 % of the generator's FLOPs), the five ResBlockDecoders and the Output block; the encoders are left out."""
    from oracle import ref_ops as O
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    src, ref, m = r(1, 128, 32, 32), r(1, 128, 32, 32), torch.rand(1, 1, 32, 32, generator=g)
    chans = [(256, 256, 256), (256, 256, 256), (256, 128, 128), (128, 64, 64), (64, 32, 32)]

    def step():
        with torch.no_grad():
            x = O.example_guided_attention(m, src, ref, r(32, 128, 1, 1) * 0.1)
            for i, (ci, ch, co) in enumerate(chans):
                x = O.res_block_decoder(x, r(ch, ci, 3, 3) / (3 * ci ** 0.5), None, r(ch, co, 3, 3) / (3 * ch ** 0.5), None,
                                        r(ci, co, 3, 3) / (3 * ci ** 0.5), None)
                if i == 1:
                    x = O.auto_attn(x, r(64, 256, 1, 1) * 0.02, torch.zeros(64), torch.ones(1))[0]
            return O.output_block(x, r(3, 32, 3, 3) * 0.05, torch.zeros(3))

    step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    return 1.0 / (ms * 1e-3), ms, (f"{args.steps} steps x 1 image: oracle/ref_ops.py restatement of EGA@32^2 + Auto_Attn@128^2 + 5 "
                                   f"ResBlockDecoder + Output (encoders left out; baseline/_ref absent), PyTorch CPU fp32, {cores} threads")


def run_reference_cpu(args, rank):
    """The reference arm: the unmodified reference on the box's host cores, all threads, same config / metric / unit. Each step
    is the full per-GPU batch when K + W steps fit in ~3 minutes, otherwise a bounded sample of it (fewer images per step)."""
    if rank != 0:
        return
    from baseline import reference as R
    name = args.workload
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    base = {"impl": "reference", "metric": "images/sec", "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(name, args.gpus)}
    if not R.available():
        # no copy of the reference travelled to this box: the oracle's restatement of the generator pieces stands in (kind "port")
        value, ms, sample = _oracle_port_cpu(args, cores)
        base.update({"value": value, "ms_per_step": ms,
                     "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
                     "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base), file=sys.__stdout__, flush=True)
        return
    b_full = WORKLOADS[name]["batch"]
    try:
        step, what, kind = _reference_step(name, torch.device("cpu"), b_full)
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter() - t0
        b = b_full
        total = args.steps + args.warmup
        if t1 * total > 200.0 and b_full > 1:
            b = max(1, int(b_full * 200.0 / (t1 * total)))
            step, what, kind = _reference_step(name, torch.device("cpu"), b)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        ms = (time.perf_counter() - t0) / args.steps * 1e3
    except Exception as ex:  # noqa: BLE001
        base.update({"unavailable": f"{type(ex).__name__}: {str(ex)[:300]}"})
        print(json.dumps(base), file=sys.__stdout__, flush=True)
        return
    value = b / (ms * 1e-3)
    sample = (f"{args.steps} steps x {b} of the {b_full} images of a per-GPU batch; {what}; PyTorch CPU fp32, {cores} threads")
    base.update({"value": value, "ms_per_step": ms,
                 "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base), file=sys.__stdout__, flush=True)


def run_reference_gpu(args):
    """The unmodified reference on the same B200 (single GPU, its own eager formulation) — the number to beat."""
    name = args.workload
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    b = WORKLOADS[name]["batch"]
    out = {"impl": "gpu_reference", "workload": name, "per_gpu_batch": b}
    try:
        step, what, kind = _reference_step(name, dev, b)
        for _ in range(max(3, args.warmup)):
            step()
        torch.cuda.synchronize()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / args.steps
        out.update({"value": b / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms, "steps": args.steps, "what": what,
                    "precision": "fp32 tensors, PyTorch defaults (cuDNN convolutions may use TF32, matmul / bmm strict fp32)",
                    "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30})
    except Exception as ex:  # noqa: BLE001
        out["unavailable"] = f"{type(ex).__name__}: {str(ex)[:300]}"
    print(json.dumps(out), file=sys.__stdout__, flush=True)


def _subprocess_json(argv, timeout):
    """Run bench.py in a fresh process (the unmodified reference cannot share a process with the installed drop-ins) and
    return its last JSON line."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT",
                                                             "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "TORCHELASTIC_RUN_ID")}
    try:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py")] + argv, capture_output=True, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout} s"}
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if not lines:
        return {"unavailable": f"rc {r.returncode}: {r.stderr[-300:]}"}
    return json.loads(lines[-1])


# ------------------------------------------------------------------------------------------------------------ main
def run_ours(args):
    h = Harness(args)
    name = args.workload
    names = [name] + ([n for n in ("refpsp", "train_picnet", "train_psp") if n != name] if (name == "picnet_ref" and not args.no_extras) else [])
    records = []
    for n in names:
        try:
            rec = OURS[n](h)
        except Exception as ex:  # noqa: BLE001
            if n == name:
                raise
            rec = {"config": workload_config(n, h.world), "error": f"{type(ex).__name__}: {str(ex)[:400]}"}
        torch.cuda.empty_cache()
        # the unmodified reference on the same GPU / on the host cores, rank 0 only, in fresh processes; the other ranks wait
        if h.rank == 0 and h.world == 1 and not args.no_reference:
            k = min(args.steps, 10)
            rec["gpu_reference"] = _subprocess_json(["--impl", "gpu_reference", "--workload", n, "--steps", str(k), "--warmup", "3"], 900)
            if n in ("picnet_ref", "refpsp"):
                cb = _subprocess_json(["--impl", "reference", "--workload", n, "--steps", "2", "--warmup", "1"], 900)
                rec["cpu_baseline"] = cb.get("cpu_baseline", {"unavailable": cb.get("unavailable", "?")})
        h.barrier()
        records.append(rec)
    if h.world > 1:
        h.dist.destroy_process_group()
    if h.rank != 0:
        return
    line = records[0]
    if len(records) > 1:
        line["records"] = records[1:]
    print(json.dumps(line), file=sys.__stdout__, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu_reference"])
    ap.add_argument("--workload", default="picnet_ref", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="only the named workload (no refpsp / train records)")
    ap.add_argument("--no-reference", action="store_true", help="skip the gpu_reference / cpu_baseline sub-runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    sys.stdout = sys.stderr      # library / reference chatter (constructors print) goes to stderr: stdout carries the JSON line only
    if args.impl == "reference":
        run_reference_cpu(args, rank)
    elif args.impl == "gpu_reference":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
