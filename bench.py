#!/usr/bin/env python
"""bench.py -- raw-pixel encode/decode GB/s of the llcomp hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     the reference's own CPU code on the host cores

Workload (BASELINE.json configs[3]): a batch of 1024 x 1024x1024 RGB8 synthetic images (smooth gradient +
uniform +-4 noise), ONE slice per image, so that every stream is byte-identical to the reference's llcompc
output and bits/pixel cost of slicing is 0.  One "step" = one pass of the encoder over the whole batch.
With N GPUs every rank codes its own 1024-image shard (independent slices, no data-path collective):
weak scaling, value = total raw bytes of all ranks / max-over-ranks device time.

One JSON line on stdout (rank 0).  `value`: encode, inputs resident in HBM.  `e2e`: the same through the
C-ABI host-buffer call (pinned host pixels in, pinned host streams out, copies inside the timed region).
`decode`: the same two numbers for the decoder.  `roofline`: the kernel that dominates the step;
`kernels`: every kernel of the step with its share, algorithmic bytes and HBM fraction.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "raw-pixel encode GB/s (decode GB/s and bits/pixel reported alongside)"
UNIT = "GB/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._p = index, [], None

    def start(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                        "--format=csv,noheader,nounits", "-lms", "200"],
                                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self._p = None

    def _read(self):
        for line in self._p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self._p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self._p.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_batch(torch, n, w, h, c, noise, seed, device):
    """Gradient + uniform noise of SURVEY.md appendix C's shape, drawn with torch's RNG on `device`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = (torch.arange(w, device=device, dtype=torch.int32) * 255 // w).view(1, 1, w, 1)
    y = (torch.arange(h, device=device, dtype=torch.int32) * 255 // h).view(1, h, 1, 1)
    ch = (torch.arange(c, device=device, dtype=torch.int32) * 10).view(1, 1, 1, c)
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=device)
    step = max(1, min(n, 64))
    for i in range(0, n, step):
        m = min(step, n - i)
        if noise < 0:                      # high-entropy variant: uniform random bytes (BASELINE configs[2])
            out[i:i + m] = torch.randint(0, 256, (m, h, w, c), generator=g, device=device, dtype=torch.int32).to(torch.uint8)
            continue
        v = (x + y) // 2 + ch
        if noise > 0:
            v = v + torch.randint(-noise, noise + 1, (m, h, w, c), generator=g, device=device, dtype=torch.int32)
        else:
            v = v.expand(m, h, w, c)
        out[i:i + m] = v.clamp(0, 255).to(torch.uint8)
    return out


def cpu_reference_encode(images_np, threads):
    """Times the reference's own compressImage/decompressImage on `threads` host threads.  Uses oracle/_ref
    (the unmodified header) when it was built, else the C port."""
    import ctypes as C

    import numpy as np

    import oracle
    n, h, w, c = images_np.shape
    kind = "reference" if oracle.have_ref() else "port"
    if kind == "reference":
        fn_e, fn_d = oracle.ref().ref_compress_batch_mt, oracle.ref().ref_decompress_batch_mt
    else:
        fn_e, fn_d = oracle.lib().llo_compress_batch_mt, oracle.lib().llo_decompress_batch_mt
    t0 = time.perf_counter()
    total = int(fn_e(images_np.ctypes.data, n, w, h, c, threads))
    t_enc = time.perf_counter() - t0
    # streams for the decode leg and the bits/pixel check (restatement == reference byte for byte)
    streams = [oracle.compress(images_np[k]) for k in range(min(n, 4))]
    sizes = [len(s) for s in streams]
    reps = (n + len(streams) - 1) // len(streams)
    blob = np.frombuffer(b"".join(streams * reps), dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([sizes[k % len(sizes)] for k in range(n)])
    t0 = time.perf_counter()
    got = int(fn_d(blob.ctypes.data, offs.ctypes.data, n, threads))
    t_dec = time.perf_counter() - t0
    assert got == images_np.size, "CPU decode failed"
    return {"kind": kind, "encode_s": t_enc, "decode_s": t_dec, "stream_bytes": total, "first_sizes": sizes}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1024, help="images per GPU (1024 = BASELINE configs[3])")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--noise", type=int, default=4)
    ap.add_argument("--tile", type=int, default=0, help="tile edge; 0 = one slice per image")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="tuning only: time the encoder alone")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = H = args.size
    C, n_img = args.channels, args.images
    cores = os.cpu_count() or 1
    content = "uniform random bytes" if args.noise < 0 else f"gradient + uniform noise +-{args.noise}"
    workload = (f"{'configs[3]: ' if (n_img, W, C, args.tile, args.noise) == (1024, 1024, 3, 0, 4) else ''}batch of {n_img} x {W}x{H} "
                f"RGB{8 if C == 3 else ''} (C={C}) per GPU, {content}, "
                + ("1 slice per image" if not args.tile else f"{args.tile}^2 tiles"))
    config = {"workload": workload, "images_per_gpu": n_img, "width": W, "height": H, "channels": C,
              "noise": args.noise, "tile": args.tile or None, "slices_per_gpu": None, "sharding": f"dp{world}",
              "l2": "inputs (>= 3 GB per step) are larger than the 126 MB L2"}

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return
        import numpy as np
        import torch
        sample = args.cpu_sample or max(8, min(64, 2 * cores))
        imgs = synth_batch(torch, sample, W, H, C, args.noise, 1234, "cpu").numpy()
        times, res = [], None
        for i in range(args.warmup + args.steps):
            res = cpu_reference_encode(imgs, cores)
            if i >= args.warmup:
                times.append(res["encode_s"])
        ms = 1e3 * sum(times) / len(times)
        val = imgs.size / (ms / 1e3) / 1e9
        line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
                "config": dict(config, sample=f"{sample} of the {n_img} images per step"),
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": res["kind"],
                                 "sample": f"{sample} images of {W}x{H}x{C}, all {cores} host threads, one image per "
                                           "thread at a time (the reference has no internal threading)"},
                "decode": {"value": imgs.size / res["decode_s"] / 1e9, "unit": UNIT},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch

    import llcomp_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; llcomp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    codec = llcomp_b200.Codec(local_rank)
    g = codec.geometry(W, H, C, args.tile, args.tile, n_img)
    n_slices = codec.slice_count(g)
    config["slices_per_gpu"] = n_slices
    raw = n_img * W * H * C
    px = synth_batch(torch, n_img, W, H, C, args.noise, 1234 + 7919 * rank, dev)
    payload = torch.empty(codec.payload_capacity(g), dtype=torch.uint8, device=dev)
    offsets = torch.empty(n_slices + 1, dtype=torch.int64, device=dev)
    out_px = torch.empty_like(px)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profiled=False):
        """warmup, then EXACTLY `steps` calls between barrier+sync; device time from CUDA events, max over ranks."""
        for _ in range(warmup):
            fn()
        codec.finish()
        stage = {}
        codec.set_profiling(profiled)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = codec.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
            if profiled:
                for k, v in codec.stage_times().items():
                    stage[k] = stage.get(k, 0.0) + v
        e1.record()
        barrier()
        codec.finish()
        codec.set_profiling(False)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, {k: v / steps for k, v in stage.items()}, codec.launch_count() - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    enc_ms, enc_stage, enc_launches = timed(lambda: codec.encode_device(px, g, payload, offsets), args.steps,
                                            args.warmup, profiled=True)
    if args.no_decode:
        dec_ms, dec_stage, dec_launches = float("nan"), {}, 0
        out_px.copy_(px)
    else:
        dec_ms, dec_stage, dec_launches = timed(lambda: codec.decode_device(payload, offsets, g, out_px), args.steps,
                                                args.warmup, profiled=True)
    clocks = sampler.stop() if rank == 0 else None
    n_bins = codec.last_bin_count()
    ok = bool(torch.equal(out_px, px))
    off_host = offsets.cpu().numpy()
    stream_bytes = int(off_host[-1])
    if dist is not None:
        t = torch.tensor([stream_bytes, int(ok)], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_stream, ok = int(t[0].item()), int(t[1].item()) == world
    else:
        total_stream = stream_bytes
    hdr = 6 if n_slices == n_img else 24 + 4 * (n_slices // n_img)
    bpp = 8.0 * (total_stream + hdr * n_img * world) / (world * n_img * W * H)

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = dec_e2e = None
    if not args.no_e2e:
        h_px = torch.empty((n_img, H, W, C), dtype=torch.uint8, pin_memory=True)
        h_px.copy_(px)
        out_cap = (raw if args.noise >= 0 else 2 * raw) + 384 * n_slices + hdr * n_img
        h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
        h_off = torch.zeros(n_img + 1, dtype=torch.int64, pin_memory=True)
        h_back = torch.empty((n_img, H, W, C), dtype=torch.uint8, pin_memory=True)

        def e2e_timed(fn):
            for _ in range(max(1, args.warmup // 2)):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            if dist is not None:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        dt = e2e_timed(lambda: codec.encode_batch_ptr(h_px.data_ptr(), g, h_out.data_ptr(), out_cap, h_off.data_ptr()))
        e2e_stream = int(h_off[n_img].item())
        e2e = {"value": world * raw / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": raw,
               "d2h_bytes_per_step": e2e_stream + 8 * (n_slices + 1), "ms_per_step": dt * 1e3,
               "api": "llcomp_b200_encode_batch (host pixels -> host streams)"}
        # what the copies alone cost on this box (the pipelined encode cannot start its last group before the
        # upload is through, and a slice then still needs its full serial coding time)
        def copy_ms(dst, src):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            a.record()
            dst.copy_(src, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)
        e2e["h2d_alone_ms"] = copy_ms(px, h_px)
        k = min(e2e_stream, raw, payload.numel())
        e2e["d2h_alone_ms"] = copy_ms(h_back.view(-1)[:k], payload[:k]) * (e2e_stream / k)
        dt = e2e_timed(lambda: codec.decode_batch_ptr(h_out.data_ptr(), h_off.data_ptr(), n_img, h_back.data_ptr(), raw))
        ok = ok and bool(torch.equal(h_back, h_px))
        dec_e2e = {"value": world * raw / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": e2e_stream + 8 * (n_slices + 1),
                   "d2h_bytes_per_step": raw, "ms_per_step": dt * 1e3,
                   "api": "llcomp_b200_decode_batch (host streams -> host pixels)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel rooflines (HBM; nothing here is a dense contraction)
    peak, peak_src = peaks()
    n_samples = raw
    alg = {  # algorithmic bytes per launch = per-sample figure of SURVEY.md 8(d) x samples of one launch
        "frontend": 5 * n_samples,                          # 1 B pixel read + 4 B record written
        "slice_coder": 4 * n_samples + stream_bytes,        # fused coder: records read + payload written to scratch
        "model_pass": 4 * n_samples + 2 * n_bins,           # (split path only) records read + queue entries written
        "scan": 4 * n_slices + 8 * (n_slices + 1),
        "compact": 2 * stream_bytes,                        # scratch read + contiguous stream written
        "slice_decoder": stream_bytes + n_samples,          # payload read + pixels written
    }
    kernels = []
    for name, ms in list(enc_stage.items()) + list(dec_stage.items()):
        if ms <= 0:
            continue
        step_ms = dec_ms if name == "slice_decoder" else enc_ms
        ach = alg[name] / (ms / 1e3) / 1e9
        kernels.append({"name": name, "ms": ms, "share_of_step": ms / step_ms, "algorithmic_bytes": alg[name],
                        "achieved_GBps": ach, "frac_of_hbm_peak": ach / peak})
    dom = max((k for k in kernels if k["name"] != "slice_decoder"), key=lambda k: k["ms"])
    # DRAM bytes per launch from the ncu --set full capture of this exact workload (profiles/r01_v9_ncu_encode_summary.json:
    # dram__bytes_read.sum + dram__bytes_write.sum); only quoted when the run IS that workload.
    ncu_traffic = {"slice_coder": 13.659265e9 + 3.194658e9, "frontend": 3.222452e9 + 12.827327e9}
    is_profiled_workload = (n_img, W, H, C, args.tile, args.noise) == (1024, 1024, 1024, 3, 0, 4)
    for k in kernels:
        k["ncu_dram_bytes"] = ncu_traffic.get(k["name"]) if is_profiled_workload else None
    roofline = {"kernel": dom["name"], "bound": "hbm", "achieved": dom["achieved_GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac_of_hbm_peak"], "traffic": dom["ncu_dram_bytes"], "peak_source": peak_src,
                "note": "slice_coder holds one serial dependency chain per slice (issue/latency-bound, not HBM-bound); "
                        "the HBM-bound kernel of the path is `frontend`, listed under kernels[]",
                # fractions against the 8 TB/s spec as well, and the whole encode as raw + stream bytes
                "frac_of_spec_8000GBps": dom["achieved_GBps"] / 8000.0,
                "pipeline": {"bytes": raw + stream_bytes, "achieved": (raw + stream_bytes) / (enc_ms / 1e3) / 1e9,
                             "frac": (raw + stream_bytes) / (enc_ms / 1e3) / 1e9 / peak}}

    line = {"metric": METRIC, "value": world * raw / (enc_ms / 1e3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": enc_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic", "config": config,
            "impl": "ours", "round_trip_exact": ok, "bits_per_pixel": bpp, "bins_per_sample": (n_bins / raw) if n_bins else None,
            "decode": {"value": world * raw / (dec_ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": dec_ms, "e2e": dec_e2e},
            "e2e": e2e, "gpu_launches": enc_launches + dec_launches,
            "gpu_launches_detail": {"encode_steps": enc_launches, "decode_steps": dec_launches},
            "roofline": roofline, "kernels": kernels, "clocks": clocks}

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same images
    if world == 1 and not args.no_cpu:
        sample = args.cpu_sample or max(8, min(64, 2 * cores))
        sample = min(sample, n_img)
        imgs = px[:sample].cpu().numpy()
        res = cpu_reference_encode(imgs, cores)
        gpu_sizes = [int(off_host[(k + 1) * (n_slices // n_img)] - off_host[k * (n_slices // n_img)]) + hdr
                     for k in range(len(res["first_sizes"]))]
        line["cpu_baseline"] = {"value": imgs.size / res["encode_s"] / 1e9, "unit": UNIT, "cores": cores,
                                "kind": res["kind"],
                                "sample": f"first {sample} images of the batch, all {cores} host threads, whole images "
                                          "per thread", "decode_value": imgs.size / res["decode_s"] / 1e9,
                                "bits_per_pixel": 8.0 * res["stream_bytes"] / (sample * W * H)}
        line["bpp_vs_reference"] = {"gpu_stream_bytes": gpu_sizes, "cpu_stream_bytes": res["first_sizes"],
                                    "identical": gpu_sizes == res["first_sizes"] if not args.tile else None}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
