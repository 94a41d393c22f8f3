#!/usr/bin/env python
"""bench.py -- headline benchmark of the LZ77 hot path (BASELINE.json metric:
"encode+decode GB/s at 1/2/4/8 B200; ratio; CPU ref GB/s same run").

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

Workload (N = 1 and per rank for N > 1, weak scaling): BASELINE.json configs[1],
256 MiB of seeded synthetic Zipf text, -s 4095 -l 15.  One step = one encode of
the whole input followed by one decode of the stream that encode produced.

  value   uncompressed GB/s of the encode+decode roundtrip, all ranks together,
          inputs and outputs resident in HBM (device entry points of the C ABI),
          timed with CUDA events on the stream the kernels run on, max over ranks
  e2e     the same roundtrip through the host-buffer entry points of the C ABI
          (lz77_gpu_encode / lz77_gpu_decode, what the reference-shaped
          encode()/decode() wrappers call) from pinned host buffers: H2D of the
          input and D2H of the result inside the timed region, every step
  roofline      the dominant kernel (longest-match search) against measured HBM
  roofline_decode  the decode match-copy kernel (the north_star's 70 % target)
  cpu_baseline  the compiled reference (oracle/_ref/lz77) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(name="configs[1]: 256 MiB synthetic Zipf text, -s 4095 -l 15",
                kind="zipf_text", n=256 << 20, sb=4095, la=15, seed=1234)
METRIC = "encode+decode GB/s"
UNIT = "GB/s"


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #

def csrc_sha16() -> str:
    """Hash of the kernel sources: the ncu figures in profiles/traffic.json are only
    reported while they describe the kernels that are actually running."""
    import hashlib
    h = hashlib.sha256()
    # the files that hold the codec kernels (capi.cu / comm.cu / context.cuh are host code)
    for name in ("common.cuh", "kernels.cuh", "match.cuh", "search_bucket.cu", "search_bigwin.cu",
                 "encode.cu", "decode.cu", "decode_jump.cu"):
        f = ROOT / "lz77_b200" / "csrc" / name
        h.update(name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_figure(key: str, n_bytes: int):
    """A per-launch figure (DRAM bytes, executed warp instructions) of the committed ncu
    capture, profiles/traffic.json: valid only for the workload size AND the kernel
    sources (csrc_sha16) it was taken on -- otherwise None (stale numbers are not reported)."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        t = json.loads(p.read_text())
        if int(t.get("workload_bytes", -1)) == int(n_bytes) and t.get("csrc_sha16") == csrc_sha16():
            return int(t[key])
    except Exception:
        pass
    return None


def issue_bound(kernel: str, n_bytes: int, tokens: int, measured_ms: float, clocks: dict):
    """The second bound of a kernel whose tokens are a few bytes each: instruction issue.
    warp instructions per launch (ncu smsp__inst_executed.sum, profiles/traffic.json) over
    148 SMs x 4 schedulers x the SM clock sampled during the run = the time the kernel
    needs if every scheduler issued every cycle."""
    inst = ncu_figure(kernel + ".inst_executed", n_bytes)
    mhz = (clocks or {}).get("sm_mhz")
    if not inst or not mhz:
        return None
    slots_per_ms = 148 * 4 * mhz * 1e3
    min_ms = inst / slots_per_ms
    return {"warp_inst_per_launch": inst, "warp_inst_per_token": inst / max(tokens, 1),
            "issue_slots_per_ms": slots_per_ms, "min_ms": min_ms,
            "frac": min_ms / measured_ms if measured_ms else None}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_near_gpu(local_rank: int) -> dict:
    """One process per GPU: run on the CPUs next to this rank's GPU, so that the pinned
    host buffers of the end-to-end leg are allocated on the GPU's NUMA node (with several
    ranks on one socket's memory the H2D / D2H copies share its bandwidth).  Best effort:
    any failure leaves the affinity as it was.  Returns what happened (on a box that
    reports every GPU next to the same CPUs the binding changes nothing)."""
    before = len(os.sched_getaffinity(0))
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        after = len(os.sched_getaffinity(0))
        return {"cpus_before": before, "cpus_after": after, "changed": after != before}
    except Exception as e:  # noqa: BLE001
        print(f"[bench] no NUMA binding for rank {local_rank}: {e}", file=sys.stderr)
        return {"cpus_before": before, "cpus_after": before, "changed": False, "error": str(e)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
# CPU reference arm
# --------------------------------------------------------------------------- #

def _ref_roundtrip(args):
    """Encode + decode one slice with the compiled reference CLI (or the oracle
    port when the binary did not travel); returns (bytes, compressed bytes)."""
    path, use_binary = args
    if use_binary:
        exe = str(ROOT / "oracle" / "_ref" / "lz77")
        subprocess.run([exe, "-c", "-i", path, "-o", path + ".lz"], check=True,
                       capture_output=True)
        subprocess.run([exe, "-d", "-i", path + ".lz", "-o", path + ".out"], check=True,
                       capture_output=True)
        n, c = os.path.getsize(path), os.path.getsize(path + ".lz")
        ok = open(path, "rb").read() == open(path + ".out", "rb").read()
    else:
        from oracle import oracle
        orc = oracle()
        data = open(path, "rb").read()
        enc = orc.ref_encode(data, WORKLOAD["sb"], WORKLOAD["la"])
        ok = orc.decode(enc) == data
        n, c = len(data), len(enc)
    if not ok:
        raise RuntimeError("CPU reference roundtrip mismatch")
    return n, c


class CpuReference:
    """The reference's own CPU implementation of the path on all host cores:
    one independent single-threaded process per core, each on its own slice of
    the workload (the reference has no threading of its own)."""

    def __init__(self, slice_bytes: int):
        from oracle import build_oracle, ref_binary
        from lz77_b200 import synth
        build_oracle()
        self.use_binary = ref_binary() is not None
        self.kind = "reference" if self.use_binary else "port"
        self.cores = os.cpu_count() or 1
        self.slice_bytes = slice_bytes
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.dir = tempfile.TemporaryDirectory(dir=tmp)
        data = synth.make(WORKLOAD["kind"], slice_bytes * self.cores, seed=WORKLOAD["seed"],
                          device="cpu").numpy()
        self.paths = []
        for i in range(self.cores):
            p = os.path.join(self.dir.name, f"slice{i}")
            data[i * slice_bytes:(i + 1) * slice_bytes].tofile(p)
            self.paths.append(p)

    def step(self):
        """one bounded sample: every core encodes + decodes its slice; returns
        (seconds, bytes, compressed bytes)"""
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        if self.use_binary:
            with ThreadPoolExecutor(self.cores) as ex:   # threads only wait on subprocesses
                res = list(ex.map(_ref_roundtrip, [(p, True) for p in self.paths]))
        else:
            from concurrent.futures import ProcessPoolExecutor
            with ProcessPoolExecutor(self.cores) as ex:
                res = list(ex.map(_ref_roundtrip, [(p, False) for p in self.paths]))
        dt = time.perf_counter() - t0
        return dt, sum(r[0] for r in res), sum(r[1] for r in res)

    def sample_desc(self):
        return (f"{self.cores} x {self.slice_bytes >> 20} MiB slices of the Zipf-text workload, "
                f"one single-threaded {'oracle/_ref/lz77' if self.use_binary else 'oracle port'} "
                f"process per core, encode+decode, files on tmpfs")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(slice_bytes=args.ref_slice_mib << 20)
    for _ in range(args.warmup):
        ref.step()
    total_t, total_b, total_c = 0.0, 0, 0
    for _ in range(args.steps):
        dt, b, c = ref.step()
        total_t += dt
        total_b += b
        total_c += c
    value = total_b / total_t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "sb": WORKLOAD["sb"], "la": WORKLOAD["la"],
                   "step": ref.sample_desc()},
        "ratio": total_b / total_c,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                         "sample": ref.sample_desc()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# this framework
# --------------------------------------------------------------------------- #

def run_cli(src_tensor, sb: int, la: int, gpus: int):
    """The drop-in surface (main.c:141-169): `lz77 -c` / `lz77 -d` file to file on tmpfs
    for the bench workload.  Two figures each way: the whole process (CUDA context
    creation and pinned allocations included) and the codec loop alone as the program
    reports it with -v (read + GPU + write overlapped)."""
    import re
    exe = ROOT / "lz77_b200" / "bin" / "lz77"
    if not exe.exists():
        return {"unavailable": "lz77_b200/bin/lz77 is not built"}
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else None
    rec = {"gpus": gpus, "files": "tmpfs"}
    with tempfile.TemporaryDirectory(dir=tmp) as d:
        fin, flz, fout = (os.path.join(d, x) for x in ("in.bin", "out.lz", "back.bin"))
        src_tensor.cpu().numpy().tofile(fin)
        n = os.path.getsize(fin)
        # (several GPUs: one piece = the whole file, so that every GPU gets 1/n of it per call)
        extra = ["-G", str(gpus), "-p", str(max(32, n >> 20))] if gpus > 1 else []
        for mode, a, b, key in (("-c", fin, flz, "encode"), ("-d", flz, fout, "decode")):
            best_wall, steady = None, None
            # (the second run has the files in the page cache; several GPUs: one run -- the
            # start-up of n contexts and the communicator dominates either way)
            for _ in range(2 if gpus == 1 else 1):
                t0 = time.perf_counter()
                r = subprocess.run([str(exe), mode, "-i", a, "-o", b, "-s", str(sb), "-l", str(la),
                                    "-v", *extra], capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    return {"unavailable": f"lz77 {mode} failed: {r.stderr.strip()[:200]}"}
                m = re.search(r"in ([0-9.]+) s \(([0-9.]+) GB/s", r.stderr)
                if best_wall is None or dt < best_wall:
                    best_wall, steady = dt, (float(m.group(2)) if m else None)
            rec[key + "_wall_s"] = best_wall
            rec[key + "_gbs_process"] = n / best_wall / 1e9
            rec[key + "_gbs_codec_loop"] = steady
        # the same with the output thrown away: what the file system's write path costs
        for mode, a, key in ((("-c", fin, "encode"), ("-d", flz, "decode")) if gpus == 1 else ()):
            r = subprocess.run([str(exe), mode, "-i", a, "-o", "/dev/null", "-s", str(sb), "-l", str(la),
                                "-v", *extra], capture_output=True, text=True)
            m = re.search(r"in ([0-9.]+) s \(([0-9.]+) GB/s", r.stderr)
            rec[key + "_gbs_codec_loop_to_devnull"] = float(m.group(2)) if (r.returncode == 0 and m) else None
        rec["stream_bytes"] = os.path.getsize(flz)
        rec["roundtrip_exact"] = open(fin, "rb").read() == open(fout, "rb").read()
    return rec


def sharded_workload(world: int):
    """ONE input resident on rank 0, cut into runs of whole blocks over all ranks (NCCL
    scatter -> per-rank encode -> gather into ONE stream -> sharded decode)."""
    if world == 4:   # BASELINE.json configs[3]
        return dict(name="configs[3]: 4 GiB log-like (4 KiB period), -s 4095 -l 15, 4 GPUs",
                    kind="log_like", n=4 << 30, sb=4095, la=15, seed=1234)
    if world == 8:   # configs[4]'s generator and parameters, 1 GiB per GPU: the worst-case
        #              stream bound of the full 32 GiB (128 GiB) does not fit root's HBM
        return dict(name="configs[4] generator: 8 GiB mixed corpus, -s 65535 -l 255, 8 GPUs",
                    kind="mixed", n=8 << 30, sb=65535, la=255, seed=1234)
    return dict(name=f"{world} x 256 MiB synthetic Zipf text, -s 4095 -l 15",
                kind="zipf_text", n=world * (256 << 20), sb=4095, la=15, seed=1234)


def run_sharded(args, dist, dev, rank, world, barrier):
    """Returns the `sharded` record (rank 0) -- timed with CUDA events around the
    collective calls, max over ranks; the NCCL transfers are inside the timed region."""
    import hashlib
    import torch
    import lz77_b200
    from lz77_b200 import api, synth

    wl = sharded_workload(world)
    if args.sharded_bytes:
        wl = dict(wl, n=args.sharded_bytes, name=wl["name"] + f" (cut to {args.sharded_bytes} B)")
    n, sb, la = wl["n"], wl["sb"], wl["la"]
    api.comm_init_torch(dev)
    src = out_stream = out_plain = None
    if rank == 0:
        src = synth.make(wl["kind"], n, seed=wl["seed"], device=dev)
        cap = (api.encode_bound(n, sb, la) + 15) & ~15
        out_stream = torch.empty(cap, dtype=torch.uint8, device=dev)
        out_plain = torch.empty((n + 15) & ~15, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, args.sharded_steps))
    stream_t = torch.cuda.current_stream(dev)
    merged = back = None
    for _ in range(2):   # warm-up: allocations, NCCL channels
        merged, k = api.encode_sharded_tensor(src, la=la, sb=sb, out=out_stream)
        back = api.decode_sharded_tensor(merged, out=out_plain)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    enc_stats, dec_stats = [], []
    barrier()
    ev[0].record(stream_t)
    for i in range(steps):
        merged, k = api.encode_sharded_tensor(src, la=la, sb=sb, out=out_stream)
        enc_stats.append(api.comm_stats())
        ev[2 * i + 1].record(stream_t)
        back = api.decode_sharded_tensor(merged, out=out_plain)
        dec_stats.append(api.comm_stats())
        ev[2 * i + 2].record(stream_t)
    barrier()
    enc_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)) / steps
    dec_ms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)) / steps
    times = torch.tensor([enc_ms, dec_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
    enc_ms, dec_ms = times.tolist()
    rec = None
    if rank == 0:
        # bit identity: the merged stream against what ONE GPU writes for the whole input
        roundtrip = bool(torch.equal(back, src))
        del out_plain, back
        single, k1 = lz77_b200.encode_tensor(src, la=la, sb=sb)
        same = bool(single.numel() == merged.numel() and torch.equal(single, merged) and k1 == k)
        sha_m = hashlib.sha256(merged.cpu().numpy()).hexdigest()
        sha_s = hashlib.sha256(single.cpu().numpy()).hexdigest()
        med = lambda rows, key: statistics.median(r[key] for r in rows)  # noqa: E731
        nv_enc = enc_stats[-1]["sent_bytes"] + enc_stats[-1]["recv_bytes"]
        nv_dec = dec_stats[-1]["sent_bytes"] + dec_stats[-1]["recv_bytes"]
        phases = {"enc_scatter_ms": med(enc_stats, "scatter_ms"),
                  "enc_compute_ms_root": med(enc_stats, "compute_ms"),
                  "enc_gather_ms": med(enc_stats, "gather_ms"),
                  "dec_scatter_ms": med(dec_stats, "scatter_ms"),
                  "dec_compute_ms_root": med(dec_stats, "compute_ms"),
                  "dec_gather_ms": med(dec_stats, "gather_ms")}
        worst = max(("enc_scatter_ms", "enc_gather_ms", "dec_scatter_ms", "dec_gather_ms"),
                    key=lambda k_: phases[k_])
        names = {"enc_scatter_ms": "scatter of the input runs (grouped ncclSend/ncclRecv, root -> ranks)",
                 "enc_gather_ms": "gather of the token payloads (grouped ncclSend/ncclRecv, ranks -> root)",
                 "dec_scatter_ms": "scatter of the token slices (grouped ncclSend/ncclRecv, root -> ranks)",
                 "dec_gather_ms": "gather of the plaintext (grouped ncclSend/ncclRecv, ranks -> root)"}
        rec = {"workload": wl["name"], "bytes": n, "ranks": world, "steps": steps,
               "bit_identical": same and sha_m == sha_s, "roundtrip_exact": roundtrip,
               "sha256": sha_m, "stream_bytes": int(merged.numel()), "tokens": int(k),
               "gbs": n / ((enc_ms + dec_ms) * 1e-3) / 1e9,
               "encode_gbs": n / (enc_ms * 1e-3) / 1e9, "decode_gbs": n / (dec_ms * 1e-3) / 1e9,
               "encode_ms": enc_ms, "decode_ms": dec_ms,
               "nvlink_bytes": int(nv_enc + nv_dec),
               "nvlink_bytes_encode": int(nv_enc), "nvlink_bytes_decode": int(nv_dec),
               "phases_ms": phases, "limiting_collective": names[worst],
               "transport": "NCCL point-to-point inside liblz77b200.so (comm.cu), one rank per GPU"}
        assert rec["bit_identical"], "sharded stream differs from the single-GPU stream"
        assert roundtrip, "sharded roundtrip mismatch"
    api.comm_destroy()
    return rec


def run_native(args):
    import torch
    import lz77_b200
    from lz77_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this codec has no CPU path")
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the first communicator is
        # created; stdout must carry exactly one JSON line, so park fd 1 on stderr
        # until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    lz77_b200.init(local_rank)
    sb, la, n = WORKLOAD["sb"], WORKLOAD["la"], args.bytes or WORKLOAD["n"]
    T = lz77_b200.token_bits(sb, la)

    # every rank owns its own shard of whole blocks (weak scaling): no data-path
    # collective, the ranks only meet at the timing barrier
    src = synth.make(WORKLOAD["kind"], n, seed=WORKLOAD["seed"] + rank, device=dev)
    cap = (api.encode_bound(n, sb, la) + 15) & ~15
    stream_buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    out_buf = torch.empty((n + 15) & ~15, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    # the library runs its kernels on this stream, and the timing events are recorded on it
    # (torch's default stream has handle 0, which the library would replace by its own)
    stream_t = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream_t)
    api.set_stream(stream_t.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        s, k = api.encode_tensor(src, la=la, sb=sb, out=stream_buf)
        t_enc = api.last_timing()
        back = api.decode_tensor(s, out=out_buf)
        t_dec = api.last_timing()
        return s, k, back, t_enc, t_dec

    # nvidia-smi needs a moment to start streaming: launch it before the warm-up so
    # it is sampling (every 50 ms) throughout both timed regions
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        s, k, back, _, _ = device_step()
    c_bytes = s.numel()
    assert torch.equal(back, src), "roundtrip mismatch"

    # ---- timed region: device-resident ------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    per = {"enc_search_ms": [], "enc_scan_ms": [], "enc_pack_ms": [], "dec_scan_ms": [],
           "dec_copy_ms": []}
    launches = 0
    barrier()
    ev[0].record(stream_t)
    for i in range(args.steps):
        s, k = api.encode_tensor(src, la=la, sb=sb, out=stream_buf)
        t_enc = api.last_timing()
        ev[2 * i + 1].record(stream_t)
        back = api.decode_tensor(s, out=out_buf)
        t_dec = api.last_timing()
        ev[2 * i + 2].record(stream_t)
        launches += t_enc["launches"] + t_dec["launches"]
        for key in ("enc_search_ms", "enc_scan_ms", "enc_pack_ms"):
            per[key].append(t_enc[key])
        for key in ("dec_scan_ms", "dec_copy_ms"):
            per[key].append(t_dec[key])
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    enc_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)) / args.steps
    dec_ms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)) / args.steps

    # ---- timed region: end to end through the host entry points -----------
    h_in = api.PinnedBuffer(n)
    h_stream = api.PinnedBuffer(cap)
    h_out = api.PinnedBuffer(n + 16)
    h_in.array[:] = src.cpu().numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        c = api.encode_into(h_in.ptr, n, h_stream.ptr, cap, la=la, sb=sb)
        m = api.decode_into(h_stream.ptr, c, h_out.ptr, n)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_enc_s = e2e_dec_s = 0.0
    e0.record(stream_t)
    for _ in range(e2e_steps):
        # (both calls are host-synchronous: the result is in host memory on return)
        t0 = time.perf_counter()
        c = api.encode_into(h_in.ptr, n, h_stream.ptr, cap, la=la, sb=sb)
        t1 = time.perf_counter()
        m = api.decode_into(h_stream.ptr, c, h_out.ptr, n)
        t2 = time.perf_counter()
        e2e_enc_s += t1 - t0
        e2e_dec_s += t2 - t1
    e1.record(stream_t)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_enc_ms, e2e_dec_ms = e2e_enc_s / e2e_steps * 1e3, e2e_dec_s / e2e_steps * 1e3
    assert m == n and bytes(h_out.array[:4096]) == bytes(h_in.array[:4096])
    assert (h_out.array[:n] == h_in.array[:n]).all(), "e2e roundtrip mismatch"

    # ---- the ceiling of that leg: the same bytes as plain pinned copies, H2D and D2H
    #      side by side on two streams (what a codec that cost nothing would take) ----
    p_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    p_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def plain_copies():
        with torch.cuda.stream(s_up):      # encode reads n, decode reads c
            d_a.copy_(p_in, non_blocking=True)
            d_a[:c_bytes].copy_(p_in[:c_bytes], non_blocking=True)
        with torch.cuda.stream(s_down):    # encode returns c, decode returns n
            p_out[:c_bytes].copy_(d_b[:c_bytes], non_blocking=True)
            p_out.copy_(d_b, non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()

    plain_copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plain_copies()
    copy_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    barrier()
    del p_in, p_out, d_a, d_b

    # ---- history mode: the window slides across block seams (the reference's ratio);
    #      such streams decode by pointer jumping, like streams of the reference encoder ----
    api.set_history(True)
    hs, hk = api.encode_tensor(src, la=la, sb=sb, out=stream_buf)
    hist_bytes = hs.numel()
    hb = api.decode_tensor(hs, out=out_buf)
    assert torch.equal(hb, src), "history-mode roundtrip mismatch"
    hist_enc, hist_dec = [], []
    for _ in range(max(1, min(args.steps, 3))):
        a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a0.record(stream_t)
        hs, hk = api.encode_tensor(src, la=la, sb=sb, out=stream_buf)
        a1.record(stream_t)
        hb = api.decode_tensor(hs, out=out_buf)
        a2.record(stream_t)
        torch.cuda.synchronize()
        hist_enc.append(a0.elapsed_time(a1))
        hist_dec.append(a1.elapsed_time(a2))
    api.set_history(False)
    hist_enc_ms, hist_dec_ms = statistics.median(hist_enc), statistics.median(hist_dec)
    clocks = sampler.stop()

    # ---- max over ranks ----------------------------------------------------
    times = torch.tensor([total_ms, enc_ms, dec_ms, e2e_ms, e2e_enc_ms, e2e_dec_ms, copy_ms,
                          hist_enc_ms, hist_dec_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    (total_ms, enc_ms, dec_ms, e2e_ms, e2e_enc_ms, e2e_dec_ms, copy_ms, hist_enc_ms,
     hist_dec_ms) = times.tolist()

    cli = None
    if rank == 0 and not args.no_cli:
        cli = run_cli(src, sb, la, 1)
        if world > 1:
            cli = {"one_gpu": cli, "all_gpus": run_cli(src, sb, la, world)}
    sharded = None
    if dist is not None and not args.no_sharded:
        del src, stream_buf, out_buf, s, back
        h_in.free(), h_stream.free(), h_out.free()
        torch.cuda.empty_cache()
        sharded = run_sharded(args, dist, dev, rank, world, barrier)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        med = {k_: statistics.median(v) for k_, v in per.items()}
        step_ms = total_ms / args.steps
        alg_bytes = n + c_bytes     # N + C per direction, per launch
        search_gbs = alg_bytes / (med["enc_search_ms"] * 1e-3) / 1e9
        copy_gbs = alg_bytes / (med["dec_copy_ms"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * n / (step_ms * 1e-3) / 1e9, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "bytes_per_gpu": n, "sb": sb, "la": la,
                       "token_bits": T, "block_bytes": lz77_b200.block_size(sb),
                       "segment_bytes": lz77_b200.segment_size(sb, la),
                       "sharding": f"{world} rank(s), whole blocks per rank, no collective",
                       "l2": "inputs (256 MiB) larger than L2 (126 MB); no flush needed"},
            "encode_gbs": world * n / (enc_ms * 1e-3) / 1e9,
            "decode_gbs": world * n / (dec_ms * 1e-3) / 1e9,
            "ratio": n / c_bytes, "tokens": k,
            "kernel_ms": med,
            "e2e": {"value": world * n / (e2e_ms / e2e_steps * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": n + c_bytes, "d2h_bytes_per_step": c_bytes + n,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "encode_gbs": world * n / (e2e_enc_ms * 1e-3) / 1e9,
                    "decode_gbs": world * n / (e2e_dec_ms * 1e-3) / 1e9,
                    "encode_ms": e2e_enc_ms, "decode_ms": e2e_dec_ms,
                    # the same H2D + D2H bytes as plain pinned copies on two streams, all
                    # ranks at once: what the host memory system / PCIe allow this leg
                    "copy_ceiling_gbs": world * n / (copy_ms * 1e-3) / 1e9,
                    "copy_ceiling_ms": copy_ms,
                    "frac_of_copy_ceiling": copy_ms / (e2e_ms / e2e_steps),
                    "numa_binding": numa},
            "history_mode": {"what": "lz77_gpu_set_history(1): window slides across block seams "
                                     "(lz77.c:101-105); decode = pointer jumping (decode_jump.cu), "
                                     "the path streams of the reference encoder take",
                             "encode_gbs": world * n / (hist_enc_ms * 1e-3) / 1e9,
                             "decode_gbs": world * n / (hist_dec_ms * 1e-3) / 1e9,
                             "encode_ms": hist_enc_ms, "jump_decode_ms": hist_dec_ms,
                             "ratio": n / hist_bytes,
                             "jump_decode_hbm_frac": (n + hist_bytes) / (hist_dec_ms * 1e-3) / 1e9 / peak},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "lz77_parse_bucket_kernel (longest-match search + greedy parse)",
                         "bound": "hbm", "achieved": search_gbs, "peak": peak, "unit": "GB/s",
                         "frac": search_gbs / peak,
                         "traffic": ncu_figure("lz77_parse_bucket_kernel", n),
                         "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes,
                         "issue_bound": issue_bound("lz77_parse_bucket_kernel", n, k,
                                                    med["enc_search_ms"], clocks)},
            "roofline_decode": {"kernel": "lz77_decode_tile_kernel (match copy)",
                                "bound": "hbm", "achieved": copy_gbs, "peak": peak,
                                "unit": "GB/s", "frac": copy_gbs / peak,
                                "traffic": ncu_figure("lz77_decode_tile_kernel", n),
                                "algorithmic_bytes": alg_bytes,
                                "issue_bound": issue_bound("lz77_decode_tile_kernel", n, k,
                                                           med["dec_copy_ms"], clocks)},
        }
        if sharded is not None:
            line["sharded"] = sharded
        if cli is not None:
            line["cli"] = cli
        if not args.no_cpu_baseline:
            try:
                ref = CpuReference(slice_bytes=args.ref_slice_mib << 20)
                dt, b, c_ref = ref.step()
                line["cpu_baseline"] = {"value": b / dt / 1e9, "unit": UNIT, "cores": ref.cores,
                                        "kind": ref.kind, "sample": ref.sample_desc(),
                                        "ratio": b / c_ref}
            except Exception as exc:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                        "sample": f"failed: {exc}"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--bytes", type=int, default=0, help="override the per-GPU input size")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-slice-mib", type=int, default=8,
                    help="per-core slice the CPU reference encodes+decodes per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the file-to-file CLI leg")
    ap.add_argument("--no-sharded", action="store_true",
                    help="N > 1: skip the one-input sharded leg (NCCL scatter/gather)")
    ap.add_argument("--sharded-steps", type=int, default=3)
    ap.add_argument("--sharded-bytes", type=int, default=0, help="override the sharded input size")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
