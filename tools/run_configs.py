#!/usr/bin/env python
"""Runs the BASELINE.json configurations at their full per-GPU sizes on one GPU
and prints one JSON line per configuration: device-resident encode and decode
GB/s, ratio, and the size-independent checks (decode(encode(x)) == x on the
device, stream size == header + tokens * T, decode_size == n).

    python tools/run_configs.py [--quick]

configs[3] (4 GiB over 4 GPUs) and configs[4] (32 GiB over 8 GPUs) are sharded runs of
whole blocks: here one GPU's share (1 GiB / 4 GiB of the same generator), i.e. what a rank
executes; the sharded one-stream runs themselves are `bench.py --gpus 4|8` (`sharded`
record).  Every line also carries the history-mode figures (window across block seams,
pointer-jumping decode) and `ref_decoder_ok`: the first 8 MiB of the stream decoded by the
compiled reference (`oracle/_ref/lz77 -d`; the full first / middle / last 64 MiB check is
tests/test_gpu_parity_at_size.py).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import lz77_b200  # noqa: E402
from lz77_b200 import api, synth  # noqa: E402

CONFIGS = [
    ("configs[0] 1 MiB zeros -s 4095 -l 15", "zeros", 1 << 20, 4095, 15),
    ("configs[1] 256 MiB Zipf text -s 4095 -l 15", "zipf_text", 256 << 20, 4095, 15),
    ("configs[2] 1 GiB random -s 65535 -l 255", "random", 1 << 30, 65535, 255),
    ("configs[3] 4 GiB log-like / 4 GPUs: one rank's 1 GiB, -s 4095 -l 15", "log_like", 1 << 30,
     4095, 15),
    ("configs[4] 32 GiB mixed / 8 GPUs: one rank's 4 GiB, -s 65535 -l 255", "mixed", 4 << 30,
     65535, 255),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="1/8 of every size")
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    lz77_b200.init(0)
    dev = torch.device("cuda", 0)
    peak = 6442.9
    try:
        peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json")
                          .read_text())["hbm_gbs"]
    except Exception:
        pass
    for idx, (name, kind, n, sb, la) in enumerate(CONFIGS):
        if args.only >= 0 and idx != args.only:
            continue
        if args.quick:
            n = max(n >> 3, 1 << 20)
        if kind == "mixed":
            src = synth.mixed(n, seed=1234, device=dev)
        else:
            src = synth.make(kind, n, seed=1234, device=dev)
        cap = (api.encode_bound(n, sb, la) + 15) & ~15
        stream_buf = torch.empty(cap, dtype=torch.uint8, device=dev)
        out_buf = torch.empty((n + 15) & ~15, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        best_enc, best_dec = 1e30, 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            s, k = api.encode_tensor(src, la=la, sb=sb, out=stream_buf)
            t_enc = api.last_timing()
            t1 = time.perf_counter()
            back = api.decode_tensor(s, out=out_buf)
            t_dec = api.last_timing()
            t2 = time.perf_counter()
            best_enc = min(best_enc, t1 - t0)
            best_dec = min(best_dec, t2 - t1)
        T = lz77_b200.token_bits(sb, la)
        c = s.numel()
        ok = bool(torch.equal(back, src)) and c == 4 + (k * T + 7) // 8 and \
            api.decode_size_tensor(s) == n
        copy_ms = t_dec["dec_copy_ms"]
        # an independent decoder on a block-aligned cut of the stream
        ref_ok = None
        try:
            from oracle import ref_binary, ref_run
            if ref_binary() is not None:
                cut = min(n, 8 << 20) // lz77_b200.block_size(sb) * lz77_b200.block_size(sb)
                if cut > 0:
                    k_hi, p_hi = lz77_b200.token_at_tensor(s, cut)
                    part = lz77_b200.slice_tokens_tensor(s, 0, k_hi).cpu().numpy().tobytes()
                    ref_ok = p_hi == cut and ref_run("-d", part) == src[:cut].cpu().numpy().tobytes()
        except Exception as exc:  # noqa: BLE001
            ref_ok = f"failed: {exc}"
        # history mode: the reference's sliding window, pointer-jumping decode
        api.set_history(True)
        t0 = time.perf_counter()
        hs, hk = api.encode_tensor(src, la=la, sb=sb, out=stream_buf)
        t1 = time.perf_counter()
        hback = api.decode_tensor(hs, out=out_buf)
        t2 = time.perf_counter()
        api.set_history(False)
        hist = {"ratio": n / hs.numel(), "encode_gbs": n / (t1 - t0) / 1e9,
                "jump_decode_gbs": n / (t2 - t1) / 1e9, "roundtrip_bit_exact": bool(torch.equal(hback, src))}
        line = {
            "config": name, "bytes": n, "sb": sb, "la": la, "token_bits": T, "tokens": k,
            "ratio": n / c, "roundtrip_bit_exact": ok, "ref_decoder_ok": ref_ok, "history_mode": hist,
            "encode_gbs": n / best_enc / 1e9, "decode_gbs": n / best_dec / 1e9,
            "kernel_ms": {"search": t_enc["enc_search_ms"], "pack": t_enc["enc_pack_ms"],
                          "dec_scan": t_dec["dec_scan_ms"], "dec_copy": copy_ms},
            "decode_copy_roofline_frac": (n + c) / (copy_ms * 1e-3) / 1e9 / peak if copy_ms else None,
        }
        print(json.dumps(line), flush=True)
        del src, stream_buf, out_buf, s, back
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
