#!/usr/bin/env python
"""Regenerate tests/golden/golden.json and the *.lz fixtures from the compiled,
unmodified reference (oracle/_ref/lz77, built from /root/reference by
oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are rebuilt from (kind, n, seed) by tests/_cases.py, so only the
reference OUTPUT is stored: full bytes for the small known-answer vectors
 and the streams under tests/golden/streams/, and
size + sha256 for the larger ones.
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from _cases import GOLDEN_CASES, case_input  # noqa: E402
from oracle import ref_run  # noqa: E402

OUT = Path(__file__).resolve().parent


def main() -> None:
    streams = OUT / "streams"
    streams.mkdir(exist_ok=True)
    table = []
    for case in GOLDEN_CASES:
        data = case_input(case)
        sb, la = case.get("sb"), case.get("la")
        enc = ref_run("-c", data, sb=sb, la=la)
        dec = ref_run("-d", enc)
        entry = dict(case)
        entry["n_out"] = len(enc)
        entry["sha256"] = hashlib.sha256(enc).hexdigest()
        entry["input_sha256"] = hashlib.sha256(data).hexdigest()
        # Appendix B2: the reference itself corrupts data for power-of-two SB
        entry["ref_roundtrip_ok"] = dec == data
        if len(enc) <= 64:
            entry["hex"] = enc.hex()
        if case.get("store"):
            (streams / f"{case['name']}.lz").write_bytes(enc)
            entry["stream"] = f"streams/{case['name']}.lz"
        table.append(entry)
        print(f"{case['name']:28s} in={len(data):8d} out={len(enc):8d} "
              f"roundtrip={'ok' if dec == data else 'CORRUPT (reference bug B2)'}")
    (OUT / "golden.json").write_text(json.dumps(table, indent=1) + "\n")


if __name__ == "__main__":
    main()
