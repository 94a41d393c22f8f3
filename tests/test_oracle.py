"""Pins the CPU oracle (oracle/lz77_oracle.c) against the reference:
 - the known-answer vectors of SURVEY.md Appendix C (bytes hard-coded here),
 - every entry of tests/golden/golden.json (outputs of the compiled reference),
 - the live reference binary oracle/_ref/lz77 when it is present.
No GPU involved."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

from _cases import GOLDEN_CASES, PARAM_SETS, case_input

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

# SURVEY.md Appendix C (generated from the compiled reference).
APPENDIX_C = [
    (b"", None, None, "ff0f0f00"),
    (b"a", None, None, "ff0f0f00000061"),
    (b"abcabcabcabcX", None, None, "ff0f0f00000061000062000063039058"),
    (b"abcabcabcabcX", 65535, 255, "ffffff0000000061000000620000006303000958"),
    (b"abcabcabcabcX", 1000, 20, "e803140000803000801800606c808405"),
    (b"abcabcabcabcX", 1, 2, "01000200c288191346cc983062c6841133060b"),
    (b"a" * 100, None, None,
     "ff0f0f0000006101e06110e0611fe0612ee0613de0614ce0615b8061"),
]


@pytest.mark.parametrize("data,sb,la,hexout", APPENDIX_C)
def test_appendix_c_known_answers(orc, data, sb, la, hexout):
    enc = orc.ref_encode(data, -1 if sb is None else sb, -1 if la is None else la)
    assert enc.hex() == hexout
    assert orc.decode(enc) == data


def test_appendix_c_large_vectors(orc):
    z = bytes(1 << 20)
    enc = orc.ref_encode(z)
    assert len(enc) == 209_722
    assert hashlib.sha256(enc).hexdigest() == \
        "42e454e313f95e04daa9717687ab77d9ab6351c691bb1fccad1ff73269e5c4b1"
    assert enc[:19].hex() == "ff0f0f0000000001e00010e0001fe0002ee000"
    assert orc.decode(enc) == z
    r = bytes(range(256)) * 16
    enc = orc.ref_encode(r)
    assert len(enc) == 1540
    assert hashlib.sha256(enc).hexdigest() == \
        "50ae8ae550a2e71ced08401e2c7bf0fe6e1e7db9f603dd4d0db2f857535e9216"


@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES if c["name"] != "kat_zeros_1m"],
                         ids=lambda c: c["name"])
def test_golden_encode_matches_reference(orc, golden, case):
    g = golden[case["name"]]
    data = case_input(case)
    assert hashlib.sha256(data).hexdigest() == g["input_sha256"], "input generator drifted"
    enc = orc.ref_encode(data, case.get("sb", -1), case.get("la", -1))
    assert len(enc) == g["n_out"]
    assert hashlib.sha256(enc).hexdigest() == g["sha256"]
    if "hex" in g:
        assert enc.hex() == g["hex"]
    if g["ref_roundtrip_ok"]:
        assert orc.decode(enc) == data


@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES if c.get("store")],
                         ids=lambda c: c["name"])
def test_golden_streams_decode(orc, golden, case):
    g = golden[case["name"]]
    stream = (GOLDEN_DIR / g["stream"]).read_bytes()
    assert hashlib.sha256(stream).hexdigest() == g["sha256"]
    if g["ref_roundtrip_ok"]:
        assert orc.decode(stream) == case_input(case)


def test_bitof_matches_libm_formula(orc):
    import math
    for n in range(1, 65536):
        assert orc.bitof(n) == int(math.ceil(math.log(n) / math.log(2))), n
    assert orc.token_bits(4095, 15) == 24
    assert orc.token_bits(65535, 255) == 32
    assert orc.token_bits(1000, 20) == 23
    assert orc.token_bits(1, 2) == 9


def _rand_cases():
    rng = np.random.default_rng(99)
    out = []
    for i in range(24):
        n = int(rng.integers(0, 6000))
        alpha = int(rng.choice([2, 4, 27, 256]))
        sb, la = PARAM_SETS[i % len(PARAM_SETS)]
        out.append((rng.integers(0, alpha, n, dtype=np.uint8).tobytes(), sb, la))
    return out


def test_live_reference_cross_check(orc, ref_available):
    """oracle == compiled reference on seeded inputs, both directions."""
    if not ref_available:
        pytest.skip("oracle/_ref/lz77 not built on this machine")
    from oracle import ref_run
    for data, sb, la in _rand_cases():
        enc_ref = ref_run("-c", data, sb=sb, la=la)
        assert orc.ref_encode(data, sb, la) == enc_ref, (len(data), sb, la)
        assert orc.decode(enc_ref) == ref_run("-d", enc_ref), (len(data), sb, la)


@pytest.mark.parametrize("sb,la", PARAM_SETS)
@pytest.mark.parametrize("block", [0, 4096, 65536])
def test_blocked_spec_roundtrip(orc, ref_available, sb, la, block):
    """The block-parallel encoder specification produces streams the reference
    decoder restatement (and the live reference) decode bit-exact."""
    from lz77_b200 import synth
    for kind, n in (("zipf_text", 70_001), ("random", 9_000), ("zeros", 10_000)):
        data = synth.make(kind, n, seed=5).numpy().tobytes()
        enc, ntok = orc.blocked_encode(data, sb, la, block)
        assert enc[:4] == bytes([sb & 255, sb >> 8, la & 255, la >> 8])
        assert len(enc) == 4 + (ntok * orc.token_bits(sb, la) + 7) // 8
        assert orc.decode(enc) == data
        if ref_available and kind != "zeros":
            from oracle import ref_run
            assert ref_run("-d", enc) == data


def test_exhaustive_greedy_equals_reference_token_count(orc):
    """the reference BST finds the exhaustive longest match on
    text, so the unblocked specification has exactly its token count."""
    from lz77_b200 import synth
    data = synth.zipf_text(200_000, seed=3).numpy().tobytes()
    ref = orc.ref_encode(data)
    _, ntok = orc.blocked_encode(data, block=0)
    assert ntok == (len(ref) - 4) // 3


def test_decode_rejects_malformed(orc):
    with pytest.raises(ValueError):
        orc.decode(b"\xff\x0f")                       # truncated header
    with pytest.raises(ValueError):
        orc.decode(b"\x00\x00\x0f\x00")               # SB == 0
    with pytest.raises(ValueError):
        orc.decode(b"\xff\x0f\x0f\x00" + bytes([5, 0x10, 65]))  # off 5 > history 0
    # trailing bits shorter than a token are padding, not an error
    assert orc.decode(b"\xff\x0f\x0f\x00\x00\x00\x61\x00\x00") == b"a"
