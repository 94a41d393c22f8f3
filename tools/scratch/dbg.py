import sys
sys.path.insert(0, '.')
import numpy as np
import lz77_b200 as lz
from lz77_b200 import synth
from oracle import oracle
orc = oracle()
lz.init(0)
def toks(stream, T):
    bits = np.unpackbits(np.frombuffer(stream[4:], dtype=np.uint8), bitorder='little')
    k = len(bits) // T
    return bits[:k*T].reshape(k, T)
for kind, n, sb, la in [("random", 9000, 1, 2), ("random", 9000, 4095, 15), ("zipf_text", 70001, 1, 2), ("random", 9000, 15, 8)]:
    data = synth.make(kind, n, seed=5).numpy().tobytes()
    enc = lz.encode(data, la=la, sb=sb)
    spec, ntok = orc.blocked_encode(data, sb, la, lz.block_size(sb), lz.segment_size(sb, la))
    print(kind, n, sb, la, "len", len(enc), len(spec), "equal", enc == spec)
    if enc != spec:
        T = lz.token_bits(sb, la)
        a, b = toks(enc, T), toks(spec, T)
        m = min(len(a), len(b))
        d = np.nonzero((a[:m] != b[:m]).any(axis=1))[0]
        print("  tokens", len(a), len(b), "first diffs", d[:10])
        ob = T - 8 - (T - 8 - int(np.ceil(np.log2(sb + 1))) if False else 0)
        for i in d[:5]:
            print("   gpu ", ''.join(map(str, a[i])), " spec", ''.join(map(str, b[i])))
        # byte position of the first differing token in the spec
        pos = 0
        obits = lz.token_bits(sb, la) - 8
        import math
        def bitof(x):
            return max(1, int(x).bit_length())
        ob_, lb_ = bitof(sb), bitof(la)
        for i in range(int(d[0])):
            ln = int(''.join(map(str, b[i][ob_:ob_+lb_][::-1])), 2)
            pos += ln + 1
        print("  first differing token starts at byte", pos)
    try:
        print("  gpu decode ok:", lz.decode(enc) == data, " oracle decode of gpu stream ok:", orc.decode(enc) == data)
    except Exception as e:
        print("  decode error", e)
