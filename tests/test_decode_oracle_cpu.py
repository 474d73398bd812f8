"""The decode-step oracle (oracle/decode_oracle.py) against logits produced by the reference's own
inference/model.py Transformer on CPU (tests/golden/decode_golden.npz, made by tests/golden/make_decode_golden.py)."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    g = np.load(os.path.join(HERE, "golden", "decode_golden.npz"))
    L, H, Hkv, dim, inter, vocab, S = map(int, g["meta"])
    w = {k[2:]: torch.from_numpy(g[k].astype(np.float32)) for k in g.files if k.startswith("w:")}
    return g, w, (L, H, Hkv, dim, inter, vocab, S)


def test_decode_oracle_matches_reference_model_fp32():
    from oracle.decode_oracle import DecodeOracle

    g, w, (L, H, Hkv, dim, inter, vocab, S) = _load()
    o = DecodeOracle(w, L, H, Hkv, dim, S, rope_base=500000.0, half_rounding=False)
    for pos, tok in enumerate(g["tokens"]):
        logits = o.step(int(tok), pos).numpy()
        ref = g["logits"][pos]
        assert np.abs(logits - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-5, pos
        assert int(np.argmax(logits)) == int(np.argmax(ref))


def test_decode_oracle_half_rounding_stays_close():
    from oracle.decode_oracle import DecodeOracle

    g, w, (L, H, Hkv, dim, inter, vocab, S) = _load()
    o = DecodeOracle(w, L, H, Hkv, dim, S, rope_base=500000.0, half_rounding=True)
    for pos, tok in enumerate(g["tokens"]):
        logits = o.step(int(tok), pos).numpy()
        ref = g["logits"][pos]
        assert np.abs(logits - ref).max() <= 2e-2 * np.abs(ref).max(), pos
