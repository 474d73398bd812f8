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


def test_sampling_oracle_equals_reference_formulation():
    """oracle.sample_topk_scores (what the GPU tests check the kernel against) picks the same token as the reference's
    logits_to_probs + multinomial_sample_one_no_sync (generate.py:55-73) given the same Exp(1) noise."""
    import numpy as np

    from oracle.decode_oracle import sample_topk_reference, sample_topk_scores, sample_uniform

    rng = np.random.default_rng(0)
    u = sample_uniform(1234, 5, 200000)
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3 and abs(np.var(u) * 12 - 1) < 2e-2
    assert not np.array_equal(u[:100], sample_uniform(1234, 6, 100)) and not np.array_equal(u[:100], sample_uniform(1235, 5, 100))
    for t in range(100):
        V = int(rng.integers(40, 3000))
        l = (rng.standard_normal(V) * 2).astype(np.float16)
        for top_k in (None, 1, 32, V + 5):
            s, q = sample_topk_scores(l, 0.8, top_k, 99, t)
            a, probs = sample_topk_reference(l, 0.8, top_k, q)
            assert a == int(np.argmax(s))
            assert np.array_equal(np.isfinite(s), (probs > 0).numpy() | np.isfinite(s))  # kept set ⊇ non-zero probs
            if top_k is not None and top_k < V:
                assert np.isfinite(s).sum() >= top_k


def test_sharded_sampler_algorithm_equals_the_unsharded_one():
    """the two-exchange scheme of the tensor-parallel sampler (k largest keys per rank -> global pivot; per-rank best ->
    winner) picks the token the single-GPU sampler picks, for every world size, with ties at the pivot"""
    import numpy as np

    from oracle import decode_oracle as D

    rng = np.random.default_rng(3)
    for V, k, temp in ((1024, 8, 0.8), (4096, 32, 1.0), (2048, 1, 0.5), (1024, None, 1.3), (512, 5, 0.9)):
        logits = rng.standard_normal(V).astype(np.float16)
        logits[rng.integers(0, V, 40)] = np.float16(1.5)  # many ties, some of them at the pivot for small k
        for world in (2, 4, 8):
            for pos in range(4):
                s, _ = D.sample_topk_scores(logits, temp, k, 1234, pos)
                assert D.sample_topk_sharded(logits, temp, k, 1234, pos, world) == int(np.argmax(s)), (V, k, world, pos)
