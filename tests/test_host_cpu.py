"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls),
the torch-facing layer validates arguments like the reference, the op is registered with the reference's
schema, and the K-shard planner / re-packer is exact."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    from guidedquant_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "apgemv_b200.h")).read() + open(os.path.join(ROOT, "include", "apdecode_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ap[gd]_[a-z0-9_]+)\s*\(", hdr))
    assert {"apd_embed", "apd_attn_decode", "apd_lm_head", "apd_argmax_advance", "apg_gemv_fused"} <= declared
    assert {"apg_gemv", "apg_gemv_ex", "apg_dequant", "apg_version"} <= declared
    path = _lib.build()
    L = ctypes.CDLL(path)
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in include/apgemv_b200.h but not exported"
    assert set(_lib.EXPORTS) <= declared
    L.apg_version.restype = ctypes.c_int
    assert L.apg_version() >= 100
    L.apg_status_string.restype = ctypes.c_char_p
    assert b"Bitwidth" in L.apg_status_string(2)


def test_capi_argument_validation_without_gpu():
    """validation happens before any CUDA call, so it is checkable on a CPU-only box"""
    from guidedquant_b200 import _lib

    L = _lib.lib()
    buf = (ctypes.c_uint8 * 4096)()
    p = ctypes.addressof(buf)
    p = (p + 63) & ~63
    assert L.apg_gemv(None, p, p, p, 1, 4, 128, 2, None) == 1          # null
    assert L.apg_gemv(p, p, p, p, 1, 4, 128, 9, None) == 2             # bits
    assert L.apg_gemv(p, p, p, p, 9, 4, 128, 2, None) == 3             # batch
    assert L.apg_gemv(p, p, p, p, 1, 4, 100, 2, None) == 4             # K % 32
    assert L.apg_gemv(p, p, p, p, 1, 0, 128, 2, None) == 4             # N
    assert L.apg_gemv(p + 2, p, p, p, 1, 4, 128, 2, None) == 5         # alignment of x
    assert L.apg_gemv_ex(p, p, None, p, p, 1, 4, 128, 2, 0x80, 0, None) == 7   # unknown flag
    assert L.apg_dequant(p, p, p, 4, 128, 1, None) == 2


def test_python_layer_validates_like_the_reference():
    from guidedquant_b200 import ap_gemv

    x = torch.zeros((1, 1, 128), dtype=torch.float16)
    q = torch.zeros((2, 8, 4), dtype=torch.int32)
    lut = torch.zeros((8, 4), dtype=torch.float16)
    out = torch.zeros((1, 1, 8), dtype=torch.float16)
    with pytest.raises(RuntimeError, match="must be on GPU"):
        ap_gemv.anyprec_gemv(x, out, q, lut, 2)
    with pytest.raises(RuntimeError, match="Bitwidth must be between 2 and 8"):
        ap_gemv.anyprec_gemv(x, out, q, lut, 1)
    with pytest.raises(RuntimeError, match="Mismatched data types"):
        ap_gemv.anyprec_gemv(x.float(), out, q, lut, 2)
    with pytest.raises(RuntimeError, match="shape \\(batch_size, seq_len, hidden_size\\)"):
        ap_gemv.anyprec_gemv(x[0], out, q, lut, 2)
    with pytest.raises(RuntimeError):
        ap_gemv.anyprec_dequant(q, lut, 2)   # CPU tensors


def test_plugin_op_schema_matches_reference():
    import guidedquant_b200.plugin as plugin  # noqa: F401

    op = torch.ops.plugin.anyprec_gemv.default
    s = str(op._schema)
    # reference: anyprec_gemv(x, q_weight, lut, output, bitwidth) -> None, mutates output (inference/plugin.py:7-13)
    assert "plugin::anyprec_gemv(Tensor x, Tensor q_weight, Tensor lut, Tensor(a3!) output, SymInt bitwidth) -> ()" == s or \
        ("Tensor x, Tensor q_weight, Tensor lut" in s and "output" in s and "bitwidth" in s and s.endswith("-> ()")), s
    assert callable(plugin.anyprec_dequant)


def test_install_as_ap_gemv():
    import sys

    import guidedquant_b200

    guidedquant_b200.install_as_ap_gemv()
    import ap_gemv

    assert ap_gemv is sys.modules["guidedquant_b200.ap_gemv"]
    assert {"anyprec_gemv", "anyprec_dequant", "lutgemm_gemv"} <= set(dir(ap_gemv))


def test_shard_bounds_and_repack_exact():
    from guidedquant_b200 import pack as P

    rng = np.random.default_rng(0)
    for K, world, bits in ((28672, 8, 2), (14336, 4, 3), (8192, 2, 4), (11008, 2, 2), (4096, 8, 2)):
        bounds = P.shard_bounds(K, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == K
        assert all(b[1] == bounds[i + 1][0] for i, b in enumerate(bounds[:-1]))
        assert all((b - a) % 128 == 0 and b > a for a, b in bounds)
        idx = rng.integers(0, 1 << bits, size=(5, K), dtype=np.uint8)
        q = P.pack_indices(idx, bits)
        for a, b in bounds:
            qs = P.shard_k(q, a, b)
            assert qs.shape == (bits, 5, (b - a) // 32)
            assert np.array_equal(P.unpack_indices(qs, bits), idx[:, a:b])
            assert np.array_equal(P.shard_k_torch(torch.from_numpy(q), a, b).numpy(), qs)   # torch path: bit-identical


def test_checkpoint_converter_matches_reference_converter():
    """guidedquant_b200.convert vs the reference's sqllm_llama_convert_fuse.py run on the same seeded synthetic
    checkpoint (tests/golden/convert_golden.json, made by tests/golden/make_convert_golden.py): same keys, shapes,
    dtypes and bytes."""
    import json

    from guidedquant_b200.convert import convert_state_dict
    from tests.convert_fixture import digest, make_hf_checkpoint

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "convert_golden.json")))
    out = convert_state_dict(make_hf_checkpoint(), gold["bitwidth"])
    assert sorted(out.keys()) == sorted(gold["tensors"].keys())
    for k, v in out.items():
        assert digest(v) == gold["tensors"][k], k
    assert out["layers.0.attention.wqkv.qweight"].shape == (3, 64 + 32 + 32, 2)
    assert out["layers.0.feed_forward.w1w3.lut"].shape == (192, 8)
    # a float32 lut is cast to fp16 here (the reference's cast at :56-57 tests the pre-rename key and never fires)
    f32 = make_hf_checkpoint(n_layer=1)
    f32["model.layers.0.self_attn.o_proj.lut2"] = f32["model.layers.0.self_attn.o_proj.lut2"].float()
    assert convert_state_dict(f32, 2)["layers.0.attention.wo.lut"].dtype == torch.float16
    # layer count comes from the keys (the reference only accepts Llama-2 directory names)
    small = {k: v for k, v in make_hf_checkpoint(n_layer=3).items()}
    assert "layers.2.attention.wqkv.qweight" in convert_state_dict(small, 2)


def test_fast_kernel_work_decomposition_covers_every_row_once():
    """apg_plan_fast exposes the host-side plan of the fast GEMV kernel; replay the kernel's own row arithmetic
    (apgemv_fast.cuh: units -> rows -> stages -> groups/slots) for many shapes and check the invariants the device code
    relies on: rows partitioned exactly, only a CTA's last stage partial, ring slots a multiple of the groups, the
    reduction buffer large enough, shared memory within the opt-in budget, block size within the launch bound."""
    import ctypes

    from guidedquant_b200 import _lib

    L = _lib.lib()
    plan = (ctypes.c_uint32 * 16)()
    shapes = [(N, K) for N in (1, 3, 4, 7, 8, 9, 100, 1184, 1185, 4096, 6144, 8192, 10240, 14336, 28672, 57344, 128256)
              for K in (128, 1024, 1152, 2048, 3584, 4096, 7168, 8192, 11008, 14336, 28672, 32768)]
    checked = 0
    for N, K in shapes:
        for bits in (2, 3, 4):
            for ctas in (0, 1, 2, 3):
                for sms in (148, 132, 8):
                    rc = L.apg_plan_fast(N, K, bits, ctas, sms, ctypes.byref(plan))
                    if rc != 0:
                        assert rc == 8, rc  # APG_ERR_UNSUPPORTED: falls to the wide kernel
                        continue
                    cpw, nwk, G, RS, NS, stage_bytes, grid, red_rows, threads, unit, uq, urem, smem = list(plan)[:13]
                    nchunk = (K + 1023) // 1024
                    assert cpw in (1, 2) and nwk == (nchunk + cpw - 1) // cpw and 1 <= nwk <= 16
                    assert threads == (G * nwk + 1) * 32 <= 544
                    assert RS in (2, 4, 8) and unit in (RS, RS // 2) and unit >= 1
                    assert stage_bytes == RS * bits * K // 8
                    assert NS >= G and NS % G == 0
                    assert smem <= 200 * 1024
                    assert 1 <= grid <= max(1, sms * (ctas if ctas > 0 else 2))
                    # replay the kernel's partition
                    nxt, max_rows = 0, 0
                    for b in range(grid):
                        u0 = b * uq + min(b, urem)
                        n = uq + (1 if b < urem else 0)
                        r0, r1 = min(u0 * unit, N), min((u0 + n) * unit, N)
                        assert r0 == nxt, (N, K, bits, b)
                        nxt = r1
                        nrows = r1 - r0
                        max_rows = max(max_rows, nrows)
                        nstages = (nrows + RS - 1) // RS
                        for s in range(nstages):
                            rows = min(RS, r1 - (r0 + s * RS))
                            assert rows >= 1 and (rows == RS or s == nstages - 1)
                    assert nxt == N, (N, K, bits, ctas, sms)
                    assert max_rows <= red_rows
                    checked += 1
    assert checked > 2000


def test_arch_from_hf_config_and_checkpoint_reader(tmp_path):
    """the CPU half of APTransformer.from_checkpoint: config.json -> (arch, rope base, eps) for both spellings of the RoPE
    settings, loud refusal of what the kernels do not implement, and the checkpoint reader on a .bin directory."""
    import torch

    from guidedquant_b200.convert import arch_from_hf_config, convert_state_dict, read_checkpoint

    base = {"hidden_size": 256, "num_hidden_layers": 2, "num_attention_heads": 2, "num_key_value_heads": 1,
            "intermediate_size": 512, "vocab_size": 128, "rms_norm_eps": 1e-5}
    cfg, theta, eps = arch_from_hf_config({**base, "rope_theta": 500000.0})           # transformers 4.x (the reference pins 4.52)
    assert cfg == dict(dim=256, n_layer=2, n_head=2, n_kv=1, inter=512, vocab=128) and theta == 500000.0 and eps == 1e-5
    cfg2, theta2, _ = arch_from_hf_config({**base, "rope_parameters": {"rope_theta": 10000.0, "rope_type": "default"}})  # 5.x
    assert cfg2 == cfg and theta2 == 10000.0
    assert arch_from_hf_config({**base, "rope_scaling": None})[1] == 10000.0          # default base
    mha = dict(base)
    del mha["num_key_value_heads"]
    assert arch_from_hf_config(mha)[0]["n_kv"] == 2                                    # MHA checkpoints omit the key
    l3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}
    cfg3, theta3, _ = arch_from_hf_config({**base, "rope_theta": 500000.0, "rope_scaling": l3})   # Llama-3.1 (README.md:46-49)
    assert cfg3["rope_scaling"]["factor"] == 8.0 and theta3 == 500000.0
    with pytest.raises(NotImplementedError, match="rope scaling"):
        arch_from_hf_config({**base, "rope_parameters": {"rope_theta": 5e5, "rope_type": "yarn", "factor": 4.0}})
    with pytest.raises(NotImplementedError, match="head_dim"):
        arch_from_hf_config({**base, "num_attention_heads": 4})

    sd = {"model.embed_tokens.weight": torch.zeros((4, 8), dtype=torch.bfloat16)}
    for pj, n in (("q", 8), ("k", 4), ("v", 4)):
        sd[f"model.layers.0.self_attn.{pj}_proj.qweight"] = torch.zeros((4, n, 1), dtype=torch.int32)
        sd[f"model.layers.0.self_attn.{pj}_proj.lut3"] = torch.zeros((n, 8), dtype=torch.float16)
        sd[f"model.layers.0.self_attn.{pj}_proj.lut4"] = torch.zeros((n, 16), dtype=torch.float16)
    torch.save(sd, tmp_path / "pytorch_model.bin")
    got = read_checkpoint(str(tmp_path))
    assert got.keys() == sd.keys() and all(torch.equal(got[k], sd[k]) for k in sd)
    conv = convert_state_dict(got, 3)
    assert conv["layers.0.attention.wqkv.qweight"].shape == (3, 16, 1) and conv["layers.0.attention.wqkv.lut"].shape == (16, 8)
    assert conv["tok_embeddings.weight"].dtype == torch.float16
    with pytest.raises(FileNotFoundError):
        read_checkpoint(str(tmp_path / "nothing-here"))


def test_roofline_arithmetic_matches_the_survey_table():
    """the algorithmic-byte figures bench.py's roofline uses == SURVEY.md §8 (shape table and §8d)."""
    from guidedquant_b200.runtime import MODEL_CONFIGS, gemv_algo_bytes, linear_shapes

    assert gemv_algo_bytes(4096, 4096, 2) == 4_194_304 + 32_768 + 8_192 + 8_192 == 4_243_456          # §8d worked example
    s8 = linear_shapes(MODEL_CONFIGS["llama3-8b"])
    assert s8 == {"wqkv": (6144, 4096), "wo": (4096, 4096), "w1w3": (28672, 4096), "w2": (4096, 14336)}
    s70 = linear_shapes(MODEL_CONFIGS["llama3-70b"])
    assert s70 == {"wqkv": (10240, 8192), "wo": (8192, 8192), "w1w3": (57344, 8192), "w2": (8192, 28672)}
    assert linear_shapes(MODEL_CONFIGS["llama2-70b"]) == s70
    per_block = sum(N * K for N, K in s8.values())
    assert per_block == 218_103_808                                                                  # weights per block
    for bits, gb in ((2, 1.745), (3, 2.617), (4, 3.490)):                                             # planes per token
        assert abs(32 * per_block * bits / 8 / 1e9 - gb) < 1e-3
    rows = sum(N for N, _ in s8.values()) * 32
    assert rows == 1_376_256 and rows * 4 * 2 == 11_010_048                                           # 2-bit LUT bytes/token
    assert 2 * MODEL_CONFIGS["llama3-8b"]["vocab"] * 4096 == 1_050_673_152                           # fp16 lm_head
    planes = {n: 2 * N * K // 8 for n, (N, K) in s8.items()}
    assert planes == {"wqkv": 6_291_456, "wo": 4_194_304, "w1w3": 29_360_128, "w2": 14_680_064}


def test_llama3_rope_table_matches_transformers():
    """the llama3 frequency rescaling (what the reference gets from ROPE_INIT_FUNCTIONS["llama3"], inference/model.py:353)
    against transformers' own init function, and the default table against the closed form."""
    import torch
    from guidedquant_b200.convert import rope_inv_freq

    l3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}
    ours = rope_inv_freq(500000.0, l3, 128)
    base = 1.0 / (500000.0 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))
    assert torch.equal(rope_inv_freq(500000.0, None, 128), base)
    assert torch.equal(ours[:20], base[:20]) and torch.allclose(ours[-1], base[-1] / 8.0)   # high freqs kept, low ones / factor
    try:
        from transformers import LlamaConfig
        from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    except Exception:
        pytest.skip("transformers rope utilities not importable")
    try:
        cfg = LlamaConfig(hidden_size=4096, num_attention_heads=32, rope_theta=500000.0, rope_scaling=dict(l3),
                          max_position_embeddings=131072)
        ref, scale = ROPE_INIT_FUNCTIONS["llama3"](cfg, "cpu")
    except Exception as e:  # config spelling differs between transformers versions
        pytest.skip(f"transformers llama3 rope init not callable here: {e}")
    assert scale == 1.0 and torch.allclose(ours, ref.float(), rtol=1e-6, atol=0)


def test_prefill_route_selection_follows_the_measured_crossover():
    """ap_gemv.prefill_prefers_fused: fused tcgen05 kernel for few tokens / large matrices, dequant + matmul beyond the
    measured cross-over (profiles/r2_prefill_fused_vs_dequant_matmul.jsonl) and for shapes the kernel does not take"""
    import json
    import os

    import torch

    from guidedquant_b200 import ap_gemv

    def q(bits, N, K):
        return torch.empty((bits, N, K // 32), dtype=torch.int32, device="meta")

    assert ap_gemv.prefill_prefers_fused(q(2, 4096, 4096), 2, 64)
    assert not ap_gemv.prefill_prefers_fused(q(2, 4096, 4096), 2, 2048)
    assert ap_gemv.prefill_prefers_fused(q(2, 28672, 4096), 2, 1024)
    assert not ap_gemv.prefill_prefers_fused(q(4, 28672, 4096), 4, 1024)
    assert not ap_gemv.prefill_prefers_fused(q(5, 4096, 4096), 5, 16)       # bits > 4
    assert not ap_gemv.prefill_prefers_fused(q(2, 4096, 4096 + 128), 2, 16)  # K % 256 != 0
    # the table agrees with the committed measurements: wherever it picks the fused kernel, the kernel was not slower
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_prefill_fused_vs_dequant_matmul.jsonl")
    rows = [json.loads(l) for l in open(path)]
    assert len(rows) >= 90
    picked = [r for r in rows if ap_gemv.prefill_prefers_fused(q(r["bits"], r["N"], r["K"]), r["bits"], r["T"])]
    assert len(picked) >= 40
    assert all(r["speedup"] >= 0.95 for r in picked), [r for r in picked if r["speedup"] < 0.95]


def test_prefill_plan_budgets_hold_for_every_shape():
    """apg_prefill_plan (host-side planner of the tcgen05 prefill kernel): token tile, ring depth, shared memory, TMEM columns
    and split-K scratch stay within the hardware budgets for every (T, N, K, bits); unsupported shapes are refused"""
    import ctypes

    from guidedquant_b200 import _lib

    L = _lib.lib()
    plan = (ctypes.c_uint32 * 8)()
    need = ctypes.c_uint64(0)
    n_ok = 0
    for bits in (2, 3, 4):
        tbl = {2: 8192, 3: 32768, 4: 8192}[bits]
        for K in (256, 1024, 4096, 11008, 14336, 28672):
            for N in (2, 128, 200, 4096, 6144, 28672):
                for T in (1, 9, 16, 33, 64, 128, 129, 256, 257, 300, 1000, 2048, 5000):
                    assert L.apg_prefill_plan(T, N, K, bits, 148, plan, ctypes.byref(need)) == 0, (T, N, K, bits)
                    t_tile, tok_tiles, row_tiles, splits, stages, smem, tmem, sb = list(plan)
                    assert t_tile % 32 == 0 and 32 <= t_tile <= 256 and tok_tiles * t_tile >= T > (tok_tiles - 1) * t_tile
                    assert row_tiles * 128 >= N > (row_tiles - 1) * 128 and sb == K // 256
                    assert 2 <= stages <= 8 and t_tile + stages * 32 <= tmem <= 512 and tmem & (tmem - 1) == 0
                    assert 2048 + tbl + stages * t_tile * 128 <= smem <= 227 * 1024 - 1024
                    assert 1 <= splits <= max(1, sb // 2) and (splits == 1 or row_tiles * tok_tiles * splits <= 148)
                    assert need.value == (splits * T * N * 4 if splits > 1 else 0)
                    if t_tile <= 128:
                        assert tmem <= 256  # two CTAs per SM
                    n_ok += 1
    assert n_ok == 3 * 6 * 6 * 13
    for (T, N, K, bits) in ((16, 128, 128, 2), (16, 128, 4096 + 32, 3), (16, 128, 4096, 5), (0, 128, 4096, 2)):
        assert L.apg_prefill_plan(T, N, K, bits, 148, plan, ctypes.byref(need)) != 0
