"""Runs the reference's OWN converter (/root/reference/inference/sqllm_llama_convert_fuse.py) on a seeded synthetic
HF-named checkpoint and records name -> (shape, dtype, sha1) of its output as tests/golden/convert_golden.json.
Build-container only.  The directory is called Llama-2-7b-* because the reference derives the layer count (32) from the
directory name (:62-69)."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
from tests.convert_fixture import make_hf_checkpoint, digest  # noqa: E402

BITWIDTH = 3
with tempfile.TemporaryDirectory() as tmp:
    d = os.path.join(tmp, "Llama-2-7b-synthetic")
    os.makedirs(d)
    torch.save(make_hf_checkpoint(), os.path.join(d, "pytorch_model.bin"))
    subprocess.check_call([sys.executable, os.path.join(REF, "inference", "sqllm_llama_convert_fuse.py"), "--ckpt_dir", d,
                           "--bitwidth", str(BITWIDTH)], cwd=tmp)
    out = torch.load(os.path.join(d, "converted_pytorch_model.bin"), weights_only=True)
json.dump({"bitwidth": BITWIDTH, "tensors": {k: digest(v) for k, v in out.items()}},
          open(os.path.join(ROOT, "tests", "golden", "convert_golden.json"), "w"), indent=0, sort_keys=True)
print("wrote convert_golden.json with", len(out), "tensors")
