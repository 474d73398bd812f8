#!/usr/bin/env python
"""debug aid: run the persistent engine under torchrun and report the device-side watchdog word if a launch fails"""
import os, sys, traceback
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from guidedquant_b200.runtime import ApGemvChain
from guidedquant_b200.model import APTransformer
model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
objs = []
def report(tag, prog):
    e = int(prog.err[0]) & 0xFFFFFFFF
    print(f"[rank {rank}] {tag}: err word {e:#x} code {e & 0xff} cta {(e >> 8) & 0xfff} thread {e >> 20} epoch? ", flush=True)
try:
    ch = ApGemvChain(model, bits=2, n_layer=layers, world_size=world, rank=rank, process_group=dist.group.WORLD, engine="persistent")
    ch.capture()
    for i in range(steps):
        ch.step()
        if i % 20 == 19:
            ch.stream.synchronize()
            print(f"[rank {rank}] chain step {i} done", flush=True)
    ch.stream.synchronize()
    report("chain ok", ch.prog)
    dist.barrier()
    torch.cuda.synchronize()
    prog_keep = ch.prog
    ch.graph = None
    del ch
    torch.cuda.empty_cache()
except Exception:
    traceback.print_exc()
    report("chain FAILED", ch.prog)
    os._exit(1)
try:
    tf = APTransformer(model, bits=2, max_seq_len=512, n_layer=layers, world_size=world, rank=rank, process_group=dist.group.WORLD, engine="persistent").random_init()
    tf.capture()
    tf.reset(1)
    for i in range(steps):
        tf.step()
        if i % 20 == 19:
            tf.stream.synchronize()
            print(f"[rank {rank}] decode step {i} done", flush=True)
    tf.stream.synchronize()
    report("decode ok", tf.prog)
    print(f"[rank {rank}] tokens", tf.history[:6].cpu().tolist(), flush=True)
except Exception:
    traceback.print_exc()
    report("decode FAILED", tf.prog)
    os._exit(1)
sys.stdout.flush()
os._exit(0)
