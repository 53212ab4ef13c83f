"""BASELINE's small configs only (bench.py's per_config leg): us per step as one CUDA graph of 200 steps.
    python tools/bench_small.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from po_brax_b200 import envs
dev = torch.device('cuda', 0)
for name, m in (('ant_heavenhell', 128), ('ant', 4096), ('ant_gather', 16384), ('ant_tag', 65536), ('ant_heavenhell', 131072)):
    c = bench.bench_small_config(torch, envs, name, m, 70.8, dev)
    print(f"{name} {m}: {c['us_per_step']:.2f} us/step (plain launches {c['us_per_step_plain_launches']:.2f}), frac {c['frac']:.3f}", flush=True)
