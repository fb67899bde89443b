#!/usr/bin/env python
"""What does each stage cost UNDER GRAPH REPLAY?  Runs the bench's HBM-resident leg (cfg2, schedule S1) with one stage
switched off at a time and prints ms/scan.  (The map evolves differently without a stage, so this is an estimate of the
stage's share of the replayed scan, not an exact decomposition.)  Usage: python tools/exp_stage_marginals.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vofod_b200 import abi, capi, synth  # noqa: E402

W, H, K, WM = bench.W, bench.H, 40, 40


def main():
    v = capi.Vofod(0)
    p = bench.make_params()
    dirs = synth.sim_lut(W, H)
    v.reset(p, bench.VOXEL)
    v.set_sensor(W, H, dirs)
    stream = torch.cuda.ExternalStream(v.stream(), device=torch.device("cuda", 0))
    n = K + WM
    host = np.zeros((n, W * H), dtype=abi.PT_DTYPE)
    poses, rps = [], []
    for k in range(n):
        _, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs, 1.0, out=host[k])
        poses.append(pose)
        rps.append(rp)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dets = np.zeros(256, dtype=abi.DETECTION_DTYPE)
    variants = {"all": {}, "no raycast": {"do_raycast": False}, "no classify": {"do_classify": False}, "no sepclusters": {"do_sepclusters": False},
                "filter+cluster+closefar+update only": {"do_raycast": False, "do_classify": False, "do_sepclusters": False},
                "no side branch": {"overlap": 0}, "pdl in graph": {"pdl": 2}, "raycast block 128": {"rb": 128}, "raycast block 256": {"rb": 256}, "all (again)": {}}
    out = {}
    for name, kw in variants.items():
        kw = dict(kw)
        v.set_option(4, kw.pop("overlap", 1))
        v.set_option(abi.OPT_PDL, kw.pop("pdl", 1))
        v.set_option(5, kw.pop("rb", 64))
        v.reset(p, bench.VOXEL)
        for k in range(n):
            v.upload_scan(k, host[k])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda.synchronize()
        for k in range(n):
            with torch.cuda.stream(stream):
                flush.fill_(k & 0xFF)
                ev[k][0].record(stream)
                v.process_scan_resident(k, poses[k], p, abi.schedule_s1(rps[k], **kw), dets=dets)
                ev[k][1].record(stream)
        torch.cuda.synchronize()
        ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(WM, n)]
        out[name] = {"mean_ms": float(np.mean(ms)), "median_ms": float(np.median(ms))}
        print(name, out[name], flush=True)
    print(json.dumps(out))
    del flush
    torch.cuda.synchronize()
    v.close()


if __name__ == "__main__":
    main()
