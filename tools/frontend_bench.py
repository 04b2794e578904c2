#!/usr/bin/env python3
"""Stand-alone timing of the N2 front-end (bench.py's frontend_bench leg) on one GPU: python tools/frontend_bench.py [streams]"""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import sdr_receiver_dvb_t2_b200 as t2
    from sdr_receiver_dvb_t2_b200 import engine as E
    streams = [int(a) for a in sys.argv[1:]] or [1, 8, 64, 256]
    st = torch.cuda.Stream()
    for n in streams:
        print(json.dumps(bench.frontend_bench(torch, t2, E, 0, st, bench.hbm_peak()[0], n_streams=n)), flush=True)


if __name__ == '__main__':
    main()
