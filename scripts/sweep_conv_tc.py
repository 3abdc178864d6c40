"""(MB, BN) sweep of conv_tc over the small-K / small-N convolution shapes of a frame (depth stencil, 1x1 inputs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, ops
from bench_conv_tc import run

SWEEP = {  # (cin, cout, k): [(mb, bn), ...]
    (6, 192, 3): [(2, 96), (2, 64), (1, 64), (1, 96), (2, 192), (1, 192)],
    (12, 384, 3): [(2, 128), (2, 64), (1, 128), (1, 64), (2, 192), (1, 192)],
    (24, 768, 3): [(2, 128), (2, 64), (1, 128), (2, 256), (1, 256), (2, 192)],
    (48, 1536, 3): [(2, 128), (2, 64), (1, 128), (2, 256), (1, 256), (2, 192)],
    (192, 6, 3): [(2, 16), (1, 16)],
    (384, 12, 3): [(2, 16), (1, 16)],
    (768, 24, 3): [(2, 32), (1, 32)],
    (1536, 48, 3): [(2, 48), (1, 48)],
    (16, 320, 1): [(2, 64), (1, 64), (2, 160), (1, 160), (2, 32)],
    (48, 320, 1): [(2, 64), (1, 64), (2, 160), (1, 160)],
    (16, 256, 3): [(2, 128), (1, 128), (2, 64), (2, 256)],
}
if __name__ == "__main__":
    keys = list(SWEEP)
    sel = [int(a) for a in sys.argv[1:]] or range(len(keys))
    for i in sel:
        cin, cout, k = keys[i]
        for mb, bn in SWEEP[keys[i]]:
            try:
                run(cin, cout, k, 512, 512, mb, bn, act=int(os.environ.get('SWEEP_ACT', ops.ACT_ELU)), reps=3)
            except Exception as e:
                print("FAIL", cin, cout, k, mb, bn, str(e)[:100], flush=True)
