"""The SMALL convolutions of a frame (conditioning-net 2-D convs, sub-network input 1x1, LRNN mean-volume branch) at 512 x 512:
time, bytes moved and the HBM-floor they would have as pure streaming passes (`gbs` = operand + output bytes / time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_conv_tc as B
from cwfa_b200 import ops, tc

SH = [  # cin, cout, k, mb (None = library default), act
    (29, 48, 3, None, ops.ACT_PRELU), (48, 48, 3, None, ops.ACT_NONE if hasattr(ops, "ACT_NONE") else 0), (29, 48, 1, None, 0),
    (29, 6, 3, None, ops.ACT_PRELU), (6, 6, 3, None, 0), (29, 6, 1, None, 0),
    (6, 64, 1, None, 0), (64, 64, 1, None, ops.ACT_GELU), (64, 6, 1, None, 0), (6, 6, 7, None, 0), (64, 64, 7, None, 0),
]
for cin, cout, k, mb, act in SH:
    bn = min(tc.pad16(cout), 256)
    if mb is None:
        mb = 2 if bn * 2 <= 512 else 1
        if bn >= 192 and tc.pad16(cin) <= 64:
            mb = 1
    B.run(cin, cout, k, 512, 512, mb, bn, act=act, reps=4)
