"""Development aid: precise-mode forward with the CTA-pair ff conv (M2T_VAR_W2_PAIR) against the default two-CTAs-per-tile form."""
import sys, types, torch
sys.path.insert(0, ".")
from m2trans_b200 import _lib
from m2trans_b200.M2Trans_network import M2Trans
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict
def model(scale, var, nb=8):
    m = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=nb, kernel_variant=var)).cuda()
    m.load_state_dict(synthetic_state_dict(scale, 3, n_blocks=nb))
    return m
for scale, shape in ((2, (1, 3, 32, 32)), (2, (1, 3, 32, 40)), (3, (2, 3, 33, 47)), (2, (3, 3, 96, 72)), (4, (2, 3, 64, 64))):
    x = synthetic_input(shape[0], shape[2], shape[3], seed=5).cuda()
    ya = model(scale, _lib.VAR_PRECISE_ON | _lib.VAR_W2_PAIR)(x)
    yb = model(scale, _lib.VAR_PRECISE_ON)(x)
    torch.cuda.synchronize()
    print(scale, shape, "pair vs split max-abs", float((ya - yb).abs().max()), "equal", bool(torch.equal(ya, yb)), flush=True)
