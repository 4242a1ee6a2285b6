"""Host logic of the opt-in L2 weight look-ahead (ops.WeightLookahead): record a launch sequence, then hand launch i the
weights of launch i + distance through ea_gemm_args.prefetch.  No GPU: only the ctypes structs are filled."""
import torch

from editanything_b200 import _lib as L
from editanything_b200.ops import WeightLookahead


def _args(n, m):
    gs = [L.GemmArgs() for _ in range(n)]
    for g in gs:
        g.M = m
    return gs


def test_record_then_replay_wraps_and_filters():
    big = [torch.zeros(1 << 20, dtype=torch.float16) for _ in range(4)]     # 2 MB each
    small = torch.zeros(16, dtype=torch.float16)
    seq = [([big[0]], 8192), ([big[1], big[2], big[3]], 128), ([small], 128)]
    la = WeightLookahead(distance=1, max_m=512)
    for ws, m in seq:
        la.visit(_args(len(ws), m), ws)
    assert len(la.seq) == 3
    la.replay()
    # launch 0 (1 group) -> the three weights of launch 1 land in its three slots
    g0 = _args(1, 8192)
    la.visit(g0, [big[0]])
    assert [g0[0].prefetch[k] for k in range(3)] == [w.data_ptr() for w in big[1:]]
    assert [g0[0].prefetch_bytes[k] for k in range(3)] == [2 << 20] * 3
    # launch 1 (3 groups) -> launch 2's operand is below MIN_BYTES: nothing set
    g1 = _args(3, 128)
    la.visit(g1, big[1:])
    assert all(not g.prefetch[k] for g in g1 for k in range(3))
    # launch 2 wraps to launch 0, whose M = 8192 exceeds max_m: filtered
    g2 = _args(1, 128)
    la.visit(g2, [small])
    assert not g2[0].prefetch[0]


def test_targets_spread_over_group_slots():
    big = [torch.zeros(1 << 20, dtype=torch.float16) for _ in range(3)]
    la = WeightLookahead(distance=1)
    la.visit(_args(3, 128), big)
    la.visit(_args(3, 128), big)
    la.replay()
    gs = _args(3, 128)
    la.visit(gs, big)
    assert [g.prefetch[0] for g in gs] == [w.data_ptr() for w in big]     # one range per group, slot 0
    assert all(not g.prefetch[1] for g in gs)
