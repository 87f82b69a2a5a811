"""pimc_sharded_evaluate (csrc/comm.cc -> pimc_internal_evaluate_many, csrc/capi.cu): the whole-path evaluations of several
actions of one context run side by side -- the first on the context's stream, the others on side streams forked from it and
joined back -- into one buffer, then ONE all-reduce (a no-op on one rank).  Same kernels, same order of every sum: the
result must equal the action-by-action calls bit for bit, eagerly and replayed from a captured CUDA graph."""
import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pp_action", ["IlkkaPairAction", "BarePairAction", "DavidPairAction"])
def test_actions_evaluated_side_by_side_equal_the_sequential_calls(pp_action):
    import torch
    from simpimc_b200 import sharded
    cfg = S.plasma_config(Ne=40, Np=36, M=32, pp_action=pp_action)       # e-e, e-p, p-p: three kernels of different length
    C = 3
    sp = sharded.ShardedPath(cfg, C, 0, 0, 1)
    for s in range(2):
        sp.SetPositions(s, np.stack([S.synthetic_paths(cfg, s, c, 21) for c in range(C)]))
    n_act = len(sp.pair_actions)
    assert n_act == 3
    for which, name in ((1, "DActionDBeta"), (2, "Potential"), (0, "TotalAction")):
        ref = np.stack([getattr(a, name)() for a in sp.pair_actions])     # one after the other on the context's stream
        out = torch.full((n_act, C), np.nan, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()      # torch filled it on its own stream; the library works on the context's
        from simpimc_b200 import capi
        capi.check(sp.lib.pimc_sharded_evaluate(sp.path.h, sp.comm, which, sp._act_handles, n_act, out.data_ptr()))
        sp.path.Sync()
        assert np.array_equal(out.cpu().numpy(), ref), (pp_action, name, out.cpu().numpy(), ref)
    # the captured step (rho_k rebuilds, the three evaluations forked and joined, the all-reduce) replays to the same numbers
    ref = np.stack([a.DActionDBeta() for a in sp.pair_actions])
    out = torch.full((n_act, C), np.nan, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    replay = sp.CaptureStep(out)
    assert replay.n_nodes >= 2 + 3 * 2
    for _ in range(3):
        out.fill_(np.nan)
        torch.cuda.synchronize()
        replay()
        sp.path.Sync()
        assert np.array_equal(out.cpu().numpy(), ref), (pp_action, "graph replay")
    # per-kernel timing switches the side streams off (an event pair would time a kernel's wait for SMs): same numbers
    sp.path.SetTiming(True)
    out.fill_(np.nan)
    torch.cuda.synchronize()
    sp.DActionDBetaAllDevice(out)
    sp.path.Sync()
    sp.path.SetTiming(False)
    assert np.array_equal(out.cpu().numpy(), ref)
    sp.close()
