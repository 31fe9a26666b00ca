"""Streaming C-ABI (glc_stream_*): nodes submitted in chunks while the device runs time slices and finished nodes are
collected must end up exactly where the one-shot batch call puts them."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pops", [512, 0])
def test_stream_equals_batch(pops):
    """pops = 512: machine slices of 512 pops per warp; pops = 0: adaptive ticks (machine slice or lane pass with refill)."""
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params()
    n, chunk = 9000, 3000
    props, flags, t_end = synthetic.standard_nodes(p, n, seed=777)
    ev = Evolver(0)
    synthetic.install(ev, p)
    pb, fb = props.copy(), flags.copy()
    sb, ib, cb = ev.evolve_batch(pb, fb, t_end)

    ev.stream_begin(n)
    got = {}
    done_seen = 0
    for k in range(0, n, chunk):
        first = ev.stream_submit(np.ascontiguousarray(props[k:k + chunk]), flags[k:k + chunk], t_end[k:k + chunk])
        assert first == k  # tickets are consecutive
        for _ in range(3):
            done, c = ev.stream_run(pops)
            assert done >= done_seen
            done_seen = done
            t, pr, fl, st, it = ev.stream_collect(1000)
            for j, tk in enumerate(t):
                assert int(tk) not in got
                got[int(tk)] = (pr[j].copy(), fl[j], st[j], it[j])
    c = ev.stream_finish()
    while True:
        t, pr, fl, st, it = ev.stream_collect(4096)
        if len(t) == 0:
            break
        for j, tk in enumerate(t):
            assert int(tk) not in got
            got[int(tk)] = (pr[j].copy(), fl[j], st[j], it[j])
    ev.stream_end()
    assert sorted(got) == list(range(n))  # every ticket collected exactly once
    ps = np.stack([got[i][0] for i in range(n)])
    assert np.array_equal(ps, pb), "streamed records differ from the batch call"
    assert np.array_equal(np.array([got[i][1] for i in range(n)]), fb)
    assert np.array_equal(np.array([got[i][2] for i in range(n)]), sb)
    assert np.array_equal(np.array([got[i][3] for i in range(n)]), ib)
    assert c == cb  # same accepted/rejected steps, RHS evaluations, segments
    ev.close()


def test_stream_session_guards(oracle_lib):
    """A batch call during a streaming session is refused (GLC_ERR_BUSY: the session owns the arena); a session may be
    continued after glc_stream_finish (slots released by the drain kernel are re-armed) and gives the batch results."""
    from galacticus_b200.evolver import Evolver, GlcError

    p = cases.standard_params(with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 3000, seed=41)
    ev = Evolver(0)
    synthetic.install(ev, p)
    ref_p, ref_f = props.copy(), flags.copy()
    ref_s, ref_i, _ = ev.evolve_batch(ref_p, ref_f, t_end)
    ev.stream_begin(4 * props.shape[0])
    with pytest.raises(GlcError):
        ev.evolve_batch(props.copy(), flags.copy(), t_end)
    half = props.shape[0] // 2
    ev.stream_submit(props[:half], flags[:half], t_end[:half])
    ev.stream_finish()
    ev.stream_submit(props[half:], flags[half:], t_end[half:])  # after finish: the drained slots must fetch again
    ev.stream_finish()
    tickets, rows, fl, st, it = ev.stream_collect(props.shape[0])
    ev.stream_end()
    order = np.argsort(tickets)
    assert tickets.size == props.shape[0]
    np.testing.assert_array_equal(st[order], ref_s)
    np.testing.assert_array_equal(fl[order], ref_f)
    assert np.array_equal(rows[order], ref_p), "streamed-after-finish records differ from the batch call"
    ev.close()
