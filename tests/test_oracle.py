"""Pin the CPU oracle (oracle/) against the reference.

* golden fixtures produced by the unmodified reference (tests/golden/make_golden.py),
* the reference's own known-answer vectors (its tests/test_string.py:115-130,171-192
  and the docstring example at _string.py:1108-1143),
* the NIST sclite fixture (its tests/sclite/*).

Integer / dyadic costs: bit-exact.  Non-dyadic costs: 1e-6 relative (the oracle
evaluates the deletion term in the reference's operation order, so in practice these
are bit-exact too, but the bar that is enforced is the stated one).
"""
import numpy as np
import pytest

from oracle import oracle as O

FLOAT_COSTS = {(0.7, 1.1, 1.3), (0.3, 0.3, 0.3), (1.1, 0.9, 1.7)}


def _kw(p):
    return dict(eos=p["eos"], include_eos=p["include_eos"], batch_first=p["batch_first"],
                ins_cost=p["ins"], del_cost=p["del"], sub_cost=p["sub"])


def _check(exp, act, exact):
    assert exp.shape == act.shape, (exp.shape, act.shape)
    if exact:
        assert np.array_equal(exp, act), np.abs(exp - act).max()
    else:
        np.testing.assert_allclose(act, exp, rtol=1e-6, atol=0)


def test_golden_case_count(golden_sm):
    assert len(golden_sm.params) >= 80


@pytest.mark.parametrize("func", ["error_rate", "edit_distance"])
def test_final_vs_golden(golden_sm, func):
    n = 0
    for case, p in golden_sm.params.items():
        if not golden_sm.has(case, func):
            continue
        exact = (p["ins"], p["del"], p["sub"]) not in FLOAT_COSTS
        act = getattr(O, func)(golden_sm.get(case, "ref"), golden_sm.get(case, "hyp"),
                               norm=p["norm"], **_kw(p))
        _check(golden_sm.get(case, func), act, exact)
        n += 1
    assert n >= 80


@pytest.mark.parametrize("func", ["prefix_error_rates", "prefix_edit_distances"])
def test_prefix_vs_golden(golden_sm, func):
    n = 0
    for case, p in golden_sm.params.items():
        if not golden_sm.has(case, func):
            continue
        exact = (p["ins"], p["del"], p["sub"]) not in FLOAT_COSTS
        act = getattr(O, func)(golden_sm.get(case, "ref"), golden_sm.get(case, "hyp"),
                               norm=p["norm"], padding=p["padding"],
                               exclude_last=p["exclude_last"], **_kw(p))
        _check(golden_sm.get(case, func), act, exact)
        n += 1
    assert n >= 70


def test_mask_and_completion_vs_golden(golden_sm):
    n = 0
    for case, p in golden_sm.params.items():
        if not golden_sm.has(case, "mask"):
            continue
        ref, hyp = golden_sm.get(case, "ref"), golden_sm.get(case, "hyp")
        m = O.completion_mask(ref, hyp, exclude_last=p["exclude_last"], **_kw(p))
        assert np.array_equal(m, golden_sm.get(case, "mask")), case
        if golden_sm.has(case, "optimal_completion"):
            oc = O.optimal_completion(ref, hyp, padding=p["padding"],
                                      exclude_last=p["exclude_last"], **_kw(p))
            exp = golden_sm.get(case, "optimal_completion")
            assert oc.shape == exp.shape and np.array_equal(oc, exp), case
        n += 1
    assert n >= 55


def test_fast_del_matches_exact_del_on_integer_costs(golden_sm):
    """The sequential deletion update equals the reference's min-over-k form whenever
    the fp32 sums are exact (integer costs)."""
    for case, p in golden_sm.params.items():
        c = (p["ins"], p["del"], p["sub"])
        if any(x != round(x) for x in c):
            continue
        ref, hyp = golden_sm.get(case, "ref"), golden_sm.get(case, "hyp")
        a = O.edit_distance(ref, hyp, norm=False, exact_del=True, **_kw(p))
        b = O.edit_distance(ref, hyp, norm=False, exact_del=False, **_kw(p))
        assert np.array_equal(a, b)


# reference tests/test_string.py:179-192
KNOWN_PAIRS = (
    ((1, 2, 3), (1, 2, 3), 0),
    ((2, 3), (1, 2, 3), 1),
    ((1, 3), (1, 2, 3), 1),
    ((3,), (1, 2, 3), 2),
    ((1, 2, 3), (1, 3), 1),
    ((1, 2, 3), (1, 2), 1),
    ((1, 2, 3), (1,), 2),
    ((1, 3, 1, 2, 3), (1, 2, 3), 2),
    ((1, 2, 3), (4, 5, 6), 3),
    ((2, 2, 2), (2,), 2),
    (tuple(), (1,), 1),
    (tuple(), tuple(), 0),
)


def _pad(seqs, value, batch_first):
    T = max(len(s) for s in seqs)
    out = np.full((len(seqs), T), value, np.int64)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = s
    return out if batch_first else np.ascontiguousarray(out.T)


@pytest.mark.parametrize("include_eos", [0, 1])
@pytest.mark.parametrize("batch_first", [True, False])
@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("func", ["edit_distance", "error_rate"])
def test_known_pairs(include_eos, batch_first, norm, func):
    eos = 0
    ref = _pad([x[0] + (eos,) * include_eos for x in KNOWN_PAIRS], eos, batch_first)
    hyp = _pad([x[1] + (eos,) * include_eos for x in KNOWN_PAIRS], eos, batch_first)
    rl = np.array([len(x[0]) + include_eos for x in KNOWN_PAIRS], np.float32)
    hl = np.array([len(x[1]) + include_eos for x in KNOWN_PAIRS], np.float32)
    exp = np.array([x[2] for x in KNOWN_PAIRS], np.float32)
    if norm:  # empty-ref rule, tests/test_string.py:206-207
        exp = np.where(rl == 0, (hl != 0).astype(np.float32), exp / np.maximum(rl, 1))
    act = getattr(O, func)(ref, hyp, eos=eos, norm=norm, include_eos=bool(include_eos),
                           batch_first=batch_first)
    np.testing.assert_allclose(act, exp, rtol=1e-6)


# reference tests/test_string.py:120-130
TRIPLETS = (
    ("sunday#", "saturday#", ["s", "u", "un", "und", "n", "nd", "a", "y", "#", ""]),
    ("sunday#", "satrapy#", ["s", "u", "un", "und", "unda", "y", "y#", "#", ""]),
    ("abc#", "abc#", ["a", "b", "c", "#", ""]),
    ("foot#", "bot#", ["f", "fo", "o", "ot#", ""]),
    ("abc#", "def#", ["a", "ab", "abc", "abc#", ""]),
)


@pytest.mark.parametrize("include_eos", [True, False])
@pytest.mark.parametrize("batch_first", [True, False])
@pytest.mark.parametrize("exclude_last", [True, False])
def test_known_completions(include_eos, batch_first, exclude_last):
    eos, padding = ord("#"), -1
    ref = _pad([[ord(c) for c in w] for w, _, _ in TRIPLETS], padding, batch_first)
    hyp = _pad([[ord(c) for c in w] for _, w, _ in TRIPLETS], eos, batch_first)
    act = O.optimal_completion(ref, hyp, eos=eos, include_eos=include_eos,
                               batch_first=batch_first, padding=padding,
                               exclude_last=exclude_last)
    if not batch_first:
        act = act.transpose(1, 0, 2)
    for act_bt, (_, _, exp_bt) in zip(act, TRIPLETS):
        if not include_eos:
            exp_bt = [s.replace("#", "") for s in exp_bt[:-1]]
        if exclude_last:
            exp_bt = exp_bt[:-1]
        assert act_bt.shape[0] >= len(exp_bt)
        assert (act_bt[len(exp_bt):] == padding).all()
        for a, e in zip(act_bt, exp_bt):
            got = [chr(i) for i in a[a != padding].tolist()]
            assert got == sorted(set(e))  # ascending and duplicate free (SM:503-507)


def test_docstring_foot_bot():
    """_string.py:1108-1143"""
    ref = np.array([[ord(c)] for c in "foot"], np.int64)
    hyp = np.array([[ord(c)] for c in "bot"], np.int64)
    oc = O.optimal_completion(ref, hyp)[:, 0]
    got = ["".join(chr(i) for i in row[row >= 0]) for row in oc]
    assert got == ["f", "fo", "o", "ot"]


def test_uniform_cost_shortcut():
    """SM:168-174: rate callers keep mult == 1, distance callers scale by the cost."""
    ref = np.array([[ord(c)] for c in "kitten"], np.int64)
    hyp = np.array([[ord(c)] for c in "sitting"], np.int64)
    assert O.error_rate(ref, hyp, norm=False, ins_cost=2, del_cost=2, sub_cost=2)[0] == 3.0
    assert O.edit_distance(ref, hyp, ins_cost=2, del_cost=2, sub_cost=2)[0] == 6.0


def test_sclite(golden_sclite):
    """command_line.py:1124-1147 on tests/sclite: NIST costs, 3-decimal agreement."""
    z = golden_sclite
    ers = O.error_rate(z["ref"], z["hyp"], eos=-1, include_eos=False, norm=False,
                       ins_cost=3.0, del_cost=3.0, sub_cost=4.0)
    assert np.array_equal(ers, z["errs"])
    for k in range(len(ers)):
        assert f"{ers[k] / z['ref_lens'][k]:.03f}" == f"{z['per_utt'][k]:.03f}"
    assert f"{ers.sum() / z['ref_lens'].sum():.03f}" == f"{float(z['total']):.03f}"


def test_ocd_loss_vs_golden(golden_loss):
    n = 0
    for case, p in golden_loss.params.items():
        if not case.startswith("ocd"):
            continue
        g = golden_loss
        w = g.get(case, "weight") if p["weight"] else None
        go = g.get(case, "grad_output") if g.has(case, "grad_output") else None
        loss, grad = O.hard_optimal_completion_distillation_loss(
            g.get(case, "logits"), g.get(case, "ref"), g.get(case, "hyp"), eos=p["eos"],
            include_eos=p["include_eos"], batch_first=p["batch_first"], weight=w,
            reduction=p["reduction"], ignore_index=p["ignore_index"], grad_output=go)
        np.testing.assert_allclose(loss, g.get(case, "loss"), rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(grad, g.get(case, "grad"), rtol=2e-5, atol=1e-6)
        n += 1
    assert n == 24


def test_mwer_loss_vs_golden(golden_loss):
    n = 0
    for case, p in golden_loss.params.items():
        if not case.startswith("mwer"):
            continue
        g = golden_loss
        go = g.get(case, "grad_output") if g.has(case, "grad_output") else None
        loss, grad = O.minimum_error_rate_loss(
            g.get(case, "log_probs"), g.get(case, "ref"), g.get(case, "hyp"), eos=p["eos"],
            include_eos=p["include_eos"], sub_avg=p["sub_avg"],
            batch_first=p["batch_first"], norm=p["norm"], ins_cost=p["ins"],
            del_cost=p["del"], sub_cost=p["sub"], reduction=p["reduction"], grad_output=go)
        np.testing.assert_allclose(loss, g.get(case, "loss"), rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(grad, g.get(case, "grad"), rtol=2e-5, atol=1e-6)
        n += 1
    assert n == 48


def test_fill_after_eos():
    """tests/test_string.py:31-52"""
    rng = np.random.default_rng(0)
    T, N, V = 15, 12, 10
    tok = rng.integers(0, V - 1, (T, N))
    tok.reshape(-1)[:: N + 1] = V - 1
    out = O.fill_after_eos(tok, V - 1)
    logits = rng.standard_normal((T, N, V)).astype(np.float32)
    out2 = O.fill_after_eos(tok[:, :, None], V - 1, value=logits)
    for n in range(N):
        assert (out[: n + 1, n] == tok[: n + 1, n]).all()
        assert (out[n:, n] == V - 1).all()
        assert (out2[: n + 1, n] == logits[: n + 1, n]).all()
        assert (out2[n + 1:, n] == V - 1).all()


def test_seqlp_oracle_matches_reference_fixtures(golden_seqlp):
    """oracle.sequence_log_probs (float64 restatement of _decoding.py:1516-1548) against the
    reference's outputs and gradients; the fixture dtype bounds the agreement."""
    import parity_cases as PC

    n = 0
    for name, p in golden_seqlp.params.items():
        rtol, atol = PC.SEQLP_TOL[p["dtype"]]
        out, grad = O.sequence_log_probs(golden_seqlp.get(name, "logits"), golden_seqlp.get(name, "hyp"),
                                         p["dim"], p["eos"], grad_out=golden_seqlp.get(name, "grad_out"))
        np.testing.assert_allclose(out, golden_seqlp.get(name, "out"), rtol=rtol, atol=atol, err_msg=name)
        scale = max(1.0, float(np.abs(golden_seqlp.get(name, "grad_out")).max())) if out.size else 1.0
        np.testing.assert_allclose(grad, golden_seqlp.get(name, "grad"), rtol=rtol, atol=atol * scale,
                                   err_msg=name)
        n += 1
    assert n == 84


def test_ctc_greedy_oracle_against_reference_fixture(golden_ctc):
    """oracle.ctc_greedy_search reproduces every case make_golden.py ran through the
    reference: paths (whole tensor), out_lens, max_."""
    import parity_cases as PC

    for name in golden_ctc.params:
        p, logits, lens = PC._ctc_case(golden_ctc, name)
        max_, paths, out_lens = O.ctc_greedy_search(logits.numpy(), None if lens is None else lens.numpy(),
                                                    p["blank_idx"], p["batch_first"], p["is_probs"])
        np.testing.assert_array_equal(paths, golden_ctc.get(name, "paths"), err_msg=name)
        np.testing.assert_array_equal(out_lens, golden_ctc.get(name, "out_lens"), err_msg=name)
        np.testing.assert_allclose(max_, golden_ctc.get(name, "max").astype(np.float64),
                                   rtol=3e-6 if p["dtype"] == "float32" else 1e-12,
                                   atol=1e-30 if p["is_probs"] else (3e-6 if p["dtype"] == "float32" else 1e-12)
                                   * logits.shape[1 if p["batch_first"] else 0], err_msg=name)
    assert len(golden_ctc.params) == 106


def test_oracle_decode_steps_match_reference(golden_decode):
    """oracle.beam_search_advance / random_walk_advance pinned by the reference's outputs."""
    g = golden_decode
    for name in g.names("beam"):
        p = g.params[name]
        lens = g.get(name, "y_prev_lens") if g.has(name, "y_prev_lens") else None
        y, ln, lp, src = O.beam_search_advance(g.get(name, "log_probs_t"), p["width"], g.get(name, "log_probs_prev"),
                                               g.get(name, "y_prev"), lens)
        K = p["K"]
        assert np.array_equal(y[:, :, :K], g.get(name, "y_next")[:, :, :K]) and y.shape == g.get(name, "y_next").shape
        assert np.array_equal(ln, g.get(name, "y_next_lens")) and np.array_equal(src, g.get(name, "next_src"))
        assert np.array_equal(lp, g.get(name, "log_probs_next"))
    for name in g.names("walk"):
        lens = g.get(name, "y_prev_lens") if g.has(name, "y_prev_lens") else None
        y, lp = O.random_walk_advance(g.get(name, "log_probs_t"), g.get(name, "log_probs_prev"), g.get(name, "y_prev"),
                                      g.get(name, "y_t"), lens)
        assert np.array_equal(y, g.get(name, "y_next"))
        assert np.allclose(lp, g.get(name, "log_probs_next"), rtol=1e-6, atol=1e-7)
