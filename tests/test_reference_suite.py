"""The reference's OWN test module for the hot path (tests/test_string.py of
sdrobert/pydrobert-pytorch, unmodified) run against the B200 kernels through
``b200lev.install()``: nojit, ``torch.jit.trace`` and ``torch.jit.script`` variants of every
case (310 per device).

The reference's files are not part of this repository: ``oracle/make_ref.sh`` (run by
``__graft_entry__.build()`` where /root/reference exists) installs the package into
``baseline/_ref`` and copies test_string.py / conftest.py into ``oracle/_ref/tests``; both are
git-ignored and travel to the GPU box with the snapshot.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "baseline", "_ref")
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
EXPECTED_CASES = 310  # per device: what `pytest tests/test_string.py -m cpu` collects upstream
# tests/test_decoding.py: the step functions of the N-best producers AND the reference's own
# BeamSearch / RandomWalk modules running on them (install() rebinds the names their loops call)
DECODING_SELECT = "beam_search or random_walk"
EXPECTED_DECODING = 9  # per device (+ 3 cases the reference itself marks xfail under trace)
EXPECTED_ADVANCE = 2  # test_beam_search_advance_greedy, test_beam_search_advance
EXPECTED_CTC = 38  # test_ctc_greedy_search (nojit / trace / script) + test_ctc_greedy_search_ignores_padding
EXPECTED_SEQLP = 12  # test_sequence_log_probs: tensor (4 step axes) and PackedSequence inputs x nojit / trace / script


def _run(marker, test_file="test_string.py", select=None, expected=EXPECTED_CASES, workers=0):
    if not (os.path.isfile(os.path.join(REF_TESTS, test_file))
            and os.path.isdir(os.path.join(REF_PKG, "pydrobert"))):
        pytest.skip("reference copy absent (oracle/make_ref.sh needs /root/reference)")
    env = dict(os.environ)
    env.pop("B200LEV_BVSHORT_MIN_PAIRS", None)  # the reference's own suite runs at the library's defaults
    env["PYTHONPATH"] = os.pathsep.join([REF_PKG, os.path.join(ROOT, "tests")]
                                        + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_suite_plugin", "-p", "no:cacheprovider", "-q",
           "-W", "ignore", "--rootdir", REF_TESTS, "-c", os.path.join(REF_TESTS, "pytest.ini"),
           os.path.join(REF_TESTS, test_file), "-m", marker] + (["-k", select] if select else [])
    if workers > 1:
        # the emulator runs one kernel thread: spread the cases over single-threaded workers
        # (eight torch threads per worker would fight over the cores instead)
        try:
            import xdist  # noqa: F401
            cmd += ["-n", str(workers)]
            env["OMP_NUM_THREADS"] = env["MKL_NUM_THREADS"] = "1"
        except ImportError:
            pass
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=REF_TESTS, timeout=3000)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) == expected, tail
    assert not re.search(r"\b\d+ failed", r.stdout.splitlines()[-1]), tail


def test_reference_suite_on_emulator():
    """-m cpu: host layer (script / trace / shapes / messages) + kernel logic on the emulator."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the gpu-marked run below is the gate")
    _run("cpu", workers=min(6, os.cpu_count() or 1))
    # (the step-function cases only: the reference's BeamSearch / RandomWalk module cases train a
    # language model first, which takes minutes on the emulator; the GPU run below has them all)
    _run("cpu", "test_decoding.py", "beam_search_advance or random_walk_advance", EXPECTED_ADVANCE)
    _run("cpu", "test_decoding.py", "ctc_greedy", EXPECTED_CTC)
    _run("cpu", "test_decoding.py", "sequence_log_probs", EXPECTED_SEQLP)


@pytest.mark.gpu
def test_reference_suite_on_b200():
    """-m gpu: the same 310 cases with CUDA tensors on the real kernels."""
    _run("gpu")


@pytest.mark.gpu
def test_reference_decoding_steps_on_b200():
    """-m gpu: the reference's beam-search / random-walk tests (step functions and its own
    BeamSearch / RandomWalk modules) on the step kernels."""
    _run("gpu", "test_decoding.py", DECODING_SELECT, EXPECTED_DECODING)


@pytest.mark.gpu
def test_reference_ctc_greedy_search_on_b200():
    """-m gpu: the reference's CTCGreedySearch tests (plain, traced, scripted) on lev_ctc kernels."""
    _run("gpu", "test_decoding.py", "ctc_greedy", EXPECTED_CTC)


@pytest.mark.gpu
def test_reference_sequence_log_probs_on_b200():
    """-m gpu: the reference's SequenceLogProbabilities tests (tensor and PackedSequence logits; plain,
    traced, scripted) on lev_seqlp kernels."""
    _run("gpu", "test_decoding.py", "sequence_log_probs", EXPECTED_SEQLP)
