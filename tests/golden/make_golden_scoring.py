#!/usr/bin/env python
"""Golden fixture for the bulk-scoring front end (tests/golden/scoring.json).

Run in the build container only:

    python tests/golden/make_golden_scoring.py

Writes small token data directories (the reference's on-disk format, _datasets.py:64-106)
into a temporary directory, runs the UNMODIFIED reference command
``compute_torch_token_data_dir_error_rates`` (command_line.py:858-1147) on them with a
list of option sets, and stores the corpora (as plain lists) together with what the
command printed -- or the exception it raised -- per option set.  The tests rebuild the
directories from the lists and never import the reference.

The reference's ``_datasets`` module imports ``param``, which this image does not have;
nothing on this command's path uses it, so an empty stand-in module is registered first.
"""
import io
import json
import os
import sys
import tempfile
import types
import warnings

import torch

REF_SRC = os.environ.get("B200LEV_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF_SRC)


class _Param(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "Parameterized":
            return type("Parameterized", (), {})
        if name == "parameterized":
            return self
        return lambda *a, **k: None


try:
    import param  # noqa: F401
except ImportError:
    sys.modules["param"] = _Param("param")

import pydrobert.torch.command_line as CL  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_corpus(seed, n, vocab, max_len, empty_ref):
    g = torch.Generator().manual_seed(seed)
    utts = {}
    for i in range(n):
        utt = f"utt{(i * 7919) % 1000:03d}{'abc'[i % 3]}"
        rl = int(torch.randint(0 if empty_ref else 1, max_len, (), generator=g))
        ref = torch.randint(-2, vocab, (rl,), generator=g).tolist()
        hyp = []
        for t in ref:  # a noisy copy: keep / substitute / delete / insert
            r = float(torch.rand((), generator=g))
            if r < 0.6:
                hyp.append(t)
            elif r < 0.75:
                hyp.append(int(torch.randint(-2, vocab, (), generator=g)))
            elif r < 0.88:
                hyp.extend([t, int(torch.randint(-2, vocab, (), generator=g))])
        if i % 11 == 5:
            hyp = []
        kind = ["flat", "col1", "timed"][i % 3]
        utts[utt] = {"ref": ref, "hyp": hyp, "ref_kind": "flat" if i % 4 else "timed", "hyp_kind": kind}
    return utts


def tensor_of(seq, kind):
    t = torch.tensor(seq, dtype=torch.long)
    if kind == "col1":
        return t.unsqueeze(-1)
    if kind == "timed":
        n = t.shape[0]
        start = torch.arange(n) * 3
        return torch.stack([t, start, start + 2], -1) if n else torch.zeros((0, 3), dtype=torch.long)
    return t


def write_dirs(root, utts, prefix="", suffix=".pt", drop_ref=(), drop_hyp=()):
    for side, drop in (("ref", drop_ref), ("hyp", drop_hyp)):
        os.makedirs(os.path.join(root, side), exist_ok=True)
        for utt, d in utts.items():
            if utt in drop:
                continue
            torch.save(tensor_of(d[side], d[side + "_kind"]), os.path.join(root, side, prefix + utt + suffix))
        # a file the prefix/suffix filter must skip
        torch.save(torch.zeros(2, dtype=torch.long), os.path.join(root, side, "stray.bin"))


def run(dirs, opts):
    out = io.StringIO()
    rec = {}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            # the command writes to a file argument; give it a path and read it back
            with tempfile.NamedTemporaryFile("r", suffix=".txt") as f:
                rc = CL.compute_torch_token_data_dir_error_rates(dirs + [f.name] + opts)
                rec["rc"] = rc
                rec["out"] = open(f.name).read()
        except Exception as e:  # noqa: BLE001
            rec["raises"] = type(e).__name__
            rec["message"] = str(e)
    rec["warned_missing"] = sorted(str(x.message) for x in w if "does not contain" in str(x.message))
    del out
    return rec


def main():
    store = {"corpora": {}, "files": {}, "cases": []}
    A = make_corpus(1, 45, 12, 14, empty_ref=False)
    B = make_corpus(2, 30, 6, 9, empty_ref=True)
    store["corpora"] = {"A": A, "B": B}
    names = sorted(A)
    drop_ref, drop_hyp = [names[3], names[-1]], [names[0], names[10]]
    id2token = "".join(f"{i} w{abs(i) % 9 if i != 4 else 3}\n" for i in range(-2, 12))  # w3 is ambiguous
    id2token_swapped = "".join(f"w{abs(i) % 9 if i != 4 else 3} {i}\n" for i in range(-2, 12))
    id2token_short = "".join(f"{i} w{i}\n" for i in range(-2, 8))
    files = {
        "id2token.txt": id2token, "id2token_swapped.txt": id2token_swapped,
        "id2token_short.txt": id2token_short,
        "replace_int.txt": "3 5\n-1 7\n9 100\n", "ignore_int.txt": "0 5\n-2\n",
        "replace_tok.txt": "w1 w2\nw8 sil\n", "ignore_tok.txt": "w0 sil\n",
        "replace_bad.txt": "w1 w2\n",
    }
    store["files"] = files
    store["missing"] = {"drop_ref": drop_ref, "drop_hyp": drop_hyp}
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in files.items():
            open(os.path.join(tmp, name), "w").write(text)
        write_dirs(os.path.join(tmp, "A"), A)
        write_dirs(os.path.join(tmp, "B"), B)
        write_dirs(os.path.join(tmp, "Apre"), A, prefix="tok_", suffix=".tok")
        write_dirs(os.path.join(tmp, "Amiss"), A, drop_ref=drop_ref, drop_hyp=drop_hyp)
        cases = [
            ("A", []), ("A", ["--per-utt"]), ("A", ["--distances"]), ("A", ["--distances", "--per-utt"]),
            ("A", ["--nist-costs"]), ("A", ["--nist-costs", "--per-utt", "--quiet"]),
            ("A", ["--costs", "1", "2", "3", "--per-utt", "--quiet"]),
            ("A", ["--costs", "2", "2", "2"]), ("A", ["--costs", "0.5", "1", "0.25", "--per-utt", "--quiet"]),
            ("A", ["--batch-size", "7", "--per-utt"]),
            ("A", ["--replace", "@replace_int.txt", "--per-utt"]),
            ("A", ["--ignore", "@ignore_int.txt", "--per-utt"]),
            ("A", ["--replace", "@replace_int.txt", "--ignore", "@ignore_int.txt"]),
            ("A", ["--id2token", "@id2token.txt", "--per-utt"]),
            ("A", ["--id2token", "@id2token_swapped.txt", "--swap", "--per-utt"]),
            ("A", ["--id2token", "@id2token.txt", "--replace", "@replace_tok.txt", "--ignore",
                   "@ignore_tok.txt", "--per-utt"]),
            ("A", ["--id2token", "@id2token.txt", "--replace", "@replace_tok.txt", "--ignore",
                   "@ignore_tok.txt"]),
            ("A", ["--id2token", "@id2token_short.txt"]),
            ("A", ["--replace", "@replace_bad.txt"]),
            ("Apre", ["--file-prefix", "tok_", "--file-suffix", ".tok", "--per-utt"]),
            ("Amiss", []), ("Amiss", ["--warn-missing", "--per-utt"]), ("Amiss", ["--warn-missing"]),
            ("B", ["--distances"]), ("B", ["--distances", "--per-utt"]), ("B", []), ("B", ["--per-utt"]),
            ("B", ["--ignore", "@ignore_int.txt", "--distances", "--per-utt"]),
            ("A/ref+B/hyp", ["--warn-missing"]), ("A/ref+A/hyp", ["--per-utt", "--nist-costs", "--quiet"]),
        ]
        for where, opts in cases:
            if "+" in where:
                dirs = [os.path.join(tmp, x) for x in where.split("+")]
            else:
                dirs = [os.path.join(tmp, where)]
            if len(dirs) == 1:  # the command's positionals are dir [hyp] [out]
                dirs.append(os.path.join(dirs[0], "hyp"))
                dirs[0] = os.path.join(dirs[0], "ref")
            rec = run(dirs, [os.path.join(tmp, o[1:]) if o.startswith("@") else o for o in opts])
            for key in ("message",):
                if key in rec:
                    rec[key] = rec[key].replace(tmp, "{tmp}")
            rec["warned_missing"] = [m.replace(tmp, "{tmp}") for m in rec["warned_missing"]]
            rec["where"], rec["opts"] = where, opts
            store["cases"].append(rec)
            print(where, opts, "->", rec.get("raises") or rec["out"][:40].replace("\n", " | "))
    with open(os.path.join(HERE, "scoring.json"), "w") as f:
        json.dump(store, f)
    print("cases:", len(store["cases"]), "bytes:", os.path.getsize(os.path.join(HERE, "scoring.json")))


if __name__ == "__main__":
    main()
