"""Inputs of tests/golden/golden_long.json, rebuilt from (seed, kind, lengths) with the stateless generator of
SURVEY.md 8(d).  Shared by the generator (make_golden_long.py, oracle side), the GPU tests and bench.py's golden checks
(product side: pass `trialign_b200.synth`), so the GPU box needs no sequence files."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden_long.json")


def mutate(synth, seed, which, src):
    """B / C of a mutated-copy triple: src with 1/8 of its positions substituted.  Drawn from the same stateless
    generator through its public 2-bit symbols only (triple indices 1 and 3 decide where, 2 decides what)."""
    n = len(src)
    hit = (synth(seed, 1, which, n) == 0) & (synth(seed, 3, which, n) < 2)
    sub = synth(seed, 2, which, n)
    out = src.copy()
    out[hit] = sub[hit]
    return out


def build_case(synth, seed, kind, lens):
    """kind "synth": triple 0 of the seed (uniform iid symbols); "mutated": A = triple 0's A, B and C = mutated copies
    of it (exercises large scores: uniform triples score near 0)."""
    la, lb, lc = lens
    if kind == "synth":
        return tuple(synth(seed, 0, s, n) for s, n in zip(range(3), lens))
    if kind != "mutated":
        raise ValueError(kind)
    base = synth(seed, 0, 0, max(lens))
    return base[:la].copy(), mutate(synth, seed, 1, base)[:lb].copy(), mutate(synth, seed, 2, base)[:lc].copy()


def load():
    with open(PATH) as f:
        return json.load(f)["cases"]
