"""helpers shared by the tests: golden mechanisms, backend construction, error metrics"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from spitfire_b200.mechanism import ChemicalMechanismSpec  # noqa: E402


def golden_mech_names():
    return sorted(f[:-5] for f in os.listdir(os.path.join(GOLDEN, 'mech')) if f.endswith('.json'))


def load_mech_data(name):
    """`name#k` is the single-reaction mechanism made of reaction k of `name` -- how the reference's own rate tests use
    reaction_test_mechanism.yaml (tests/griffon/test_reaction_rates.py:127-180: one Solution per reaction)"""
    base, _, k = name.partition('#')
    with open(os.path.join(GOLDEN, 'mech', base + '.json')) as f:
        md = json.load(f)
    # JSON turns the reaction tuples into lists; that is all mech_data_to_extracted needs
    if k:
        md['reactions'] = [md['reactions'][int(k)]]
    return md


def single_reaction_cases(name='reaction_test_mechanism'):
    """the reference's 16 one-reaction cases (elementary / non-elementary orders below, across and above one / the five
    rate-constant forms, irreversible and reversible / third-body, Lindemann, Troe), test_reaction_rates.py:127-143"""
    return [f'{name}#{k}' for k in range(len(load_mech_data(name)['reactions']))]


def has_nasa9(md):
    return any(s['cp'][0] == 'NASA9' for s in md['species'].values())


def oracle_available(kind):
    from oracle import oracle
    return oracle.available(kind)


def build_mech(name, backend):
    """backend: 'gpu' (the product), 'port' or 'reference' (oracle/)"""
    md = load_mech_data(name)
    if backend == 'gpu':
        return ChemicalMechanismSpec(mech_data=md)
    from oracle.oracle import OracleKernels
    return ChemicalMechanismSpec(mech_data=md, griffon_factory=lambda: OracleKernels(backend))


def rel_err(a, ref, floor_scale=0.0):
    """max |a-ref| / (|ref| + floor), floor = floor_scale * max|ref| (per call)"""
    a, ref = np.asarray(a), np.asarray(ref)
    floor = floor_scale * np.max(np.abs(ref)) if ref.size else 0.0
    den = np.abs(ref) + floor
    with np.errstate(divide='ignore', invalid='ignore'):
        e = np.where(den > 0, np.abs(a - ref) / den, np.where(a == ref, 0.0, np.inf))
    return float(np.max(e)) if e.size else 0.0
