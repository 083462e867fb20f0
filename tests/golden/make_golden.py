"""
Generates the committed fixtures under tests/golden/ from the reference tree. Run HERE (needs /root/reference):

    python tests/golden/make_golden.py

 mech/<name>.json      `mech_data` dictionaries (the form the reference pickles, mechanism.py:108-115) produced by
                       spitfire_b200._yaml_ingest from the reference's fixture mechanisms
                       (tests/test_mechanisms/*.yaml, old_xmls/*.yaml); h2-burke is asserted equal to the mech_data
                       embedded by Cantera in tests/tabulation/adiabatic_slfm/gold.pkl.
 gold_*.npz            arrays of the reference's gold pickles (closed reactors, adiabatic / non-adiabatic SLFM).
 ref_*.npz             outputs of the UNMODIFIED reference C++ (oracle/_ref) on seeded inputs, for mechanisms that have
                       no gold file in the reference (GRI-3.0), so the vectors travel to the GPU box.
"""
import glob
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'
TM = os.path.join(REF, 'tests', 'test_mechanisms')

from spitfire_b200.mechanism import extract_yaml_mechanism_data, populate_griffon_mechanism_data  # noqa: E402
from spitfire_b200._yaml_ingest import load_yaml  # noqa: E402


class _Recorder(object):
    """accepts every mechanism_* setter and records nothing (we only want the mech_data dict)"""

    def __getattr__(self, name):
        if name.startswith('mechanism_'):
            return lambda *a, **k: None
        raise AttributeError(name)


def plain(o):
    if isinstance(o, dict):
        return {str(k): plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [plain(v) for v in o]
    if isinstance(o, np.ndarray):
        return [plain(v) for v in o.tolist()]
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    return o


def new_md():
    return dict(ref_pressure=None, ref_temperature=None, elements=[], species={}, reactions=[])


class _Stub(object):
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.__dict__['state'] = s


class StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith('spitfire'):
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


def load_gold(path):
    with open(path, 'rb') as f:
        return StubUnpickler(f).load()


def main():
    os.makedirs(os.path.join(HERE, 'mech'), exist_ok=True)
    files = [(os.path.join(TM, 'h2-burke.yaml'), 'h2-burke', 'h2-burke'),
             (os.path.join(TM, 'methane-gri30.yaml'), 'methane-gri30', 'methane-gri30'),
             (os.path.join(TM, 'methane-lu30.yaml'), None, 'methane-lu30'),
             (os.path.join(TM, 'heptane-liu-hewson-chen-pitsch-highT.yaml'), None, 'heptane-liu'),
             (os.path.join(TM, 'reaction_test_mechanism.yaml'), None, 'reaction_test_mechanism')]
    files += [(f, None, 'old_xmls_' + os.path.basename(f)[:-5]) for f in sorted(glob.glob(os.path.join(TM, 'old_xmls', '*.yaml')))]
    for path, phase, name in files:
        if phase is None:
            phase = load_yaml(path)['phases'][0]['name']
        ex = extract_yaml_mechanism_data(path, phase)
        md = new_md()
        populate_griffon_mechanism_data(_Recorder(), md, *ex)
        with open(os.path.join(HERE, 'mech', name + '.json'), 'w') as f:
            json.dump(plain(md), f)
        print('mech', name, len(md['species']), len(md['reactions']))

    # h2-burke must equal the Cantera-derived mech_data pickled in the reference gold library
    lib = load_gold(os.path.join(REF, 'tests', 'tabulation', 'adiabatic_slfm', 'gold.pkl'))
    gold_md = lib.state['extra_attributes']['mech_spec'].state['mech_data']
    with open(os.path.join(HERE, 'mech', 'h2-burke.json')) as f:
        mine = json.load(f)
    mine.pop('transport-model', None)
    assert plain(gold_md) == mine, 'YAML ingest differs from the Cantera-derived mech_data of the gold file'
    print('h2-burke mech_data == gold.pkl mech_data')

    # gold libraries -> npz
    for case in ('adiabatic_slfm', 'nonadiabatic_defect_steady_slfm', 'nonadiabatic_defect_transient_slfm'):
        lib = load_gold(os.path.join(REF, 'tests', 'tabulation', case, 'gold.pkl'))
        st = lib.state
        out = {}
        for dname, d in st['dimensions'].items():
            out['dim_' + dname] = np.asarray(d.state['values'] if hasattr(d, 'state') else d['values'])
        for pname, arr in st['properties'].items():
            out['prop_' + pname] = np.asarray(arr)
        np.savez_compressed(os.path.join(HERE, 'gold_' + case + '.npz'), **out)
        print('gold', case, {k: v.shape for k, v in out.items() if k.startswith('dim_')})
    with open(os.path.join(REF, 'tests', 'reactor', 'closed_reactors', 'gold.pkl'), 'rb') as f:
        g = pickle.load(f)
    out = {}
    for k, v in g.items():
        key = '_'.join(k) if isinstance(k, tuple) else str(k)
        t, T, Y = v
        out[key + '__t'], out[key + '__T'], out[key + '__Y'] = np.asarray(t), np.asarray(T), np.asarray(Y)
    np.savez_compressed(os.path.join(HERE, 'gold_closed_reactors.npz'), **out)
    print('gold closed reactors', list(g.keys()))


if __name__ == '__main__':
    main()
