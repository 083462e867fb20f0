"""
Cantera-free ingest of Cantera YAML mechanism files.

Produces exactly the tuple that the reference's `ChemicalMechanismSpec._extract_cantera_mechanism_data`
(reference: src/spitfire/chemistry/mechanism.py:388-541) builds from a `cantera.Solution`:
(element_mw_map, element_list, ref_temperature, ref_pressure, species_name_list, species_dict, reaction_list,
transport_model), with activation energies in J/kmol, pre-exponentials in kmol-m-s units and the reactions
stably sorted simple < three-body < Lindemann < Troe (mechanism.py:540).

Supported: `ideal-gas` phases; NASA7 / constant-cp / NASA9 species thermo; elementary, three-body and falloff
(Lindemann, Troe) reactions with optional `orders`; an optional top-level / per-section `units` block for
length, quantity, time and activation-energy. Anything else raises `MechanismIngestError` (no silent skips).
"""
import re

import numpy as np
import yaml

GAS_CONSTANT = 8314.46261815324  # J/kmol/K, the value Cantera >= 2.5 (CODATA 2018) hands to Griffon (mechanism.py:135)

# standard atomic weights as tabulated by Cantera 3.0 (kg/kmol); the subset exercised by the reference fixtures is
# cross-checked against the `element_mw_map` pickled inside the reference's gold files (tests/golden/...).
ELEMENT_WEIGHTS = {
    'H': 1.008, 'D': 2.0141017781, 'Tr': 3.0160492820, 'He': 4.002602, 'Li': 6.94, 'Be': 9.0121831, 'B': 10.81,
    'C': 12.011, 'N': 14.007, 'O': 15.999, 'F': 18.998403163, 'Ne': 20.1797, 'Na': 22.98976928, 'Mg': 24.305,
    'Al': 26.9815384, 'Si': 28.085, 'P': 30.973761998, 'S': 32.06, 'Cl': 35.45, 'Ar': 39.95, 'K': 39.0983,
    'Ca': 40.078, 'Ti': 47.867, 'Cr': 51.9961, 'Mn': 54.938043, 'Fe': 55.845, 'Ni': 58.6934, 'Cu': 63.546,
    'Zn': 65.38, 'Br': 79.904, 'Kr': 83.798, 'I': 126.90447, 'Xe': 131.293, 'U': 238.02891, 'E': 5.48579909065e-4,
}


class MechanismIngestError(Exception):
    pass


class _NoBoolLoader(yaml.SafeLoader):
    """YAML 1.1 resolves the species name `NO` (and `ON`, `YES`, ...) to booleans; Cantera files mean strings."""


_NoBoolLoader.yaml_implicit_resolvers = {
    k: [(tag, rx) for tag, rx in v if tag != 'tag:yaml.org,2002:bool']
    for k, v in yaml.SafeLoader.yaml_implicit_resolvers.items()}
# PyYAML does not resolve floats written like `1.0e+05` without a sign/dot variants; make the float regex Cantera-like
_NoBoolLoader.add_implicit_resolver(
    'tag:yaml.org,2002:float',
    re.compile(r'''^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)
                    |\.(?:nan|NaN|NAN))$''', re.X),
    list('-+0123456789.'))


def load_yaml(path):
    with open(path, 'r') as f:
        return yaml.load(f, Loader=_NoBoolLoader)


# ---- units -----------------------------------------------------------------------------------------------------
_EA_TO_J_PER_KMOL = {'J/kmol': 1.0, 'J/mol': 1.0e3, 'kJ/mol': 1.0e6, 'kJ/kmol': 1.0e3, 'cal/mol': 4184.0,
                     'kcal/mol': 4184.0e3, 'cal/kmol': 4.184, 'K': None, 'eV': 96485332.12331001}
_LENGTH_TO_M = {'m': 1.0, 'cm': 1.0e-2, 'mm': 1.0e-3}
_QUANTITY_TO_KMOL = {'kmol': 1.0, 'mol': 1.0e-3, 'molec': 1.0 / 6.02214076e26}
_TIME_TO_S = {'s': 1.0, 'ms': 1.0e-3, 'min': 60.0}


class _Units(object):
    def __init__(self, block=None, parent=None):
        self.length = parent.length if parent else 'm'
        self.quantity = parent.quantity if parent else 'kmol'
        self.time = parent.time if parent else 's'
        self.activation_energy = parent.activation_energy if parent else 'J/kmol'
        self.explicit_ea = parent.explicit_ea if parent else False
        if block:
            for k, v in block.items():
                if k == 'length':
                    self.length = v
                elif k == 'quantity':
                    self.quantity = v
                elif k == 'time':
                    self.time = v
                elif k == 'activation-energy':
                    self.activation_energy = v
                    self.explicit_ea = True
                elif k in ('mass', 'energy', 'pressure', 'temperature', 'current'):
                    if (k, v) not in (('mass', 'kg'), ('energy', 'J'), ('pressure', 'Pa'), ('temperature', 'K'),
                                      ('current', 'A')):
                        raise MechanismIngestError(f'unsupported default unit {k}: {v}')
            if not self.explicit_ea and ('quantity' in block or 'energy' in block):
                self.activation_energy = 'J/' + self.quantity
        for name, table in (('length', _LENGTH_TO_M), ('quantity', _QUANTITY_TO_KMOL), ('time', _TIME_TO_S)):
            if getattr(self, name) not in table:
                raise MechanismIngestError(f'unsupported {name} unit {getattr(self, name)}')

    def ea(self, value):
        """activation energy -> J/kmol, multiplying the way Cantera does (value * factor)"""
        if isinstance(value, str):
            parts = value.split()
            v, u = float(parts[0]), (parts[1] if len(parts) > 1 else self.activation_energy)
        else:
            v, u = float(value), self.activation_energy
        if u not in _EA_TO_J_PER_KMOL:
            raise MechanismIngestError(f'unsupported activation-energy unit {u}')
        if u == 'K':
            return v * GAS_CONSTANT
        return v * _EA_TO_J_PER_KMOL[u]

    def pre_exponential(self, value, concentration_order):
        """A in (length^3/quantity)^(order-1)/time -> (m^3/kmol)^(order-1)/s"""
        if isinstance(value, str):
            parts = value.split()
            if len(parts) > 1:
                raise MechanismIngestError(f'explicit units on a pre-exponential are not supported: {value}')
            value = parts[0]
        v = float(value)
        lf, qf, tf = _LENGTH_TO_M[self.length], _QUANTITY_TO_KMOL[self.quantity], _TIME_TO_S[self.time]
        if lf == 1.0 and qf == 1.0 and tf == 1.0:
            return v
        return v * (lf ** 3 / qf) ** (concentration_order - 1.0) / tf


def _quantity(value, allowed):
    """'298.15 K' -> 298.15 with a unit check"""
    if isinstance(value, str):
        parts = value.split()
        if len(parts) > 1 and parts[1] not in allowed:
            raise MechanismIngestError(f'unsupported unit in "{value}" (allowed {allowed})')
        return float(parts[0])
    return float(value)


# ---- reaction equations ----------------------------------------------------------------------------------------
_ARROWS = (' <=> ', ' => ', ' = ')


def _parse_side(side):
    """'2 OH + M' -> ({'OH': 2.0}, has_M, falloff_collider)"""
    stoich = dict()
    has_m = False
    collider = None
    m = re.search(r'\(\s*\+\s*([^\s)]+)\s*\)', side)
    if m:
        collider = m.group(1)
        side = side[:m.start()] + side[m.end():]
    for tok in side.split(' + '):
        tok = tok.strip()
        if not tok:
            continue
        if tok == 'M':
            has_m = True
            continue
        parts = tok.split()
        if len(parts) == 2:
            coeff, name = float(parts[0]), parts[1]
        elif len(parts) == 1:
            mm = re.match(r'^(\d+\.?\d*)([A-Za-z(].*)$', parts[0])
            if mm and False:
                coeff, name = float(mm.group(1)), mm.group(2)
            else:
                coeff, name = 1.0, parts[0]
        else:
            raise MechanismIngestError(f'cannot parse species token "{tok}"')
        stoich[name] = stoich.get(name, 0.0) + coeff
    return stoich, has_m, collider


def parse_equation(eq):
    for arrow in _ARROWS:
        if arrow in eq:
            lhs, rhs = eq.split(arrow)
            reversible = arrow != ' => '
            r, rm, rc = _parse_side(lhs)
            p, pm, pc = _parse_side(rhs)
            if rm != pm or rc != pc:
                raise MechanismIngestError(f'unbalanced third body in "{eq}"')
            return r, p, reversible, rm, rc
    raise MechanismIngestError(f'no reaction arrow in "{eq}"')


# ---- species -----------------------------------------------------------------------------------------------------
def _species_entry(sp):
    th = sp['thermo']
    model = th['model']
    atoms = {str(k): float(v) for k, v in sp['composition'].items()}
    if model == 'NASA7':
        tr = [_quantity(t, ('K',)) for t in th['temperature-ranges']]
        data = th['data']
        if len(tr) == 2 and len(data) == 1:
            tr = [tr[0], tr[1], tr[1]]
            data = [data[0], data[0]]
        if len(tr) != 3 or len(data) != 2 or len(data[0]) != 7 or len(data[1]) != 7:
            raise MechanismIngestError(f'NASA7 species {sp["name"]} needs one or two 7-coefficient ranges')
        cp = dict({'type': 'NASA7', 'Tmin': tr[0], 'Tmid': np.float64(tr[1]), 'Tmax': tr[2],
                   'low-coeffs': np.array([float(x) for x in data[0]]),
                   'high-coeffs': np.array([float(x) for x in data[1]])})
    elif model == 'constant-cp':
        cp = dict({'type': 'constant',
                   'Tmin': _quantity(th.get('T-min', 0.0), ('K',)),
                   'Tmax': _quantity(th.get('T-max', 1.0e30), ('K',)),
                   'T0': _quantity(th.get('T0', 298.15), ('K',)),
                   'h0': _quantity(th.get('h0', 0.0), ('J/kmol',)),
                   's0': _quantity(th.get('s0', 0.0), ('J/kmol/K',)),
                   'cp': _quantity(th.get('cp0', 0.0), ('J/kmol/K',))})
    elif model == 'NASA9':
        tr = [_quantity(t, ('K',)) for t in th['temperature-ranges']]
        data = th['data']
        if len(data) != len(tr) - 1:
            raise MechanismIngestError(f'NASA9 species {sp["name"]}: ranges and data do not match')
        coeffs = [float(len(data))]
        for k, d in enumerate(data):  # layout of cantera's Nasa9PolyMultiTempRegion.coeffs: [nzones, (Tlo, Thi, a0..a8)*]
            coeffs += [tr[k], tr[k + 1]] + [float(x) for x in d]
        cp = dict({'type': 'NASA9', 'Tmin': tr[0], 'Tmax': tr[-1], 'coeffs': np.array(coeffs)})
    else:
        raise MechanismIngestError(f'unsupported thermo model "{model}" for species {sp["name"]}')
    return dict({'atoms': atoms, 'heat-capacity': cp})


def _troe_roundtrip(x):
    """Cantera stores 1/T3 and 1/T1 and reports their reciprocals (TroeRate::getParameters)"""
    return 1.0 / (1.0 / x) if x != 0.0 else x


def extract_yaml_mechanism_data(path, group_name='gas'):
    doc = load_yaml(path)
    phases = doc.get('phases', [])
    phase = None
    for ph in phases:
        if ph.get('name') == group_name:
            phase = ph
    if phase is None:
        names = [ph.get('name') for ph in phases]
        raise MechanismIngestError(f'phase "{group_name}" not found in {path}; phases present: {names}')
    if phase.get('thermo') != 'ideal-gas':
        raise MechanismIngestError(f'phase "{group_name}" is "{phase.get("thermo")}", only ideal-gas is supported')
    top_units = _Units(doc.get('units'))

    elem_list = [str(e) for e in phase.get('elements', [])]
    for e in elem_list:
        if e not in ELEMENT_WEIGHTS:
            raise MechanismIngestError(f'no atomic weight known for element "{e}"')
    element_mw_map = {e: np.float64(ELEMENT_WEIGHTS[e]) for e in elem_list}

    species_section = {str(s['name']): s for s in doc.get('species', [])}
    spec_names = phase.get('species', 'all')
    if spec_names == 'all':
        spec_name_list = list(species_section.keys())
    else:
        spec_name_list = []
        for item in spec_names:
            if isinstance(item, dict):  # {species: [A, B]} form
                for sec, lst in item.items():
                    if sec != 'species':
                        raise MechanismIngestError(f'species from section "{sec}" are not supported')
                    spec_name_list += list(species_section.keys()) if lst == 'all' else [str(x) for x in lst]
            else:
                spec_name_list.append(str(item))
    spec_dict = dict()
    for s in spec_name_list:
        if s not in species_section:
            raise MechanismIngestError(f'species "{s}" of phase "{group_name}" has no definition')
        spec_dict[s] = _species_entry(species_section[s])

    kin = phase.get('kinetics', None)
    rxn_spec = phase.get('reactions', 'all' if kin else 'none')
    rxn_docs = []
    if kin is not None and rxn_spec != 'none':
        if rxn_spec in ('all', 'declared-species'):
            rxn_docs = [(r, top_units) for r in doc.get('reactions', [])]
        else:
            for item in rxn_spec:
                sec = item if isinstance(item, str) else list(item.keys())[0]
                rxn_docs += [(r, top_units) for r in doc.get(sec, [])]
    species_set = set(spec_name_list)

    reac_temporary_list = list()
    for rx, units in rxn_docs:
        if 'units' in rx:
            units = _Units(rx['units'], units)
        reactants, products, reversible, has_m, collider = parse_equation(rx['equation'])
        involved = set(reactants) | set(products)
        if not involved <= species_set:
            if rxn_spec == 'declared-species':
                continue
            raise MechanismIngestError(f'reaction "{rx["equation"]}" uses undeclared species {involved - species_set}')
        rtype = rx.get('type', 'three-body' if has_m else 'elementary')
        if rtype == 'elementary' and has_m:
            rtype = 'three-body'
        order = sum(reactants.values())
        orders = {str(k): float(v) for k, v in rx['orders'].items()} if 'orders' in rx else None
        if orders is not None:
            order = sum(orders.get(k, v) for k, v in reactants.items())

        def _eff():
            eff = {str(k): float(v) for k, v in rx.get('efficiencies', {}).items()}
            default = float(rx.get('default-efficiency', 1.0))
            if collider is not None and collider != 'M':
                eff, default = {collider: 1.0}, 0.0
            return {k: v for k, v in eff.items() if k in species_set}, default

        if rtype == 'elementary':
            rc = rx['rate-constant']
            d = dict({'type': 'simple', 'reversible': reversible, 'reactants': reactants, 'products': products,
                      'A': units.pre_exponential(rc['A'], order), 'b': float(rc['b']), 'Ea': units.ea(rc['Ea'])})
            key = 0
        elif rtype == 'three-body':
            rc = rx['rate-constant']
            eff, default = _eff()
            d = dict({'type': 'three-body', 'reversible': reversible, 'reactants': reactants, 'products': products,
                      'default-eff': default, 'efficiencies': eff,
                      'A': units.pre_exponential(rc['A'], order + 1.0), 'b': float(rc['b']),
                      'Ea': units.ea(rc['Ea'])})
            key = 1
        elif rtype == 'falloff':
            hi, lo = rx['high-P-rate-constant'], rx['low-P-rate-constant']
            eff, default = _eff()
            d = dict({'reversible': reversible, 'reactants': reactants, 'products': products,
                      'default-eff': default, 'efficiencies': eff,
                      'fwd-A': units.pre_exponential(hi['A'], order), 'fwd-b': float(hi['b']),
                      'fwd-Ea': units.ea(hi['Ea']),
                      'flf-A': units.pre_exponential(lo['A'], order + 1.0), 'flf-b': float(lo['b']),
                      'flf-Ea': units.ea(lo['Ea'])})
            if 'Troe' in rx:
                t = rx['Troe']
                params = [float(t['A']), _troe_roundtrip(float(t['T3'])), _troe_roundtrip(float(t['T1']))]
                params.append(float(t['T2']) if 'T2' in t else 0.0)
                d['type'] = 'Troe'
                d['Troe-params'] = np.array(params)
                key = 3
            elif 'SRI' in rx or 'Tsang' in rx:
                raise MechanismIngestError(f'falloff form of "{rx["equation"]}" is not supported by Griffon')
            else:
                d['type'] = 'Lindemann'
                key = 2
        else:
            raise MechanismIngestError(f'reaction type "{rtype}" of "{rx["equation"]}" is not supported by Griffon')
        if orders is not None:
            d['orders'] = orders
        reac_temporary_list.append((key, d))

    reac_list = [y[1] for y in sorted(reac_temporary_list, key=lambda x: x[0])]
    ref_temperature = 298.15
    ref_pressure = 101325.0  # cantera's OneAtm, the reference pressure of every NASA/const-cp species thermo
    transport_model = None
    return element_mw_map, elem_list, ref_temperature, ref_pressure, spec_name_list, spec_dict, reac_list, \
        transport_model
