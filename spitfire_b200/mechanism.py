"""
Chemical mechanism loading and stream mixing, Cantera-free.

Host-side mirror of the reference's `spitfire.chemistry.mechanism` (reference:
src/spitfire/chemistry/mechanism.py:52-757): same class name, constructor arguments, properties and methods,
same `mech_data` dictionary (so pickles interoperate), but the mechanism is read by `_yaml_ingest` instead of a
`cantera.Solution`, and streams are `spitfire_b200.streams.Stream` objects (duck-typing `cantera.Quantity`)
whose thermodynamics is evaluated by Griffon's NASA7 kernels.
"""
from numpy import sum, array, abs  # noqa: F401  (names kept as in the reference module)
import numpy as np

from spitfire_b200._yaml_ingest import extract_yaml_mechanism_data, GAS_CONSTANT, MechanismIngestError  # noqa: F401


def populate_griffon_mechanism_data(griffon, mech_data, ct_element_mw_map, elem_list, ref_temperature, ref_pressure,
                                    spec_name_list, spec_dict, reac_list, transport_model,
                                    gas_constant=GAS_CONSTANT):
    """Feed a Griffon-like object (anything with the 18 `mechanism_*` setters) and fill `mech_data`.

    Follows mechanism.py:122-258 call for call, so the same routine drives the product binding and (in tests) the
    oracle bindings."""
    griffon.mechanism_set_ref_pressure(ref_pressure)
    mech_data['ref_pressure'] = ref_pressure
    griffon.mechanism_set_ref_temperature(ref_temperature)
    mech_data['ref_temperature'] = ref_temperature
    griffon.mechanism_set_gas_constant(gas_constant)
    mech_data['gas_constant'] = gas_constant

    griffon.mechanism_set_element_mw_map(ct_element_mw_map)
    mech_data['element_mw_map'] = ct_element_mw_map

    for e in elem_list:
        griffon.mechanism_add_element(e)
        mech_data['elements'].append(e)

    for s in spec_name_list:
        griffon.mechanism_add_species(s, spec_dict[s]['atoms'])
        mech_data['species'][s] = dict()
        mech_data['species'][s]['atom_map'] = spec_dict[s]['atoms']

    griffon.mechanism_resize_heat_capacity_data()

    for s in spec_dict:
        cp = spec_dict[s]['heat-capacity']
        if cp['type'] == 'constant':
            griffon.mechanism_add_const_cp(s, cp['Tmin'], cp['Tmax'], cp['T0'], cp['h0'], cp['s0'], cp['cp'])
            mech_data['species'][s]['cp'] = ('constant', cp['Tmin'], cp['Tmax'], cp['T0'], cp['h0'], cp['s0'],
                                             cp['cp'])
        elif cp['type'] == 'NASA7':
            lo, hi = np.asarray(cp['low-coeffs']).tolist(), np.asarray(cp['high-coeffs']).tolist()
            griffon.mechanism_add_nasa7_cp(s, cp['Tmin'], cp['Tmid'], cp['Tmax'], lo, hi)
            mech_data['species'][s]['cp'] = ('NASA7', cp['Tmin'], cp['Tmid'], cp['Tmax'], lo, hi)
        elif cp['type'] == 'NASA9':
            c = np.asarray(cp['coeffs']).tolist()
            griffon.mechanism_add_nasa9_cp(s, cp['Tmin'], cp['Tmax'], c)
            mech_data['species'][s]['cp'] = ('NASA9', cp['Tmin'], cp['Tmax'], c)
        if transport_model is not None and 'transport-data' in spec_dict[s]:
            mech_data['species'][s]['transport-data'] = dict(spec_dict[s]['transport-data'])

    mech_data['transport-model'] = transport_model

    R = gas_constant
    for rx in reac_list:
        t = rx['type']
        special = 'orders' in rx
        if t == 'simple':
            if special:
                griffon.mechanism_add_reaction_simple_with_special_orders(
                    rx['reactants'], rx['products'], rx['reversible'], rx['A'], rx['b'], rx['Ea'] / R, rx['orders'])
                mech_data['reactions'].append(('simple-special', rx['reactants'], rx['products'], rx['reversible'],
                                               rx['A'], rx['b'], rx['Ea'], rx['orders']))
            else:
                griffon.mechanism_add_reaction_simple(
                    rx['reactants'], rx['products'], rx['reversible'], rx['A'], rx['b'], rx['Ea'] / R)
                mech_data['reactions'].append(('simple', rx['reactants'], rx['products'], rx['reversible'],
                                               rx['A'], rx['b'], rx['Ea']))
        elif t == 'three-body':
            if special:
                griffon.mechanism_add_reaction_three_body_with_special_orders(
                    rx['reactants'], rx['products'], rx['reversible'], rx['A'], rx['b'], rx['Ea'] / R,
                    rx['efficiencies'], rx['default-eff'], rx['orders'])
                mech_data['reactions'].append(('three-body-special', rx['reactants'], rx['products'],
                                               rx['reversible'], rx['A'], rx['b'], rx['Ea'], rx['efficiencies'],
                                               rx['default-eff'], rx['orders']))
            else:
                griffon.mechanism_add_reaction_three_body(
                    rx['reactants'], rx['products'], rx['reversible'], rx['A'], rx['b'], rx['Ea'] / R,
                    rx['efficiencies'], rx['default-eff'])
                mech_data['reactions'].append(('three-body', rx['reactants'], rx['products'], rx['reversible'],
                                               rx['A'], rx['b'], rx['Ea'], rx['efficiencies'], rx['default-eff']))
        elif t == 'Lindemann':
            if special:
                griffon.mechanism_add_reaction_Lindemann_with_special_orders(
                    rx['reactants'], rx['products'], rx['reversible'], rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'] / R,
                    rx['efficiencies'], rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea'] / R, rx['orders'])
                mech_data['reactions'].append(('Lindemann-special', rx['reactants'], rx['products'],
                                               rx['reversible'], rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'],
                                               rx['efficiencies'], rx['default-eff'], rx['flf-A'], rx['flf-b'],
                                               rx['flf-Ea'], rx['orders']))
            else:
                griffon.mechanism_add_reaction_Lindemann(
                    rx['reactants'], rx['products'], rx['reversible'], rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'] / R,
                    rx['efficiencies'], rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea'] / R)
                mech_data['reactions'].append(('Lindemann', rx['reactants'], rx['products'], rx['reversible'],
                                               rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'], rx['efficiencies'],
                                               rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea']))
        elif t == 'Troe':
            troe = np.asarray(rx['Troe-params']).tolist()
            if special:
                griffon.mechanism_add_reaction_Troe_with_special_orders(
                    rx['reactants'], rx['products'], rx['reversible'], rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'] / R,
                    rx['efficiencies'], rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea'] / R, troe,
                    rx['orders'])
                mech_data['reactions'].append(('Troe-special', rx['reactants'], rx['products'], rx['reversible'],
                                               rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'], rx['efficiencies'],
                                               rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea'], troe,
                                               rx['orders']))
            else:
                griffon.mechanism_add_reaction_Troe(
                    rx['reactants'], rx['products'], rx['reversible'], rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'] / R,
                    rx['efficiencies'], rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea'] / R, troe)
                mech_data['reactions'].append(('Troe', rx['reactants'], rx['products'], rx['reversible'],
                                               rx['fwd-A'], rx['fwd-b'], rx['fwd-Ea'], rx['efficiencies'],
                                               rx['default-eff'], rx['flf-A'], rx['flf-b'], rx['flf-Ea'], troe))
        else:
            raise ValueError(f'unknown reaction type {t}')


def mech_data_to_extracted(gdata):
    """Invert `mech_data` (the pickled form, mechanism.py:108-115) back into the extraction tuple.

    The reference rebuilds a cantera.Solution from it (`_build_cantera_solution`, mechanism.py:260-373); no
    Cantera is needed here because `mech_data` already holds everything Griffon is fed."""
    spec_name_list = list(gdata['species'].keys())
    spec_dict = dict()
    for s in spec_name_list:
        spec = gdata['species'][s]
        cp = spec['cp']
        if cp[0] == 'constant':
            hc = dict(zip(('Tmin', 'Tmax', 'T0', 'h0', 's0', 'cp'), cp[1:]))
            hc['type'] = 'constant'
        elif cp[0] == 'NASA7':
            hc = dict({'type': 'NASA7', 'Tmin': cp[1], 'Tmid': cp[2], 'Tmax': cp[3],
                       'low-coeffs': np.array(cp[4]), 'high-coeffs': np.array(cp[5])})
        elif cp[0] == 'NASA9':
            hc = dict({'type': 'NASA9', 'Tmin': cp[1], 'Tmax': cp[2], 'coeffs': np.array(cp[3])})
        else:
            raise ValueError(f'unknown cp type {cp[0]}')
        spec_dict[s] = dict({'atoms': spec['atom_map'], 'heat-capacity': hc})
        if 'transport-data' in spec:
            spec_dict[s]['transport-data'] = spec['transport-data']
    reac_list = list()
    for rxn in gdata['reactions']:
        t = rxn[0]
        base = t.replace('-special', '')
        d = dict({'type': base, 'reactants': rxn[1], 'products': rxn[2], 'reversible': rxn[3]})
        if base in ('simple', 'three-body'):
            d['A'], d['b'], d['Ea'] = rxn[4:7]
            if base == 'three-body':
                d['efficiencies'], d['default-eff'] = rxn[7], rxn[8]
        else:
            d['fwd-A'], d['fwd-b'], d['fwd-Ea'] = rxn[4:7]
            d['efficiencies'], d['default-eff'] = rxn[7], rxn[8]
            d['flf-A'], d['flf-b'], d['flf-Ea'] = rxn[9:12]
            if base == 'Troe':
                d['Troe-params'] = np.array(rxn[12])
        if 'special' in t:
            d['orders'] = rxn[-1]
        reac_list.append(d)
    return (gdata['element_mw_map'], list(gdata['elements']), gdata['ref_temperature'], gdata['ref_pressure'],
            spec_name_list, spec_dict, reac_list, gdata.get('transport-model', None)), gdata.get('gas_constant',
                                                                                                GAS_CONSTANT)


class ChemicalMechanismSpec(object):
    """A class that loads chemical mechanisms and mixes streams (mirror of mechanism.py:52-757, Cantera-free).

    **Constructor**: specify a chemical mechanism file in cantera YAML format

    Parameters
    ----------
    cantera_input : str
        a cantera YAML file describing the thermochemistry
    group_name : str
        the phase to use
    mech_data : dict
        (extension) a `mech_data` dictionary as pickled by the reference (`ChemicalMechanismSpec.mech_data`); when
        given, the file arguments are ignored. This is what `__setstate__` uses instead of rebuilding a
        cantera.Solution (mechanism.py:111-113).
    griffon_factory : callable
        (extension, used by tests) builds the Griffon-like object to populate; defaults to the CUDA
        `spitfire_b200.griffon.PyCombustionKernels`.
    """

    def __init__(self, cantera_input=None, group_name='gas', cantera_solution=None, cantera_xml=None, mech_data=None,
                 griffon_factory=None):
        if cantera_xml is not None:
            cantera_input = cantera_input if cantera_input is not None else cantera_xml
            print('Deprecation warning: the "cantera_xml" input argument to ChemicalMechanismSpec is deprecated and '
                  'will be removed.\nUse the "cantera_input" argument instead.')
        if cantera_solution is not None:
            raise NotImplementedError('spitfire_b200 reads mechanisms without Cantera; pass a YAML file or mech_data')
        self._mech_file_path = cantera_input if cantera_input is not None else 'cantera-input-not-given'
        self._group_name = group_name if group_name is not None else 'cantera-group-not-given'

        self._mech_data = dict()
        self._mech_data['ref_pressure'] = None
        self._mech_data['ref_temperature'] = None
        self._mech_data['elements'] = list()
        self._mech_data['species'] = dict()
        self._mech_data['reactions'] = list()
        self._mech_data['transport-model'] = None

        self._element_stoichiometry = {'O': -1.0, 'H': 0.5, 'C': 2.0, 'Al': 1.5, 'U': 1.0, 'Ar': 0.0, 'N': 0.0,
                                       'He': 0.0}
        if griffon_factory is None:
            from spitfire_b200.griffon import PyCombustionKernels
            griffon_factory = PyCombustionKernels
        self._griffon_factory = griffon_factory
        self._griffon = griffon_factory()
        if mech_data is not None:
            extracted, gas_constant = mech_data_to_extracted(mech_data)
        else:
            extracted, gas_constant = extract_yaml_mechanism_data(cantera_input, group_name), GAS_CONSTANT
        populate_griffon_mechanism_data(self._griffon, self._mech_data, *extracted, gas_constant=gas_constant)
        self._species_names = list(extracted[4])
        self._species_index = {s: i for i, s in enumerate(self._species_names)}
        self._element_names = list(extracted[1])
        self._atom_maps = [dict(extracted[5][s]['atoms']) for s in self._species_names]
        mw_map = extracted[0]
        self._molecular_weights = array([sum([mw_map[a] * n for a, n in sorted(am.items())])
                                         for am in self._atom_maps])
        self._gas_constant = gas_constant

    @property
    def mech_data(self):
        return self._mech_data

    @property
    def element_stoichiometry(self):
        return self._element_stoichiometry

    @element_stoichiometry.setter
    def element_stoichiometry(self, custom_stoichiometry):
        if custom_stoichiometry is not None:
            self._element_stoichiometry = custom_stoichiometry

    def __getstate__(self):
        return dict({'mech_data': self._mech_data, 'element_stoichiometry': self._element_stoichiometry})

    def __setstate__(self, state):
        self.__init__(mech_data=state['mech_data'])
        if 'element_stoichiometry' in state:
            self._element_stoichiometry = state['element_stoichiometry']

    @property
    def griffon(self):
        """Obtain this mechanism's griffon PyCombustionKernels object"""
        return self._griffon

    @property
    def mech_file_path(self):
        return self._mech_file_path

    @property
    def group_name(self):
        return self._group_name

    @property
    def n_species(self):
        return len(self._species_names)

    @property
    def n_reactions(self):
        return len(self._mech_data['reactions'])

    @property
    def species_names(self):
        return list(self._species_names)

    @property
    def element_names(self):
        return list(self._element_names)

    @property
    def gas_constant(self):
        return self._gas_constant

    def species_index(self, name):
        return self._species_index[name]

    @property
    def molecular_weights(self):
        return array(self._molecular_weights)

    def molecular_weight(self, ni):
        if isinstance(ni, str):
            return self._molecular_weights[self._species_index[ni]]
        elif isinstance(ni, (int, np.integer)):
            return self._molecular_weights[ni]
        else:
            raise TypeError('ChemicalMechanismSpec.molecular_weight(ni) takes a string or integer, given ' + str(ni))

    def n_atoms(self, species, element):
        k = species if isinstance(species, (int, np.integer)) else self._species_index[species]
        return self._atom_maps[k].get(element, 0.0)

    # ---- streams (mechanism.py:602-757), Cantera-free -------------------------------------------------------------------
    def stream(self, properties=None, values=None, stp_air=False):
        """Build a mixture of species with certain properties, e.g. stream('TPX', (300., 101325., 'O2:1, N2:3.76')),
        stream('TPY', ...), stream('HPY', ...), stream('X', 'H2:1') or stream(stp_air=True)
        (3.74 mol N2 per mol O2 at 300 K and one atmosphere, mechanism.py:618-622)."""
        from spitfire_b200.streams import Stream
        q = Stream(self)
        if stp_air:
            if properties is not None or values is not None:
                print('Warning in building a stream of air at standard conditions!'
                      'The properties and values arguments will be ignored because stp_air=True was set.')
            q.TPX = 300., 101325., 'o2:1 n2:3.74'
        else:
            if properties is None:
                raise ValueError('ChemicalMechanismSpec.stream() was called improperly.\n'
                                 'There are two ways to build streams:\n'
                                 ' 1)  stream(stp_air=True)\n'
                                 ' 2)  stream(properties, values), e.g. stream(\'X\', \'O2:1, N2:1\')\n'
                                 '     or stream(\'TPY\', (300., 101325., \'O2:1, N2:1\'))\n')
            if values is None:
                raise ValueError('ChemicalMechanismSpec.stream() expects two arguments '
                                 'if properties are set in the construction')
            if not hasattr(type(q), properties) or not isinstance(getattr(type(q), properties), property):
                raise ValueError(f'unsupported stream property set "{properties}"')
            setattr(q, properties, values)
        return q

    def copy_stream(self, stream):
        """Make a duplicate of a stream - use this to avoid inadvertently modifying a stream by reference."""
        from spitfire_b200.streams import Stream
        q = Stream(self)
        q.TPX = stream.TPX
        return q

    @staticmethod
    def mix_streams(streams, basis, constant='HP'):
        """Mix a number of streams by mass/mole and at constant HP (default), TP or UV (mechanism.py:651-677)"""
        q_list = []
        for stream, amount in streams:
            if basis == 'mass':
                stream.mass = amount
            elif basis == 'mole':
                stream.moles = amount
            stream.constant = constant
            q_list.append(stream)
        mix = q_list[0]
        for q in q_list[1:]:
            mix = mix + q
        return mix if len(q_list) > 1 else mix.copy()

    def _get_atoms_in_stream(self, stream, atom_names):
        atom_amounts = {atom: 0 for atom in atom_names}
        X = stream.X
        # (species in index order as in the reference; absent species add exact zeros and are skipped)
        for i in np.nonzero(X)[0].tolist():
            for atom in atom_names:
                atom_amounts[atom] += X[i] * self.n_atoms(i, atom)
        return atom_amounts

    def stoich_molar_fuel_to_oxy_ratio(self, fuel_stream, oxy_stream):
        """molar ratio of fuel to oxidizer at stoichiometric conditions (mechanism.py:696-713)"""
        present = self._element_names
        for atom in present:
            if atom not in self._element_stoichiometry:
                raise KeyError(f'Error computing stoichiometric fuel/oxidizer ratio. Atom "{atom}" is not present in '
                               f'the element stoichiometry map, {self._element_stoichiometry}. All atoms must be '
                               f'included in the stoichiometry map. Atoms present in your mechanism are {present}.')
        atom_names = [a for a in self._element_stoichiometry.keys() if a in present]
        fuel_atoms = self._get_atoms_in_stream(fuel_stream, atom_names)
        oxy_atoms = self._get_atoms_in_stream(oxy_stream, atom_names)
        return -sum([self._element_stoichiometry[a] * oxy_atoms[a] for a in atom_names]) / \
            sum([self._element_stoichiometry[a] * fuel_atoms[a] for a in atom_names])

    def stoich_mass_fuel_to_oxy_ratio(self, fuel_stream, oxy_stream):
        mf = fuel_stream.mean_molecular_weight
        mx = oxy_stream.mean_molecular_weight
        return mf / mx * self.stoich_molar_fuel_to_oxy_ratio(fuel_stream, oxy_stream)

    def stoich_mixture_fraction(self, fuel_stream, oxy_stream):
        """mixture fraction at stoichiometric conditions (mechanism.py:715-722)"""
        eta = self.stoich_mass_fuel_to_oxy_ratio(fuel_stream, oxy_stream)
        return eta / (1. + eta)

    def mix_for_equivalence_ratio(self, phi, fuel, oxy):
        """Mix a stream of fuel and oxidizer such that the mixture has a specified equivalence ratio (mole basis)"""
        from spitfire_b200.streams import Stream
        r_st = self.stoich_molar_fuel_to_oxy_ratio(fuel, oxy)
        X = phi * r_st * fuel.X + oxy.X
        q = Stream(self)
        q.TPX = oxy.T, oxy.P, X / np.sum(X)
        return q

    def mix_for_normalized_equivalence_ratio(self, normalized_phi, fuel, oxy):
        return self.mix_for_equivalence_ratio(normalized_phi / (1. - normalized_phi), fuel, oxy)
