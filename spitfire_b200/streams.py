"""
Cantera-free gas streams.

`Stream` duck-types the subset of `cantera.Quantity` that the reference's chemistry layer uses on the host side of the
hot path (reference: src/spitfire/chemistry/mechanism.py:602-757, flamelet.py:375-586, reactors.py:208-262,
tabulation.py): `T`, `P`, `Y`, `X`, `mean_molecular_weight`, `enthalpy_mass`/`H`, `density`, `species_names`,
`n_atoms`, the `TPY/TPX/TP/HPY/HPX/HP` property pairs, `mass`/`moles`/`constant` and `a + b` mixing at constant
(H,P), (T,P) or (U,V), and `equilibrate('HP'|'TP')`.

Thermodynamics comes from the mechanism's Griffon object (NASA7 / constant-cp kernels), so a stream built here and the
kernels agree to rounding. Enthalpy inversions T(h, Y) use Newton's method on Griffon's enthalpy and cp.
"""
import numpy as np


def _parse_composition(spec, species_names):
    """'O2:1, N2:3.74' or 'O2:1 N2:3.74' or a dict or an array -> array over species"""
    n = len(species_names)
    if isinstance(spec, str):
        out = np.zeros(n)
        index = {s.upper(): i for i, s in enumerate(species_names)}
        for token in spec.replace(',', ' ').split():
            name, _, value = token.partition(':')
            key = name.strip().upper()
            if key not in index:
                raise ValueError(f'unknown species "{name}" in composition "{spec}"')
            out[index[key]] = float(value) if value else 1.0
        return out
    if isinstance(spec, dict):
        out = np.zeros(n)
        index = {s.upper(): i for i, s in enumerate(species_names)}
        for name, value in spec.items():
            out[index[name.upper()]] = float(value)
        return out
    out = np.array(spec, dtype=float).ravel()
    if out.size != n:
        raise ValueError('composition array has the wrong size')
    return out


class Stream(object):
    def __init__(self, mechanism, T=300., P=101325., Y=None, mass=1.):
        self._mech = mechanism
        self._T, self._P = float(T), float(P)
        n = mechanism.n_species
        self._Y = np.zeros(n) if Y is None else np.array(Y, dtype=float)
        if Y is None:
            self._Y[0] = 1.
        self.mass = float(mass)
        self.constant = 'UV'

    # -- basic state ------------------------------------------------------------------------------------------------
    @property
    def mechanism(self):
        return self._mech

    @property
    def species_names(self):
        return self._mech.species_names

    @property
    def n_species(self):
        return self._mech.n_species

    def n_atoms(self, species, element):
        return self._mech.n_atoms(species, element)

    def species_index(self, name):
        return self._mech.species_index(name)

    @property
    def T(self):
        return self._T

    @property
    def P(self):
        return self._P

    @property
    def Y(self):
        return self._Y

    @property
    def X(self):
        w = self._mech.molecular_weights
        moles = self._Y / w
        return moles / np.sum(moles)

    @property
    def mean_molecular_weight(self):
        return 1. / np.sum(self._Y / self._mech.molecular_weights)

    @property
    def moles(self):
        return self.mass / self.mean_molecular_weight

    @moles.setter
    def moles(self, n):
        self.mass = float(n) * self.mean_molecular_weight

    @property
    def density(self):
        return self._P * self.mean_molecular_weight / (self._mech.gas_constant * self._T)

    density_mass = density

    @property
    def enthalpy_mass(self):
        return self._mech.griffon.enthalpy_mix(self._T, np.ascontiguousarray(self._Y))

    @property
    def int_energy_mass(self):
        return self._mech.griffon.energy_mix(self._T, np.ascontiguousarray(self._Y))

    @property
    def cp_mass(self):
        return self._mech.griffon.cp_mix(self._T, np.ascontiguousarray(self._Y))

    @property
    def cv_mass(self):
        return self._mech.griffon.cv_mix(self._T, np.ascontiguousarray(self._Y))

    @property
    def H(self):
        """total enthalpy of the stream (J), as cantera.Quantity.H"""
        return self.mass * self.enthalpy_mass

    # -- setters ----------------------------------------------------------------------------------------------------
    def _set_Y(self, Y):
        Y = _parse_composition(Y, self.species_names)
        s = np.sum(Y)
        if s <= 0.:
            raise ValueError('composition sums to zero')
        self._Y = Y / s

    def _set_X(self, X):
        X = _parse_composition(X, self.species_names)
        Y = X * self._mech.molecular_weights
        self._Y = Y / np.sum(Y)

    def _set_T_from_h(self, h, T_guess=None):
        """Newton on h(T) = h with Griffon's cp (dh/dT = cp)"""
        g = self._mech.griffon
        y = np.ascontiguousarray(self._Y)
        T = self._T if T_guess is None else T_guess
        for _ in range(100):
            dT = (h - g.enthalpy_mix(T, y)) / g.cp_mix(T, y)
            dT = max(min(dT, 500.), -500.)
            T += dT
            if abs(dT) < 1.e-10 * max(T, 1.):
                break
        self._T = T

    def _set_T_from_u(self, u, T_guess=None):
        g = self._mech.griffon
        y = np.ascontiguousarray(self._Y)
        T = self._T if T_guess is None else T_guess
        for _ in range(100):
            dT = (u - g.energy_mix(T, y)) / g.cv_mix(T, y)
            dT = max(min(dT, 500.), -500.)
            T += dT
            if abs(dT) < 1.e-10 * max(T, 1.):
                break
        self._T = T

    TP = property(lambda self: (self._T, self._P))
    TPY = property(lambda self: (self._T, self._P, self._Y))
    TPX = property(lambda self: (self._T, self._P, self.X))
    HP = property(lambda self: (self.enthalpy_mass, self._P))
    HPY = property(lambda self: (self.enthalpy_mass, self._P, self._Y))
    HPX = property(lambda self: (self.enthalpy_mass, self._P, self.X))

    @TP.setter
    def TP(self, v):
        self._T, self._P = float(v[0]), float(v[1])

    @TPY.setter
    def TPY(self, v):
        self._T, self._P = float(v[0]), float(v[1])
        self._set_Y(v[2])

    @TPX.setter
    def TPX(self, v):
        self._T, self._P = float(v[0]), float(v[1])
        self._set_X(v[2])

    @HP.setter
    def HP(self, v):
        self._P = float(v[1])
        self._set_T_from_h(float(v[0]))

    @HPY.setter
    def HPY(self, v):
        self._P = float(v[1])
        self._set_Y(v[2])
        self._set_T_from_h(float(v[0]))

    @HPX.setter
    def HPX(self, v):
        self._P = float(v[1])
        self._set_X(v[2])
        self._set_T_from_h(float(v[0]))

    @Y.setter
    def Y(self, v):
        self._set_Y(v)

    @X.setter
    def X(self, v):
        self._set_X(v)

    def copy(self):
        s = Stream(self._mech, self._T, self._P, np.copy(self._Y), self.mass)
        s.constant = self.constant
        return s

    # -- mixing (cantera.Quantity.__add__) ---------------------------------------------------------------------------
    def __add__(self, other):
        if other == 0:  # so that sum([...]) works
            return self.copy()
        if self.constant != other.constant:
            raise ValueError('streams must be mixed at the same `constant` pair')
        mass = self.mass + other.mass
        Y = (self.mass * self._Y + other.mass * other._Y) / mass
        out = Stream(self._mech, self._T, self._P, Y, mass)
        out.constant = self.constant
        T_guess = (self.mass * self._T + other.mass * other._T) / mass
        if self.constant == 'HP':
            h = (self.mass * self.enthalpy_mass + other.mass * other.enthalpy_mass) / mass
            out._set_T_from_h(h, T_guess)
        elif self.constant == 'TP':
            out._T = self._T
        elif self.constant == 'UV':
            u = (self.mass * self.int_energy_mass + other.mass * other.int_energy_mass) / mass
            vol = self.mass / self.density + other.mass / other.density
            out._set_T_from_u(u, T_guess)
            out._P = mass / vol * self._mech.gas_constant * out._T / out.mean_molecular_weight
        else:
            raise ValueError(f'unsupported constant pair "{self.constant}"')
        return out

    __radd__ = __add__

    def equilibrate(self, XY='HP', **kwargs):
        """chemical equilibrium at constant (H,P) or (T,P) by element-potential Gibbs minimisation"""
        from spitfire_b200.equilibrium import equilibrate
        equilibrate(self, XY)

    def __repr__(self):
        major = np.argsort(self._Y)[::-1][:4]
        comp = ', '.join(f'{self.species_names[i]}:{self._Y[i]:.4g}' for i in major if self._Y[i] > 0)
        return f'Stream(T={self._T:.2f} K, P={self._P:.1f} Pa, Y=[{comp}], mass={self.mass:g})'
