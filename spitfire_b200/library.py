"""
Structured tabulated-chemistry containers: `Dimension` and `Library`.

Mirror of the reference's `spitfire.chemistry.library` (reference: src/spitfire/chemistry/library.py:19-452): named
N-D property arrays on a tensor grid of named dimensions, slicing into sub-libraries, and pickle persistence with the
same state dictionary (`dimensions`, `dim_ordering`, `properties`, `extra_attributes`). A pickle names the class by
its module path, so a file written by one code base loads in the other only through a module alias
(`sys.modules['spitfire.chemistry.library'] = spitfire_b200.library`, and likewise for the mechanism class held in
`extra_attributes['mech_spec']`); the state dictionaries themselves are interchangeable. This is the output format of
the flamelet sweeps; it holds no numerics.
"""
import os
import pickle
import shutil
from copy import copy, deepcopy

import numpy as np


class Dimension(object):
    """A named independent variable of a structured library (library.py:19-90)"""

    def __init__(self, name, values, log_scaled=False):
        values = np.asarray(values)
        if not str(name).isidentifier():
            raise ValueError(f'Error in building Dimension "{name}", the name cannot contain hyphens or spaces '
                             f'(it must be a valid Python variable name, check with name.isidentifier())')
        if values.ndim != 1:
            raise ValueError(f'Error in building Dimension "{name}", the values object must be one-dimensional. '
                             f'Use the ravel() method to flatten your data.')
        if values.size != np.unique(values).size:
            raise ValueError(f'Error in building structured dimension "{name}", duplicate values were identified!')
        self._name = name
        self._values = np.copy(values)
        self._min = np.min(values)
        self._max = np.max(values)
        self._npts = values.size
        self._log_scaled = log_scaled

    def __str__(self):
        return f'Dimension "{self._name}" spanning [{self._min}, {self._max}] with {self._npts} points'

    def __repr__(self):
        return f'Spitfire Dimension(name="{self._name}", min={self._min}, max={self._max}, npts={self._npts})'

    name = property(lambda self: self._name)
    values = property(lambda self: self._values)
    min = property(lambda self: self._min)
    max = property(lambda self: self._max)
    npts = property(lambda self: self._npts)
    log_scaled = property(lambda self: self._log_scaled)

    def _get_dict_for_file_save(self):
        return {'name': self._name, 'values': self._values, 'log_scaled': self._log_scaled}


class LibraryIndexError(IndexError):
    pass


class Library(object):
    """Property arrays over the tensor product of `Dimension`s (library.py:97-452).

    `lib['name'] = array` sets a property, `lib['name']` reads it, `lib[:, 2:5]` slices every property into a new
    library (dimensionality preserved), `lib.<dim>_values` / `lib.<dim>_grid` expose the grid."""

    def __init__(self, *dimensions):
        dims = []
        for d in dimensions:
            if isinstance(d, Dimension):
                dims.append(Dimension(d.name, d.values, d.log_scaled))
            else:
                dims.append(Dimension(d[0], d[1], False if len(d) == 2 else d[2]))
        self._dims = {d.name: d for d in dims}
        self._dims_ordering = {i: d.name for i, d in enumerate(dims)}
        self._props = dict()
        self._extra_attributes = dict()
        if dims:
            self._set_grid()

    def _set_grid(self):
        ordered = self.dims
        grids = np.meshgrid(*[d.values for d in ordered], indexing='ij')
        self._grid_shape = grids[0].shape
        self._grid_size = grids[0].size
        for d, g in zip(ordered, grids):
            setattr(self, d.name, d.name)
            setattr(self, d.name + '_grid', np.copy(g))
            for attr in ('_name', '_values', '_min', '_max', '_npts', '_log_scaled'):
                setattr(self, d.name + attr, getattr(d, attr))

    # -- dimensions ------------------------------------------------------------------------------------------------
    @property
    def dims(self):
        return [self._dims[self._dims_ordering[i]] for i in sorted(self._dims_ordering)]

    @property
    def dim_names(self):
        return [d.name for d in self.dims]

    def dim(self, name):
        return self._dims[name]

    def scale_dimension(self, dim_name, multiplier):
        self.remap_dimension(dim_name, lambda x: multiplier * x)

    def remap_dimension(self, dim_name, mapping):
        if dim_name not in self._dims:
            raise KeyError(f'Invalid dimension name "{dim_name}" provided to remap_dimension() on library {self}.')
        self._dims[dim_name] = Dimension(dim_name, mapping(self._dims[dim_name].values))
        self._set_grid()

    # -- properties ------------------------------------------------------------------------------------------------
    size = property(lambda self: self._grid_size)
    shape = property(lambda self: self._grid_shape)
    extra_attributes = property(lambda self: self._extra_attributes)

    @property
    def props(self):
        return list(self._props.keys())

    def get_empty_dataset(self):
        return np.ndarray(self._grid_shape)

    def add_empty_property(self, name):
        self._props[name] = self.get_empty_dataset()

    def remove(self, *quantities):
        for q in quantities:
            self._props.pop(q)

    def __contains__(self, prop):
        return prop in self._props

    def __setitem__(self, quantity, values):
        if isinstance(values, np.ndarray):
            if values.shape != self._grid_shape:
                raise ValueError(f'The shape of the "{quantity}" array does not conform to that of the library. '
                                 f'Given shape = {values.shape}, grid shape = {self._grid_shape}')
            if quantity in self._props:
                self._props[quantity][:] = values
            else:
                self._props[quantity] = values.view()
        elif isinstance(values, (float, int)):
            if quantity not in self._props:
                self._props[quantity] = self.get_empty_dataset()
            self._props[quantity].fill(float(values))
        else:
            raise TypeError(f'In Library[arg] = values, values must be a np.ndarray or float, received {values}')

    def __getitem__(self, *slices):
        first = slices[0]
        if isinstance(first, str):
            return self._props[first]
        if isinstance(first, slice):
            if first == slice(None, None, None) and len(self._dims) > 1:
                slices = tuple([slice(None, None, None)] * len(self._dims))
        else:
            slices = first
        ordered = self.dims
        if len(slices) != len(ordered):
            raise LibraryIndexError(f'Library[...] slicing must be given the same number of arguments as there are '
                                    f'dimensions, you provided {len(slices)} slices to a Library of dimension '
                                    f'{len(ordered)}')
        new_dims = []
        for d, s in zip(ordered, slices):
            if not isinstance(s, (slice, int)):
                raise LibraryIndexError(f'Library[...] can either take a single string or standard Python slices, '
                                        f'you provided it {slices}')
            v = d.values[s]
            new_dims.append(Dimension(d.name, np.array([v]) if np.ndim(v) == 0 else v, d.log_scaled))
        sub = Library(*new_dims)
        for p in self._props:
            sub[p] = self._props[p][slices].reshape(sub.shape)
        sub._extra_attributes.update(self._extra_attributes)
        return sub

    def __str__(self):
        lines = [f'\nSpitfire Library with {len(self._dims)} dimensions and {len(self._props)} properties',
                 '-' * 42]
        lines += [f'{i + 1}. {d}' for i, d in enumerate(self.dims)]
        lines.append('-' * 42)
        lines += [f'{k:20}, min = {np.min(v)} max = {np.max(v)}' for k, v in self._props.items()]
        lines.append(f'Extra attributes: {self._extra_attributes}')
        lines.append('-' * 42 + '\n')
        return '\n'.join(lines)

    def __repr__(self):
        return (f'\nSpitfire Library(ndim={len(self._dims)}, nproperties={len(self._props)})\n' +
                '\n'.join(f'{i + 1}. {d}' for i, d in enumerate(self.dims)) +
                f'\nProperties: [{", ".join(self._props.keys())}]\nExtra attributes: {self._extra_attributes}')

    # -- persistence (pickle state identical to library.py:166-183) ---------------------------------------------------
    def __getstate__(self):
        return dict(dimensions={n: d._get_dict_for_file_save() for n, d in self._dims.items()},
                    dim_ordering=self._dims_ordering,
                    properties=self._props,
                    extra_attributes=self._extra_attributes)

    def __setstate__(self, state):
        ordered = [None] * len(state['dimensions'])
        for index, name in state['dim_ordering'].items():
            d = state['dimensions'][name]
            ordered[index] = Dimension(d['name'], d['values'], log_scaled=d.get('log_scaled', False))
        self.__init__(*ordered)
        for prop, arr in state['properties'].items():
            self[prop] = arr
        self._extra_attributes.update(state.get('extra_attributes', dict()))

    def save_to_file(self, file_name):
        with open(file_name, 'wb') as f:
            pickle.dump(self, f)

    @classmethod
    def load_from_file(cls, file_name):
        with open(file_name, 'rb') as f:
            data = pickle.load(f)
        if isinstance(data, dict):  # v1.0 files were pickled as plain dictionaries
            lib = Library()
            lib.__setstate__(data)
            return lib
        return data

    def save_to_text_directory(self, output_directory, ravel_order='F', format='%.14e'):
        """one text file per dimension / property plus a metadata file (library.py:185-243)"""
        if os.path.isdir(output_directory):
            shutil.rmtree(output_directory)
        os.mkdir(output_directory)
        with open(os.path.join(output_directory, 'metadata_independent_variables.txt'), 'w') as f:
            for d in self.dims:
                f.write(d.name + '\n')
                np.savetxt(os.path.join(output_directory, f'bulkdata_ivar_{d.name}.txt'), d.values, fmt=format)
        with open(os.path.join(output_directory, 'metadata_dependent_variables.txt'), 'w') as f:
            for p in self._props:
                f.write(p + '\n')
                np.savetxt(os.path.join(output_directory, f'bulkdata_dvar_{p.replace(" ", "_")}.txt'),
                           self._props[p].ravel(order=ravel_order), fmt=format)
        with open(os.path.join(output_directory, 'metadata_user_defined_attributes.txt'), 'w') as f:
            f.write(str(self._extra_attributes))

    # -- copies and reshapes -----------------------------------------------------------------------------------------
    def __copy__(self):
        new = Library(*self.dims)
        for p in self._props:
            new[p] = self._props[p]
        new._extra_attributes.update(self._extra_attributes)
        return new

    def __deepcopy__(self, *args, **kwargs):
        new = Library(*[Dimension(d.name, np.copy(d.values), d.log_scaled) for d in self.dims])
        for p in self._props:
            new[p] = np.copy(self._props[p])
        new._extra_attributes.update(self._extra_attributes)
        return new

    @classmethod
    def copy(cls, library):
        """shallow copy: new Library and Dimension objects that share the arrays (library.py:282-285)"""
        return copy(library)

    @classmethod
    def deepcopy(cls, library):
        """deep copy (library.py:287-290)"""
        return deepcopy(library)

    @classmethod
    def squeeze(cls, library):
        kept = [d for d in library.dims if d.values.size > 1]
        if not kept:
            return dict(properties={p: np.squeeze(library[p]) for p in library.props},
                        dimensions={d.name: (np.squeeze(d.values), d.log_scaled) for d in library.dims},
                        extra_attributes=library.extra_attributes)
        new = Library(*kept)
        for p in library.props:
            new[p] = np.squeeze(library[p])
        new._extra_attributes.update(library.extra_attributes)
        return new

    @classmethod
    def swapaxes(cls, library, idx1, idx2):
        names = library.dim_names
        if len(names) < 2:
            raise ValueError('Cannot perform swap on library with one dimension.')
        names[idx1], names[idx2] = names[idx2], names[idx1]
        new = Library(*[Dimension(n, library.dim(n).values, library.dim(n).log_scaled) for n in names])
        for p in library.props:
            new[p] = np.swapaxes(library[p], idx1, idx2)
        new._extra_attributes.update(library.extra_attributes)
        return new
