"""SimulationParameters with the API of pyphysim.simulations.parameters (host-side sweep bookkeeping:
named parameters, some of them marked to be "unpacked" into the cartesian product of their values).
Reference: simulations/parameters.py:113-1011.  Config-file loading needs configobj/validate, which
the reference also imports optionally; it raises if they are absent."""
import copy
import itertools
import json
import pickle
from collections.abc import Iterable

import numpy as np

__all__ = ['SimulationParameters', 'combine_simulation_parameters']


class SimulationParameters:
    def __init__(self):
        self.parameters = {}
        self._unpacked_parameters_set = set()
        self._unpack_index = -1              # >= 0: this object is one variation of _original_sim_params
        self._original_sim_params = None

    # ---- construction --------------------------------------------------------------------------
    @staticmethod
    def _create(params_dict, unpack_index=-1, original_sim_params=None):
        sp = SimulationParameters()
        sp.parameters = copy.deepcopy(params_dict)
        sp._unpack_index = unpack_index if unpack_index >= 0 else -1
        sp._original_sim_params = original_sim_params
        return sp

    @staticmethod
    def create(params_dict):
        """parameters.py:233-262."""
        return SimulationParameters._create(params_dict)

    def add(self, name, value):
        self.parameters[name] = value

    def remove(self, name):
        del self.parameters[name]
        self._unpacked_parameters_set.discard(name)

    def set_unpack_parameter(self, name, unpack_bool=True):
        """parameters.py:325-358."""
        if name not in self.parameters:
            raise ValueError("Unknown parameter: `{0}`".format(name))
        if not isinstance(self.parameters[name], Iterable):
            raise ValueError("Parameter {0} is not iterable".format(name))
        if unpack_bool is True:
            self._unpacked_parameters_set.add(name)
        else:
            self._unpacked_parameters_set.remove(name)

    # ---- access --------------------------------------------------------------------------------
    unpack_index = property(lambda self: self._unpack_index)

    @property
    def unpacked_parameters(self):
        return sorted(self._unpacked_parameters_set)

    @property
    def fixed_parameters(self):
        return [n for n in self.parameters if n not in self._unpacked_parameters_set]

    def __getitem__(self, name):
        return self.parameters[name]

    def __setitem__(self, key, value):
        self.parameters[key] = value

    def __len__(self):
        return len(self.parameters)

    def __iter__(self):
        return iter(self.parameters)

    def __repr__(self):
        items = ("'{0}{1}': {2}".format(n, '*' if n in self._unpacked_parameters_set else '', v)
                 for n, v in self.parameters.items())
        return '{%s}' % ', '.join(items)

    def __eq__(self, other):
        """parameters.py:392-424: 'rep_max' is ignored so a run can be continued with more reps."""
        if self is other:
            return True
        if not isinstance(other, self.__class__):
            return False
        if self._unpacked_parameters_set != other._unpacked_parameters_set:
            return False
        if set(self.parameters) != set(other.parameters):
            return False
        if self._unpack_index != other._unpack_index:
            return False
        return all(key == "rep_max" or not np.any(self.parameters[key] != other.parameters[key])
                   for key in self.parameters)

    def __ne__(self, other):
        return not self.__eq__(other)

    # ---- unpacking -----------------------------------------------------------------------------
    def get_num_unpacked_variations(self):
        """parameters.py:429-451."""
        if self._original_sim_params is not None:
            return self._original_sim_params.get_num_unpacked_variations()
        n = 1
        for name in self._unpacked_parameters_set:
            n *= len(self.parameters[name])
        return n

    def get_pack_indexes(self, fixed_params_dict=None):
        """Indexes (in the unpacked list) of the variations with the given fixed values
        (parameters.py:551-652)."""
        fixed = {} if fixed_params_dict is None else fixed_params_dict
        names = self.unpacked_parameters
        dims = [len(self.parameters[n]) for n in names]
        grid = np.arange(self.get_num_unpacked_variations()).reshape(dims)
        sel = tuple(list(self.parameters[n]).index(fixed[n]) if n in fixed else slice(None) for n in names)
        return grid[sel].flatten()

    def get_unpacked_params_list(self):
        """Cartesian product over the unpacked parameters in sorted-name order, last name varying
        fastest (parameters.py:654-754)."""
        if not self._unpacked_parameters_set:
            return [self]
        names = sorted(self._unpacked_parameters_set)
        regular = [n for n in self.parameters if n not in self._unpacked_parameters_set]
        out = []
        for i, comb in enumerate(itertools.product(*(self.parameters[n] for n in names))):
            d = dict(zip(names, comb))
            for n in regular:
                d[n] = self.parameters[n]
            out.append(SimulationParameters._create(d, i, self))
        return out

    # ---- persistence ---------------------------------------------------------------------------
    def save_to_pickled_file(self, filename):
        with open(filename, 'wb') as fh:
            pickle.dump(self, fh, protocol=2)

    @staticmethod
    def load_from_pickled_file(filename):
        with open(filename, 'rb') as fh:
            return pickle.load(fh)

    def to_dict(self):
        """The JSON schema of the reference (parameters.py:`_to_dict` + util/serialize.py:36-66): arrays as
        {data, dtype, _is_numpy_array, shape}, sets as {data, _is_set} — files stay readable by its
        `SimulationResults.load_from_file`."""
        def conv(v):
            if isinstance(v, np.ndarray):
                return {'data': v.tolist(), 'dtype': str(v.dtype), '_is_numpy_array': True, 'shape': list(v.shape)}
            if isinstance(v, (np.integer,)):
                return int(v)
            if isinstance(v, (np.floating,)):
                return float(v)
            return v
        return {'parameters': {k: conv(v) for k, v in self.parameters.items()},
                'unpacked_parameters_set': {'data': sorted(self._unpacked_parameters_set), '_is_set': True},
                'unpack_index': self._unpack_index,
                'original_sim_params': None if self._original_sim_params is None
                else self._original_sim_params.to_dict()}

    @staticmethod
    def from_dict(d):
        def conv(v):
            if isinstance(v, dict) and v.get('_is_numpy_array') is True:
                return np.array(v['data'])               # like json_numpy_or_set_obj_hook (serialize.py:87-91)
            if isinstance(v, dict) and v.get('_is_set') is True:
                return set(v['data'])
            if isinstance(v, dict) and '__ndarray__' in v:      # files written by pyphysim_b200 0.1.0
                return np.array(v['__ndarray__'], dtype=v['dtype'])
            return v
        sp = SimulationParameters()
        if not d:
            return sp
        sp.parameters = {k: conv(v) for k, v in d['parameters'].items()}
        sp._unpacked_parameters_set = set(conv(d['unpacked_parameters_set']))
        sp._unpack_index = d['unpack_index']
        if d.get('original_sim_params') is not None:
            sp._original_sim_params = SimulationParameters.from_dict(d['original_sim_params'])
        return sp

    def to_json(self):
        return json.dumps(self.to_dict())

    @staticmethod
    def from_json(data):
        return SimulationParameters.from_dict(json.loads(data))

    @staticmethod
    def load_from_config_file(filename, spec=None, save_parsed_file=False):
        """parameters.py:790-940 — needs the third-party configobj + validate packages."""
        try:
            import configobj  # noqa: F401
            import validate  # noqa: F401
        except ImportError as e:
            raise ImportError("load_from_config_file needs the 'configobj' and 'validate' packages "
                              "(as in the reference): %s" % e)
        raise NotImplementedError("config-file loading is outside the hot-path scope (SURVEY.md §2)")


def combine_simulation_parameters(params1, params2):
    """Union of two parameter sets that differ only in the values of their unpacked parameters
    (parameters.py:943-1011)."""
    if set(params1.parameters) != set(params2.parameters) or \
            params1._unpacked_parameters_set != params2._unpacked_parameters_set or \
            not params1._unpacked_parameters_set:
        raise RuntimeError("Both SimulationParameters objects must have the same parameters and the "
                           "same (non-empty) set of unpacked parameters")
    for name in params1.fixed_parameters:
        if np.any(params1[name] != params2[name]):
            raise RuntimeError("Fixed parameter '%s' differs" % name)
    out = SimulationParameters()
    for name in params1.fixed_parameters:
        out.add(name, params1[name])
    for name in params1.unpacked_parameters:
        out.add(name, np.array(sorted(set(params1[name]).union(params2[name]))))
        out.set_unpack_parameter(name)
    return out
