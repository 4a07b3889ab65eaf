"""Result / SimulationResults with the API of pyphysim.simulations.results: the containers the error
counters land in.  Reference: simulations/results.py:128-786 (Result), :795-1624 (SimulationResults).
"""
import json
import os
import pickle
from collections.abc import Iterable

import numpy as np

from .parameters import SimulationParameters

__all__ = ['Result', 'SimulationResults', 'calc_confidence_interval', 'counters_to_results']

_CRITICAL = {50: 0.674, 60: 0.842, 70: 1.036, 80: 1.282, 90: 1.645, 95: 1.960, 98: 2.326, 99: 2.576,
             99.5: 2.807, 99.8: 3.090, 99.9: 3.291}


def calc_confidence_interval(mean, std, n, P=95.0):
    """util/misc.py:807-867: mean -/+ C(P) * std / sqrt(n)."""
    half = _CRITICAL[P] * std / np.sqrt(n)
    return mean - half, mean + half


_OWN, _REF = 'pyphysim_b200.simulations', 'pyphysim.simulations'


class _ReferencePathPickler(pickle._Pickler):
    """Pickle files carry class paths.  The reference writes `pyphysim.simulations.results.SimulationResults`
    (results.py:1454-1473, protocol 2); so does this pickler for the classes of this package, which keeps
    result files — final and partial — loadable by the reference's own `load_from_file` and
    `bin/combine_results.py`.  Pure-Python pickler: result files are a few KB."""

    def save_global(self, obj, name=None):
        mod = getattr(obj, '__module__', None) or ''
        if mod == _OWN or mod.startswith(_OWN + '.'):
            qual = getattr(obj, '__qualname__', None) or obj.__name__
            self.write(pickle.GLOBAL + (_REF + mod[len(_OWN):]).encode('utf-8') + b'\n' +
                       qual.encode('utf-8') + b'\n')
            self.memoize(obj)
            return
        super().save_global(obj, name)


class _ReferencePathUnpickler(pickle.Unpickler):
    """...and files written by the reference (or by the pickler above) resolve to this package's classes,
    whether or not the `pyphysim` alias package is importable."""

    def find_class(self, module, name):
        if module == _REF or module.startswith(_REF + '.'):
            module = _OWN + module[len(_REF):]
        return super().find_class(module, name)


class Result:
    """One named statistic with an update rule (results.py:128-786)."""

    (SUMTYPE, RATIOTYPE, MISCTYPE, CHOICETYPE) = range(4)
    _all_types_names = {0: "SUMTYPE", 1: "RATIOTYPE", 2: "MISCTYPE", 3: "CHOICETYPE"}

    def __init__(self, name, update_type_code, accumulate_values=False, choice_num=None):
        self.name = name
        self._update_type_code = update_type_code
        self._value = 0
        self._total = 0
        self._result_sum = 0.0
        self._result_squared_sum = 0.0
        self.num_updates = 0
        if update_type_code == Result.CHOICETYPE:
            if not isinstance(choice_num, int):
                raise RuntimeError("'choice_num' argument for the Result object must be "
                                   "an integer for the CHOICETYPE type.")
            self._value = np.zeros(choice_num, dtype=int)
        self._accumulate_values_bool = accumulate_values
        self._value_list = []
        self._total_list = []

    @staticmethod
    def create(name, update_type, value, total=0, accumulate_values=False):
        """results.py:227-330."""
        if update_type == Result.CHOICETYPE:
            if total == 0:
                raise RuntimeError("When creating a new Result of CHOICETYPE you must "
                                   "provide the 'total' as well as the 'value.")
            r = Result(name, update_type, accumulate_values, choice_num=total)
            r.update(value)
        else:
            r = Result(name, update_type, accumulate_values)
            r.update(value, total)
        return r

    accumulate_values_bool = property(lambda self: self._accumulate_values_bool)
    type_code = property(lambda self: self._update_type_code)
    type_name = property(lambda self: Result._all_types_names[self._update_type_code])

    def __repr__(self):
        if self._update_type_code == Result.RATIOTYPE:
            v, t = self._value, self._total
            return "Result -> {0}: {1}/{2} -> {3}".format(self.name, v, t, v / t if t != 0 else "NaN")
        return "Result -> {0}: {1}".format(self.name, self.get_result())

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, self.__class__):
            return False
        for att in ('name', '_update_type_code', '_total', '_accumulate_values_bool', '_value_list',
                    '_total_list', '_result_squared_sum', '_result_sum'):
            if getattr(self, att) != getattr(other, att):
                return False
        if self._update_type_code == Result.CHOICETYPE:
            return bool(np.array_equal(self._value, other._value))
        return self._value == other._value

    def __ne__(self, other):
        return not self.__eq__(other)

    def update(self, value, total=None):
        """results.py:469-581."""
        self.num_updates += 1
        code = self._update_type_code
        if code == Result.SUMTYPE:
            self._value += value
            self._result_sum += value
            self._result_squared_sum += value ** 2
        elif code == Result.RATIOTYPE:
            if total is None:
                raise ValueError("A 'p_value' and a 'p_total' are required when "
                                 "updating a Result object of the RATIOTYPE type.")
            self._value += value
            self._total += total
            ratio = value / total
            self._result_sum += ratio
            self._result_squared_sum += ratio ** 2
            if self._accumulate_values_bool:
                self._total_list.append(total)
        elif code == Result.MISCTYPE:
            self._value = value
        elif code == Result.CHOICETYPE:
            # (the reference asserts against np.int, which NumPy >= 1.24 no longer has)
            assert isinstance(value, (int, np.integer)), "Value for the CHOICETYPE must be an integer."
            self._value[value] += 1
            self._total += 1
        else:
            self.num_updates -= 1
            raise ValueError("Can't update a Result object of type '{0}'".format(code))
        if self._accumulate_values_bool:
            self._value_list.append(value)

    def merge(self, other):
        """results.py:583-623."""
        assert isinstance(other, self.__class__)
        assert self._update_type_code == other._update_type_code and self.name == other.name, \
            "Can only merge two objects with the same name and type"
        if self.accumulate_values_bool:
            assert other.accumulate_values_bool, \
                "The merged Result also must have been set to accumulate values."
            self._value_list.extend(other._value_list)
            self._total_list.extend(other._total_list)
        if self._update_type_code == Result.MISCTYPE:
            self.num_updates = other.num_updates
            self._value, self._total = other._value, other._total
            self._result_sum, self._result_squared_sum = other._result_sum, other._result_squared_sum
        else:
            self.num_updates += other.num_updates
            self._value += other._value
            self._total += other._total
            self._result_sum += other._result_sum
            self._result_squared_sum += other._result_squared_sum

    def get_result(self):
        if self.num_updates == 0:
            return "Nothing yet"
        if self._update_type_code in (Result.RATIOTYPE, Result.CHOICETYPE):
            return self._value / self._total
        return self._value

    def get_result_accumulated_values(self):
        return self._value_list

    def get_result_accumulated_totals(self):
        return self._total_list

    def get_result_mean(self):
        return self._result_sum / self.num_updates

    def get_result_var(self):
        return (self._result_squared_sum / self.num_updates) - self.get_result_mean() ** 2

    def get_confidence_interval(self, P=95.0):
        if self._update_type_code == Result.MISCTYPE:
            raise RuntimeError("Calling get_confidence_interval is not valid for the MISC update type.")
        return calc_confidence_interval(self.get_result_mean(), np.sqrt(self.get_result_var()),
                                        self.num_updates, P)

    # ---- persistence ---------------------------------------------------------------------------
    def to_dict(self):
        def plain(v):
            if isinstance(v, np.ndarray):
                return v.tolist()
            if isinstance(v, np.integer):
                return int(v)
            if isinstance(v, np.floating):
                return float(v)
            return v
        return dict(name=self.name, update_type_code=self._update_type_code, value=plain(self._value),
                    total=plain(self._total), result_sum=plain(self._result_sum),
                    result_squared_sum=plain(self._result_squared_sum), num_updates=self.num_updates,
                    accumulate_values_bool=self._accumulate_values_bool,
                    value_list=[plain(v) for v in self._value_list],
                    total_list=[plain(v) for v in self._total_list])

    @staticmethod
    def from_dict(d):
        if d['update_type_code'] == Result.CHOICETYPE and isinstance(d['value'], Iterable):
            r = Result(d['name'], d['update_type_code'], d['accumulate_values_bool'],
                       choice_num=len(d['value']))
            r._value = np.array(d['value'], dtype=int)
            r._total = d['total']
        else:
            r = Result(d['name'], d['update_type_code'], d['accumulate_values_bool'])
            r._value, r._total = d['value'], d['total']
        r._value_list, r._total_list = d['value_list'], d['total_list']
        r.num_updates = d['num_updates']
        r._result_sum, r._result_squared_sum = d['result_sum'], d['result_squared_sum']
        return r

    def to_json(self):
        return json.dumps(self.to_dict())

    @staticmethod
    def from_json(data):
        return Result.from_dict(json.loads(data))


def _array_label(v, filename_mode):
    """Compact text for an array inside a file name ("0_(5)_30" for an arithmetic progression,
    util/misc.py:911-960)."""
    a = np.asarray(v)
    if a.ndim == 1 and a.size >= 4:
        step = a[1] - a[0]
        if np.allclose(a[1:] - step, a[:-1]):
            step = int(step) if a.dtype.kind in 'iu' else round(float(step), 12)
            return ("{0}_({1})_{2}" if filename_mode else "{0}:{1}:{2}").format(a[0], step, a[-1])
    return ','.join(str(x) for x in a.reshape(-1))


class SimulationResults:
    """Named lists of Result objects, one entry per parameter variation (results.py:795-1624)."""

    def __init__(self):
        self._results = {}
        self._params = SimulationParameters()
        self.runned_reps = None
        self.original_filename = None
        self.current_rep = -1

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, self.__class__):
            return False
        if self._params != other._params or self.current_rep != other.current_rep:
            return False
        if np.any(np.asarray(self.runned_reps, dtype=object) != np.asarray(other.runned_reps, dtype=object)):
            return False
        if self._results.keys() != other._results.keys():
            return False
        return all(self[k] == other[k] for k in self._results if k != 'elapsed_time')

    def __ne__(self, other):
        return not self.__eq__(other)

    params = property(lambda self: self._params)

    def set_parameters(self, params):
        if not isinstance(params, SimulationParameters):
            raise ValueError('params must be a SimulationParameters object')
        self._params = params

    def __repr__(self):
        return "SimulationResults: {0}".format(sorted(self._results.keys()))

    def add_result(self, result):
        self._results[result.name] = [result]

    def add_new_result(self, name, update_type, value, total=0):
        self.add_result(Result.create(name, update_type, value, total))

    def append_result(self, result):
        if result.name in self._results:
            if self._results[result.name][0].type_code != result.type_code:
                raise ValueError("Can only append to results of the same type")
            self._results[result.name].append(result)
        else:
            self.add_result(result)

    def append_all_results(self, other):
        for results in other:
            for result in results:
                self.append_result(result)

    def merge_all_results(self, other):
        """Merge the LAST entry of every name; 'num_skipped_reps' is special-cased
        (results.py:1103-1159)."""
        if len(self) == 0:
            for name in other.get_result_names():
                self._results[name] = other[name]
            return
        for name in self.get_result_names():
            if name != 'num_skipped_reps':
                self._results[name][-1].merge(other[name][-1])
        if 'num_skipped_reps' in other.get_result_names():
            if 'num_skipped_reps' not in self.get_result_names():
                self.add_new_result('num_skipped_reps', Result.SUMTYPE, 0)
            self._results['num_skipped_reps'][-1].merge(other['num_skipped_reps'][-1])

    def get_result_names(self):
        return list(self._results.keys())

    def _select(self, result_name, fixed_params):
        if fixed_params:
            idx = set(np.atleast_1d(self.params.get_pack_indexes(fixed_params)).tolist())
            return [v for i, v in enumerate(self[result_name]) if i in idx]
        return list(self[result_name])

    def get_result_values_list(self, result_name, fixed_params=None):
        return [v.get_result() for v in self._select(result_name, fixed_params)]

    def get_result_values_confidence_intervals(self, result_name, P=95.0, fixed_params=None):
        return [v.get_confidence_interval(P) for v in self._select(result_name, fixed_params)]

    def __getitem__(self, key):
        return self._results[key]

    def __len__(self):
        return len(self._results)

    def __iter__(self):
        return iter(self._results.values())

    def get_filename_with_replaced_params(self, filename):
        """'{SNR}'-style placeholders are filled from the parameters (results.py:1386-1416)."""
        try:
            labels = {n: "[{0}]".format(_array_label(v, True)) if isinstance(v, np.ndarray) else v
                      for n, v in self.params.parameters.items()}
            return filename.format(**labels)
        except (KeyError, IndexError, ValueError):
            return filename

    # ---- persistence: pickle (protocol 2) or JSON, by extension --------------------------------
    def to_dict(self):
        return {'params': self._params.to_dict(), 'runned_reps': self.runned_reps,
                'original_filename': self.original_filename, 'current_rep': self.current_rep,
                'results': {n: [r.to_dict() for r in v] for n, v in self._results.items()}}

    @staticmethod
    def from_dict(d):
        sr = SimulationResults()
        sr._params = SimulationParameters.from_dict(d['params'])
        sr.runned_reps = d['runned_reps']
        sr.original_filename = d['original_filename']
        sr.current_rep = d.get('current_rep', -1)
        sr._results = {n: [Result.from_dict(r) for r in v] for n, v in d['results'].items()}
        return sr

    def to_json(self):
        return json.dumps(self.to_dict())

    @staticmethod
    def from_json(data):
        return SimulationResults.from_dict(json.loads(data))

    def save_to_file(self, filename):
        ext = os.path.splitext(filename)[-1]
        if ext == '':
            filename, ext = filename + '.pickle', '.pickle'
        self.original_filename = filename
        filename = self.get_filename_with_replaced_params(filename)
        if ext == '.json':
            with open(filename, 'w') as fh:
                fh.write(self.to_json())
        elif ext == '.pickle':
            with open(filename, 'wb') as fh:
                _ReferencePathPickler(fh, protocol=2).dump(self)
        else:
            raise KeyError(ext)
        return filename

    @staticmethod
    def load_from_file(filename):
        ext = os.path.splitext(filename)[-1]
        if ext == '':
            filename, ext = filename + '.pickle', '.pickle'
        if ext == '.json':
            with open(filename, 'r') as fh:
                return SimulationResults.from_json(fh.read())
        with open(filename, 'rb') as fh:
            obj = _ReferencePathUnpickler(fh).load()
        assert isinstance(obj, SimulationResults)
        return obj

    def to_dataframe(self):
        """results.py:1598-1615."""
        import pandas as pd
        data = {}
        all_params = self.params.get_unpacked_params_list()
        for name in self.params:
            data[name] = [a[name] for a in all_params]
        for res in self:
            data[res[0].name] = [r.get_result() for r in res]
        if self.runned_reps is not None:
            data['runned_reps'] = self.runned_reps
        return pd.DataFrame(data)


def counters_to_results(counters, extra=None):
    """The six Results every hot-path simulator of the reference returns
    (apps/awgn_modulators/simulate_psk.py:90-112), from the 4 device counters
    [symbol_errors, bit_errors, num_symbols, num_bits] of one fused-link call."""
    se, be, ns, nb = (int(v) for v in counters)
    sr = SimulationResults()
    sr.add_new_result("symbol_errors", Result.SUMTYPE, se)
    sr.add_new_result("num_symbols", Result.SUMTYPE, ns)
    sr.add_new_result("bit_errors", Result.SUMTYPE, be)
    sr.add_new_result("num_bits", Result.SUMTYPE, nb)
    sr.add_new_result("ber", Result.RATIOTYPE, be, nb)
    sr.add_new_result("ser", Result.RATIOTYPE, se, ns)
    for name, (typ, value, total) in (extra or {}).items():
        sr.add_new_result(name, typ, value, total)
    return sr
