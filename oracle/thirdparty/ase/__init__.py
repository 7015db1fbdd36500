"""ORACLE / TEST INFRASTRUCTURE — minimal stand-in for ase==3.20.0: just the Atom / Atoms / data / io.read(xyz) /
formula.Formula.count surface the reference hot path and its tests touch (molgym/spaces.py:33-41,55-74;
molgym/agents/covariant/tools.py:8-15; molgym/agents/internal/agent.py:95,175-176; molgym/tools/util.py:21-23)."""
import numpy as np

from . import data  # noqa: F401


class Atom:
    def __init__(self, symbol='X', position=(0.0, 0.0, 0.0)):
        if isinstance(symbol, (int, np.integer)):
            symbol = data.chemical_symbols[int(symbol)]
        self.symbol = symbol
        self.position = np.array(position, dtype=float)

    @property
    def number(self):
        return data.atomic_numbers[self.symbol]

    def __repr__(self):
        return f"Atom('{self.symbol}', {self.position.tolist()})"


class Atoms:
    def __init__(self, symbols=None, positions=None):
        self._symbols = []
        self._positions = np.zeros((0, 3), dtype=float)
        if symbols is not None:
            if isinstance(symbols, str):
                from .formula import Formula
                symbols = [s for s, c in Formula(symbols).count().items() for _ in range(c)]
            if positions is None:
                positions = np.zeros((len(symbols), 3))
            for s, p in zip(symbols, positions):
                self.append(Atom(s, p))

    def append(self, atom):
        self._symbols.append(atom.symbol)
        self._positions = np.concatenate([self._positions, np.asarray(atom.position, dtype=float)[None]], axis=0)

    def copy(self):
        new = Atoms()
        new._symbols = list(self._symbols)
        new._positions = self._positions.copy()
        return new

    @property
    def positions(self):
        return self._positions

    @positions.setter
    def positions(self, value):
        value = np.asarray(value, dtype=float)
        assert value.shape == self._positions.shape
        self._positions = value.copy()

    def get_positions(self):
        return self._positions.copy()

    @property
    def symbols(self):
        return list(self._symbols)

    @property
    def numbers(self):
        return np.array([data.atomic_numbers[s] for s in self._symbols], dtype=int)

    def get_atomic_numbers(self):
        return self.numbers

    def get_chemical_symbols(self):
        return list(self._symbols)

    def get_number_of_atoms(self):
        return len(self)

    def __len__(self):
        return len(self._symbols)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            n = len(self)
            if i < -n or i >= n:
                raise IndexError('Index out of range.')
            return Atom(self._symbols[i], self._positions[i])
        idx = np.arange(len(self))[i] if not (isinstance(i, np.ndarray) and i.dtype == bool) else np.nonzero(i)[0]
        new = Atoms()
        new._symbols = [self._symbols[k] for k in np.atleast_1d(idx)]
        new._positions = self._positions[np.atleast_1d(idx)].copy()
        return new

    def __repr__(self):
        return f'Atoms(symbols={"".join(self._symbols)!r})'
