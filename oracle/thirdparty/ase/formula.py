"""ORACLE / TEST INFRASTRUCTURE — Formula('C3H5NO3').count() -> ordered {symbol: count}."""
import re


class Formula:
    def __init__(self, formula=''):
        self._formula = formula
        self._count = {}
        for symbol, num in re.findall(r'([A-Z][a-z]?)(\d*)', formula):
            self._count[symbol] = self._count.get(symbol, 0) + (int(num) if num else 1)

    def count(self):
        return dict(self._count)

    def __len__(self):
        return sum(self._count.values())

    def __str__(self):
        return self._formula
