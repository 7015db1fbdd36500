"""ORACLE / TEST INFRASTRUCTURE — ase.io.read for plain .xyz files only."""
from . import Atom, Atoms


def read(filename, index=None, format=None):
    frames = []
    if hasattr(filename, 'read'):
        lines = filename.read().splitlines()
    else:
        with open(filename) as f:
            lines = f.read().splitlines()
    pos = 0
    while pos < len(lines) and lines[pos].strip():
        n = int(lines[pos].split()[0])
        atoms = Atoms()
        for line in lines[pos + 2:pos + 2 + n]:
            parts = line.split()
            atoms.append(Atom(parts[0], [float(v) for v in parts[1:4]]))
        frames.append(atoms)
        pos += 2 + n
    if index is None:
        index = -1
    if isinstance(index, str):
        index = int(index) if index != ':' else slice(None)
    return frames[index]
