"""Mirror of the reference's py/pauli.py `Pauli` (the part its three scripts touch): a batch of Pauli
strings on the window [begin, end) with signs.  A plain container — all algebra runs in libsgs.so."""
import numpy as np

_LETTER = {'I': (0, 0), 'X': (0, 1), 'Z': (1, 0), 'Y': (1, 1)}     # (z, x); array4 = 2z + x  (py/pauli.py:7-12)


class Pauli:
    def __init__(self, z, x, begin=0, phase=None):
        z = np.asarray(z, dtype=np.int64) % 2
        x = np.asarray(x, dtype=np.int64) % 2
        assert z.shape == x.shape
        self.z, self.x, self.begin = z, x, int(begin)
        self.length = z.shape[-1]
        self.end = self.begin + self.length
        self.batch_size = z.shape[:-1]
        self.phase = np.ones(self.batch_size, dtype=np.int64) if phase is None else np.asarray(phase, dtype=np.int64)

    @classmethod
    def from_string(cls, string, sites):                       # py/pauli.py:97-110
        begin, end = min(sites), max(sites) + 1
        z = np.zeros(end - begin, dtype=np.int64)
        x = np.zeros(end - begin, dtype=np.int64)
        for letter, site in zip(string, sites):
            if letter not in _LETTER:
                raise KeyError(letter)
            z[site - begin], x[site - begin] = _LETTER[letter]
        return cls(z, x, begin=begin)

    @property
    def array4(self):
        return self.z * 2 + self.x

    def copy(self):
        return Pauli(self.z.copy(), self.x.copy(), self.begin, self.phase.copy())

    def __mul__(self, c):                                      # py/pauli.py:89-94
        r = self.copy()
        r.phase = r.phase * np.asarray(c, dtype=np.int64)
        return r

    def __getitem__(self, key):
        return Pauli(self.z[key], self.x[key], self.begin, self.phase[key])

    def __len__(self):
        return self.batch_size[0]

    def range(self, begin, end, allow_truncation=False):       # py/pauli.py:133-150
        if not allow_truncation:
            assert begin <= self.begin and end >= self.end
        z = np.zeros(self.batch_size + (end - begin,), dtype=np.int64)
        x = np.zeros_like(z)
        lo, hi = max(begin, self.begin), min(end, self.end)
        if hi > lo:
            z[..., lo - begin:hi - begin] = self.z[..., lo - self.begin:hi - self.begin]
            x[..., lo - begin:hi - begin] = self.x[..., lo - self.begin:hi - self.begin]
        return Pauli(z, x, begin, self.phase)

    @classmethod
    def stack(cls, paulis):
        begin = min(p.begin for p in paulis)
        end = max(p.end for p in paulis)
        ps = [p.range(begin, end) for p in paulis]
        return cls(np.stack([p.z for p in ps]), np.stack([p.x for p in ps]), begin, np.stack([p.phase for p in ps]))

    def to_string(self):                                       # py/pauli.py:112-119
        title = f'qubit {self.begin}-{self.end - 1} size {self.batch_size} '
        if int(np.prod(self.batch_size)) == 0:
            return title
        a4 = self.array4.reshape(-1, self.length)
        ph = np.asarray(self.phase).reshape(-1)
        return title + ' '.join(('+' if ph[i] == 1 else '-') + ''.join('IXZY'[v] for v in a4[i]) for i in range(a4.shape[0]))

    __str__ = to_string
    __repr__ = to_string

    # ---- packed form for the C ABI: bit 2j = z_j, bit 2j+1 = x_j at absolute site j
    def packed(self, n):
        W = (2 * n + 63) // 64
        a4 = self.array4.reshape(-1, self.length)
        rows = []
        for r in a4:
            words = [0] * W
            for j, v in enumerate(r):
                zz, xx = int(v) >> 1, int(v) & 1
                b = 2 * (self.begin + j)
                words[b >> 6] |= (zz | (xx << 1)) << (b & 63)
            rows.append(words)
        return rows

    @classmethod
    def from_packed(cls, rows, signs, n, begin=0, end=None):
        end = n if end is None else end
        z = np.zeros((len(rows), end - begin), dtype=np.int64)
        x = np.zeros_like(z)
        for i, words in enumerate(rows):
            for j in range(begin, end):
                v = (words[(2 * j) >> 6] >> ((2 * j) & 63)) & 3
                z[i, j - begin], x[i, j - begin] = v & 1, v >> 1
        return cls(z, x, begin, np.asarray(signs, dtype=np.int64))
