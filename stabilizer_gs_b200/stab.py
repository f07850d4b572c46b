"""Mirror of py/stab.py `StabilizerGroup` (what the reference scripts use): canonical form and
membership/sign queries run as CUDA kernels through sgs_group_canonicalize / sgs_group_energy."""
import ctypes as C

import numpy as np

from . import _capi
from .pauli import Pauli


class StabilizerGroup:
    def __init__(self, Ps: Pauli, canonicalized=False, device=0):
        self.device = device
        self.n = Ps.end
        self.Ps = Ps if canonicalized else self._canonicalize(Ps)

    @property
    def begin(self):
        return self.Ps.begin

    @property
    def end(self):
        return self.Ps.end

    def _canonicalize(self, Ps):                      # py/stab.py:45-51
        n, m = Ps.end, int(Ps.batch_size[0])
        W = (2 * n + 63) // 64
        rows = Ps.packed(n)
        buf = (C.c_uint64 * max(m * W, 1))(*[w for r in rows for w in r])
        sg = (C.c_int8 * max(m, 1))(*[int(s) for s in np.asarray(Ps.phase).reshape(-1)])
        rank = C.c_int()
        _capi.check(_capi.lib().sgs_group_canonicalize(self.device, n, m, buf, sg, C.byref(rank)))
        out = [[buf[i * W + k] for k in range(W)] for i in range(rank.value)]
        return Pauli.from_packed(out, [sg[i] for i in range(rank.value)], n, Ps.begin, Ps.end)

    def get_energy(self, P: Pauli):                   # py/stab.py:147-153: 0 if not in ±S, else the relative sign
        n = max(self.Ps.end, P.end)
        W = (2 * n + 63) // 64
        m = int(self.Ps.batch_size[0])
        rows = self.Ps.packed(n)
        buf = (C.c_uint64 * max(m * W, 1))(*[w for r in rows for w in r])
        sg = (C.c_int8 * max(m, 1))(*[int(s) for s in np.asarray(self.Ps.phase).reshape(-1)])
        single = len(P.batch_size) == 0
        Pb = P if not single else Pauli(P.z[None], P.x[None], P.begin, np.asarray(P.phase).reshape(1))
        prow = Pb.packed(n)
        B = len(prow)
        pbuf = (C.c_uint64 * (B * W))(*[w for r in prow for w in r])
        psg = (C.c_int8 * B)(*[int(s) for s in np.asarray(Pb.phase).reshape(-1)])
        out = (C.c_int8 * B)()
        _capi.check(_capi.lib().sgs_group_energy(self.device, n, m, buf, sg, B, pbuf, psg, out))
        return float(out[0]) if single else np.array([float(v) for v in out])

    def __repr__(self):
        return repr(self.Ps)
