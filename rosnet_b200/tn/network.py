"""Tensor-network workload generators (index labels are ints; tensors are small host ndarrays).

``rand_equation`` restates the semantics of ``opt_einsum.helpers.rand_equation`` used by
``benchmark/bench_random_tn.py:49-56``; ``qft_amplitude_network`` follows the gate sequence of
``benchmark/bench_qft.py:47-59`` (H then controlled-phase... the script uses CZ) closed with
|0> kets and <0| bras; ``sycamore_amplitude_network`` builds the 53-qubit, m-cycle supremacy
circuit (BASELINE.json config 5; the reference's own script is a different 16x10 lattice).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


@dataclass
class TensorNetwork:
    inputs: List[Tuple[int, ...]]      # index labels of every tensor
    output: Tuple[int, ...]            # open indices, in output order
    sizes: Dict[int, int]              # extent of every index
    arrays: List[np.ndarray] = field(default_factory=list)  # optional data, aligned with inputs

    @property
    def n_tensors(self):
        return len(self.inputs)

    def shapes(self):
        return [tuple(self.sizes[i] for i in ix) for ix in self.inputs]

    def random_arrays(self, dtype, seed=0):
        rng = np.random.default_rng(seed)
        out = []
        for shp in self.shapes():
            x = rng.standard_normal(shp)
            if np.dtype(dtype).kind == "c":
                x = x + 1j * rng.standard_normal(shp)
            # keep magnitudes O(1) through long contraction chains
            out.append((x / np.sqrt(max(np.prod(shp), 1)) ** 0.5).astype(dtype))
        self.arrays = out
        return out

    def equation(self) -> str:
        """einsum string (only for networks small enough for the 52 letters)."""
        import string

        letters = string.ascii_lowercase + string.ascii_uppercase
        labels = sorted(self.sizes)
        if len(labels) > len(letters):
            raise ValueError("too many indices for an einsum string")
        m = {l: letters[i] for i, l in enumerate(labels)}
        return ",".join("".join(m[i] for i in ix) for ix in self.inputs) + "->" + "".join(m[i] for i in self.output)


def rand_equation(n: int, reg: int, n_out: int = 0, d_min: int = 2, d_max: int = 9, seed=None) -> TensorNetwork:
    """Random contraction: ``n`` tensors, ``n * reg // 2`` bonds placed on two random distinct
    tensors each plus ``n_out`` open indices; every tensor gets at least one index.  Extents are
    uniform in [d_min, d_max]."""
    rng = np.random.RandomState(seed)
    num_inds = n * reg // 2 + n_out
    sizes = {i: int(rng.randint(d_min, d_max + 1)) for i in range(num_inds)}
    inputs: List[List[int]] = [[] for _ in range(n)]
    output = []
    placements = []
    for i in range(num_inds):
        if i < n_out:
            output.append(i)
            placements.append(i)
        else:
            placements += [i, i]
    order = rng.permutation(len(placements))
    for pos, k in enumerate(order):
        ix = placements[k]
        if pos < n:
            where = pos  # make sure every tensor has an index
            if ix in inputs[where]:
                where = int(rng.randint(0, n))
        else:
            where = int(rng.randint(0, n))
        while ix in inputs[where]:
            where = int(rng.randint(0, n))
        inputs[where].append(ix)
    return TensorNetwork([tuple(ix) for ix in inputs], tuple(output), sizes)


# ---- circuits -------------------------------------------------------------------------------------
class _CircuitBuilder:
    def __init__(self, n_qubits, dtype):
        self.dtype = np.dtype(dtype)
        self.next = 0
        self.wire = []
        self.inputs, self.arrays, self.sizes = [], [], {}
        zero = np.array([1, 0], dtype=self.dtype)
        for _ in range(n_qubits):
            i = self._new()
            self.wire.append(i)
            self._add((i,), zero)

    def _new(self):
        i = self.next
        self.next += 1
        self.sizes[i] = 2
        return i

    def _add(self, ix, arr):
        self.inputs.append(tuple(ix))
        self.arrays.append(np.asarray(arr, dtype=self.dtype))

    def gate1(self, q, u):
        o = self._new()
        self._add((o, self.wire[q]), u)  # u[out, in]
        self.wire[q] = o

    def gate2(self, q0, q1, u):
        o0, o1 = self._new(), self._new()
        self._add((o0, o1, self.wire[q0], self.wire[q1]), np.asarray(u).reshape(2, 2, 2, 2))
        self.wire[q0], self.wire[q1] = o0, o1

    def close(self, bits=None):
        for q, w in enumerate(self.wire):
            b = 0 if bits is None else bits[q]
            bra = np.zeros(2, dtype=self.dtype)
            bra[b] = 1
            self._add((w,), bra)
        return TensorNetwork(self.inputs, (), self.sizes, self.arrays)


H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
CZ = np.diag([1, 1, 1, -1])


def qft_amplitude_network(n: int = 30, dtype="complex128") -> TensorNetwork:
    """<0...0| C |0...0> for the circuit of ``benchmark/bench_qft.py:49-54``:
    for i in 0..n-2: H(i); CZ(j, i) for all j > i.  (29 H and 435 CZ for n = 30.)"""
    c = _CircuitBuilder(n, dtype)
    for i in range(n - 1):
        c.gate1(i, H)
        for j in range(i + 1, n):
            c.gate2(j, i, CZ)
    return c.close()


def _sycamore_layout():
    """53 qubits on the 54-site diagonal grid (one site missing), couplers in four classes."""
    sites = [(r, c) for r in range(9) for c in range(6)]
    missing = (0, 5)
    sites = [s for s in sites if s != missing]
    index = {s: i for i, s in enumerate(sites)}
    couplers = {"A": [], "B": [], "C": [], "D": []}
    for (r, c) in sites:
        # rotated-lattice neighbours: (r+1, c) and (r+1, c +- 1 depending on row parity)
        for dr, dc, even_cls, odd_cls in ((1, 0, "A", "C"), (1, 1 if r % 2 else -1, "B", "D")):
            t = (r + dr, c + dc)
            if t in index:
                couplers[even_cls if r % 2 == 0 else odd_cls].append((index[(r, c)], index[t]))
    return len(sites), couplers


def sycamore_amplitude_network(cycles: int = 14, seed: int = 0, dtype="complex64") -> TensorNetwork:
    """Amplitude <0|C|0> of a Sycamore-style random circuit: 53 qubits, ``cycles`` cycles of
    (random single-qubit gate from {sqrt X, sqrt Y, sqrt W} on every qubit, not repeating the
    previous one, then fSim(pi/2, pi/6) on the couplers of the cycle's class, pattern ABCDCDAB),
    closing layer of single-qubit gates."""
    rng = np.random.default_rng(seed)
    nq, couplers = _sycamore_layout()
    sx = np.array([[1, -1j], [-1j, 1]]) / np.sqrt(2)
    sy = np.array([[1, -1], [1, 1]]) / np.sqrt(2)
    sw = np.array([[1, -np.sqrt(1j)], [np.sqrt(-1j), 1]]) / np.sqrt(2)
    singles = [sx, sy, sw]
    theta, phi = np.pi / 2, np.pi / 6
    fsim = np.array([[1, 0, 0, 0], [0, np.cos(theta), -1j * np.sin(theta), 0], [0, -1j * np.sin(theta), np.cos(theta), 0],
                     [0, 0, 0, np.exp(-1j * phi)]])
    pattern = "ABCDCDAB"
    c = _CircuitBuilder(nq, dtype)
    last = [-1] * nq
    for m in range(cycles):
        for q in range(nq):
            choices = [g for g in range(3) if g != last[q]]
            g = int(rng.choice(choices))
            last[q] = g
            c.gate1(q, singles[g])
        for (q0, q1) in couplers[pattern[m % 8]]:
            c.gate2(q0, q1, fsim)
    for q in range(nq):
        choices = [g for g in range(3) if g != last[q]]
        c.gate1(q, singles[int(rng.choice(choices))])
    return c.close()


def simplify(tn: TensorNetwork) -> TensorNetwork:
    """Absorb rank-1 and rank-2 tensors into a neighbour (host NumPy on 2x2-sized data; what
    quimb's ``full_simplify`` does before the reference's benchmarks contract,
    ``benchmark/bench_qft.py:66-72``).  Keeps the network's value and open indices."""
    inputs = [list(ix) for ix in tn.inputs]
    arrays = [np.asarray(a) for a in tn.arrays]
    out_set = set(tn.output)
    alive = [True] * len(inputs)
    changed = True
    while changed:
        changed = False
        owners: Dict[int, List[int]] = {}
        for t, ix in enumerate(inputs):
            if alive[t]:
                for i in ix:
                    owners.setdefault(i, []).append(t)
        for t, ix in enumerate(inputs):
            if not alive[t] or len(ix) > 2 or len(ix) == 0:
                continue
            partner = None
            for i in ix:
                if i in out_set:
                    continue
                others = [u for u in owners[i] if u != t and alive[u]]
                if len(others) == 1 and len(owners[i]) == 2:
                    partner = (others[0], i)
                    break
            if partner is None:
                continue
            u, i = partner
            shared = [j for j in ix if j in inputs[u] and j not in out_set and len(owners[j]) == 2]
            ax_t = [ix.index(j) for j in shared]
            ax_u = [inputs[u].index(j) for j in shared]
            new = np.tensordot(arrays[u], arrays[t], (ax_u, ax_t))
            new_ix = [j for j in inputs[u] if j not in shared] + [j for j in ix if j not in shared]
            if len(set(new_ix)) != len(new_ix):
                continue
            inputs[u], arrays[u] = new_ix, new
            alive[t] = False
            changed = True
            break
    keep = [t for t in range(len(inputs)) if alive[t]]
    used = {i for t in keep for i in inputs[t]} | out_set
    return TensorNetwork([tuple(inputs[t]) for t in keep], tn.output, {i: s for i, s in tn.sizes.items() if i in used},
                         [arrays[t] for t in keep])
