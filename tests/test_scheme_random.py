"""Scheme compilers (artensor_b200/scheme.py) on random small tensor networks, without the
reference package: a stub with the attributes the compilers use of the reference's
ContractionTree (contraction_tree.py: `tree`, `all_tensors`, `tn.tensor_bonds`,
`tn.final_qubits`, `mark_rep_tensor`, `tree_order_dfs`; vertices with `left`, `right`, `sc`,
`contain_bonds`, `contain_tensors`, `rep_tensor`) over a random binary tree.  The compiled
scheme is executed by the numpy oracle and by the host emulation of the lowered plan, and
compared with one brute-force `np.einsum` over the whole network (per requested bitstring in
sparse mode)."""
import itertools

import numpy as np
import pytest
import torch

from artensor_b200 import scheme as S
from artensor_b200.backend import ContractionPlan, PlanOptions
from oracle import tn_oracle as O
import emulate


class Vertex:
    def __init__(self, tensors, left=None, right=None):
        self.contain_tensors = frozenset(tensors)
        self.left, self.right = left, right
        self.contain_bonds = set()
        self.sc = 0
        self.rep_tensor = None


class Net:
    def __init__(self, tensor_bonds, final_qubits):
        self.tensor_bonds = tensor_bonds
        self.final_qubits = final_qubits


class StubTree:
    """Random binary contraction tree over the tensors of `tensor_bonds` (bond -> appears in one
    tensor = open, in two = internal)."""

    def __init__(self, tensor_bonds, final_qubits, rng):
        self.tn = Net({t: list(b) for t, b in tensor_bonds.items()}, final_qubits)
        count = {}
        for bl in tensor_bonds.values():
            for b in bl:
                count[b] = count.get(b, 0) + 1
        nodes = [Vertex([t]) for t in tensor_bonds]
        for v in nodes:
            (t,) = v.contain_tensors
            v.contain_bonds = set(tensor_bonds[t])
            v.sc = len(v.contain_bonds)
        self.tree = {v.contain_tensors: v for v in nodes}
        while len(nodes) > 1:
            i, j = sorted(rng.choice(len(nodes), 2, replace=False))
            b, a = nodes.pop(j), nodes.pop(i)
            v = Vertex(a.contain_tensors | b.contain_tensors, a, b)
            inside = {}
            for t in v.contain_tensors:
                for x in tensor_bonds[t]:
                    inside[x] = inside.get(x, 0) + 1
            v.contain_bonds = {x for x, c in inside.items() if c < count[x] or count[x] == 1}
            v.sc = len(v.contain_bonds)
            self.tree[v.contain_tensors] = v
            nodes.append(v)
        self.all_tensors = nodes[0].contain_tensors

    def _post_order(self):
        out, stack = [], [(self.tree[self.all_tensors], False)]
        while stack:
            v, done = stack.pop()
            if done or not (v.left and v.right):
                out.append(v)
            else:
                stack += [(v, True), (v.right, False), (v.left, False)]
        return out

    def mark_rep_tensor(self):      # contraction_tree.py:305-314
        for v in self._post_order():
            if v.left and v.right:
                v.rep_tensor = v.left.rep_tensor if v.left.sc > v.right.sc else v.right.rep_tensor
            else:
                v.rep_tensor = min(v.contain_tensors)

    def tree_order_dfs(self):       # contraction_tree.py:334-357 (any children-first order is valid)
        self.mark_rep_tensor()
        order = []
        for v in self._post_order():
            if v.left and v.right:
                keep, other = (v.left, v.right) if v.rep_tensor == v.left.rep_tensor else (v.right, v.left)
                order.append((keep.rep_tensor, other.rep_tensor))
        return order


def random_network(rng, n_tensors, n_internal, n_open):
    """tensor_bonds of a connected random network: a random spanning tree of internal bonds plus
    extra internal bonds, plus open bonds; every bond has extent 2."""
    bonds = {t: [] for t in range(n_tensors)}
    name = iter(f"b{i}" for i in itertools.count())
    for t in range(1, n_tensors):
        u = int(rng.randint(0, t))
        b = next(name)
        bonds[t].append(b), bonds[u].append(b)
    for _ in range(n_internal):
        t, u = rng.choice(n_tensors, 2, replace=False)
        b = next(name)
        bonds[int(t)].append(b), bonds[int(u)].append(b)
    for _ in range(n_open):
        bonds[int(rng.randint(0, n_tensors))].append(next(name))
    for t in bonds:
        rng.shuffle(bonds[t])
    return bonds


def rnd(rng, shape):
    return torch.from_numpy((rng.randn(*shape) + 1j * rng.randn(*shape)).astype(np.complex64))


def brute_force(tensor_bonds, leaves, out_bonds):
    labels = {}
    for bl in tensor_bonds.values():
        for b in bl:
            labels.setdefault(b, len(labels))
    args = []
    for t, bl in tensor_bonds.items():
        args += [leaves[t].numpy().astype(np.complex128), [labels[b] for b in bl]]
    return np.einsum(*args, [labels[b] for b in out_bonds], optimize=True)


@pytest.mark.parametrize("seed", range(8))
def test_normal_scheme_on_random_networks(seed):
    rng = np.random.RandomState(seed)
    n = int(rng.randint(4, 9))
    tb = random_network(rng, n, int(rng.randint(1, 5)), int(rng.randint(1, 5)))
    leaves = {t: rnd(rng, (2,) * len(bl)) for t, bl in tb.items()}
    tree = StubTree(tb, [], rng)
    scheme, out_bonds = S.contraction_scheme(tree)
    assert len(scheme) == n - 1
    want = brute_force(tb, leaves, out_bonds)
    got = O.tensor_contraction(dict(leaves), scheme)
    scale = np.abs(want).max()
    assert np.abs(np.asarray(got) - want).max() < 2e-5 * scale
    plan = ContractionPlan(scheme, {t: tuple(v.shape) for t, v in leaves.items()}, False, build_native=False)
    emu = emulate.run_plan(plan, plan.pack_leaves(leaves).numpy(), [0]).reshape(want.shape)
    assert np.abs(emu - want).max() < 2e-5 * scale


@pytest.mark.parametrize("seed,sc_target", [(0, 30), (1, 30), (2, 4), (3, 5), (4, 3), (5, 6), (6, 4), (7, 2)])
def test_sparse_scheme_on_random_networks(seed, sc_target):
    """Closed networks whose last tensors are final-qubit leaves `[bit, bond]`; a random subset of
    bitstrings; small sc_target values force subset outer steps and chunked batched steps."""
    rng = np.random.RandomState(100 + seed)
    n_body = int(rng.randint(3, 7))
    n_fq = int(rng.randint(2, 6))
    tb = random_network(rng, n_body, int(rng.randint(1, 4)), 0)
    final_qubits = []
    for q in range(n_fq):                         # one final-qubit leaf per qubit, hanging off a body tensor
        t = n_body + q
        b = f"q{q}"
        tb[t] = [b]
        tb[int(rng.randint(0, n_body))].append(b)
        final_qubits.append(t)
    leaves = {t: rnd(rng, (2,) * len(bl)) for t, bl in tb.items() if t < n_body}
    for t in final_qubits:
        leaves[t] = rnd(rng, (2, 2))              # [bit value, in bond]  (tensor_network.py:143-145)
    all_bits = ["".join(map(str, v)) for v in itertools.product((0, 1), repeat=n_fq)]
    k = int(rng.randint(1, len(all_bits) + 1))
    bits = sorted(rng.choice(all_bits, k, replace=False).tolist())
    tree = StubTree(tb, list(final_qubits), rng)
    scheme, rest, ordered = S.contraction_scheme_sparse(tree, bits, sc_target=sc_target)
    assert rest == [] and sorted(ordered) == bits
    for st in scheme:                             # chunks cover every row once
        if len(st) == 5 and st[3] is None:
            sizes = [len(c) for c in st[2][0]]
            assert 0 not in sizes and sum(sizes) == st[4][0] and sizes == [len(c) for c in st[2][1]]
    # brute force: fix the final-qubit bit values of every requested bitstring
    want = {}
    body = {t: bl for t, bl in tb.items()}
    for s in bits:
        fixed = dict(leaves)
        for q, t in enumerate(final_qubits):
            fixed[t] = leaves[t][int(s[q])]
        want[s] = complex(brute_force(body, fixed, []))
    got = np.asarray(O.tensor_contraction_sparse(dict(leaves), scheme)).reshape(-1)
    scale = max(abs(v) for v in want.values())
    assert max(abs(got[i] - want[s]) for i, s in enumerate(ordered)) < 2e-5 * scale
    plan = ContractionPlan(scheme, {t: tuple(v.shape) for t, v in leaves.items()}, True, build_native=False,
                           options=PlanOptions(stem_min_elems=1 << 3))
    emu = emulate.run_plan(plan, plan.pack_leaves(leaves).numpy(), [0]).reshape(-1)
    assert max(abs(emu[i] - want[s]) for i, s in enumerate(ordered)) < 2e-5 * scale


@pytest.mark.parametrize("seed", range(4))
def test_sparse_scheme_single_and_repeated_bitstrings(seed):
    """Edge cases of the bitstring list: a single bitstring (every row mode has extent 1 after the
    first merge) and a list with repeats (the compiler works on the distinct strings, like
    `np.unique` in contraction.py:260, and reports them once)."""
    rng = np.random.RandomState(500 + seed)
    n_body, n_fq = 4, 4
    tb = random_network(rng, n_body, 2, 0)
    final_qubits = []
    for q in range(n_fq):
        t = n_body + q
        tb[t] = [f"q{q}"]
        tb[int(rng.randint(0, n_body))].append(f"q{q}")
        final_qubits.append(t)
    leaves = {t: rnd(rng, (2,) * len(bl)) for t, bl in tb.items() if t < n_body}
    for t in final_qubits:
        leaves[t] = rnd(rng, (2, 2))
    one = "".join(map(str, rng.randint(0, 2, n_fq)))
    other = "".join(map(str, rng.randint(0, 2, n_fq)))
    for bits in ([one], [one, other, one, one, other]):
        tree = StubTree(tb, list(final_qubits), np.random.RandomState(seed))
        scheme, rest, ordered = S.contraction_scheme_sparse(tree, bits, sc_target=3)
        assert rest == [] and ordered == sorted(set(bits))
        got = np.asarray(O.tensor_contraction_sparse(dict(leaves), scheme)).reshape(-1)
        assert got.shape == (len(ordered),)
        for i, s_ in enumerate(ordered):
            fixed = dict(leaves)
            for q, t in enumerate(final_qubits):
                fixed[t] = leaves[t][int(s_[q])]
            want = complex(brute_force(tb, fixed, []))
            assert abs(got[i] - want) < 2e-5 * max(abs(want), 1e-3)
        plan = ContractionPlan(scheme, {t: tuple(v.shape) for t, v in leaves.items()}, True, build_native=False)
        emu = emulate.run_plan(plan, plan.pack_leaves(leaves).numpy(), [0]).reshape(-1)
        assert np.abs(emu - got).max() < 1e-5 * max(np.abs(got).max(), 1e-3)
