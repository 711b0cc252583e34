"""CPU gate for the dataflow Cholesky (csrc/chol_dag.cu): its claim order, through the test-support symbol
fgp_test_dag_order.  The kernel's CTAs claim tile tasks in number order and spin-wait for their inputs; that cannot
deadlock if and only if every input of a task is produced by a task with a smaller number (a waiting CTA then only
waits for tasks that are running or done, by induction over the numbers).  Checked exhaustively for a range of tile
counts, batch sizes, claim-group widths and both matrix orders -- plus that the order is a bijection onto the lower
triangle of tiles of every matrix."""
import ctypes as C
import itertools

import pytest

from fungp_b200 import _lib


def _order(T, B, G, mm):
    lib = _lib.load()
    n = B * T * (T + 1) // 2
    i, j, b = C.c_int(), C.c_int(), C.c_int()
    out = []
    for t in range(n):
        assert lib.fgp_test_dag_order(T, B, G, mm, t, C.byref(i), C.byref(j), C.byref(b)) == 0
        out.append((i.value, j.value, b.value))
    assert lib.fgp_test_dag_order(T, B, G, mm, n, C.byref(i), C.byref(j), C.byref(b)) == -1
    return out


@pytest.mark.parametrize("T,B,G,mm", [(T, B, G, mm) for T, B, mm in itertools.product((1, 2, 3, 4, 8, 12, 16), (1, 2, 5, 13), (0, 1))
                                      for G in (1, 2, 4) if T % G == 0])
def test_claim_order_is_a_topological_order(T, B, G, mm):
    order = _order(T, B, G, mm)
    num = {}
    for t, key in enumerate(order):
        i, j, b = key
        assert 0 <= j <= i < T and 0 <= b < B
        assert key not in num                     # bijection
        num[key] = t
    assert len(num) == B * T * (T + 1) // 2
    for (i, j, b), t in num.items():
        deps = [(i, k, b) for k in range(j)] + [(j, k, b) for k in range(j)]       # L(i, 0:j), L(j, 0:j)
        if i > j:
            deps.append((j, j, b))                                                  # W_j
        else:
            deps += [(k, k, b) for k in range(j)]                                   # the y entries of the columns before
        for d in deps:
            assert num[d] < t, (T, B, G, mm, (i, j, b), d)


def test_column_order_starts_every_column_with_its_diagonal_tiles():
    """G = 1: the diagonal tiles of all matrices open a column (the chain diagonal block -> solves -> next diagonal block
    is what a single factorisation waits for), then the matrices follow one after the other."""
    T, B = 6, 3
    order = _order(T, B, 1, 1)
    pos = 0
    for j in range(T):
        assert order[pos:pos + B] == [(j, j, b) for b in range(B)]
        pos += B
        for b in range(B):
            assert order[pos:pos + T - j - 1] == [(i, j, b) for i in range(j + 1, T)]
            pos += T - j - 1
